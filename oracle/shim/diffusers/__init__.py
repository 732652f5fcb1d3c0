"""TEST INFRASTRUCTURE -- a minimal stand-in named `diffusers` so that the reference's own files under
/root/reference (models/*.py, utils/guidance.py) import VERBATIM in this container, where diffusers==0.27.2
(requirements.txt:5 of the reference) is not installed and cannot be fetched.  Only the symbols the hot-path files
import are provided; the arithmetic-bearing ones (ResnetBlock2D, TemporalConvLayer, Downsample2D, Upsample2D,
Timesteps, TimestepEmbedding) restate the published diffusers 0.27.2 semantics.  Nothing in
the product package imports this."""
__version__ = "0.27.2-shim"
