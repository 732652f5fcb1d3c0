"""lvd_b200 -- B200-native (sm_100a) denoising hot path of LLM-grounded Video Diffusion.

Only what the hot path needs lives here: `csrc/` (hand-written CUDA kernels + the C-ABI), `_lib` (ctypes
binding), `ops` (launch helpers) and the host-side mirror of the reference interface (`models/`, `utils/`).
Import as `lvd_b200`.
"""
__version__ = "0.1.0"
