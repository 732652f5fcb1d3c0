"""Host-side mirror of the reference's `utils` pieces that sit on the hot path (utils/utils.py, utils/guidance.py)."""
from .utils import get_hw_from_attn_dim, scale_proportion  # noqa: F401
