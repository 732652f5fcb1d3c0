"""TEST INFRASTRUCTURE: import the reference's hot-path modules VERBATIM from /root/reference through the shim.
Works only where /root/reference exists (the builder container); never on the GPU box."""
import os
import sys
import types

REF_ROOT = os.environ.get("LVD_REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


_loaded = None


def stub_host_modules():
    """Empty stand-ins for the plotting / LLM-client imports of utils/parse.py and utils/llm.py (matplotlib, imageio,
    easydict, pyjson5): none carries arithmetic of the path."""
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    for name in ("matplotlib", "imageio", "easydict", "pyjson5"):
        try:
            __import__(name)
        except ImportError:
            pass
    mp = mod("matplotlib")
    mp.patches = mod("matplotlib.patches", Polygon=object)
    mp.collections = mod("matplotlib.collections", PatchCollection=object)
    mp.pyplot = mod("matplotlib.pyplot")
    mod("imageio", RETURN_BYTES="<bytes>")
    mod("easydict", EasyDict=dict)
    import json as _json
    mod("pyjson5", loads=_json.loads, Json5Exception=ValueError, Json5EOF=ValueError, Json5DecoderException=ValueError,
        Json5IllegalCharacter=ValueError)


def load():
    """Returns a namespace with the reference classes/functions (UNet3DConditionModel, latent_backward_guidance,
    guidance module, parse helpers)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch

    for p in (SHIM, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # utils/guidance.py hard-codes device="cuda" / .cuda(); on a CPU-only box redirect them
    if not torch.cuda.is_available():
        _orig_zeros = torch.zeros

        def zeros(*a, **k):
            if k.get("device") == "cuda":
                k["device"] = "cpu"
            return _orig_zeros(*a, **k)

        torch.zeros = zeros
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    # avoid executing models/__init__.py (pulls LMD leftovers); register the package path only
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF_ROOT, "models")]
    sys.modules["models"] = pkg
    import utils as ref_utils  # noqa: F401  (reference utils package)
    from models.unet_3d_condition import UNet3DConditionModel
    from models.attention_processor import AttnProcessor
    from utils import guidance as ref_guidance

    ns = types.SimpleNamespace()
    ns.UNet3DConditionModel = UNet3DConditionModel
    ns.AttnProcessor = AttnProcessor
    ns.guidance = ref_guidance
    ns.utils = ref_utils
    # latent_backward_guidance lives in models/pipelines.py whose imports drag LMD-only modules; load it in isolation
    import importlib.util

    mm = types.ModuleType("models.models")
    mm.process_input_embeddings = None
    mm.torch_device = "cpu"
    sys.modules["models.models"] = mm
    sched = types.ModuleType("utils.schedule")
    sys.modules.setdefault("utils.schedule", sched)
    setattr(ref_utils, "schedule", sys.modules["utils.schedule"])
    spec = importlib.util.spec_from_file_location("models.pipelines", os.path.join(REF_ROOT, "models", "pipelines.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["models.pipelines"] = mod
    spec.loader.exec_module(mod)
    ns.latent_backward_guidance = mod.latent_backward_guidance
    _loaded = ns
    return ns
