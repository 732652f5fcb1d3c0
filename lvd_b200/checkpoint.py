"""Local diffusers-format checkpoint directories (there is no hub access: names that are not local paths raise).
`<root>/<subfolder>/config.json` + one weight file (`*.safetensors` preferred, then `*.bin`)."""
import json
import os

import torch

_WEIGHT_NAMES = ("diffusion_pytorch_model.safetensors", "model.safetensors", "diffusion_pytorch_model.bin",
                 "pytorch_model.bin")


def load_component(root, subfolder=None):
    """-> (config dict, state_dict) of one pipeline component."""
    d = os.path.join(root, subfolder) if subfolder else root
    cfg_path = os.path.join(d, "config.json")
    if not os.path.isfile(cfg_path):
        raise OSError(f"{cfg_path} not found: lvd_b200 loads local checkpoints only (no hub access)")
    cfg = json.load(open(cfg_path))
    for name in _WEIGHT_NAMES:
        p = os.path.join(d, name)
        if os.path.isfile(p):
            if name.endswith(".safetensors"):
                from safetensors.torch import load_file
                return cfg, load_file(p)
            return cfg, torch.load(p, map_location="cpu", weights_only=True)
    raise OSError(f"no weight file ({', '.join(_WEIGHT_NAMES)}) under {d}")


def init_kwargs(cls, cfg):
    """The config entries that are constructor arguments of `cls`."""
    code = cls.__init__.__code__
    allowed = code.co_varnames[1:code.co_argcount]
    return {k: v for k, v in cfg.items() if k in allowed}
