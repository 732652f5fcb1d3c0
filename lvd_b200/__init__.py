"""Import alias: the product package lives in `llm-groundedvideodiffusion_b200/` (a directory name Python cannot
import directly because of the hyphen).  `import lvd_b200` maps onto that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "llm-groundedvideodiffusion_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
