"""Import shim: `differentiable-renderer_b200/` is not a valid Python identifier,
so load it by path under the module name `differentiable_renderer_b200` and
re-export it.  `import drt_b200 as drt` is the supported spelling."""
import importlib.util as _ilu
import sys as _sys
from pathlib import Path as _Path

_root = _Path(__file__).resolve().parent / "differentiable-renderer_b200"
_name = "differentiable_renderer_b200"
if _name not in _sys.modules:
    _spec = _ilu.spec_from_file_location(_name, _root / "__init__.py",
                                         submodule_search_locations=[str(_root)])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_pkg = _sys.modules[_name]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
