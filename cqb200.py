"""Import shim: the package directory is named `sha2-on-cq-halo2_b200` (not a Python identifier), so this module loads
it under the importable alias `sha2_on_cq_halo2_b200` and re-exports it as `cqb200`."""
import importlib.util
import os
import sys

_NAME = "sha2_on_cq_halo2_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sha2-on-cq-halo2_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(pkg).items() if not k.startswith("__")})
