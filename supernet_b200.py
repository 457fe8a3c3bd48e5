"""Import shim: loads the package directory (whose name, fixed by the project layout, is not a valid
Python identifier) under the module name `supernet_b200`."""
import importlib.util
import os
import sys

_PKG = "super-net-bayesian-image-segmentation-with-uncertainty-propagation_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), _PKG)

if "supernet_b200" not in sys.modules or getattr(sys.modules["supernet_b200"], "__path__", None) is None:
    _spec = importlib.util.spec_from_file_location("supernet_b200", os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["supernet_b200"] = _mod
    _spec.loader.exec_module(_mod)
