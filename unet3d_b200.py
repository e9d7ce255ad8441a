"""Import shim: the package directory name (3d-unet-renal-anatomy-extraction_b200) is not a valid
Python identifier, so it is loaded here under the importable name ``unet3d_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3d-unet-renal-anatomy-extraction_b200")
_spec = importlib.util.spec_from_file_location("unet3d_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["unet3d_b200"] = _mod
_spec.loader.exec_module(_mod)
