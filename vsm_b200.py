"""Import shim: the package directory is named `visual-slam-pipeline_b200` (hyphens), which
Python cannot import by name.  `import vsm_b200` loads it under this module's name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "visual-slam-pipeline_b200")
_spec = importlib.util.spec_from_file_location("vsm_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vsm_b200"] = _mod
_spec.loader.exec_module(_mod)
