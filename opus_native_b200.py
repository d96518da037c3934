"""Import shim: the package directory is named `opus-native_b200` (hyphen, project layout),
which Python cannot import by name.  `import opus_native_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "opus-native_b200")
_spec = importlib.util.spec_from_file_location(
    "opus_native_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["opus_native_b200"] = _mod
_spec.loader.exec_module(_mod)
