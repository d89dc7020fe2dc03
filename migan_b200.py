"""Import shim: the package directory is `mi-gan_b200/` (not a valid Python identifier), so
`import migan_b200` loads it from there under the name `migan_b200`."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "mi-gan_b200")
_spec = importlib.util.spec_from_file_location(
    "migan_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["migan_b200"] = _mod
_spec.loader.exec_module(_mod)
