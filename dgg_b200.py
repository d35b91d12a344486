"""Import alias: the package directory is ``learning-adaptive-neighborhoods-for-gnns_b200`` (not a
valid Python identifier), so this module loads it under the name ``dgg_b200``."""
import importlib.util
import os
import sys

_real = os.path.join(os.path.dirname(os.path.abspath(__file__)), "learning-adaptive-neighborhoods-for-gnns_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
