"""Import helper: the package directory is `scalable-ccd_b200` (hyphen), registered as the
module `scalable_ccd_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
NAME = "scalable_ccd_b200"


def load_package():
    if NAME in sys.modules:
        return sys.modules[NAME]
    pkg_dir = os.path.join(ROOT, "scalable-ccd_b200")
    spec = importlib.util.spec_from_file_location(
        NAME, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
