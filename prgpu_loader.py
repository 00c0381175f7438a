"""Imports the package directory `pr-ompl_b200/` (a hyphen is not a valid module name) as
module `pr_ompl_b200`."""
import importlib.util
import os
import sys

_NAME = "pr_ompl_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pr-ompl_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(path, "__init__.py"),
                                                  submodule_search_locations=[path])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
