"""Import helper: the product package lives in the directory `raytracer.jl_b200/` (the reference's name plus
`_b200`); a dot is not legal in a Python module name, so it is registered as `raytracer_jl_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "raytracer.jl_b200")
NAME = "raytracer_jl_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def load_build():
    spec = importlib.util.spec_from_file_location(NAME + "_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
