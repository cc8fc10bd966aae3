"""Import helper: the product package directory is named after the reference
(`cuda-aho-corasick-wu-manber_b200/`), which is not a valid Python identifier, so it
is loaded here under the module name `cuda_aho_corasick_wu_manber_b200`."""
import importlib.util
import os
import sys

PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda-aho-corasick-wu-manber_b200")
MOD_NAME = "cuda_aho_corasick_wu_manber_b200"


def load():
    if MOD_NAME in sys.modules:
        return sys.modules[MOD_NAME]
    spec = importlib.util.spec_from_file_location(MOD_NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[MOD_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def submodule(name: str):
    load()
    return importlib.import_module(f"{MOD_NAME}.{name}")
