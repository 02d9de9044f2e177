"""Import shim: the product package lives in the directory ``tinyllama.cpp_b200`` (the name the task
fixes), which is not a valid Python identifier.  ``import gtb`` registers it as ``tinyllama_cpp_b200``."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "tinyllama.cpp_b200"
NAME = "tinyllama_cpp_b200"

if NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(NAME, PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[NAME]
