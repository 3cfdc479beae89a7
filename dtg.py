"""Import shim: the product package lives in ``domain-transfer-gan_b200/`` (not an importable
identifier), so ``import dtg`` registers it as the module ``dtg_b200`` and re-exports it."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "domain-transfer-gan_b200")

if "dtg_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location("dtg_b200", os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["dtg_b200"] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules["dtg_b200"]
