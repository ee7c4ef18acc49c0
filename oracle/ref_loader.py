"""Test infrastructure ONLY: import the UNMODIFIED reference prototype for differential checks.

Only usable in the development container, where the read-only reference tree is mounted at
/root/reference.  It does not exist on the GPU box, so nothing in `-m gpu` tests, smoke() or
bench.py may call this.  The reference imports `soundfile` and `matplotlib.pyplot` at module
scope (center_extraction.py:28-29); neither does arithmetic and neither is installed here, so
empty stand-in modules are registered before the import.  No reference source is copied.
"""
import importlib
import os
import sys
import types

REF_DIR = "/root/reference/python-prototype"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "center_extraction.py"))


def load_reference():
    """Return the reference `center_extraction` module, imported from where it lies."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at /root/reference")
    for name in ("soundfile", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location(
        "_upmix_reference_center_extraction", os.path.join(REF_DIR, "center_extraction.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
