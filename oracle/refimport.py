"""Import the reference's own Python (read-only, /root/reference) in THIS container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py to produce the golden
vectors under tests/golden/ and by the few CPU tests that are skipped when the
reference tree is absent (it does not exist on the GPU box).

The reference imports third-party packages that are not installable offline
(xgboost, sklearn_crfsuite, matplotlib, seaborn, calibration, allel); SURVEY.md
Appendix C.1 verified that permissive stub modules make every `src.*` module
importable and that Base / slide_window / string_kernel / gnofix / postprocess run
unmodified on numpy 2.3.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GNOMIX_REFERENCE", "/root/reference")

_STUBS = [
    "xgboost", "sklearn_crfsuite", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
    "matplotlib.animation", "seaborn", "calibration", "allel",
]


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def import_reference():
    """Returns the reference's `src` package (with stubbed third-party imports)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__getattr__ = lambda attr, _n=name: _Dummy  # type: ignore[attr-defined]
        mod.__path__ = []  # behave like a package for dotted imports
        sys.modules[name] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the repo may ship its own `src` alias package; make sure the reference's wins here
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        f = getattr(sys.modules[k], "__file__", "") or ""
        if not f.startswith(REFERENCE_ROOT):
            del sys.modules[k]
    import src  # noqa: F401
    import src.Base.base  # noqa: F401
    import src.Base.models  # noqa: F401
    import src.Base.string_kernel  # noqa: F401
    import src.Smooth.utils  # noqa: F401
    import src.Smooth.smooth  # noqa: F401
    import src.Smooth.models  # noqa: F401
    import src.Gnofix.gnofix  # noqa: F401
    import src.postprocess  # noqa: F401
    return sys.modules["src"]
