"""Import the real reference scripts (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` is absent on the GPU box, so nothing in
the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  It is used by
``oracle/make_golden.py`` (golden-vector generator) and by the optional
``-m "not gpu"`` cross-checks that skip when the reference is not mounted.

The three scripts import ``tensorflow``, ``gdal`` and ``skimage`` at module level;
none is installed.  They are replaced by ``MagicMock`` so that the
pure-NumPy host functions run unmodified.  ``np.int`` (removed in NumPy >= 1.24)
is restored as an alias of ``int`` because the scripts use it as a dtype.
"""
import os
import sys
from unittest.mock import MagicMock

REFERENCE_DIR = os.environ.get("DRS_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "isprs_dilated_random.py"))


def load():
    """Return (isprs, contest, coffee) reference modules."""
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_DIR)
    import numpy as np
    import numpy.ma  # noqa: F401  (must be imported before np.int is patched)
    import scipy
    import scipy.ndimage  # noqa: F401  the scripts call scipy.ndimage.rotate
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import scipy.misc  # noqa: F401
    if not hasattr(np, "int"):
        np.int = int
    for m in ("tensorflow", "gdal", "skimage"):
        if m not in sys.modules:
            sys.modules[m] = MagicMock()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import isprs_dilated_random as isprs
        import contest_dilated_random as contest
        import coffee_dilated_random as coffee
    return isprs, contest, coffee
