"""CPU: the C-ABI shared library loads, exports exactly what include/drs.h declares, and its host-only entry
point (sliding-window visiting order) matches the reference-generated vectors.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "drs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drs_[a-z0-9_]+)\s*\(", src)) - {"drs_allreduce_fn"})


def test_library_exports_every_declared_symbol(drs):
    from drs_b200 import lib
    l = lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(l, s), "libdrs.so does not export %s" % s
    assert sorted(lib.EXPORTS) == syms, "lib.py prototypes and include/drs.h disagree"
    assert l.drs_version() >= 100


def test_config_struct_matches_header(drs):
    from drs_b200 import lib
    assert ctypes.sizeof(lib.Config) == 14 * 4


def test_no_cpu_fallback(drs):
    """Without a GPU every compute entry point must fail loudly (this container has none)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from drs_b200 import lib
    with pytest.raises(lib.DrsError, match="no CUDA device|no CPU fallback"):
        drs.Session("dilated_grsl", 4, 6)
    with pytest.raises(ValueError, match="Net type not identified"):
        drs.Session("resnet50", 4, 6)


def test_grid_positions_through_abi(golden, drs):
    variant = {0: "isprs", 1: "contest", 2: "coffee"}
    for n, (v, h, w, crop, batch) in enumerate(golden["grid_cases"]):
        pos = drs.grid_positions(int(h), int(w), int(crop), int(batch), variant[int(v)])
        assert np.array_equal(pos, golden["grid_%d_pos" % n]), (n, variant[int(v)])
    assert len(drs.grid_positions(2000, 2500, 25, 16)) == 166 * 208
    assert len(drs.grid_positions(6000, 6000, 65, 16)) == 187 * 187
    from drs_b200 import lib
    with pytest.raises(lib.DrsError, match="does not fit"):
        drs.grid_positions(20, 20, 25, 4)
