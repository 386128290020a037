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


def _schedule(lib, num_units, mt, grid):
    """[(unit, is_pair), ...] per CTA through drs_debug_conv_schedule (the kernel's own ConvSched, compiled for the host)."""
    import ctypes as C
    out = []
    for b in range(grid):
        n = C.c_int32()
        units = (C.c_int32 * 64)()
        pairs = (C.c_uint8 * 64)()
        assert lib.drs_debug_conv_schedule(num_units, mt, grid, b, units, pairs, 64, C.byref(n)) == 0
        assert n.value <= 64
        out.append([(units[i], bool(pairs[i])) for i in range(n.value)])
    return out


def test_conv_tile_schedule_is_a_partition(drs):
    """conv_tc's tile schedule (csrc/conv_tc.cuh: ConvSched) hands every 128-pixel unit to exactly one CTA, for single tiles and
    for pairs on one filter slice: full rounds of pairs in the round-robin sweep, then one mixed round in which no CTA gets
    more than a pair.  A unit that is skipped is a hole in a layer's output, one that is taken twice a wasted tile -- neither
    shows up as a numerical error bound, so the schedule is enumerated for every unit count around the round boundaries."""
    from drs_b200 import lib
    l = lib.load()
    for grid in (1, 2, 3, 7, 148, 296):
        counts = set(range(0, min(5 * grid + 3, 40))) | {grid - 1, grid, grid + 1, 2 * grid - 1, 2 * grid, 2 * grid + 1,
                                                         3 * grid, 3 * grid + 1, 4 * grid - 1, 4 * grid, 4 * grid + 1, 5 * grid + 2, 685, 5919}
        for U in sorted(c for c in counts if 0 <= c <= 64 * grid):
            for mt in (1, 2):
                sched = _schedule(l, U, mt, grid)
                seen = []
                for b, items in enumerate(sched):
                    for i, (u, pair) in enumerate(items):
                        assert 0 <= u and u + (2 if pair else 1) <= U, (grid, U, mt, b, i, u, pair)
                        assert pair is False or mt == 2
                        seen += [u, u + 1] if pair else [u]
                assert sorted(seen) == list(range(U)), (grid, U, mt)
                if mt == 2 and U:
                    per_cta = [sum(2 if p else 1 for _, p in items) for items in sched]
                    full_rounds = U // (2 * grid)
                    # the mixed last round gives a CTA at most one pair on top of the full rounds
                    assert max(per_cta) <= 2 * full_rounds + 2 and max(per_cta) - min(per_cta) <= 2, (grid, U, per_cta[:8])
                    # all full rounds are pairs, swept round-robin: CTA b's i-th pair starts at unit 2 * (i * grid + b)
                    for b, items in enumerate(sched):
                        assert items[:full_rounds] == [(2 * (i * grid + b), True) for i in range(full_rounds)]
                if mt == 1:
                    for b, items in enumerate(sched):
                        assert items == [(b + i * grid, False) for i in range(len(items))]
    # batch 64, crop 37 on 148 SMs: 685 units = 2 rounds of pairs + 93 singles (not 3 rounds of pairs)
    s = _schedule(l, 685, 2, 148)
    assert [len(x) for x in s] == [3] * 93 + [2] * 55 and s[0][2] == (592, False) and s[92][2] == (684, False)
