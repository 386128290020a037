"""CPU property tests (hypothesis) of the host-side integer logic that has to be bit-exact (SURVEY section 8, rows a13,
a16, e): the C++ visiting order behind the C-ABI against the NumPy oracle for random geometry, coverage and stripe
invariants, and the contract of the reference arm of bench.py."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import host_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@settings(max_examples=120, deadline=None)
@given(h=st.integers(25, 140), w=st.integers(25, 140), crop=st.integers(7, 25), batch=st.integers(1, 40),
       variant=st.sampled_from(["isprs", "contest", "coffee"]))
def test_grid_positions_abi_equals_oracle_for_random_geometry(drs, h, w, crop, batch, variant):
    """drs_grid_positions (C++, host-only) == the NumPy restatement of create_patches_per_map (isprs:337-400,
    contest:257-328 incl. the offset_h bug, coffee:296-349) for arbitrary scene / patch / batch sizes."""
    if variant == "coffee":
        w = h                                   # coffee tiles are square (coffee:302-307 derive both counts from h)
    ref = np.array(host_np.all_patch_positions(h, w, crop, batch, variant), dtype=np.int64).reshape(-1, 2)
    got = np.asarray(drs.grid_positions(h, w, crop, batch, variant)).reshape(-1, 2)
    assert np.array_equal(got, ref)


@settings(max_examples=60, deadline=None)
@given(h=st.integers(25, 200), w=st.integers(25, 200), crop=st.integers(6, 25), batch=st.integers(1, 64))
def test_isprs_grid_covers_every_pixel_inside_the_scene(drs, h, w, crop, batch):
    pos = np.asarray(drs.grid_positions(h, w, crop, batch, "isprs")).reshape(-1, 2)
    assert pos.min() >= 0 and (pos[:, 0] + crop).max() == h and (pos[:, 1] + crop).max() == w
    cover = np.zeros((h, w), dtype=np.int32)
    for r, c in np.unique(pos, axis=0):
        cover[r:r + crop, c:c + crop] += 1
    assert cover.min() >= 1
    stride = crop // 2
    assert len(pos) == host_np.grid_count(h, crop, stride) * host_np.grid_count(w, crop, stride)


@settings(max_examples=60, deadline=None)
@given(h=st.integers(60, 400), crop=st.integers(8, 49), world=st.integers(1, 8))
def test_stripes_partition_rows_and_hold_every_patch_that_touches_them(drs, h, crop, world):
    """Row stripes: disjoint cover of [0, H); the rows a rank keeps resident contain every patch that
    intersects its stripe, so each pixel's contributions are all local (no halo exchange)."""
    from drs_b200 import dist as ddist
    if crop > h:
        return
    cuts = ddist.stripe_bounds(h, world)
    assert cuts[0] == 0 and cuts[-1] == h and all(a <= b for a, b in zip(cuts[:-1], cuts[1:]))
    stride = crop // 2
    origins = sorted(set(list(range(0, h - crop + 1, stride)) + [h - crop]))
    for r in range(world):
        r0, r1 = ddist.stripe_bounds(h, world, r)
        if r0 == r1:
            continue
        u0, u1 = ddist.stripe_rows_needed(h, crop, r0, r1)
        for o in origins:
            if o < r1 and o + crop > r0:
                assert u0 <= o and o + crop <= u1
        assert 0 <= u0 <= r0 and r1 <= u1 <= h


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: one JSON line with the contract's keys, timed on the host cores through the oracle port."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "train patches/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
