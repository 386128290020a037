"""CPU: the native step planner (csrc/host_plan.cpp) against NumPy's legacy generator and the Python planner.

The planner restates MT19937 / randint / legacy_gauss from NumPy's sources; these tests pin it bit for bit against
``np.random`` itself (values AND the generator state left behind), and against ``host.plan_isprs_batch`` -- the function
whose RNG consumption is pinned against the reference's ``dynamically_create_patches`` by tests/test_oracle_golden.py.
"""
import numpy as np
import pytest

from drs_b200 import host


@pytest.fixture(scope="module")
def planner():
    p = host.NativePlanner(threads=3)
    yield p
    p.close()


def _same_state(a, b):
    return a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2:] == b[2:]


@pytest.mark.parametrize("seed", [0, 1, 77, 12345])
def test_randint_and_normal_match_numpy(planner, seed):
    # interleave the calls the planner makes, with odd counts so that the cached Gaussian crosses call boundaries
    np.random.seed(seed)
    want = [np.random.randint(0, 2, size=5), np.random.normal(0, 0.01, 7), np.random.randint(0, 3, size=9),
            np.random.normal(0.5, 2.0, 1), np.random.normal(0, 0.01, 10001), np.random.randint(0, 3, size=700),
            np.random.normal(-1.0, 0.3, 2)]
    want_state = np.random.get_state()
    np.random.seed(seed)
    got = [planner.randint(2, 5), planner.normal(0, 0.01, 7), planner.randint(3, 9), planner.normal(0.5, 2.0, 1),
           planner.normal(0, 0.01, 10001), planner.randint(3, 700), planner.normal(-1.0, 0.3, 2)]
    for w, g in zip(want, got):
        assert np.array_equal(np.asarray(w, dtype=g.dtype), g)
    assert _same_state(want_state, np.random.get_state())


def test_generator_crosses_many_twists(planner):
    np.random.seed(5)
    a = np.random.normal(0, 1, 300001)        # > 1200 regenerations of the 624-word state, odd count
    b = np.random.randint(0, 3, size=11)
    np.random.seed(5)
    assert np.array_equal(a, planner.normal(0, 1, 300001))
    assert np.array_equal(b, planner.randint(3, 11))


def _scenes(rs, shapes, C):
    data = [rs.rand(h, w, C) for h, w in shapes]
    labels = [rs.randint(0, 6, size=(h, w)).astype(np.uint8) for h, w in shapes]
    return data, labels


@pytest.mark.parametrize("crop,C,B", [(25, 4, 16), (37, 4, 64), (26, 5, 9), (49, 5, 7), (31, 3, 1)])
def test_plan_matches_python_planner(planner, crop, C, B):
    rs = np.random.RandomState(crop * 100 + C)
    shapes = [(120, 150), (90, 200)]
    data, labels = _scenes(rs, shapes, C)
    inst = np.zeros((B, 4), dtype=np.int64)
    inst[:, 0] = rs.randint(0, 2, size=B)
    for i in range(B):
        h, w = shapes[inst[i, 0]]
        inst[i, 1], inst[i, 2] = rs.randint(0, h), rs.randint(0, w)      # some windows stick out: the border rule moves them back
    inst[:, 3] = rs.randint(0, 360, size=B)
    hw = np.asarray(shapes, dtype=np.int32)
    for rep, seed in enumerate((3, 4, 5)):
        np.random.seed(seed)
        if rep == 1:
            np.random.normal(0, 1, 3)          # leave a cached Gaussian behind: the plan must consume it first
        st0 = np.random.get_state()
        want = host.plan_isprs_batch(data, labels, inst, crop, is_train=True, rotate_on_device=True)
        after_py = np.random.get_state()
        tail_py = np.random.randint(0, 1 << 30, size=4)
        np.random.set_state(st0)
        got = planner.plan(hw, inst, crop, C)
        assert _same_state(after_py, np.random.get_state())
        assert np.array_equal(tail_py, np.random.randint(0, 1 << 30, size=4))
        assert np.array_equal(want.inst, got.inst) and np.array_equal(want.flips, got.flips)
        assert np.array_equal(want.noise_on, got.noise_on)
        rot_on = want.rot_on if want.rot_on is not None else np.zeros(B, dtype=np.uint8)
        assert np.array_equal(rot_on, got.rot_on)
        for b in range(B):
            if rot_on[b]:
                assert np.array_equal(want.rot[b], got.rot[b])
            if want.noise_on[b]:
                k = got.noise_slot[b]
                blk = got.noise[k * crop * crop * C:(k + 1) * crop * crop * C].reshape(crop, crop, C)
                assert np.array_equal(want.noise[b], blk)        # bit-identical float64 noise
            else:
                assert got.noise_slot[b] == -1


def test_rank_slice_scans_everything_but_transforms_its_share(planner):
    crop, C, B = 33, 4, 32
    rs = np.random.RandomState(9)
    inst = np.zeros((B, 4), dtype=np.int64)
    inst[:, 1], inst[:, 2], inst[:, 3] = rs.randint(0, 80, size=B), rs.randint(0, 80, size=B), rs.randint(0, 360, size=B)
    hw = np.asarray([(120, 120)], dtype=np.int32)
    np.random.seed(21)
    full = planner.plan(hw, inst, crop, C)
    full_noise = full.noise.copy()
    full_slots = full.noise_slot.copy()
    st_full = np.random.get_state()
    for b0, b1 in ((0, 8), (8, 16), (24, 32)):
        np.random.seed(21)
        part = planner.plan(hw, inst, crop, C, slot=host.PlanSlot(B, crop, C), own=(b0, b1))
        assert _same_state(st_full, np.random.get_state())        # every rank stays on the same stream
        assert np.array_equal(part.noise_slot, full_slots)
        for b in range(b0, b1):
            k = full_slots[b]
            if k >= 0:
                n = crop * crop * C
                assert np.array_equal(part.noise[k * n:(k + 1) * n], full_noise[k * n:(k + 1) * n])


def test_window_that_cannot_fit_is_an_error(planner):
    hw = np.asarray([(20, 20)], dtype=np.int32)
    with pytest.raises(ValueError):
        planner.plan(hw, np.asarray([[0, 3, 3, 10]], dtype=np.int64), 25, 4)
