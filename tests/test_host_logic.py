"""CPU: the product's host-side mirror (dynamic-rs-segmentation_b200/host.py) against reference-generated vectors.

host.py keeps the reference's L3 policy and the *decisions* of the L2 data path (which window, which flip,
which noise) on the host with the reference's RNG consumption; pixels are moved by the GPU gather kernel.
Here the decisions are checked by applying them with the NumPy checker (oracle.host_np.apply_plan).
"""
import random

import numpy as np

from oracle import host_np


def test_policy_functions_match_reference(golden, drs):
    from drs_b200 import host
    random.seed(7)
    total, bs = 50, 8
    shuffle = np.asarray(random.sample(range(total), total))
    it = 0
    for n in range(20):
        shuffle, batch, it = host.select_batch(shuffle, bs, it, total)
        assert np.array_equal(batch, golden["select_batch_batches"][n]) and it == golden["select_batch_its"][n]
    assert np.array_equal(host.define_multinomial_probs([25, 29, 33, 37, 41, 45, 49]), golden["probs_25_49"])
    for case in range(8):
        dist = ["multi_fixed", "uniform"][case % 2]
        upd = ["acc", "loss"][(case // 2) % 2]
        values = [25, 33, 41, 49] if dist == "multi_fixed" else [25, 32]
        pal = golden["best_%d_in" % case][0].astype(np.float32)
        occ = golden["best_%d_in" % case][1].astype(np.int32)
        chosen = np.zeros(len(occ), dtype=np.int32)
        assert int(host.select_best_patch_size(dist, values, pal, occ, upd, chosen)) == int(golden["best_vals"][case])
        assert np.array_equal(occ, golden["best_%d_occ_out" % case])
        assert np.array_equal(chosen, golden["best_%d_chosen_out" % case])


def test_isprs_batch_plan_reproduces_reference_augmentation(golden, drs):
    """dynamically_create_patches (isprs:245-334) incl. rotation, noise and flips under a fixed np.random seed."""
    from drs_b200 import host
    scenes, labs, inst = golden["gather_scenes"], golden["gather_labels"], golden["gather_inst"]
    for crop in (9, 12):
        np.random.seed(100 + crop)
        plan = host.plan_isprs_batch(scenes, labs, inst, crop, is_train=True)
        x, y = host_np.apply_plan(scenes, labs, plan.inst, plan.flips, crop, np.zeros(4), np.ones(4), plan.noise,
                                  plan.noise_on, plan.over_x, plan.over_y, plan.over_on, cast=False)
        assert np.array_equal(x, golden["gather_train_%d_p" % crop])
        assert np.array_equal(y.astype(np.int64), golden["gather_train_%d_l" % crop])
        am = np.ones(y.shape, dtype=bool) if plan.acc_mask is None else plan.acc_mask.astype(bool)
        assert np.array_equal(am, golden["gather_train_%d_m" % crop])
        # eval mode: no RNG consumed, no augmentation
        state = np.random.get_state()[1].copy()
        plan = host.plan_isprs_batch(scenes, labs, inst, crop, is_train=False)
        assert np.array_equal(np.random.get_state()[1], state)
        x, y = host_np.apply_plan(scenes, labs, plan.inst, plan.flips, crop, np.zeros(4), np.ones(4), cast=False)
        assert np.array_equal(x, golden["gather_eval_%d_p" % crop])


def test_index_flip_plans(golden, drs):
    from drs_b200 import host
    img, lab = golden["contest_scene"], golden["contest_labels"]
    plan = host.plan_index_flip_batch([tuple(r) for r in golden["contest_distr"]], golden["contest_shuf"], 11,
                                      [img.shape[:2]])
    x, y = host_np.apply_plan([img], [lab], plan.inst, plan.flips, 11, np.zeros(3), np.ones(3), cast=False)
    assert np.array_equal(x, golden["contest_gather_p"])
    assert np.array_equal(y.astype(np.int8), golden["contest_gather_l"])
    imgs, labs = golden["coffee_scenes"], golden["coffee_labels"]
    cd = [(int(m), (int(a), int(b))) for m, a, b in golden["coffee_distr"]]
    plan = host.plan_index_flip_batch(cd, golden["coffee_shuf"], 9, [im.shape[:2] for im in imgs], with_map=True)
    x, y = host_np.apply_plan(imgs, labs, plan.inst, plan.flips, 9, np.zeros(3), np.ones(3), cast=False)
    assert np.array_equal(x.astype(np.float16), golden["coffee_gather_p"])


def test_acc_norm_from_confusion(golden, drs):
    from drs_b200 import host
    assert host.acc_norm_from_cm(golden["cm_out"], 6) == golden["cm_acc"][1]
    assert host.acc_norm_from_cm(golden["cm_nomask_out"], 6) == golden["cm_nomask_acc"][1]
    assert host.sliding_stride(25) == 12 and host.sliding_stride(30) == 15


def test_rotate_affine_reproduces_scipy_rotate():
    """The index arithmetic the gather kernel uses for the isprs rotation augmentation, restated in NumPy:
    cc = (off + i*m_0) + j*m_1 in float64, floor(cc + 0.5), constant 0 outside [0, crop-1] -- equal to
    scipy.ndimage.rotate(order=0, reshape=False) for every integer angle the reference can draw (isprs:489)."""
    import scipy.ndimage
    from drs_b200 import host
    rs = np.random.RandomState(0)
    for crop in (25, 26, 37, 49):
        patch = rs.rand(crop, crop, 2) + 1.0
        ii, jj = np.meshgrid(np.arange(crop, dtype=np.float64), np.arange(crop, dtype=np.float64), indexing="ij")
        for angle in range(0, 360, 1 if crop == 25 else 7):
            m = host.rotate_affine(angle, crop)
            c0 = (m[4] + ii * m[0]) + jj * m[1]
            c1 = (m[5] + ii * m[2]) + jj * m[3]
            ok = ~((c0 < 0) | (c0 > crop - 1) | (c1 < 0) | (c1 > crop - 1))
            out = np.zeros_like(patch)
            out[ok] = patch[np.floor(c0 + 0.5).astype(np.int64)[ok], np.floor(c1 + 0.5).astype(np.int64)[ok]]
            assert np.array_equal(out, scipy.ndimage.rotate(patch, angle, order=0, reshape=False)), (crop, angle)


def test_score_files_are_byte_identical_to_numpy_save(tmp_path):
    """isprs:1800-1802 / coffee:1341-1343: the per-patch-size score arrays are written with np.save under the reference's
    file names; the bytes on disk are those np.save produces for the same arrays (dtype float32 / int32 as isprs:2054-2064)."""
    from drs_b200 import host, loops
    pal, occ, chosen = host.init_score_arrays("multinomial", [25, 29, 33])
    pal[:] = np.arange(len(pal), dtype=np.float32) * 0.37
    occ[3], chosen[2] = 5, 1
    out = str(tmp_path) + "/"
    loops._save_scores(out, 1000, pal, occ, chosen)
    loops._save_scores(out, 1000, pal, occ, chosen, names=("errorAcc", "errorOccur", "chosenValues"))
    for name, arr in (("patch_acc_loss", pal), ("patch_occur", occ), ("patch_chosen_values", chosen), ("errorAcc", pal),
                      ("errorOccur", occ), ("chosenValues", chosen)):
        ref = tmp_path / ("ref_" + name + ".npy")
        np.save(ref, arr)
        assert open(out + name + "_step_1000.npy", "rb").read() == open(ref, "rb").read()
        assert arr.dtype == (np.float32 if arr is pal else np.int32)
