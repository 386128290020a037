"""CPU: the drop-in training loop (dynamic-rs-segmentation_b200/loops.py) against traces produced by the reference's OWN
``isprs.train`` driven through a closed-form fake ``sess.run`` (oracle/make_golden.py section 8).

What is pinned: the RNG interleaving of patch-size draws / batch selection / augmentation, the instance sampling, the
normalised patches fed at every step (sum of the float64 feed), the labels, the per-step score update and the saved
``patch_*_step_N.npy`` arrays -- for all four distribution types and both update types.
"""
import io
import os
import random
from contextlib import redirect_stdout

import numpy as np
import pytest

from oracle import host_np
from oracle.fake_net import fake_train_fetches

DIST = {0: "single_fixed", 1: "multi_fixed", 2: "uniform", 3: "multinomial"}
UPD = {0: "acc", 1: "loss"}


class FakeBackend:
    """Closed-form stand-in for the device: same fetches the fake session of the golden run returned."""

    def __init__(self, scenes, labels, mean, std, C, K):
        self.scenes, self.labels, self.mean, self.std, self.C, self.K = scenes, labels, mean, std, C, K
        self.log = []
        self.saved = []

    def train_on_plan(self, plan, loss_mask=None):
        x, y = host_np.apply_plan(self.scenes, self.labels, plan.inst, plan.flips, plan.crop, self.mean, self.std, plan.noise,
                                  plan.noise_on, plan.over_x, plan.over_y, plan.over_on, cast=False)
        B = len(plan.inst)
        bx = np.reshape(x, (-1, plan.crop * plan.crop * self.C))
        by = np.reshape(y, (-1, plan.crop * plan.crop))
        self.log.append((1, plan.crop, B, float(np.sum(bx.astype(np.float64))), float(np.sum(by.astype(np.float64)))))
        loss, pred = fake_train_fetches(bx, plan.crop, self.C, self.K)
        masks = None if plan.acc_mask is None else plan.acc_mask.astype(bool)
        acc, _, cm = host_np.confusion_by_crop(y.astype(np.int64), pred, self.K, masks)
        return loss, cm, acc

    def save(self, path):
        self.saved.append(path)

    def restore(self, path):
        raise AssertionError("not expected")


class FakeDeviceBackend(FakeBackend):
    """The stand-in with the product backend's contract: rotation happens on the "device" (the plan carries scipy's affine map
    per patch, the stand-in applies the kernel's index arithmetic in NumPy), noise arrives compact (native planner), and the
    step is asynchronous (submit_train / train_result), so the loop's prefetching, one-step-late result handling and flush
    points are all exercised against the reference's trace."""
    rotate_on_device = True

    def __init__(self, *a):
        super().__init__(*a)
        self.outstanding = {}
        self.ticket = 0
        self.max_outstanding = 0

    def _as_overrides(self, plan):
        from drs_b200 import host
        B, crop = len(plan.inst), plan.crop
        q = host.BatchPlan()
        q.crop, q.inst, q.flips = crop, np.array(plan.inst), np.array(plan.flips)
        q.noise_on = None if plan.noise_on is None else np.array(plan.noise_on)
        q.noise = None
        if plan.noise is not None:
            q.noise = np.zeros((B, crop, crop, self.C))
            for b in range(B):
                if plan.noise_on[b]:
                    if plan.noise_slot is not None:
                        k, n = int(plan.noise_slot[b]), crop * crop * self.C
                        q.noise[b] = plan.noise[k * n:(k + 1) * n].reshape(crop, crop, self.C)
                    else:
                        q.noise[b] = plan.noise[b]
        q.over_x = q.over_y = q.over_on = q.acc_mask = None
        if plan.rot_on is not None and np.any(plan.rot_on):
            q.over_x = np.zeros((B, crop, crop, self.C))
            q.over_y = np.zeros((B, crop, crop), dtype=np.uint8)
            q.over_on = np.array(plan.rot_on)
            q.acc_mask = np.ones((B, crop, crop), dtype=np.uint8)
            ii, jj = np.meshgrid(np.arange(crop, dtype=np.float64), np.arange(crop, dtype=np.float64), indexing="ij")
            for b in np.nonzero(plan.rot_on)[0]:
                m = plan.rot[b]
                c0 = (m[4] + ii * m[0]) + jj * m[1]
                c1 = (m[5] + ii * m[2]) + jj * m[3]
                ok = ~((c0 < 0) | (c0 > crop - 1) | (c1 < 0) | (c1 > crop - 1))
                r0, r1 = np.floor(c0 + 0.5).astype(np.int64), np.floor(c1 + 0.5).astype(np.int64)
                sm, sr, sc = int(plan.inst[b][0]), int(plan.inst[b][1]), int(plan.inst[b][2])
                src = self.scenes[sm][sr:sr + crop, sc:sc + crop]
                lab = self.labels[sm][sr:sr + crop, sc:sc + crop]
                q.over_x[b][ok] = src[r0[ok], r1[ok]]
                q.over_y[b][ok] = lab[r0[ok], r1[ok]]
                q.acc_mask[b] = ok
            for b in range(B):          # the accuracy mask is flipped together with the patch (isprs:308-317)
                if q.flips[b] == 1:
                    q.acc_mask[b] = np.flipud(q.acc_mask[b])
                elif q.flips[b] == 2:
                    q.acc_mask[b] = np.fliplr(q.acc_mask[b])
        return q

    def submit_train(self, plan, loss_mask=None):
        # evaluated at submission (the plan's buffers are recycled later, like pinned slots behind an async upload)
        self.outstanding[self.ticket] = FakeBackend.train_on_plan(self, self._as_overrides(plan), loss_mask)
        self.max_outstanding = max(self.max_outstanding, len(self.outstanding))
        self.ticket += 1
        return self.ticket - 1

    def train_result(self, ticket):
        return self.outstanding.pop(ticket)

    def train_on_plan(self, plan, loss_mask=None):
        return self.train_result(self.submit_train(plan, loss_mask))

    def own_rows(self, n):
        return 0, n


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
@pytest.mark.parametrize("native", [True, False])
def test_pipelined_loop_with_device_rotation_reproduces_reference_trace(golden, drs, tmp_path, ci, native, monkeypatch):
    """The product configuration of the loop -- native planner (csrc/host_plan.cpp), rotation on the device, asynchronous
    steps with results taken one step late -- fed to the closed-form stand-in must reproduce the trace of the reference's
    own ``isprs.train`` bit for bit: same sizes, same fed patches (rotated, noisy, flipped, normalised), same score files."""
    from drs_b200 import host, loops
    monkeypatch.setenv("DRS_NATIVE_PLAN", "1" if native else "0")
    case = golden["train_cases"][ci]
    dist, upd = DIST[int(case[0])], UPD[int(case[1])]
    values = [int(v) for v in case[2:] if v > 0]
    tr_d, tr_l = golden["train_scenes"], golden["train_labels"]
    te_d, te_l = golden["test_scenes"], golden["test_labels"]
    monkeypatch.chdir(tmp_path)
    np.random.seed(1000 + ci)
    random.seed(2000 + ci)
    with redirect_stdout(io.StringIO()):
        tr_distr = host.create_distributions_over_classes(tr_l, 25, 5, 6)
        te_distr = host.create_distributions_over_classes(te_l, 25, 5, 6)
        rot = host.create_rotation_distribution(tr_distr)
        mean_f, std_f = host.dynamically_calculate_mean_and_std(tr_d, tr_distr, 25)
    pal, occ, chosen = host.init_score_arrays(dist, values)
    probs = host.define_multinomial_probs(values) if dist == "multinomial" else None
    be = FakeDeviceBackend(tr_d, tr_l, mean_f, std_f, 4, 6)
    out = io.StringIO()
    with redirect_stdout(out):
        loops.isprs_train(be, tr_d, tr_l, tr_distr, rot, te_d, te_l, te_distr, ["1"], 4, 14, upd, dist, values, pal, occ,
                          chosen, probs, 20, str(tmp_path) + "/", 5, "vaihingen", "", final_validation=False)
    ref = golden["train_%d_log" % ci]
    got = np.array(be.log, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.array_equal(got[:, :3], ref[:, :3])
    assert np.array_equal(got[:, 4], ref[:, 4])
    assert np.allclose(got[:, 3], ref[:, 3], rtol=1e-12, atol=1e-9)
    if dist != "single_fixed":
        assert np.array_equal(np.load(tmp_path / "patch_acc_loss_step_14.npy"), golden["train_%d_pal" % ci])
        assert np.array_equal(np.load(tmp_path / "patch_occur_step_14.npy"), golden["train_%d_occ" % ci])
        assert np.array_equal(np.load(tmp_path / "patch_chosen_values_step_14.npy"), golden["train_%d_chosen" % ci])
    assert be.max_outstanding == 2 and not be.outstanding        # one step in flight behind the one being submitted
    assert out.getvalue().count("Training Minibatch") == 2       # display_step 5: flushed at steps 5 and 10


def test_data_parallel_local_planning_shares_sizes_and_batches(golden, drs, tmp_path, monkeypatch):
    """Data-parallel host rules: with rank-local augmentation streams every rank still draws the same patch
    size and the same global batch per step (shared streams) and plans only ITS slice; the augmentation of different ranks
    comes from different generators.  DRS_DP_PLAN=shared plans the whole global batch on the shared stream instead."""
    from drs_b200 import host, loops
    tr_d, tr_l = golden["train_scenes"], golden["train_labels"]
    te_d, te_l = golden["test_scenes"], golden["test_labels"]
    monkeypatch.chdir(tmp_path)

    class RankBackend(FakeDeviceBackend):
        def __init__(self, rank, world, *a):
            super().__init__(*a)
            self.rank, self.world, self.plans = rank, world, []

        def own_rows(self, n):
            per = n // self.world
            return self.rank * per, (self.rank + 1) * per

        def submit_train(self, plan, loss_mask=None):
            self.plans.append((plan.crop, bool(plan.local), np.array(plan.inst), np.array(plan.flips), np.array(plan.noise_on)))
            if not plan.local:         # shared mode: the backend slices (GpuBackend._gather); emulate it for the stand-in
                return super().submit_train(plan, loss_mask)
            return super().submit_train(plan, loss_mask)

    def run(rank, mode):
        monkeypatch.setenv("DRS_DP_PLAN", mode)
        d = tmp_path / ("%s_%d" % (mode, rank))      # own working directory: the loop caches its test instances there, and a
        d.mkdir()                                    # run that loads the cache consumes fewer draws than the one that made it
        monkeypatch.chdir(d)                         # (under torchrun: dist.rank0_first + dist.sync_host_rng)
        np.random.seed(77)
        random.seed(78)
        with redirect_stdout(io.StringIO()):
            tr_distr = host.create_distributions_over_classes(tr_l, 25, 5, 6)
            te_distr = host.create_distributions_over_classes(te_l, 25, 5, 6)
            rot = host.create_rotation_distribution(tr_distr)
        values = [13, 17, 21]
        pal, occ, chosen = host.init_score_arrays("multi_fixed", values)
        be = RankBackend(rank, 2, tr_d, tr_l, np.full(4, 0.5), np.full(4, 0.3), 4, 6)
        with redirect_stdout(io.StringIO()):
            loops.isprs_train(be, tr_d, tr_l, tr_distr, rot, te_d, te_l, te_distr, ["1"], 8, 6, "acc", "multi_fixed", values, pal, occ,
                              chosen, None, 20, str(tmp_path) + "/", 50, "dp", "", final_validation=False)
        return be.plans

    a, b = run(0, "local"), run(1, "local")
    full = run(0, "shared")
    assert len(a) == len(b) == len(full) == 6
    differs = False
    for (ca, la, ia, fa, na), (cb, lb, ib, fb, nb), (cf, lf, i_f, ff, nf) in zip(a, b, full):
        assert la and lb and not lf
        assert ca == cb                                     # same patch size on every rank
        assert len(ia) == len(ib) == 4 and len(i_f) == 8     # a rank plans its half of the global batch of 8
        differs = differs or not (np.array_equal(fa, fb) and np.array_equal(na, nb))
    assert differs                                          # different augmentation draws on different ranks


@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_isprs_train_loop_reproduces_reference_trace(golden, drs, tmp_path, ci):
    from drs_b200 import host, loops
    case = golden["train_cases"][ci]
    dist, upd = DIST[int(case[0])], UPD[int(case[1])]
    values = [int(v) for v in case[2:] if v > 0]
    tr_d, tr_l = golden["train_scenes"], golden["train_labels"]
    te_d, te_l = golden["test_scenes"], golden["test_labels"]
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        np.random.seed(1000 + ci)
        random.seed(2000 + ci)
        with redirect_stdout(io.StringIO()):
            tr_distr = host.create_distributions_over_classes(tr_l, 25, 5, 6)
            te_distr = host.create_distributions_over_classes(te_l, 25, 5, 6)
            rot = host.create_rotation_distribution(tr_distr)
            mean_f, std_f = host.dynamically_calculate_mean_and_std(tr_d, tr_distr, 25)
        assert np.array_equal(mean_f, golden["train_%d_mean" % ci]) and np.array_equal(std_f, golden["train_%d_std" % ci])
        pal, occ, chosen = host.init_score_arrays(dist, values)
        probs = host.define_multinomial_probs(values) if dist == "multinomial" else None
        be = FakeBackend(tr_d, tr_l, mean_f, std_f, 4, 6)
        with redirect_stdout(io.StringIO()):
            loops.isprs_train(be, tr_d, tr_l, tr_distr, rot, te_d, te_l, te_distr, ["1"], 4, 14, upd, dist, values, pal, occ,
                              chosen, probs, 20, str(tmp_path) + "/", 50, "vaihingen", "", final_validation=False)
        ref = golden["train_%d_log" % ci]
        got = np.array(be.log, dtype=np.float64)
        assert got.shape == ref.shape
        assert np.array_equal(got[:, :3], ref[:, :3])                      # is_training, patch size, batch size per step
        assert np.array_equal(got[:, 4], ref[:, 4])                        # labels fed
        assert np.allclose(got[:, 3], ref[:, 3], rtol=1e-12, atol=1e-9)    # normalised float64 patches fed
        if dist != "single_fixed":
            assert np.array_equal(np.load(tmp_path / "patch_acc_loss_step_14.npy"), golden["train_%d_pal" % ci])
            assert np.array_equal(np.load(tmp_path / "patch_occur_step_14.npy"), golden["train_%d_occ" % ci])
            assert np.array_equal(np.load(tmp_path / "patch_chosen_values_step_14.npy"), golden["train_%d_chosen" % ci])
        assert be.saved == [str(tmp_path) + "/model-14"]
    finally:
        os.chdir(cwd)


def test_metrics_from_confusion_match_sklearn():
    from drs_b200 import loops
    from sklearn.metrics import cohen_kappa_score, f1_score
    rs = np.random.RandomState(0)
    t = rs.randint(0, 6, size=5000)
    p = np.where(rs.rand(5000) < 0.6, t, rs.randint(0, 6, size=5000))
    cm = loops.confusion_counts(t, p, 6)
    assert cm.sum() == 5000 and cm[2, 3] == int(((t == 2) & (p == 3)).sum())
    assert abs(loops.kappa_from_cm(cm) - cohen_kappa_score(t, p)) < 1e-12
    assert np.allclose(loops.f1_per_class_from_cm(cm), f1_score(t, p, average=None), atol=1e-12)
    cm_i = loops.confusion_counts(np.where(rs.rand(5000) < 0.1, 6, t), p, 7)[:6, :6]     # eroded label 6 dropped
    assert cm_i.sum() < 5000


def test_multiscale_validation_reproduces_reference_label_maps(tmp_path):
    """loops.isprs_validate_test_multiscale against the label maps the reference's own validate_test_multiscale
    (isprs:1347-1474) produced through the closed-form fake sess.run (oracle/make_golden_multiscale.py)."""
    from drs_b200 import loops
    from oracle.fake_net import fake_logits
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multiscale_golden.npz"))
    img, lab, mean, std, values = g["ms_scene"], g["ms_gt"], g["ms_mean"], g["ms_std"], g["ms_values"]

    class Backend:
        crops = []

        def scene_mean_logits(self, scene_id, crop, batch, variant="isprs"):
            self.crops.append(crop)
            h, w = img.shape[:2]
            pos = host_np.all_patch_positions(h, w, crop, batch)
            inst = np.array([[0, r, c] for r, c in pos], dtype=np.int32)
            x, _ = host_np.apply_plan([img], [lab], inst, None, crop, mean, std, cast=False)
            logits = fake_logits(x.reshape(len(inst), -1), crop, 4, 6)
            _, m = host_np.accumulate_argmax(logits, pos, h, w, crop, return_mean=True)
            return m

    for ci in (0, 1):
        out = str(tmp_path) + "/"
        np.save(out + "patch_acc_loss_step_5.npy", g["ms_%d_pal" % ci])
        np.save(out + "patch_occur_step_5.npy", g["ms_%d_occ" % ci])
        be = Backend()
        be.crops = []
        with redirect_stdout(io.StringIO()):
            maps = loops.isprs_validate_test_multiscale(be, [img], [lab], ["1"], 7, 5, "multi_fixed", values.copy(),
                                                        str(g["ms_updates"][ci]), 3, False, out)
        ref_crops = g["ms_%d_crops" % ci]
        order = [int(c) for i, c in enumerate(ref_crops) if i == 0 or c != ref_crops[i - 1]]
        assert be.crops == order
        assert np.array_equal(maps[0].astype(np.uint8), g["ms_%d_labels" % ci])


class FakeContestBackend:
    """Closed-form stand-in for the contest loop: float32 scene, flips by index range, loss mask label != 7 (contest:236-239),
    asynchronous interface like the product backend."""

    def __init__(self, scene, labels, mean, std):
        self.scene, self.labels, self.mean, self.std = scene, labels, mean, std
        self.log, self.saved, self.out = [], [], {}
        self.ticket = 0

    def submit_train(self, plan, loss_mask=None):
        assert loss_mask == "label!=7"
        x, y = host_np.apply_plan([self.scene], [self.labels], plan.inst, plan.flips, plan.crop, self.mean, self.std, cast=False)
        B = len(plan.inst)
        bx = np.reshape(x, (-1, plan.crop * plan.crop * 3))
        self.log.append((1, plan.crop, B, float(np.sum(bx.astype(np.float64))), float(np.sum(y.astype(np.float64)))))
        loss, pred = fake_train_fetches(bx, plan.crop, 3, 7)
        cm = np.zeros((7, 7), dtype=np.uint32)
        self.out[self.ticket] = (loss, cm, 0)
        self.ticket += 1
        return self.ticket - 1

    def train_result(self, ticket):
        return self.out.pop(ticket)

    def save(self, path):
        self.saved.append(path)


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_contest_train_loop_reproduces_reference_trace(drs, tmp_path, ci, monkeypatch):
    """loops.contest_train (pipelined) against the trace of the reference's own contest.train (contest:970-1158) through the
    closed-form fake sess.run (oracle/make_golden_contest_train.py): patch sizes, batches incl. the reference's own ``it``
    bookkeeping and epoch reshuffles, flipped + normalised float32 patches, labels, score arrays after the final
    select_best_patch_size."""
    from drs_b200 import host, loops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "contest_train_golden.npz"))
    case = g["cases"][ci]
    dist, upd = {0: "single_fixed", 1: "multi_fixed", 2: "uniform"}[int(case[0])], UPD[int(case[1])]
    values = [int(v) for v in case[2:] if v > 0]
    img, lab, tlab = g["scene"], g["labels"], g["test_labels"]
    monkeypatch.chdir(tmp_path)
    np.random.seed(3000 + ci)
    random.seed(4000 + ci)
    with redirect_stdout(io.StringIO()):
        distr = host.contest_create_distributions_over_classes(lab, 9, 17, 7)
        mean_f, std_f = host.contest_create_mean_and_std(img, distr, 9)
    assert np.array_equal(np.asarray(distr, dtype=np.int64), g["c%d_distr" % ci])
    assert np.array_equal(mean_f, g["c%d_mean" % ci]) and np.array_equal(std_f, g["c%d_std" % ci])
    pal, occ, chosen = host.init_score_arrays(dist, values, occur_init=1)
    if pal is None:                                   # single_fixed: the reference still passes (unused) arrays
        pal, occ, chosen = np.zeros(1, dtype=np.float32), np.ones(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
    be = FakeContestBackend(img, lab, mean_f, std_f)
    with redirect_stdout(io.StringIO()):
        loops.contest_train(be, img, lab, tlab, distr, str(tmp_path) + "/", "", 4, 23, dist, upd, pal, occ, chosen, None, values, 7,
                            final_test=False)
    ref = g["c%d_log" % ci]
    got = np.array(be.log, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.array_equal(got[:, :3], ref[:, :3])
    assert np.array_equal(got[:, 4], ref[:, 4])
    assert np.allclose(got[:, 3], ref[:, 3], rtol=1e-6, atol=1e-4)     # float32 patches summed in float64
    if dist != "single_fixed":
        assert np.array_equal(pal, g["c%d_pal" % ci]) and np.array_equal(occ, g["c%d_occ" % ci])
        assert np.array_equal(chosen, g["c%d_chosen" % ci])
    assert be.saved == [str(tmp_path) + "/model-23"] and not be.out
