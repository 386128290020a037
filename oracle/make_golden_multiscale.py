"""Golden vectors for the multi-scale evaluation loop (TEST INFRASTRUCTURE, build container only).

Drives the reference's own ``validate_test_multiscale`` (/root/reference/isprs_dilated_random.py:1347-1474) through the
closed-form fake ``sess.run`` of oracle/make_golden.py and records the label map it hands to its metrics, for
``multi_fixed`` score files with both update types.  Output: tests/golden/multiscale_golden.npz.

    python -m oracle.make_golden_multiscale
"""
import io
import os
import tempfile
from contextlib import redirect_stdout

import numpy as np

from oracle import ref_import
from oracle.make_golden import FakeSession, synth_scene

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "multiscale_golden.npz")


def main():
    isprs, _, _ = ref_import.load()
    G = {}
    captured = {}
    isprs.cohen_kappa_score = lambda a, b, **kw: captured.__setitem__("pred_noerode", np.asarray(b).copy()) or 0.0
    isprs.f1_score = lambda a, b, average=None, **kw: np.zeros(6) if average is None else 0.0
    real_argmax = np.argmax

    def spy_argmax(a, axis=None, **kw):
        r = real_argmax(a, axis=axis, **kw)
        if axis == 2 and np.ndim(a) == 3:
            captured["labels"] = np.asarray(r).copy()
        return r

    rs = np.random.RandomState(77)
    img, lab = synth_scene(rs, 90, 110, 4, 6)
    lab = lab.copy()
    lab[::9, ::4] = 6                                       # eroded pixels (isprs:1294)
    mean = np.array([0.4, 0.5, 0.45, 0.3])
    std = np.array([0.2, 0.25, 0.21, 0.3])
    values = np.array([25, 30, 33, 41])
    cases = []
    for ci, (update, pal, occ) in enumerate((("acc", [3.1, 4.6, 4.0, 1.0], [5, 6, 5, 2]),
                                             ("loss", [9.0, 2.0, 3.3, 7.0], [4, 2, 3, 0]))):
        with tempfile.TemporaryDirectory() as d:
            d = d + "/"
            np.save(d + "patch_acc_loss_step_5.npy", np.asarray(pal, dtype=np.float32))
            np.save(d + "patch_occur_step_5.npy", np.asarray(occ, dtype=np.int32))
            sess = FakeSession(4, 6)
            isprs.np.argmax = spy_argmax
            try:
                with redirect_stdout(io.StringIO()):
                    isprs.validate_test_multiscale(sess, np.asarray([img]), np.asarray([lab]), ["1"], 7, mean, std, "x", "y",
                                                   "crop", "keep", "is_training", "pred_up", "logits", 5, "multi_fixed",
                                                   values.copy(), update, 3, False, d)
            finally:
                isprs.np.argmax = real_argmax
        G["ms_%d_labels" % ci] = captured["labels"].astype(np.uint8)
        G["ms_%d_pal" % ci] = np.asarray(pal, dtype=np.float32)
        G["ms_%d_occ" % ci] = np.asarray(occ, dtype=np.int32)
        G["ms_%d_crops" % ci] = np.array([c for (_, c, _, _, _) in sess.log], dtype=np.int64)
        cases.append(update)
    G["ms_scene"], G["ms_gt"], G["ms_mean"], G["ms_std"], G["ms_values"] = img, lab, mean, std, values
    G["ms_updates"] = np.array(cases)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, {k: v.shape for k, v in G.items()})


if __name__ == "__main__":
    main()
