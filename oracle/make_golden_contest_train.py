"""Generate tests/golden/contest_train_golden.npz by running the REAL ``contest_dilated_random.train`` (contest:970-1158) through
the closed-form fake ``sess.run`` of oracle/make_golden.py (TEST INFRASTRUCTURE; needs /root/reference, build container only):

    python -m oracle.make_golden_contest_train

Pins, against the reference's own loop: the interleaving of patch-size draws and batch selection incl. the reference's extra
``it`` bookkeeping on top of select_batch (contest:1141-1146), the flip-by-index-range gather, float32 normalisation, the fed
labels, and the score arrays after the final select_best_patch_size -- for the pipelined drop-in loop ``loops.contest_train``.
"""
import io
import os
import random
import tempfile
from contextlib import redirect_stdout

import numpy as np

from oracle import ref_import
from oracle.make_golden import FakeSession, synth_scene

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "contest_train_golden.npz")


def main():
    import sys
    _, contest, _ = ref_import.load()
    tf = sys.modules["tensorflow"]
    G = {}
    rs = np.random.RandomState(55)
    img, lab = synth_scene(rs, 60, 52, 3, 8, block=7, dtype=np.float32, cycle=True)      # labels 0..7, 7 = unlabelled
    timg, tlab = synth_scene(rs, 40, 36, 3, 8, block=7, dtype=np.float32, cycle=True)
    G["scene"], G["labels"], G["test_scene"], G["test_labels"] = img, lab, timg, tlab
    cases = [("multi_fixed", [9, 13], "loss"), ("uniform", [9, 12], "loss"), ("single_fixed", [11], "acc")]
    cwd = os.getcwd()
    for ci, (dist, values, upd) in enumerate(cases):
        tmp = tempfile.mkdtemp()
        os.chdir(tmp)
        try:
            np.random.seed(3000 + ci)
            random.seed(4000 + ci)
            with redirect_stdout(io.StringIO()):
                distr = contest.create_distributions_over_classes(lab, 9, 17)
                mean_f, std_f = contest.create_mean_and_std(img, distr, 9)
            n = len(values) if dist == "multi_fixed" else values[-1] - values[0] + 1
            pal = np.zeros(n, dtype=np.float32)
            occ = np.ones(n, dtype=np.int32)              # contest starts patch_occur at ones (contest:1275-1279)
            chosen = np.zeros(n, dtype=np.int32)
            sess = FakeSession(3, 7)
            tf.placeholder.side_effect = lambda *a, **k: object()
            tf.Session.return_value.__enter__.return_value = sess
            try:
                with redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                    contest.train(img, lab, timg, tlab, distr, mean_f, std_f, tmp + "/", "", 0.01, 0.005, 4, 23, "dilated_grsl",
                                  dist, upd, pal, occ, chosen, None, values)
            except TypeError as e:      # the final test() uses a Python-2 ``/`` inside range (contest:919); the loop is complete
                print("contest.train tail raised (expected, py2 range):", e)
            G["c%d_distr" % ci] = np.asarray(distr, dtype=np.int64)
            G["c%d_mean" % ci], G["c%d_std" % ci] = mean_f, std_f
            G["c%d_log" % ci] = np.array(sess.log, dtype=np.float64)
            G["c%d_pal" % ci], G["c%d_occ" % ci], G["c%d_chosen" % ci] = pal, occ, chosen
        finally:
            os.chdir(cwd)
    G["cases"] = np.array([[{"single_fixed": 0, "multi_fixed": 1, "uniform": 2}[d], {"acc": 0, "loss": 1}[u]] + v + [0] * (2 - len(v))
                           for d, v, u in cases], dtype=np.int64)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main()
