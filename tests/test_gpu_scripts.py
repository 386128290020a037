"""GPU: the three drop-in command lines end to end on tiny synthetic datasets (same positional argv as the reference)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(rs, h, w, c, k, dtype, block=25, extra_label=None):
    img = (rs.randint(0, 256, size=(h, w, c)).astype(np.uint8) / 255.0).astype(dtype)
    nb_h, nb_w = (h + block - 1) // block, (w + block - 1) // block
    cls = (np.arange(nb_h * nb_w).reshape(nb_h, nb_w) + rs.randint(0, k)) % k
    if extra_label is not None:
        cls[rs.rand(nb_h, nb_w) < 0.2] = extra_label
    lab = np.repeat(np.repeat(cls, block, axis=0), block, axis=1)[:h, :w].astype(np.uint8)
    return img, lab


def _run(script, args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, script)] + [str(a) for a in args], cwd=cwd, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_isprs_cli_train_validate_and_final_maps(tmp_path):
    rs = np.random.RandomState(0)
    data = tmp_path / "vaihingen"
    out = tmp_path / "out"
    data.mkdir()
    out.mkdir()
    for name in ("1", "3", "5"):
        img, lab = _scene(rs, 150, 175, 4, 6, np.float64)
        np.save(data / (name + "_image.npy"), img)
        np.save(data / (name + "_labels.npy"), lab)
    common = [str(data) + "/", str(out) + "/"]
    log = _run("isprs_dilated_random.py", common + ["", "1,3", "5", 0.01, 0.005, 8, 12, 25, 25, "dilated_grsl", "multinomial",
                                                    "13,17,21", "acc", "training"], tmp_path)
    assert "Optimization Finished!" in log and "Validation: Overall Accuracy=" in log
    assert (out / "model-12.npz").exists() and (out / "patch_acc_loss_step_12.npy").exists()
    occ = np.load(out / "patch_occur_step_12.npy")
    assert occ.sum() == 12 and len(occ) == 9
    log = _run("isprs_dilated_random.py", common + [str(out) + "/model-12", "1,3", "5", 0.01, 0.005, 8, 12, 25, 25,
                                                    "dilated_grsl", "multinomial", "13,17,21", "acc", "validate_test"], tmp_path)
    assert "Test ALL MAPS: Overall Accuracy=" in log
    _run("isprs_dilated_random.py", common + [str(out) + "/model-12", "1,3", "5", 0.01, 0.005, 8, 12, 25, 25, "dilated_grsl",
                                              "single_fixed", "25", "acc", "generate_final_maps"], tmp_path)
    assert any(f.startswith("top_mosaic_09cm_area5_class") for f in os.listdir(out))


def test_contest_cli_train_with_unlabelled_pixels(tmp_path):
    rs = np.random.RandomState(1)
    data = tmp_path / "contest"
    out = tmp_path / "out"
    data.mkdir()
    out.mkdir()
    for name, (h, w) in (("train", (130, 110)), ("test", (120, 100))):
        img, lab = _scene(rs, h, w, 3, 7, np.float32, extra_label=7)
        np.save(data / (name + "_image.npy"), img)
        np.save(data / (name + "_labels.npy"), lab)
    log = _run("contest_dilated_random.py", [str(data) + "/", str(out) + "/", "", 0.01, 0.005, 8, 10, 25, 10,
                                             "dilated_grsl_rate8", "multi_fixed", "13,19,25", "loss", "train"], tmp_path)
    assert "Optimization Finished!" in log and "-- Test: Overall Accuracy=" in log
    assert (out / "model-10.npz").exists()


def test_coffee_cli_train(tmp_path):
    rs = np.random.RandomState(2)
    out = tmp_path / "out"
    out.mkdir()
    for name in ("train", "test"):
        d = tmp_path / name
        d.mkdir()
        tiles = [_scene(rs, 100, 100, 3, 2, np.float32) for _ in range(2)]
        np.save(d / "tiles_image.npy", np.stack([t[0] for t in tiles]))
        np.save(d / "tiles_labels.npy", np.stack([t[1] for t in tiles]))
    log = _run("coffee_dilated_random.py", [str(tmp_path / "train"), str(tmp_path / "test"), str(out) + "/", "", 0.01, 0.005, 8, 10,
                                            25, 15, "dilated_icpr_rate6_densely", "uniform", "15,21", "acc"], tmp_path)
    assert "Optimization Finished!" in log and "-- Test: Overall Accuracy=" in log
    assert (out / "errorAcc_step_10.npy").exists()


def test_isprs_cli_two_ranks_when_two_gpus(tmp_path):
    """Data-parallel training + stripe-sharded validate_test under torchrun (skipped on a single-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rs = np.random.RandomState(3)
    data = tmp_path / "vaihingen"
    out = tmp_path / "out"
    data.mkdir()
    out.mkdir()
    for name in ("1", "3", "5"):
        img, lab = _scene(rs, 150, 175, 4, 6, np.float64)
        np.save(data / (name + "_image.npy"), img)
        np.save(data / (name + "_labels.npy"), lab)
    env = dict(os.environ, PYTHONPATH=ROOT)
    base = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
            "--master-port", "29533", os.path.join(ROOT, "isprs_dilated_random.py"), str(data) + "/", str(out) + "/"]
    tail = ["1,3", "5", "0.01", "0.005", "8", "10", "25", "25", "dilated_icpr_rate6_densely", "uniform", "13,17", "loss"]
    r = subprocess.run(base + [""] + tail + ["training"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "Optimization Finished!" in r.stdout and (out / "model-10.npz").exists()
    r = subprocess.run(base + [str(out) + "/model-10"] + tail + ["validate_test"], cwd=tmp_path, env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "Test ALL MAPS: Overall Accuracy=" in r.stdout


def test_data_parallel_sync_bn_equals_single_process_when_two_gpus():
    """2-rank sharded training step with sync_bn == single-process step on the whole batch (fp32 exact-order mode)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, PYTHONPATH=ROOT)
    for net in ("dilated_grsl", "dilated_icpr_rate6_densely"):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                            "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dp_parity.py"), net, "fp32"],
                           env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "DP_PARITY ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
