"""CPU, world_size 2 over gloo: the host-side logic of the two shardings.

The device path cannot run here (no GPU, no CPU fallback), so the oracle stands in for the per-rank compute:
  * inference: each rank evaluates the patches that intersect its row stripe in the reference's visiting order and
    the gathered stripes must equal the un-sharded label map bit for bit (no halo exchange needed);
  * training: both ranks draw the same patch size / batch from the same seeds, take disjoint rank slices, and one
    sum-allreduce over [gradients ++ loss ++ confusion] leaves identical variables on every rank.
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stripe_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import drs_b200  # noqa: F401
    from drs_b200 import dist as ddist
    from oracle import host_np
    from oracle.fake_net import fake_logits
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(5)
    H, W, C, K, crop, batch = 83, 70, 4, 6, 25, 16
    scene = rs.randint(0, 256, size=(H, W, C)).astype(np.uint8) / 255.0
    mean, std = np.full(4, 0.5), np.full(4, 0.29)
    pos = host_np.all_patch_positions(H, W, crop, batch, "isprs")

    def labels_for(r0, r1):
        mine = [(r, c) for r, c in pos if r < r1 and r + crop > r0]          # order preserved
        x, _ = host_np.apply_plan([scene], None, [(0, r, c) for r, c in mine], None, crop, mean, std, cast=False)
        lg = fake_logits(x.reshape(len(mine), -1), crop, C, K)
        return host_np.accumulate_argmax(lg, mine, H, W, crop)[r0:r1].astype(np.uint8)

    r0, r1 = ddist.stripe_bounds(H, world, rank)
    full = ddist.gather_label_stripes(labels_for(r0, r1), H, W, rank, world)
    if rank == 0:
        np.save(os.path.join(out_dir, "ok.npy"), np.array([int(np.array_equal(full, labels_for(0, H)))]))
    else:
        assert full is None
    dist.destroy_process_group()


def _train_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    import drs_b200  # noqa: F401
    from drs_b200 import dist as ddist, host
    from oracle import nets_torch
    import random
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    net, C, K, gb = "dilated_icpr_rate6_densely", 4, 6, 4
    orc = nets_torch.OracleNet(net, C, K, nets_torch.init_params(net, C, K, seed=1))
    np.random.seed(11)
    random.seed(11)
    values = [7, 9, 11]
    total = 40
    shuffle = np.asarray(random.sample(range(total), total))
    it = 0
    data = np.random.RandomState(3).randn(total, 11 * 11 * C).astype(np.float32)
    log = []
    for step in range(3):
        crop, idx = host.draw_patch_size("multi_fixed", values)
        shuffle, batch, it = host.select_batch(shuffle, gb, it, total)
        mine = ddist.rank_slice(batch, rank, world)
        log.append((crop, list(batch), list(mine)))
        x = torch.from_numpy(data[mine][:, :crop * crop * C])
        y = torch.from_numpy((data[mine][:, :crop * crop] > 0).astype(np.float32))
        names = orc.trainable()
        leaf = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in orc.p.items()}
        logits = orc.forward(x, crop, True, p=leaf)
        ce = orc.loss(logits, y, 0.0, p=leaf)
        grads = torch.autograd.grad(ce, [leaf[k] for k in names], allow_unused=True)
        flat = torch.cat([(g if g is not None else torch.zeros_like(orc.p[k])).reshape(-1) for k, g in zip(names, grads)] +
                         [ce.detach().reshape(1)])
        dist.all_reduce(flat)                      # the ONE exchange step of the data-parallel path
        flat /= world
        off = 0
        with torch.no_grad():
            for k in names:
                n = orc.p[k].numel()
                orc.p[k] -= 0.01 * flat[off:off + n].reshape(orc.p[k].shape)
                off += n
    np.save(os.path.join(out_dir, "w%d.npy" % rank), orc.p["conv_classifier/weights"].numpy())
    np.save(os.path.join(out_dir, "log%d.npy" % rank), np.array([(c, sum(b), sum(m)) for c, b, m in log]))
    dist.destroy_process_group()


def _spawn(fn, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(fn, args=(2, port, str(tmp_path)), nprocs=2, join=True)


@pytest.mark.timeout(300)
def test_stripe_sharded_inference_is_bit_exact(tmp_path):
    _spawn(_stripe_worker, tmp_path)
    assert int(np.load(tmp_path / "ok.npy")[0]) == 1


@pytest.mark.timeout(300)
def test_data_parallel_training_keeps_ranks_identical(tmp_path):
    _spawn(_train_worker, tmp_path)
    w0, w1 = np.load(tmp_path / "w0.npy"), np.load(tmp_path / "w1.npy")
    assert np.array_equal(w0, w1)
    l0, l1 = np.load(tmp_path / "log0.npy"), np.load(tmp_path / "log1.npy")
    assert np.array_equal(l0[:, :2], l1[:, :2])            # same patch size and same global batch on both ranks
    assert np.array_equal(l0[:, 2] + l1[:, 2], l0[:, 1])   # disjoint rank slices cover the global batch


def test_stripe_bounds_cover_the_scene():
    import drs_b200  # noqa: F401
    from drs_b200 import dist as ddist
    for H in (1, 7, 83, 6000):
        for world in (1, 2, 3, 8):
            cuts = ddist.stripe_bounds(H, world)
            assert cuts[0] == 0 and cuts[-1] == H and all(a <= b for a, b in zip(cuts[:-1], cuts[1:]))
            assert [ddist.stripe_bounds(H, world, r) for r in range(world)] == list(zip(cuts[:-1], cuts[1:]))


def _rng_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import random
    import torch.distributed as dist
    import drs_b200  # noqa: F401
    from drs_b200 import dist as ddist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    # (1) a harness seeded every rank alike: nothing may be touched (the draws stay those of the single-process run)
    np.random.seed(5); random.seed(5)
    res.append(int(ddist.sync_host_rng()))
    res.append(int(np.random.randint(0, 1 << 30)))
    # (2) unseeded / diverged ranks (rank 0 created a cache file and consumed draws the others did not): re-seeded from rank 0
    np.random.seed(100 + rank); random.seed(200 + rank)
    res.append(int(ddist.sync_host_rng()))
    res += [int(np.random.randint(0, 1 << 30)), int(random.randrange(1 << 30))]
    # (3) only the python stream differs
    random.seed(300 + rank)
    res.append(int(ddist.sync_host_rng()))
    res += [int(np.random.randint(0, 1 << 30)), int(random.randrange(1 << 30))]
    # (4) cache files: rank 0 creates (atomically), the others wait and load -- one table for everybody
    cache = os.path.join(out_dir, "table.npy")

    def table():
        if os.path.isfile(cache):
            return np.load(cache)
        t = np.random.RandomState(rank + 1).randint(0, 360, size=1000)
        np.save(cache + ".tmp.npy", t)
        os.replace(cache + ".tmp.npy", cache)
        return t

    t = ddist.rank0_first(table)
    res.append(int(t.sum()))
    # (5) the per-step plan check
    ddist.check_same_plan(37, 12345)
    try:
        ddist.check_same_plan(37 + rank, 12345)
        res.append(0)
    except RuntimeError:
        res.append(1)
    np.save(os.path.join(out_dir, "rng_%d.npy" % rank), np.array(res, dtype=np.int64))
    dist.destroy_process_group()


def test_host_rng_sync_and_cache_files_world2(tmp_path):
    """Data-parallel host rules: same streams on every rank, cache files created once, plan agreement checked."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_rng_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rng_0.npy"), np.load(tmp_path / "rng_1.npy")
    assert a[0] == 0 and b[0] == 0 and a[1] == b[1]                  # already equal: untouched
    np.random.seed(5)
    assert a[1] == int(np.random.randint(0, 1 << 30))                # ... and still the single-process draw
    assert a[2] == 1 and b[2] == 1 and a[3] == b[3] and a[4] == b[4]  # diverged: re-seeded, both streams equal afterwards
    assert a[5] == 1 and b[5] == 1 and a[6] == b[6] and a[7] == b[7]
    assert a[8] == b[8]                                              # one cache table
    assert a[9] == 1 and b[9] == 1                                   # disagreement detected on every rank
