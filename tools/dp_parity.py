"""Data-parallel parity (run under torchrun, 2+ ranks): with sync_bn the sharded step equals the single-process step on the
whole batch.  Prints 'DP_PARITY ok' on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import drs_b200
from drs_b200 import dist as ddist, nets

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
net, C, K, GB, crop = sys.argv[1] if len(sys.argv) > 1 else "dilated_grsl", 4, 6, 8, 19
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
comm = sys.argv[3] if len(sys.argv) > 3 else "nccl"      # nccl: the library's own communicator; torch: callback exchange
rs = np.random.RandomState(5)
x = rs.randn(GB, crop * crop * C).astype(np.float32)
y = rs.randint(0, K, size=(GB, crop * crop)).astype(np.float32)
variables = nets.initial_variables(net, C, K, seed=3)


def run(world_size, rows, sync_bn, perturb=0.0):
    s = drs_b200.Session(net, C, K, precision=prec, device=local, weight_decay=0.005, lr_initial=0.01)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    s.load_variables(variables)
    if world_size > 1:
        if comm == "nccl":
            ddist.attach_nccl(s, sync_bn=sync_bn)
        else:
            ddist.attach_allreduce(s, sync_bn=sync_bn)
    xs = x[rows] if perturb == 0.0 else (x[rows] * (1.0 + perturb * np.random.RandomState(9).randn(*x[rows].shape))).astype(np.float32)
    xd = torch.from_numpy(xs).cuda()
    yd = torch.from_numpy(y[rows]).cuda()
    B = xd.shape[0]
    cm = torch.zeros(K * K + 1, dtype=torch.int32, device="cuda")
    losses = []
    for _ in range(2):
        losses.append(float(s.train_step_dev(xd, yd, B, crop, cm_dev=cm)))
    if os.environ.get("DP_DEBUG"):
        print("rank", rank, "world", world_size, "losses", losses, flush=True)
    out = (losses, None, None, None, cm.cpu().numpy().copy(), s.variables())
    s.close()
    return out

per = GB // world
dp = run(world, slice(rank * per, (rank + 1) * per), True)
if rank == 0:
    ref = run(1, slice(0, GB), False)
    # how ill-conditioned is this step?  the same single-process step with the input perturbed by 1e-7 relative -- the size of
    # the only difference data parallelism introduces (summation order of the BN statistics): in the pooling nets a handful of
    # max-pool winners / LeakyReLU signs flip and the early-layer gradients, heavily cancelling sums over all pixels, move by
    # per cents (the PyTorch-CPU oracle shows the same: fp32 vs fp64 of one graph differ by 0.6-2 % there)
    ptb = run(1, slice(0, GB), False, perturb=1e-7)
    tol = 2e-5 if prec == "fp32" else 3e-2
    assert all(abs(a - b) < tol * max(1, abs(b)) for a, b in zip(dp[0], ref[0])), (dp[0], ref[0])
    worst, worst_p = ("", 0.0), ("", 0.0)
    for name, b in ref[5].items():          # every variable: weights, biases, moving statistics, momentum slots, global_step
        den = np.abs(b).max() + 1e-30
        err = float(np.abs(dp[5][name] - b).max() / den)
        err_p = float(np.abs(ptb[5][name] - b).max() / den)
        if err > worst[1]:
            worst = (name, err)
        if err_p > worst_p[1]:
            worst_p = (name, err_p)
        bound = max(5e-4 if prec == "fp32" else 5e-2, 3.0 * err_p)
        assert err < bound, (name, err, "bound", bound, "1e-7 perturbation moves it by", err_p)
    print("DP_PARITY ok", prec, net, comm, "world", world, "losses", dp[0], ref[0], "worst variable", worst,
          "| a 1e-7 input perturbation of the single-process step moves", worst_p, "| variables", len(ref[5]), flush=True)
dist.barrier()
dist.destroy_process_group()
