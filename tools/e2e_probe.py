"""Time Session.train_step (host feeds/fetches) per step under torchrun; prints per-phase wall times on rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import drs_b200
from drs_b200 import dist as ddist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, C, K, crop = 64, 4, 6, 37
s = drs_b200.Session("dilated_grsl", C, K, precision="bf16", device=local, seed=1)
s.set_stream(torch.cuda.current_stream().cuda_stream)
if world > 1:
    ddist.attach_allreduce(s)
rs = np.random.RandomState(0)
x = torch.from_numpy(rs.randn(B, crop * crop * C).astype(np.float32)).pin_memory()
y = torch.from_numpy(rs.randint(0, K, size=(B, crop * crop)).astype(np.float32)).pin_memory()
xd, yd = x.cuda(), y.cuda()
pred = torch.empty(B * crop * crop, dtype=torch.uint8, device="cuda")
cm = torch.zeros(K * K + 1, dtype=torch.int32, device="cuda")
for name, fn in (("dev", lambda: s.train_step_dev(xd, yd, B, crop, pred_dev=pred, cm_dev=cm)),
                 ("host", lambda: s.train_step(x.numpy(), y.numpy(), crop, want_cm=True))):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    if rank == 0:
        print("world %d %s path: %.3f ms/step" % (world, name, (time.perf_counter() - t0) / 20 * 1e3), flush=True)

# varying patch sizes, as in bench.py's e2e leg
crops = [34, 40, 27, 41, 32, 41, 25, 37, 41, 45, 30, 49, 26, 33]
hb = {}
for c in sorted(set(crops)):
    hb[c] = (rs.randn(B, c * c * C).astype(np.float32), rs.randint(0, K, size=(B, c * c)).astype(np.float32))
for rep in range(2):
    times = []
    for c in crops:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.train_step(hb[c][0], hb[c][1], c, want_cm=True)
        times.append((time.perf_counter() - t0) * 1e3)
    if rank == 0:
        print("world %d varying pass %d: %s" % (world, rep, " ".join("%.1f" % t for t in times)), flush=True)
if world > 1:
    dist.destroy_process_group()
