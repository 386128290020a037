"""Time the training step eager vs CUDA-graph replay for one patch size (steady state)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import drs_b200
B, C, K = 64, 4, 6
for crop in (25, 37, 49):
    for mode in ("eager", "graph", "graph+prof"):
        if mode == "eager":
            os.environ.pop("DRS_GRAPHS", None)
        else:
            os.environ["DRS_GRAPHS"] = "1"
        s = drs_b200.Session("dilated_grsl", C, K, precision="bf16", seed=1)
        st = torch.cuda.Stream()
        torch.cuda.set_stream(st)
        s.set_stream(st.cuda_stream)
        s.set_profiling(mode == "graph+prof")
        x = torch.randn(B * crop * crop * C, device="cuda")
        y = torch.randint(0, K, (B * crop * crop,), device="cuda").float()
        pred = torch.empty(B * crop * crop, dtype=torch.uint8, device="cuda")
        cm = torch.zeros(K * K + 1, dtype=torch.int32, device="cuda")
        for _ in range(5):
            s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 30
        for _ in range(n):
            s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n * 1e3
        print("crop %d %-10s %.3f ms/step" % (crop, mode, dt), flush=True)
        s.close()
