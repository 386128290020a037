"""Deterministic fingerprint + timing of a few training steps: A/B two builds or two environment settings.

    python tools/step_hash.py [net] [crop,crop,...] [steps]

Prints a SHA-1 of all variables after `steps` steps per patch size (bit-identity check) and ms/step of 20 further steps."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import drs_b200
net = sys.argv[1] if len(sys.argv) > 1 else "dilated_grsl"
crops = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "25,37,49").split(",")]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
B, C, K = 64, 4, 6
tag = os.environ.get("AB_TAG", "")
for crop in crops:
    s = drs_b200.Session(net, C, K, precision="bf16", seed=1)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(crop)
    x = torch.randn(B * crop * crop * C, device="cuda", generator=g)
    y = torch.randint(0, K, (B * crop * crop,), device="cuda", generator=g).float()
    pred = torch.empty(B * crop * crop, dtype=torch.uint8, device="cuda")
    cm = torch.zeros(K * K + 1, dtype=torch.int32, device="cuda")
    losses = [s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm) for _ in range(steps)]
    torch.cuda.synchronize()
    hsh = hashlib.sha1()
    for name, _ in sorted(s.variable_names()):
        hsh.update(np.ascontiguousarray(s.get_variable(name)).tobytes())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm, want_loss=False)
    e1.record()
    torch.cuda.synchronize()
    print("%s %s crop %d: sha1 %s loss %.6f  %.3f ms/step" % (tag, net, crop, hsh.hexdigest()[:16], losses[-1], e0.elapsed_time(e1) / 20), flush=True)
    s.close()
