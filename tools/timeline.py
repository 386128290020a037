"""Kernel timeline of steady-state training steps through CUPTI (torch.profiler): start, duration, stream of every kernel,
gaps on the main stream and overlap with the side streams.  ncu serialises kernels and flushes caches; this does not.

    python tools/timeline.py [crop] [net] [steps]          (DRS_GRAPHS=0/1 both work)"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import drs_b200
crop = int(sys.argv[1]) if len(sys.argv) > 1 else 37
net = sys.argv[2] if len(sys.argv) > 2 else "dilated_grsl"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
B, C, K = int(os.environ.get("TL_B", "64")), int(os.environ.get("TL_C", "4")), 6
s = drs_b200.Session(net, C, K, precision="bf16", seed=1)
s.set_stream(torch.cuda.current_stream().cuda_stream)
x = torch.randn(B * crop * crop * C, device="cuda")
y = torch.randint(0, K, (B * crop * crop,), device="cuda").float()
pred = torch.empty(B * crop * crop, dtype=torch.uint8, device="cuda")
cm = torch.zeros(K * K + 1, dtype=torch.int32, device="cuda")
for _ in range(6):
    s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm, want_loss=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm, want_loss=False)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "drs_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
if not ev:
    print("no kernel records")
    sys.exit(0)
# last step: kernels after the last memset-like boundary = split by the largest start gaps
names = [e["name"] for e in ev]
first = [i for i, n in enumerate(names) if "pack_conv1" in n or "pad_cast8" in n]
start = first[-2] if len(first) >= 2 and "pad_cast8" in names[first[-1]] and first[-1] == first[-2] + 1 else first[-1]
last = ev[start:]
t0 = last[0]["ts"]
streams = sorted({e["args"].get("stream") for e in last})
print("step: %d kernels, %.1f us wall, streams %s" % (len(last), last[-1]["ts"] + last[-1]["dur"] - t0, streams))
main = max(streams, key=lambda st: sum(1 for e in last if e["args"].get("stream") == st))
busy = {st: sum(e["dur"] for e in last if e["args"].get("stream") == st) for st in streams}
print("busy per stream:", {k: round(v, 1) for k, v in busy.items()})
prev_end = None
gap_total = 0.0
for e in last:
    st = e["args"].get("stream")
    tag = "M" if st == main else "s%d" % streams.index(st)
    gap = ""
    if st == main:
        if prev_end is not None:
            g = e["ts"] - prev_end
            gap_total += max(g, 0.0)
            gap = "gap %5.1f" % g
        prev_end = e["ts"] + e["dur"]
    n = e["name"].replace("void ", "")
    n = n[:60]
    print("%8.1f %7.1f %-3s %-60s %s" % (e["ts"] - t0, e["dur"], tag, n, gap))
print("main-stream gaps total %.1f us" % gap_total)
