"""Per-kernel time of the LAST training step in an ncu launch list of tools/train_probe.py (4 steps, no gather launches).
usage: python tools/kernel_times.py launches.csv [launches_per_step]"""
import collections, csv, re, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from launch_summary import load, agg
seq = load(sys.argv[1])
# a step starts at the first kernel of the forward: find the period by the positions of ce_fwd_bwd_kernel
pos = [i for i, s in enumerate(seq) if "ce_fwd_bwd" in s[0]]
per = pos[-1] - pos[-2]
start = len(seq) - per
agg(seq, start, len(seq), "last step (%d launches)" % per)
if len(sys.argv) > 2:
    for n, us, g in seq[start:]:
        print("   %-44s %8.1f us  grid %s" % (n, us, g))
