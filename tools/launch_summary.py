"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv [start_kernel_substring [occurrence [count]]]"""
import collections, csv, re, sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v / 1000 if u.startswith("n") else (v if u.startswith("u") else v * 1000)
        n = re.sub(r"\(.*", "", row["Kernel Name"])
        n = re.sub(r"<.*", "", n).replace("void ", "")
        seq.append((n, us, row["Grid Size"]))
    return seq


def agg(seq, a, b, title):
    d = collections.defaultdict(lambda: [0, 0.0])
    for n, us, g in seq[a:b]:
        d[n][0] += 1
        d[n][1] += us
    tot = sum(v[1] for v in d.values())
    print("== %s: launches %d, kernel time %.1f us" % (title, b - a, tot))
    for n, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print("  %-34s n=%4d  %10.1f us  %5.1f%%" % (n, c, t, 100 * t / tot))


if __name__ == "__main__":
    seq = load(sys.argv[1])
    g = [i for i, s in enumerate(seq) if "gather_kernel" in s[0]]
    print("gather launches at", g[:14])
    # training steps: between consecutive gathers while the gap looks like a train step
    for j in (3, 4, 5):
        if j + 1 < len(g):
            agg(seq, g[j], g[j + 1], "train step %d" % j)
    inf = [i for i in g if i > g[5] + 10]
    if inf:
        agg(seq, inf[1] if len(inf) > 1 else inf[0], len(seq), "inference (timed part)")
