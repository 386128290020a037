"""SASS mnemonic counts per kernel of libdrs.so (evidence that the tensor-core / TMA / PDL instructions are really there).
usage: python tools/sass_summary.py [libdrs.so] > profiles/<round>_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dynamic-rs-segmentation_b200", "libdrs.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
GROUPS = [("UTCHMMA", r"^UTC[HQ]?MMA"), ("UTMALDG.IM2COL", r"^UTMALDG.*IM2COL"), ("UTMALDG (tiled)", r"^UTMALDG(?!.*IM2COL)"),
          ("UTMASTG", r"^UTMASTG"), ("LDTM", r"^LDTM"), ("UTCBAR (commit)", r"^UTCBAR"), ("SYNCS (mbarrier)", r"^SYNCS"),
          ("ELECT", r"^ELECT"), ("R2UR (vector -> uniform register moves)", r"^R2UR"), ("ATOM/RED .64", r"^(ATOMG|REDG|RED|ATOM).*64"), ("HMNMX2/HSET2 (packed bf16)", r"^(HMNMX2|HSET2)"),
          ("HSETP2 (packed code compare)", r"^HSETP2"), ("ACQBULK/PDL (griddepcontrol)", r"^(ACQBULK|PREEXIT|DEPBAR\.LE SB0)")]
names, cur = [], None
counts = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        names.append(cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["instr"] += 1
        for label, pat in GROUPS:
            if re.search(pat, op):
                counts[cur][label] += 1
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print("SASS mnemonic counts per kernel of libdrs.so (cuobjdump -sass, sm_100a; tools/sass_summary.py).  UTCHMMA = tcgen05.mma, UTMALDG = TMA load "
      "(IM2COL = im2col mode), UTMASTG = TMA store, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, "
      "HSETP2 = setp.eq.f16x2 (winner-code compares of the lean pool backward), ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents.\n")
for n, d in zip(names, dem):
    c = counts[n]
    print("%-112s instr %6d | %s" % (d[:112], c["instr"], "  ".join("%s=%d" % (l, c[l]) for l, _ in GROUPS if c[l])))
