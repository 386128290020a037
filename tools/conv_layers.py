"""Per-layer conv kernel times of one inference chunk from an ncu launch list."""
import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from launch_summary import load
seq = load(sys.argv[1])
g = [i for i, s in enumerate(seq) if 'gather_kernel' in s[0]]
a, b = g[3], g[4]
print(" ".join("%s=%.0f" % (n[:8], us) for n, us, _ in seq[a:b] if 'conv' in n))
