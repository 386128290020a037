"""Localise a backward mismatch: compare per-layer dA / dZ / dW of one fp32 train step with autograd."""
import os, sys
os.environ["DRS_DEBUG_KEEP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import drs_b200
from oracle import nets_torch
net = sys.argv[1] if len(sys.argv) > 1 else "dilated_icpr_original"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
crop = int(sys.argv[4]) if len(sys.argv) > 4 else 13
C, K = 4, 6
params = nets_torch.init_params(net, C, K, seed=5)
o = nets_torch.OracleNet(net, C, K, params)
s = drs_b200.Session(net, C, K, precision=prec)
s.load_variables(params)
rs = np.random.RandomState(6)
x = rs.randn(B, crop * crop * C).astype(np.float32)
y = rs.randint(0, K, size=(B, crop * crop)).astype(np.float32)
names = o.trainable()
leaf = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in o.p.items()}
zt = {}
logits = o.forward(torch.from_numpy(x), crop, True, p=leaf, update_stats=False, ztaps=zt)
loss = o.loss(logits, torch.from_numpy(y), 0.005, p=leaf)
scopes = [p_[0] for p_ in o.plan]
dzs = torch.autograd.grad(loss, [zt[sc] for sc in scopes] + [leaf[sc + "/weights"] for sc in scopes])
lg = s.train_step(x, y, crop)[0]
print("loss", lg, float(loss))
for i, sc in enumerate(scopes):
    co = o.plan[i][4]
    ref = dzs[i].permute(0, 2, 3, 1).numpy()
    got = s.debug_activation("dz:" + sc, B, crop, co)
    den = np.abs(ref).max()
    e = np.abs(got - ref).reshape(-1, co).max(0) / den
    wref = dzs[len(scopes) + i].numpy()
    wgot = s.get_gradient(sc + "/weights", wref.shape)
    we = np.abs(wgot - wref).reshape(-1, co).max(0) / np.abs(wref).max()
    print("%-12s dZ rel-err max %.3e (worst ch %s)   dW rel-err max %.3e (worst ch %s)" %
          (sc, e.max(), np.argsort(-e)[:4].tolist(), we.max(), np.argsort(-we)[:4].tolist()))
    if e.max() > 1e-3:
        c_ = int(np.argmax(e))
        bad = np.argwhere(np.abs(got[..., c_] - ref[..., c_]) > 1e-3 * den)
        print("     channel %d: %d bad pixels of %d, first %s" % (c_, len(bad), B * crop * crop, bad[:6].tolist()))
        print("     got", got[tuple(bad[0])][c_], "ref", ref[tuple(bad[0])][c_])
