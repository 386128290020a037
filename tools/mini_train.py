"""Localise a backward mismatch: compare per-layer dA / dZ / dW of one fp32 train step with autograd, and report
whether each mismatching element sits on an activation kink (|x_hat| ~ 0) or a max-pool tie in the oracle."""
import os, sys
os.environ["DRS_DEBUG_KEEP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import drs_b200
from oracle import nets_torch
net = sys.argv[1] if len(sys.argv) > 1 else "dilated_icpr_original"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
crop = int(sys.argv[4]) if len(sys.argv) > 4 else 13
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 6
C, K = 4, 6
params = nets_torch.init_params(net, C, K, seed=5)
o = nets_torch.OracleNet(net, C, K, params)
s = drs_b200.Session(net, C, K, precision=prec)
s.load_variables(params)
rs = np.random.RandomState(seed)
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
print("net", net, prec, "B", B, "crop", crop, "loss", lg, float(loss))
pool = o.spec["pool"]
for i, sc in enumerate(scopes):
    co = o.plan[i][4]
    z = zt[sc].detach()
    mean = z.mean(dim=(0, 2, 3), keepdim=True)
    var = z.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    xh = ((z - mean) * torch.rsqrt(var + 1e-3))
    xh64 = ((z.double() - z.double().mean(dim=(0, 2, 3), keepdim=True)) * torch.rsqrt(z.double().var(dim=(0, 2, 3), unbiased=False, keepdim=True) + 1e-3))
    xh_n = xh.permute(0, 2, 3, 1).numpy()
    near = int((np.abs(xh_n) < 1e-5).sum())
    ties = 0
    if pool:
        a = torch.maximum(0.1 * xh, xh)
        u = F.unfold(F.pad(a, (1, 1, 1, 1), value=-1e30), 3).reshape(B, co, 9, -1)
        top2 = torch.topk(u, 2, dim=2).values
        ties = int(((top2[:, :, 0] - top2[:, :, 1]) < 1e-6).sum())
    ref = dzs[i].permute(0, 2, 3, 1).numpy()
    got = s.debug_activation("dz:" + sc, B, crop, co)
    den = np.abs(ref).max()
    err = np.abs(got - ref) / den
    wref = dzs[len(scopes) + i].numpy()
    wgot = s.get_gradient(sc + "/weights", wref.shape)
    we = np.abs(wgot - wref).max() / np.abs(wref).max()
    bad = np.argwhere(err > 1e-3)
    print("%-12s dZ max rel-err %.3e  bad elems %d/%d   dW max rel-err %.3e   oracle: |xh|<1e-5: %d  pool ties<1e-6: %d" %
          (sc, err.max(), len(bad), err.size, we, near, ties))
    for bidx in bad[:4]:
        t = tuple(bidx)
        print("     bad %s got %.4e ref %.4e  oracle xh(fp32) %.3e xh(fp64) %.3e" %
              (list(t), got[t], ref[t], xh_n[t], float(xh64.permute(0, 2, 3, 1)[t])))
