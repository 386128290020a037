"""One patch size, steady state: used under ncu to list the kernels of a training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import drs_b200
B, C, K = 64, 4, 6
crop = int(sys.argv[1]) if len(sys.argv) > 1 else 37
net = sys.argv[2] if len(sys.argv) > 2 else "dilated_grsl"
s = drs_b200.Session(net, C, K, precision="bf16", seed=1)
s.set_stream(torch.cuda.current_stream().cuda_stream)
x = torch.randn(B * crop * crop * C, device="cuda")
y = torch.randint(0, K, (B * crop * crop,), device="cuda").float()
pred = torch.empty(B * crop * crop, dtype=torch.uint8, device="cuda")
cm = torch.zeros(K * K + 1, dtype=torch.int32, device="cuda")
for _ in range(4):
    s.train_step_dev(x, y, B, crop, pred_dev=pred, cm_dev=cm)
torch.cuda.synchronize()
print("done")
