"""One scene pass of dilated_grsl_rate8 (f16) on a 1500x1500x5 Potsdam-shaped tile: used under ncu to capture the kernels of an
inference chunk (0.76 M patch-pixels per chunk, as in the full 6000x6000 pass)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import drs_b200
from drs_b200 import synth
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
img, _ = synth.scene("potsdam", H=H, W=W)
mean, std = synth.normalisation(img)
s = drs_b200.Session("dilated_grsl_rate8", 5, 6, precision="f16", seed=9)
s.set_stream(torch.cuda.current_stream().cuda_stream)
s.set_normalization(mean, std)
s.upload_scene(0, img, None)
for _ in range(2):
    lab = s.scene_infer(0, 25, 64, H, W)
torch.cuda.synchronize()
print("done", lab.shape)
