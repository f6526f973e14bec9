"""Short profiling workload for ncu: one fused rank-5 JVP pass, one rank-5 VJP pass and one plain
B=1 forward and one B=8 forward of the DDPM-256 U-Net (the launch programs every edit is made of).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import DDPM256, random_state_dict

dev = torch.device("cuda:0")
net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=1234), device=dev)
k = int(os.environ.get("K", "5"))
x = torch.randn(1 + k, 3, 256, 256, device=dev)
g = torch.randn(k, 3, 256, 256, device=dev)
p = net.plan(1, k, k)
p1 = net.plan(1)
p8 = net.plan(8)
x8 = torch.randn(8, 3, 256, 256, device=dev)
reps = int(os.environ.get("REPS", "1"))
for _ in range(reps):
    p.forward(x, 595.3636)
    p.vjp(g)
    p1.forward(x[:1].contiguous(), 595.3636)
    p8.forward(x8, 595.3636)
torch.cuda.synchronize()
print("done")
