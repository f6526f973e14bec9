"""profile_step.py for the P2 / guided-diffusion U-Net (BASELINE config 2): one rank-3 JVP pass, one rank-3
VJP pass, one B=1 and one B=8 forward."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import P2_256, random_state_dict

dev = torch.device("cuda:0")
net = B200UNet(P2_256, random_state_dict(P2_256, seed=1234), device=dev)
k = 3
x = torch.randn(1 + k, 3, 256, 256, device=dev)
g = torch.randn(k, 3, 256, 256, device=dev)
p, p1, p8 = net.plan(1, k, k), net.plan(1), net.plan(8)
x8 = torch.randn(8, 3, 256, 256, device=dev)
p.forward(x, 199.8)
p.vjp(g)
p1.forward(x[:1].contiguous(), 199.8)
p8.forward(x8, 199.8)
torch.cuda.synchronize()
print("done")
