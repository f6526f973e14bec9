"""Device time of the fused Jacobian passes of the DDPM-256 U-Net (primal + k tangents, k cotangents) for k = 5, 10.
Used for A/B runs of plan-level switches (LOCO_JVP_STATS=0|1, ...)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import DDPM256, random_state_dict

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=1234), device=dev)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("LOCO_")}}
g = torch.Generator(device=dev).manual_seed(5)
for k in (5, 10):
    plan = net.plan(1, k, k)
    xin = torch.randn(1 + k, 3, 256, 256, device=dev, generator=g)
    gin = torch.randn(k, 3, 256, 256, device=dev, generator=g)
    for _ in range(3):
        y = plan.forward(xin, 595.3636); plan.vjp(gin)
    torch.cuda.synchronize()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    reps = 10
    a.record()
    for _ in range(reps):
        plan.forward(xin, 595.3636)
    b.record()
    for _ in range(reps):
        plan.vjp(gin)
    c.record()
    torch.cuda.synchronize()
    out["k%d" % k] = {"jvp_pass_ms": a.elapsed_time(b) / reps, "vjp_pass_ms": b.elapsed_time(c) / reps,
                      "checksum": float(y.double().abs().sum())}
print(json.dumps(out))
