"""Stress of concurrent plan slots at full size: three streams replay Jacobian passes of the DDPM-256 U-Net with
different ranks (so their split-K grids differ in size and interleave differently) plus a B = 1 forward chain on a
fourth stream.  A split-K exchange that starves would hit the kernel's wait bound (a launch error); results are compared
with the sequential ones."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import DDPM256, random_state_dict

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=1234), device=dev)
ks = (10, 5, 3)
gen = torch.Generator(device=dev).manual_seed(5)
plans = [net.plan(1, k, k, slot=i) for i, k in enumerate(ks)]
p1 = net.plan(1)
xs = [torch.randn(1 + k, 3, 256, 256, device=dev, generator=gen) for k in ks]
gs = [torch.randn(k, 3, 256, 256, device=dev, generator=gen) for k in ks]
x1 = torch.randn(1, 3, 256, 256, device=dev, generator=gen)
ref = []
for i in range(len(ks)):
    e = plans[i].forward(xs[i], 595.3636).clone()
    ref.append((e, plans[i].vjp(gs[i]).clone()))
r1 = p1.forward(x1, 595.3636).clone()
torch.cuda.synchronize()
streams = [torch.cuda.Stream(dev) for _ in range(len(ks) + 1)]
for s_ in streams:
    s_.wait_stream(torch.cuda.current_stream(dev))
reps = int(os.environ.get("REPS", "60"))
t0 = time.time()
outs = [None] * len(ks)
for rep in range(reps):
    for i in range(len(ks)):
        with torch.cuda.stream(streams[i]):
            e = plans[i].forward(xs[i], 595.3636)
            outs[i] = (e, plans[i].vjp(gs[i]))
    with torch.cuda.stream(streams[-1]):
        for _ in range(4):
            o1 = p1.forward(x1, 595.3636)
torch.cuda.synchronize()
rel = lambda a, b: float((a - b).norm() / b.norm())
res = {"reps": reps, "ranks": ks, "wall_s": time.time() - t0,
       "rel_err_eps": [rel(outs[i][0], ref[i][0]) for i in range(len(ks))],
       "rel_err_vjp": [rel(outs[i][1], ref[i][1]) for i in range(len(ks))],
       "rel_err_b1": rel(o1, r1)}
print(json.dumps(res))
assert max(res["rel_err_eps"] + res["rel_err_vjp"] + [res["rel_err_b1"]]) < 5e-3
