"""Does a B = 8 forward chain run faster as two B = 4 chains on two streams (own plan slots)?  Device time per step."""
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
out = {}
for B, parts in ((8, 1), (8, 2), (8, 4), (40, 1), (40, 2)):
    b = B // parts
    plans = [net.plan(b, slot=i) for i in range(parts)]
    xs = [torch.randn(b, 3, 256, 256, device=dev) for _ in range(parts)]
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    main = torch.cuda.current_stream(dev)

    def step():
        for s_ in streams:
            s_.wait_stream(main)
        for i in range(parts):
            with torch.cuda.stream(streams[i]):
                for _ in range(4):
                    plans[i].forward(xs[i], 595.3636)
        for s_ in streams:
            main.wait_stream(s_)
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    out["B%d_x%d" % (B, parts)] = e0.elapsed_time(e1) / 12
    for p in plans:
        p.release()
    net._plans.clear()
print(json.dumps(out))
