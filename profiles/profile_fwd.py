"""ncu workload: plain forwards (the DDIM-loop programs) of the DDPM-256 U-Net at the batch sizes an
edit uses (1: inversion / forward-to-t of a single edit, 5: its final stage, 8 and 40: the batch-edit
bench), on fp16 plans (default) or tf32 plans (HALF=0).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv \
        python profiles/profile_fwd.py
    python profiles/summarize_launches.py gpurun_out/x.csv fwd_b1 fwd_b5 fwd_b8 fwd_b40
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import DDPM256, random_state_dict

dev = torch.device("cuda:0")
half = os.environ.get("HALF", "1") != "0"
net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=1234), device=dev)
for b in [int(v) for v in os.environ.get("BATCHES", "1,5,8,40").split(",")]:
    p = net.plan(b, half=half)
    x = torch.randn(b, 3, 256, 256, device=dev)
    for _ in range(int(os.environ.get("REPS", "1"))):
        p.forward(x, 595.3636)
torch.cuda.synchronize()
print("done")
