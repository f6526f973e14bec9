"""Short profiling workload for ncu: one fused rank-5 primal + tangent pass and one rank-5 cotangent pass of
the SD-shaped VAE decoder (latent 4 x 64 x 64 -> image 3 x 512 x 512), the network inside every Jacobian
product of the latent-space twin.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sd.csv \
        python profiles/profile_sd.py
    python profiles/summarize_by_kernel.py gpurun_out/launches_sd.csv
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from loco_edit_b200.unet import B200VAEDecoder
from loco_edit_b200.weights import SD_VAE_DECODER, random_state_dict

dev = torch.device("cuda:0")
vae = B200VAEDecoder(SD_VAE_DECODER, random_state_dict(SD_VAE_DECODER, seed=4321), device=dev)
k = 5
z = torch.randn(1 + k, 4, 64, 64, device=dev)
g = torch.randn(k, 3, 512, 512, device=dev)
p = vae.plan(1, k, k)
p.forward(z, 0.0)
p.vjp(g)
torch.cuda.synchronize()
print("done")
