"""Device timings of the latent-space twin (bench.py's `sd_latent` leg alone)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402

if __name__ == "__main__":
    torch.cuda.set_device(0)
    print(json.dumps(bench.measure_sd_latent(torch.device("cuda:0")), indent=1))
