#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2D_tests.log
timeout 600 python - > gpurun_out/r2D_dropin.json 2> gpurun_out/r2D_dropin.err <<'PY'
import json, sys, torch
sys.path.insert(0, '.')
import bench
from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import DDPM256, random_state_dict
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=1234), device=dev)
print(json.dumps(bench.measure_dropin(net, dev, 256)))
PY
