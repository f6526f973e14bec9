#!/bin/bash
# repeat the tests added in the closing session to see their margins / flakiness
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout 300 python -m pytest tests/test_gpu_unet.py tests/test_gpu_driver.py -q -m gpu -k "concurrent or fused_power or converging_pair" 2>&1 | tail -2 >> gpurun_out/r2F_flaky.log
done
timeout 300 python - >> gpurun_out/r2F_flaky.log 2>&1 <<'PY'
import os, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from gpu_util import principal_angles_deg, rel_err
from loco_edit_b200.pipeline import EditPipeline
from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import random_state_dict
dev = torch.device('cuda:0')
g = torch.load('tests/golden/pullback_tiny.pt', weights_only=False)
net = B200UNet(g['arch'], random_state_dict(g['arch'], seed=g['seed'], perturb_norm=g['perturb_norm']), device=dev)
R = g['arch']['resolution']
gen = torch.Generator().manual_seed(11)
x0 = (0.5 * torch.randn(3, 3, R, R, generator=gen)).clamp(-1, 1).to(dev)
masks = torch.zeros(3, 3, R, R, dtype=torch.bool)
for b in range(3):
    masks[b, :, 4 + 2 * b:16 + 2 * b, 6:22] = True
masks = masks.to(dev)
for rep in range(6):
    outs = []
    for ns in (1, 2):
        pipe = EditPipeline(net, k=2, k_null=2, n_iter=3, basis_streams=ns)
        gg = torch.Generator(device=dev).manual_seed(3)
        outs.append(pipe.edit_batch_device(x0, masks, gen=gg)); torch.cuda.synchronize()
    a, b = outs
    print('rep', rep, 'xt', rel_err(a['xt'], b['xt']), 'ang', max(float(principal_angles_deg(a['vT'][i], b['vT'][i]).max()) for i in range(3)), 'img', rel_err(a['images'], b['images']))
PY
