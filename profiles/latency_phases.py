"""Single-edit latency by phase (CUDA events on the current stream + host wall clock per phase): DDIM inversion (98 B = 1
forwards), forward to t (40), the two local bases (12 fused rank-10 iterations), projection + edit batch, final stage
(59 B = 5 forwards).  A phase whose host time is close to its device time is launch-bound."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from loco_edit_b200 import ops
from loco_edit_b200.edit import local_basis_pair
from loco_edit_b200.masks import rectangle_mask
from loco_edit_b200.pipeline import EditPipeline
from loco_edit_b200.unet import B200UNet
from loco_edit_b200.weights import DDPM256, random_state_dict

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=1234), device=dev)
pipe = EditPipeline(net)
drv, sched = pipe.driver, pipe.driver.scheduler
gen = torch.Generator(device=dev).manual_seed(1)
x0 = (0.5 * torch.randn(1, 3, 256, 256, device=dev, generator=gen)).clamp(-1, 1)
mask = rectangle_mask(256).to(dev)


def run(record):
    marks = []

    def mark(name):
        if record:
            e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))
    mark("start")
    sched.set_timesteps(drv.inv_steps, device=dev, is_inversion=True)
    xt = x0
    n = len(sched._ts_host)
    for i in range(n - 1):
        t = sched._ts_host[i]
        xt = sched.step(net(xt, t), t, xt, eta=0, t_idx=i).prev_sample
    mark("inversion_98")
    xt, t, t_idx = drv.DDIMforwardsteps(xt, t_start_idx=0, t_end_idx=drv.edit_t_idx, save_image=False)
    mark("forward_to_t_40")
    vm, _, vn, _ = local_basis_pair(net, sched, xt, sched._ts_host[t_idx], 5, 5, mask, v0=pipe._v0(5, gen),
                                    v0_null=pipe._v0(5, gen), n_iter=12)
    mark("bases_12_iterations")
    vT = ops.nullspace_project(vm, vn, project=True)
    batch = drv.build_edit_batch(xt, vT[0], 2)
    mark("project_and_edit_batch")
    drv.noise_fn = lambda i, x: torch.randn(x.shape, device=x.device, dtype=x.dtype, generator=gen)
    imgs = drv.DDIMforwardsteps(batch, t_start_idx=drv.edit_t_idx, t_end_idx=-1, save_image=False, performance_boosting=True)
    mark("final_59_b5")
    host = imgs.cpu()
    mark("d2h")
    torch.cuda.synchronize()
    return marks


for _ in range(2):
    run(False)
torch.cuda.synchronize()
m = run(True)
out = {}
for (n0, e0, h0), (n1, e1, h1) in zip(m[:-1], m[1:]):
    out[n1] = {"device_ms": e0.elapsed_time(e1), "host_enqueue_ms": (h1 - h0) * 1e3}
out["total_device_ms"] = m[0][1].elapsed_time(m[-1][1])
p1, p5 = net.plan(1), net.plan(5)
x5 = torch.randn(5, 3, 256, 256, device=dev)
for pl, xx, nm in ((p1, x0, "fwd_b1_ms"), (p5, x5, "fwd_b5_ms")):
    pl.forward(xx, 300.0); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        pl.forward(xx, 300.0)
    b.record(); torch.cuda.synchronize()
    out[nm] = a.elapsed_time(b) / 10
print(json.dumps(out, indent=1))
