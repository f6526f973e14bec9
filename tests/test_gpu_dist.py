"""GPU, world_size 2, NCCL: probe-sharded power method on two B200s == single-GPU run.
Skipped on boxes with fewer than two GPUs."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from loco_edit_b200 import dist as ld
        from loco_edit_b200.edit import local_basis
        from loco_edit_b200.scheduler import YHCustomScheduler
        from loco_edit_b200.unet import B200UNet
        from loco_edit_b200.weights import random_state_dict, tiny_arch
        arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
        sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
        net = B200UNet(arch, sd, device=dev)
        sched = YHCustomScheduler(device=dev)
        sched.set_timesteps(100)
        t = sched._ts_host[40]
        g = torch.Generator().manual_seed(3)
        xt = torch.randn(1, 3, 32, 32, generator=g).to(dev)
        mask = torch.zeros(3, 32, 32, dtype=torch.bool)
        mask[:, 12:20, 8:24] = True
        k = 5
        v0, _ = torch.linalg.qr(torch.randn(xt.numel(), k, generator=g))
        v0 = v0.T.contiguous().to(dev)
        u, s, vT = ld.sharded_local_basis_cuda(net, sched, xt, t, k, v0, 3, mask=mask.to(dev))
        torch.cuda.synchronize()
        if rank == 0:
            u1, s1, v1 = local_basis(net, sched, xt, t, k, v0=v0, min_iter=10 ** 6, max_iter=3, mask=mask.to(dev),
                                     verbose=False)
            ret["s_rel"] = float(((s - s1).abs() / s1).max())
            ret["v_dot"] = float((vT * v1).sum(1).abs().min())
            ret["u_shape_ok"] = tuple(u.shape) == tuple(u1.shape)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_probe_sharded_power_method_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29711, ret), nprocs=2, join=True)
    # both runs use the same TF32 kernels; they differ by batch composition only (TF32 noise level)
    assert ret["s_rel"] < 2e-3 and ret["v_dot"] > 0.999 and ret["u_shape_ok"], dict(ret)


def test_runs_on_the_tensors_device_not_the_current_one():
    """`--device cuda:1` without torch.cuda.set_device: every ABI entry point switches to
    the device that owns its buffers, torch's stream of THAT device is used, and the per-device
    one-time setup (kernel attributes, capture stream) happens on both devices of one process."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from loco_edit_b200 import ops
    from loco_edit_b200.edit import local_basis
    from loco_edit_b200.scheduler import YHCustomScheduler
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import random_state_dict, tiny_arch
    torch.cuda.set_device(0)
    arch = tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1)
    sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
    g = torch.Generator().manual_seed(3)
    xt = torch.randn(1, 3, 32, 32, generator=g)
    v0, _ = torch.linalg.qr(torch.randn(xt.numel(), 3, generator=g))
    mask = torch.zeros(3, 32, 32, dtype=torch.bool)
    mask[:, 12:20, 8:24] = True
    res = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        assert torch.cuda.current_device() == 0
        net = B200UNet(arch, sd, device=dev)
        sched = YHCustomScheduler(device=dev)
        sched.set_timesteps(100)
        eps = net(xt.to(dev), sched._ts_host[40])
        _, s, vT = local_basis(net, sched, xt.to(dev), sched._ts_host[40], 3, v0=v0.T.contiguous().to(dev),
                               min_iter=10 ** 6, max_iter=2, mask=mask.to(dev), verbose=False)
        proj = ops.nullspace_project(vT, vT[:1].contiguous(), project=True)
        torch.cuda.synchronize(dev)
        assert eps.device == dev and vT.device == dev and torch.cuda.current_device() == 0
        res.append((eps.cpu(), s.cpu(), vT.cpu(), proj.cpu()))
    # same kernels, same inputs, different device: TF32 noise level only (GroupNorm uses atomics)
    assert float((res[0][0] - res[1][0]).abs().max()) < 1e-3 * float(res[0][0].abs().max())
    assert float(((res[0][1] - res[1][1]).abs() / res[0][1]).max()) < 1e-3
    assert float((res[0][2] * res[1][2]).sum(1).abs().min()) > 0.9999
