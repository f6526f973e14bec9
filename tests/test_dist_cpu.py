"""CPU, world_size 2, gloo: the probe-sharded power method and the image sharding/gather logic of
loco_edit_b200/dist.py with the oracle as the compute backend; results must equal the unsharded
oracle run."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loco_edit_b200 import dist as ld


def test_shard_range_is_a_partition():
    for n in (0, 1, 5, 10, 64):
        for world in (1, 2, 3, 8):
            spans = [ld.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == ld.shard_sizes(n, world)


def _worker(rank, world, port, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from loco_edit_b200.weights import random_state_dict, tiny_arch
        from oracle import ddpm_ref, pullback_ref
        arch = tiny_arch(resolution=16, ch_mult=(1,), attn_resolutions=(), num_res_blocks=1)
        sd = random_state_dict(arch, seed=5, perturb_norm=0.1)
        unet = ddpm_ref.RefUNet(arch, sd)
        sched = pullback_ref.RefScheduler()
        sched.set_timesteps(100)
        t = sched.timesteps[40]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(1, 3, 16, 16, generator=g)
        mask = torch.zeros(3, 16, 16, dtype=torch.bool)
        mask[:, 4:12, 2:10] = True
        v0, _ = torch.linalg.qr(torch.randn(x.numel(), k, generator=g))
        v0 = v0.T.contiguous()

        def probe_fn(V_rows):
            u, w, _, _ = pullback_ref.power_iteration(unet, sched, x, t, V_rows, mask=mask)
            return u, w

        def ortho_fn(W, V_prev):
            _, s, vh = torch.linalg.svd(W, full_matrices=False)
            return vh, s.sqrt()

        U, s, V = ld.sharded_local_basis(probe_fn, ortho_fn, v0, 2)
        # image sharding + ordered gather
        mine = ld.shard_images(list(range(5)))
        local = torch.tensor([[float(i), 10.0 * i] for i in mine]).reshape(len(mine), 2)
        allimgs = ld.gather_images(local, 5)
        if rank == 0:
            ru, rs, rV = pullback_ref.local_basis(unet, sched, x, t, v0, 2, mask=mask)
            ret["s_err"] = float((s - rs).abs().max() / rs.max())
            ret["v_err"] = float((1 - (V * rV).sum(1).abs()).abs().max())
            ret["u_shape"] = (tuple(U.shape), tuple(ru.T.shape))
            ret["u_err"] = float((U.abs() - ru.T.abs()).abs().max() / ru.abs().max())
            ret["imgs"] = allimgs.tolist()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k", [3, 1])
def test_sharded_power_method_equals_unsharded(k):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + k + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, k, ret), nprocs=2, join=True)
    assert ret["s_err"] < 1e-4, dict(ret)
    assert ret["v_err"] < 1e-3, dict(ret)
    assert ret["u_shape"][0] == ret["u_shape"][1]
    assert ret["u_err"] < 1e-3
    assert ret["imgs"] == [[float(i), 10.0 * i] for i in range(5)]


def _pair_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from loco_edit_b200.weights import random_state_dict, tiny_arch
        from oracle import ddpm_ref, pullback_ref
        arch = tiny_arch(resolution=16, ch_mult=(1,), attn_resolutions=(), num_res_blocks=1)
        unet = ddpm_ref.RefUNet(arch, random_state_dict(arch, seed=5, perturb_norm=0.1))
        sched = pullback_ref.RefScheduler()
        sched.set_timesteps(100)
        t = sched.timesteps[40]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(1, 3, 16, 16, generator=g)
        mask = torch.zeros(3, 16, 16, dtype=torch.bool)
        mask[:, 4:12, 2:10] = True
        k, kn = 2, 3               # 5 joint probes over 2 ranks: shards [0,3) and [3,5) straddle the boundary
        va, _ = torch.linalg.qr(torch.randn(x.numel(), k, generator=g))
        vb, _ = torch.linalg.qr(torch.randn(x.numel(), kn, generator=g))
        V0 = torch.cat([va.T, vb.T], 0).contiguous()
        d = x.numel()

        def probe_fn(V_rows, lo):
            us, ws = [], []
            for j in range(V_rows.shape[0]):        # rows below k see the mask, the others ~mask
                m = mask if lo + j < k else ~mask
                u, w, _, _ = pullback_ref.power_iteration(unet, sched, x, t, V_rows[j:j + 1], mask=m)
                uf = torch.zeros(1, d)
                uf[:, m.reshape(-1)] = u
                us.append(uf); ws.append(w)
            return torch.cat(us, 0), torch.cat(ws, 0)

        def ortho_fn(W, V_prev):
            _, s, vh = torch.linalg.svd(W, full_matrices=False)
            return vh, s.sqrt()

        _, s, V = ld.sharded_local_basis(probe_fn, ld.pair_ortho(ortho_fn, k), V0, 2)
        if rank == 0:
            _, sa, Va = pullback_ref.local_basis(unet, sched, x, t, V0[:k], 2, mask=mask)
            _, sb, Vb = pullback_ref.local_basis(unet, sched, x, t, V0[k:], 2, mask=~mask)
            rs, rV = torch.cat([sa, sb]), torch.cat([Va, Vb], 0)
            ret["s_err"] = float((s - rs).abs().max() / rs.max())
            ret["v_err"] = float((1 - (V * rV).sum(1).abs()).abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_jointly_sharded_edit_and_null_bases_equal_two_unsharded_runs(world):
    """the k + k_null probes of the edit basis (mask) and the null basis (~mask) sharded
    jointly over the ranks == two separate unsharded power methods of the oracle.  world_size 3 shards the 5
    probes 2 / 2 / 1: the ragged (padded) all-gather of the W rows."""
    ret = mp.Manager().dict()
    mp.spawn(_pair_worker, args=(world, 29650 + 7 * world + (os.getpid() % 300), ret), nprocs=world, join=True)
    assert ret["s_err"] < 1e-4 and ret["v_err"] < 1e-3, dict(ret)
