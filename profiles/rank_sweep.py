"""BASELINE config 3: numerical-rank sweep of the PMP Jacobian -- rank-64 subspace iteration
(mask = None) at several timesteps of the DDPM-256 U-Net, probe tangents sharded over the GPUs.

    python profiles/rank_sweep.py --timesteps 3 --iters 2                 # one GPU (chunks of <= 25 probes)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29650 profiles/rank_sweep.py --timesteps 3 --iters 2

Every rank holds the replicated weights and x_t, owns k/G probe columns end to end (fused JVP + VJP)
and all-gathers its rows of W = U^T J before the replicated orthonormalisation (loco_edit_b200/dist.py).
Rank 0 prints one JSON line: probes/s (one probe = one JVP + one VJP column), ms per iteration and
the leading singular values per timestep.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--timesteps", type=int, default=3, help="how many of the idx {4,14,...,94}")
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from loco_edit_b200 import dist as ld
    from loco_edit_b200.edit import local_basis
    from loco_edit_b200.scheduler import YHCustomScheduler
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import DDPM256, random_state_dict

    net = B200UNet(DDPM256, random_state_dict(DDPM256, seed=1234), device=dev)
    sched = YHCustomScheduler(device=dev)
    sched.set_timesteps(100)
    g = torch.Generator().manual_seed(0)
    xT = torch.randn(1, 3, 256, 256, generator=g).to(dev)
    k, d = a.rank, 3 * 256 * 256
    v0, _ = torch.linalg.qr(torch.randn(d, k, generator=torch.Generator().manual_seed(7)))
    v0 = v0.T.contiguous().to(dev)
    idxs = list(range(4, 100, 10))[: a.timesteps]
    out, ms_total, n_it = {}, 0.0, 0
    xt, cur = xT, 0
    for idx in idxs:
        # DDIM forward from the current latent to timestep idx (eta = 0)
        for i in range(cur, idx):
            t = sched._ts_host[i]
            xt = sched.step(net(xt, t), t, xt, eta=0, t_idx=i).prev_sample
        cur = idx
        t = sched._ts_host[idx]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):                       # rep 0 builds plans / workspaces
            e0.record()
            if world > 1:
                _, s, _ = ld.sharded_local_basis_cuda(net, sched, xt, t, k, v0, a.iters, mask=None)
            else:
                _, s, _ = local_basis(net, sched, xt, t, k, v0=v0, min_iter=10 ** 6, max_iter=a.iters,
                                      mask=None, verbose=False)
            e1.record()
            torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total += float(ms.item())
        n_it += a.iters
        out["t_idx_%d" % idx] = [round(float(v), 5) for v in s[:6].tolist()]
    if rank == 0:
        print(json.dumps({"config": "rank-%d subspace iteration, DDPM-256, mask=None, %d timesteps x %d iterations"
                                    % (k, len(idxs), a.iters),
                          "n_gpus": world, "ms_per_iteration": ms_total / n_it,
                          "probes_per_s": k * n_it / (ms_total * 1e-3), "s_head": out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
