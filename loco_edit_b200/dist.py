"""Multi-GPU layer (one process per GPU, torch.distributed; NCCL over NVLink on the B200 box, gloo in
the CPU tests).  The reference has no distributed path at all (a bash loop over images,
src/scripts/main_hf_null_space_projection_FFHQ_P2.sh); the hot path shards two ways:

  * probes: row b of W = U^T J depends only on row b of V, so rank g owns a contiguous slice of the
    k probe tangents end to end (fused JVP + VJP on replicated weights and replicated x_t); one
    all-gather of the per-rank W rows per iteration, then every rank runs the same deterministic
    orthonormalisation -> identical V everywhere, no broadcast needed;
  * images: independent edits, sharded by index, no collective until the final gather.

The compute step is injected (`probe_fn`, `ortho_fn`) so the sharding/collective logic is tested on
CPU with the oracle as the compute backend (tests/test_dist_cpu.py) and runs unchanged on CUDA.
"""
import torch
import torch.distributed as dist


def shard_range(n, world, rank):
    """Contiguous, balanced [lo, hi) slice of range(n) for `rank` (first n % world ranks get +1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n, world):
    return [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]


def all_gather_rows(local_rows, k, group=None):
    """Gather ragged row blocks [k_r, d] from all ranks into [k, d] (rank order = row order).
    Blocks are padded to the largest shard so a single fixed-size all_gather is used."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_rows
    sizes = shard_sizes(k, world)
    kmax = max(sizes)
    d = local_rows.shape[1]
    pad = torch.zeros(kmax, d, dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    out = torch.empty(world * kmax, d, dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * kmax: r * kmax + sizes[r]] for r in range(world)], 0)


def sharded_local_basis(probe_fn, ortho_fn, V0, n_iter, group=None):
    """Rank-k subspace iteration with the probe tangents sharded over the ranks.

    probe_fn(V_rows [k_r, d]) -> (U_rows [k_r, l], W_rows [k_r, d])  : masked J V^T and J^T U rows
    ortho_fn(W [k, d], V_prev [k, d]) -> (V [k, d], s [k])           : Vh / sqrt(sv) of svd(W)
    Returns (U [k, l] of the last iterate, s, V) on every rank (replicated)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    k = V0.shape[0]
    lo, hi = shard_range(k, world, rank)
    V = V0
    U_rows = s = None
    for _ in range(n_iter):
        if hi > lo:
            U_rows, W_rows = probe_fn(V[lo:hi].contiguous())
        else:   # more ranks than probes: this rank contributes nothing
            U_rows = torch.zeros(0, 1, dtype=V.dtype, device=V.device)
            W_rows = torch.zeros(0, V.shape[1], dtype=V.dtype, device=V.device)
        W = all_gather_rows(W_rows, k, group)
        V, s = ortho_fn(W, V)
    l = U_rows.shape[1] if U_rows.shape[0] else None
    if world > 1:
        # U is only returned (never iterated on): gather it once at the end
        lens = torch.tensor([U_rows.shape[1] if U_rows.shape[0] else 0], device=V.device)
        dist.all_reduce(lens, op=dist.ReduceOp.MAX, group=group)
        l = int(lens.item())
        if U_rows.shape[0] == 0:
            U_rows = torch.zeros(0, l, dtype=V.dtype, device=V.device)
        U = all_gather_rows(U_rows, k, group)
    else:
        U = U_rows
    return U, s, V


def cuda_probe_fn(unet, scheduler, xt, t, mask, noise, k_local):
    """probe_fn backed by libloco_b200.so for this rank's `k_local` tangents."""
    from . import ops
    from .edit import _pb_workspace
    ws = _pb_workspace(unet, k_local)
    at = scheduler.alpha_at(float(t))
    x = xt.to(device=unet.device, dtype=torch.float32).contiguous().reshape(1, -1)
    m8 = None if mask is None else mask.to(unet.device).reshape(-1).to(torch.uint8).contiguous()
    idx = None if m8 is None else ops.mask_indices(m8)

    def fn(V_rows):
        u_full, w = ws.probe(x, float(t), at, m8, noise, V_rows)
        u = u_full if idx is None else ops.gather_rows(u_full, idx)
        return u.clone(), w.clone()

    return fn


def cuda_ortho_fn(align_sign=True):
    from . import ops

    def fn(W, V_prev):
        return ops.orthonormalise(W.contiguous(), v_prev=V_prev if align_sign else None)

    return fn


def sharded_local_basis_cuda(unet, scheduler, xt, t, pca_rank, v0, n_iter, mask=None, noise=False,
                             align_sign=True, group=None):
    """Multi-GPU `local_encoder_decoder_pullback_xt`: same returns as edit.local_basis."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(pca_rank, world, rank)
    probe = cuda_probe_fn(unet, scheduler, xt, t, mask, noise, hi - lo) if hi > lo else None
    U, s, V = sharded_local_basis(probe, cuda_ortho_fn(align_sign), v0, n_iter, group)
    return U.T, s, V


def shard_images(indices, group=None):
    """This rank's share of a list of image indices (batch editing, no collective)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(len(indices), world, rank)
    return list(indices[lo:hi])


def gather_images(local, n_total, group=None):
    """Final gather of per-rank result tensors [n_r, ...] into [n_total, ...] in index order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    flat = local.reshape(local.shape[0], -1)
    out = all_gather_rows(flat, n_total, group)
    return out.reshape(n_total, *local.shape[1:])
