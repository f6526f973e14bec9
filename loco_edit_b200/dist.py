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


class CommTimer(object):
    """Device time spent inside the all-gathers (CUDA events on the current stream; ignored on CPU)."""

    def __init__(self):
        self.events = []

    def span(self, tensor):
        timer = self

        class _Span(object):
            def __enter__(self_inner):
                self_inner.on = tensor.is_cuda
                if self_inner.on:
                    self_inner.a = torch.cuda.Event(enable_timing=True)
                    self_inner.b = torch.cuda.Event(enable_timing=True)
                    self_inner.a.record(torch.cuda.current_stream(tensor.device))

            def __exit__(self_inner, *exc):
                if self_inner.on:
                    self_inner.b.record(torch.cuda.current_stream(tensor.device))
                    timer.events.append((self_inner.a, self_inner.b))

        return _Span()

    def total_ms(self):
        """Sum over the recorded spans (synchronises on the events)."""
        ms = 0.0
        for a, b in self.events:
            b.synchronize()
            ms += a.elapsed_time(b)
        return ms


def all_gather_rows(local_rows, k, group=None, out=None, timer=None):
    """Gather row blocks [k_r, d] from all ranks into [k, d] (rank order = row order).
    Equal shards (k % world == 0) are gathered straight into `out` with one all_gather_into_tensor;
    ragged shards are padded to the largest one and compacted afterwards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_rows
    sizes = shard_sizes(k, world)
    kmax = max(sizes)
    d = local_rows.shape[1]
    ctx = timer.span(local_rows) if timer is not None else None
    if min(sizes) == kmax:
        if out is None:
            out = torch.empty(k, d, dtype=local_rows.dtype, device=local_rows.device)
        if ctx:
            with ctx:
                dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        else:
            dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    pad = torch.zeros(kmax, d, dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    buf = torch.empty(world * kmax, d, dtype=local_rows.dtype, device=local_rows.device)
    if ctx:
        with ctx:
            dist.all_gather_into_tensor(buf, pad, group=group)
    else:
        dist.all_gather_into_tensor(buf, pad, group=group)
    if out is None:
        out = torch.empty(k, d, dtype=local_rows.dtype, device=local_rows.device)
    lo = 0
    for r in range(world):
        out[lo:lo + sizes[r]].copy_(buf[r * kmax: r * kmax + sizes[r]])
        lo += sizes[r]
    return out


def sharded_local_basis(probe_fn, ortho_fn, V0, n_iter, group=None, timer=None):
    """Rank-k subspace iteration with the probe tangents sharded over the ranks.

    probe_fn(V_rows [k_r, d], lo) -> (U_rows [k_r, l], W_rows [k_r, d])  : masked J V^T and J^T U rows
      (`lo` = index of the first row inside the full basis; single-mask callers may ignore it)
    ortho_fn(W [k, d], V_prev [k, d]) -> (V [k, d], s [k])               : Vh / sqrt(sv) of svd(W)
    Returns (U [k, l] of the last iterate, s, V) on every rank (replicated)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    k = V0.shape[0]
    lo, hi = shard_range(k, world, rank)
    V = V0
    U_rows = s = None
    Wbuf = None
    for _ in range(n_iter):
        if hi > lo:
            U_rows, W_rows = _call_probe(probe_fn, V[lo:hi].contiguous(), lo)
        else:   # more ranks than probes: this rank contributes nothing
            U_rows = torch.zeros(0, 1, dtype=V.dtype, device=V.device)
            W_rows = torch.zeros(0, V.shape[1], dtype=V.dtype, device=V.device)
        if world > 1 and Wbuf is None:
            Wbuf = torch.empty(k, V.shape[1], dtype=V.dtype, device=V.device)
        W = all_gather_rows(W_rows, k, group, out=Wbuf, timer=timer)
        V, s = ortho_fn(W, V)
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


def _call_probe(probe_fn, rows, lo):
    import inspect
    n_args = len(inspect.signature(probe_fn).parameters)
    return probe_fn(rows, lo) if n_args >= 2 else probe_fn(rows)   # single-mask probes take the rows only


def pair_ortho(ortho_fn, k):
    """Orthonormalise the edit rows [0,k) and the null rows [k,k+k_null) independently
    (src/modules/edit.py:2294 / :2307 are two separate power methods)."""
    def fn(W, V_prev):
        Va, sa = ortho_fn(W[:k].contiguous(), V_prev[:k].contiguous())
        Vb, sb = ortho_fn(W[k:].contiguous(), V_prev[k:].contiguous())
        return torch.cat([Va, Vb], 0), torch.cat([sa, sb], 0)
    return fn


def cuda_probe_fn(unet, scheduler, xt, t, mask, noise, k_local, chunk_size=25):
    """probe_fn backed by libloco_b200.so for this rank's `k_local` tangents.  More than `chunk_size`
    local rows are probed chunk by chunk (torch.chunk sizes, like src/modules/edit.py:2419, 2448): the
    1 + 2k rows of activations of one fused plan grow by ~1.45 GB per row at 256^2."""
    from . import ops
    from .edit import _pb_workspace
    at = scheduler.alpha_at(float(t))
    x = xt.to(device=unet.device, dtype=torch.float32).contiguous().reshape(1, -1)
    m8 = None if mask is None else mask.to(unet.device).reshape(-1).to(torch.uint8).contiguous()
    idx = None if m8 is None else ops.mask_indices(m8)
    n_chunk = (k_local + chunk_size - 1) // chunk_size
    sizes = [c.shape[0] for c in torch.empty(k_local, 1).chunk(n_chunk)]
    if len(sizes) == 1:
        ws = _pb_workspace(unet, k_local)

        def fn(V_rows, lo=0):
            u_full, w = ws.probe(x, float(t), at, m8, noise, V_rows)
            u = u_full.clone() if idx is None else ops.gather_rows(u_full, idx)
            return u, w          # w: view of the workspace, consumed by the all-gather before the next probe

        return fn
    d = x.numel()
    Ub = torch.empty(k_local, d if idx is None else idx.numel(), dtype=torch.float32, device=unet.device)
    Wb = torch.empty(k_local, d, dtype=torch.float32, device=unet.device)

    def fn_chunked(V_rows, lo=0):
        o = 0
        for kc in sizes:
            ws = _pb_workspace(unet, kc)
            u_full, w = ws.probe(x, float(t), at, m8, noise, V_rows[o:o + kc].contiguous())
            Ub[o:o + kc].copy_(u_full if idx is None else ops.gather_rows(u_full, idx))
            Wb[o:o + kc].copy_(w)
            o += kc
        return Ub, Wb

    return fn_chunked


def cuda_ortho_fn(align_sign=True):
    from . import ops

    def fn(W, V_prev):
        return ops.orthonormalise(W.contiguous(), v_prev=V_prev if align_sign else None)

    return fn


def sharded_local_basis_cuda(unet, scheduler, xt, t, pca_rank, v0, n_iter, mask=None, noise=False,
                             align_sign=True, group=None, timer=None):
    """Multi-GPU `local_encoder_decoder_pullback_xt`: same returns as edit.local_basis."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(pca_rank, world, rank)
    probe = cuda_probe_fn(unet, scheduler, xt, t, mask, noise, hi - lo) if hi > lo else None
    U, s, V = sharded_local_basis(probe, cuda_ortho_fn(align_sign), v0, n_iter, group, timer=timer)
    return U.T, s, V


def cuda_pair_probe_fn(unet, scheduler, xt, t, mask, noise, k, k_local):
    """probe_fn for a shard of the joint {edit, null} probe set: rows of the full [k + k_null, d]
    basis below `k` see the mask, the others its complement (loco_pullback_probe_pair)."""
    from .edit import _pb_workspace
    ws = _pb_workspace(unet, k_local)
    at = scheduler.alpha_at(float(t))
    x = xt.to(device=unet.device, dtype=torch.float32).contiguous().reshape(1, -1)
    m8 = mask.to(unet.device).reshape(-1).to(torch.uint8).contiguous()

    def fn(V_rows, lo):
        k_inv = min(max(k - lo, 0), V_rows.shape[0])
        u_full, w = ws.probe_pair(x, float(t), at, m8, noise, V_rows, k_inv)
        return u_full, w          # views of the workspace: consumed by the all-gather before the next probe

    return fn


def sharded_local_basis_pair_cuda(unet, scheduler, xt, t, k, k_null, mask, v0, v0_null, n_iter,
                                  noise=False, align_sign=True, group=None, timer=None):
    """Multi-GPU `edit.local_basis_pair`: the k + k_null probes of the edit basis (mask) and the null
    basis (~mask) of one image are sharded jointly over the ranks; one all-gather of the
    [k + k_null, d] W rows per iteration, each basis orthonormalised on its own, replicated.
    Returns (vT_modify, s_modify, vT_null, s_null) like local_basis_pair."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    kt = k + k_null
    lo, hi = shard_range(kt, world, rank)
    probe = cuda_pair_probe_fn(unet, scheduler, xt, t, mask, noise, k, hi - lo) if hi > lo else None
    V0 = torch.cat([v0.reshape(k, -1), v0_null.reshape(k_null, -1)], 0).contiguous()
    _, s, V = sharded_local_basis(probe, pair_ortho(cuda_ortho_fn(align_sign), k), V0, n_iter, group,
                                  timer=timer)
    return V[:k].contiguous(), s[:k], V[k:].contiguous(), s[k:]


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
    feat = 1
    for n in local.shape[1:]:
        feat *= n
    flat = local.reshape(local.shape[0], feat)      # (a rank may hold no image at all)
    out = all_gather_rows(flat, n_total, group)
    return out.reshape(n_total, *local.shape[1:])
