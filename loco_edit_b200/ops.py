"""Thin torch-facing wrappers over the C ABI (include/loco_b200.h).  Every function enqueues CUDA
kernels of libloco_b200.so on torch's current stream; none of them has a PyTorch fallback."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _f32(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), (t.device, t.dtype, t.is_contiguous())
    return t


def _scratch(nbytes, device):
    return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)


def _aligned(buf):
    off = (-buf.data_ptr()) % 256
    return C.c_void_p(buf.data_ptr() + off)


def pmp_forward(x, eps, at):
    """(x - eps*sqrt(1-at))/sqrt(at) -- get_x0 without mask (reference modules/edit.py:2386)."""
    out = torch.empty_like(_f32(x))
    check(_lib.load().loco_pmp_forward(ptr(x), ptr(_f32(eps)), float(at), x.numel(), ptr(out), stream_ptr(x)),
          "loco_pmp_forward")
    return out


def combine3(a, wa, b=None, wb=0.0, c=None, wc=0.0):
    """wa*a + wb*b + wc*c (classifier-free-guidance combination, reference modules/edit.py:660-673)."""
    a = _f32(a)
    out = torch.empty_like(a)
    check(_lib.load().loco_combine3(ptr(a), float(wa), ptr(_f32(b)) if b is not None else None, float(wb),
                                    ptr(_f32(c)) if c is not None else None, float(wc), a.numel(), ptr(out),
                                    stream_ptr(a)), "loco_combine3")
    return out


def pmp_jvp_epilogue(V, eps_dot, mask_u8, at, noise=False, k_invert=None):
    """(u, g_eps, gx_direct), all [k, d]: the PMP differentiated along the rows of V given the tangents
    eps_dot of the noise prediction, and the seeds of the transposed pass (pullback.cuh)."""
    V, eps_dot = _f32(V), _f32(eps_dot)
    k, d = V.shape
    u, ge, gd = torch.empty_like(V), torch.empty_like(V), torch.empty_like(V)
    check(_lib.load().loco_pmp_jvp_epilogue(ptr(V), ptr(eps_dot), ptr(mask_u8) if mask_u8 is not None else None,
                                            float(at), 1 if noise else 0, k, k if k_invert is None else int(k_invert),
                                            d, ptr(u), ptr(ge), ptr(gd), stream_ptr(V)), "loco_pmp_jvp_epilogue")
    return u, ge, gd


def orthonormalise(W, v_prev=None):
    """(V, s): V = Vh of svd(W) (rows, up to sign), s = sqrt(singular values)
    (reference modules/edit.py:2482, 2499-2502)."""
    lib = _lib.load()
    W = _f32(W)
    k, d = W.shape
    V = torch.empty_like(W)
    s = torch.empty(k, dtype=torch.float32, device=W.device)
    scr = _scratch(lib.loco_orthonormalise_scratch_bytes(k), W.device)
    check(lib.loco_orthonormalise(ptr(W), k, d, ptr(_f32(v_prev)) if v_prev is not None else None,
                                  ptr(V), ptr(s), _aligned(scr), stream_ptr(W)), "loco_orthonormalise")
    return V, s


def nullspace_project(vT_mod, vT_null, project=True):
    """normalise_rows(vT_mod - (Vn^T (Vn vT_mod^T))^T) (reference modules/edit.py:2317-2323)."""
    lib = _lib.load()
    vT_mod = _f32(vT_mod)
    k, d = vT_mod.shape
    kn = 0 if vT_null is None else vT_null.shape[0]
    out = torch.empty_like(vT_mod)
    scr = _scratch(8 * (kn * k + k), vT_mod.device)
    check(lib.loco_nullspace_project(ptr(vT_mod), k, ptr(_f32(vT_null)) if kn else None, kn, d,
                                     1 if (project and kn) else 0, ptr(out), _aligned(scr), stream_ptr(vT_mod)),
          "loco_nullspace_project")
    return out


def ddim_step(xt, et, at, at_next, eta=0.0, noise=None, want_x0=False):
    """YHCustomScheduler.step arithmetic (reference utils/utils.py:357-374)."""
    xt = _f32(xt)
    out = torch.empty_like(xt)
    x0 = torch.empty_like(xt) if want_x0 else None
    check(_lib.load().loco_ddim_step(ptr(xt), ptr(_f32(et)), ptr(_f32(noise)) if noise is not None else None,
                                     float(at), float(at_next), float(eta), xt.numel(), ptr(out),
                                     ptr(x0), stream_ptr(xt)), "loco_ddim_step")
    return (out, x0) if want_x0 else out


def axpy(x, v, scale):
    """x + scale * v (reference modules/edit.py:2623)."""
    x = _f32(x)
    out = torch.empty_like(x)
    check(_lib.load().loco_axpy(ptr(x), ptr(_f32(v)), float(scale), x.numel(), ptr(out), stream_ptr(x)),
          "loco_axpy")
    return out


def mask_indices(mask):
    """Row-major indices of True entries: the selection order of `P_xt[:, mask]`
    (reference modules/edit.py:2390).  Returns an int32 tensor (one host sync for the count)."""
    m = mask.to(torch.uint8).contiguous().reshape(-1)
    assert m.is_cuda
    idx = torch.empty(m.numel(), dtype=torch.int32, device=m.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=m.device)
    check(_lib.load().loco_mask_indices(ptr(m), m.numel(), ptr(idx), ptr(cnt), stream_ptr(m)),
          "loco_mask_indices")
    return idx[: int(cnt.item())]


def gather_rows(src, idx):
    src = _f32(src)
    rows, d = src.shape
    out = torch.empty(rows, idx.numel(), dtype=torch.float32, device=src.device)
    check(_lib.load().loco_gather_rows(ptr(src), rows, d, ptr(idx), idx.numel(), ptr(out), stream_ptr(src)),
          "loco_gather_rows")
    return out


def scatter_rows(src, idx, d):
    src = _f32(src)
    rows = src.shape[0]
    out = torch.empty(rows, d, dtype=torch.float32, device=src.device)
    check(_lib.load().loco_scatter_rows(ptr(src), rows, d, ptr(idx), idx.numel(), ptr(out), stream_ptr(src)),
          "loco_scatter_rows")
    return out


def gram(A, B):
    A, B = _f32(A), _f32(B)
    G = torch.zeros(A.shape[0], B.shape[0], dtype=torch.float64, device=A.device)
    check(_lib.load().loco_gram(ptr(A), A.shape[0], ptr(B), B.shape[0], A.shape[1], ptr(G), stream_ptr(A)),
          "loco_gram")
    return G


class PullbackWorkspace:
    """Buffers of one rank-k power method on a B200UNet (plan (1,k,k) + iteration scratch)."""

    def __init__(self, unet, k, slot=0):
        self.lib = _lib.load()
        self.unet, self.k = unet, k
        R = unet.arch["resolution"]
        self.d = 3 * R * R
        self.plan = unet.plan(1, k, k, slot=slot)
        dev = unet.device
        self.scratch = _scratch(self.lib.loco_pullback_scratch_bytes(k, self.d), dev)
        self.u_full = torch.empty(k, self.d, dtype=torch.float32, device=dev)
        self.w = torch.empty(k, self.d, dtype=torch.float32, device=dev)
        self.s = torch.empty(k, dtype=torch.float32, device=dev)
        self.V = [torch.empty(k, self.d, dtype=torch.float32, device=dev) for _ in range(2)]

    def iterate(self, xt, t, at, mask_u8, noise, V_in, V_out, align_sign=False):
        check(self.lib.loco_pullback_iteration(
            self.plan.handle, ptr(xt), float(t), float(at), ptr(mask_u8) if mask_u8 is not None else None,
            1 if noise else 0, ptr(V_in), self.k, self.d, 1 if align_sign else 0, ptr(self.u_full),
            ptr(self.w), ptr(V_out), ptr(self.s), _aligned(self.scratch), stream_ptr(xt)),
            "loco_pullback_iteration")


    def iterate_pair(self, xt, t, at, mask_u8, noise, V_in, V_out, k1, k2, align_sign=False):
        """Edit basis (rows [0,k1), mask) and null basis (rows [k1,k1+k2), ~mask) in one fused pass."""
        assert k1 + k2 == self.k and mask_u8 is not None
        check(self.lib.loco_pullback_pair_iteration(
            self.plan.handle, ptr(xt), float(t), float(at), ptr(mask_u8), 1 if noise else 0, ptr(V_in),
            k1, k2, self.d, 1 if align_sign else 0, ptr(self.u_full), ptr(self.w), ptr(V_out),
            ptr(self.s), _aligned(self.scratch), stream_ptr(xt)), "loco_pullback_pair_iteration")

    def probe(self, xt, t, at, mask_u8, noise, V_in):
        """Rows of U (masked J V^T) and W (J^T U) for the k rows of V_in, no orthonormalisation."""
        check(self.lib.loco_pullback_probe(
            self.plan.handle, ptr(xt), float(t), float(at), ptr(mask_u8) if mask_u8 is not None else None,
            1 if noise else 0, ptr(V_in), self.k, self.d, ptr(self.u_full), ptr(self.w),
            _aligned(self.scratch), stream_ptr(xt)), "loco_pullback_probe")
        return self.u_full, self.w


    def probe_pair(self, xt, t, at, mask_u8, noise, V_in, k_invert):
        """Like probe(), rows [0,k_invert) through the mask, rows [k_invert,k) through its complement."""
        check(self.lib.loco_pullback_probe_pair(
            self.plan.handle, ptr(xt), float(t), float(at), ptr(mask_u8), 1 if noise else 0, ptr(V_in),
            self.k, int(k_invert), self.d, ptr(self.u_full), ptr(self.w), _aligned(self.scratch),
            stream_ptr(xt)), "loco_pullback_probe_pair")
        return self.u_full, self.w


# ---------------------------------------------------------------------------------------------
# single layers (parity tests)
# ---------------------------------------------------------------------------------------------
def conv2d_nhwc(kind, x, w, bias=None, bias_rows=0, addend=None, accumulate=False, out=None, splitk=True):
    """kind: 0 3x3 | 1 1x1 | 2 3x3 stride-2 pad(0,1,0,1) | 3 dgrad of 0 | 4 dgrad of 2.
    x: [N,H,W,C] channels-last contiguous; w: torch conv weight [Cout,Cin,k,k] of the forward conv."""
    x, w = _f32(x), _f32(w)
    N, H, W_, Cx = x.shape
    Cout, Cin = w.shape[0], w.shape[1]
    Ho, Wo = (H // 2, W_ // 2) if kind == 2 else ((2 * H, 2 * W_) if kind == 4 else (H, W_))
    Cy = Cout if kind in (0, 1, 2) else Cin
    if out is None:
        out = torch.zeros(N, Ho, Wo, Cy, dtype=torch.float32, device=x.device)
    wpack = torch.empty(w.numel(), dtype=torch.float32, device=x.device)
    scr = torch.zeros(16 << 20, dtype=torch.uint8, device=x.device) if splitk else None
    check(_lib.load().loco_conv2d_nhwc(kind, ptr(x), N, H, W_, Cx, ptr(w), Cout, Cin, ptr(wpack),
                                       ptr(bias), bias_rows, ptr(addend), 1 if accumulate else 0,
                                       ptr(out), ptr(scr), scr.numel() if splitk else 0, stream_ptr(x)),
          "loco_conv2d_nhwc")
    return out


def conv2d_nhwc_typed(kind, x, w, bias=None, bias_rows=0, addend=None, out_dtype=None, splitk=True):
    """Like conv2d_nhwc with fp16 or fp32 tensors: x.dtype selects the tensor-core kind (fp16 ->
    kind::f16, fp32 -> kind::tf32), out_dtype (default x.dtype) the storage of the result / addend."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float16, torch.float32)
    out_dtype = out_dtype or x.dtype
    w = _f32(w)
    N, H, W_, Cx = x.shape
    Cout, Cin = w.shape[0], w.shape[1]
    Ho, Wo = (H // 2, W_ // 2) if kind == 2 else ((2 * H, 2 * W_) if kind == 4 else (H, W_))
    Cy = Cout if kind in (0, 1, 2) else Cin
    out = torch.zeros(N, Ho, Wo, Cy, dtype=out_dtype, device=x.device)
    if addend is not None:
        assert addend.dtype == out_dtype and addend.is_contiguous()
    wpack = torch.empty(w.numel(), dtype=torch.float32, device=x.device)
    scr = torch.zeros(16 << 20, dtype=torch.uint8, device=x.device) if splitk else None
    check(_lib.load().loco_conv2d_nhwc_ex(kind, ptr(x), N, H, W_, Cx, ptr(w), Cout, Cin, ptr(wpack),
                                          ptr(bias), bias_rows, ptr(addend), 0, ptr(out), ptr(scr),
                                          scr.numel() if splitk else 0, 1 if x.dtype == torch.float16 else 0,
                                          1 if out_dtype == torch.float16 else 0, stream_ptr(x)),
          "loco_conv2d_nhwc_ex")
    return out


def conv2d_fused_nhwc(x, w, x2=None, w2=None, bias=None, bias_rows=0, stat_groups=0):
    """conv3x3(x; w) + conv1x1(x2; w2) + bias in one launch (conv2 + shortcut of a ResnetBlock);
    stat_groups > 0 also returns the fused GroupNorm statistics [N, stat_groups, 2] of the result."""
    x, w = _f32(x), _f32(w)
    N, H, W_, Cin = x.shape
    Cout = w.shape[0]
    out = torch.empty(N, H, W_, Cout, dtype=torch.float32, device=x.device)
    wpack = torch.empty(w.numel(), dtype=torch.float32, device=x.device)
    C2, wpack2 = 0, None
    if x2 is not None:
        x2, w2 = _f32(x2), _f32(w2)
        C2 = x2.shape[-1]
        wpack2 = torch.empty(w2.numel(), dtype=torch.float32, device=x.device)
    stats = None
    if stat_groups:
        assert stat_groups == 32
        stats = torch.zeros(N, 32, 2, dtype=torch.float64, device=x.device)
    check(_lib.load().loco_conv2d_fused_nhwc(ptr(x), N, H, W_, Cin, ptr(w), Cout, ptr(x2), C2, ptr(w2),
                                             ptr(wpack), ptr(wpack2), ptr(bias), bias_rows, ptr(out),
                                             ptr(stats), Cout // 32 if stat_groups else 0, stream_ptr(x)),
          "loco_conv2d_fused_nhwc")
    return (out, stats) if stat_groups else out


def groupnorm_silu_fwd(x, n_primal, gamma, beta, eps, silu):
    x = _f32(x)
    N, H, W_, Cc = x.shape
    y = torch.empty_like(x)
    stats = torch.empty(64 * N, dtype=torch.float64, device=x.device)
    check(_lib.load().loco_groupnorm_silu_fwd(ptr(x), N, H, W_, Cc, n_primal, ptr(_f32(gamma)),
                                              ptr(_f32(beta)), float(eps), 1 if silu else 0, ptr(y),
                                              ptr(stats), stream_ptr(x)), "loco_groupnorm_silu_fwd")
    return y


def groupnorm_silu_vjp(xp, gy, gamma, beta, eps, silu):
    xp, gy = _f32(xp), _f32(gy)
    K, H, W_, Cc = gy.shape
    gx = torch.empty_like(gy)
    stats = torch.empty(64 * (K + 1), dtype=torch.float64, device=gy.device)
    check(_lib.load().loco_groupnorm_silu_vjp(ptr(xp), H, W_, Cc, ptr(gy), K, ptr(_f32(gamma)),
                                              ptr(_f32(beta)), float(eps), 1 if silu else 0, ptr(gx),
                                              ptr(stats), stream_ptr(gy)), "loco_groupnorm_silu_vjp")
    return gx


def groupnorm_silu_fwd_ex(x, n_primal, gamma, beta, eps, silu, y=None, stats=None, stages=3):
    """Typed GroupNorm(+SiLU) forward / JVP on an NHWC tensor of dtype float32 or float16."""
    assert x.dtype in (torch.float32, torch.float16) and x.is_contiguous()
    N, H, W_, Cc = x.shape
    y = torch.empty_like(x) if y is None else y
    stats = torch.empty(64 * N, dtype=torch.float64, device=x.device) if stats is None else stats
    check(_lib.load().loco_groupnorm_silu_fwd_ex(ptr(x), int(x.dtype == torch.float16), N, H, W_, Cc, n_primal,
                                                 ptr(_f32(gamma)), ptr(_f32(beta)), float(eps),
                                                 1 if silu else 0, ptr(y), ptr(stats), stages, stream_ptr(x)),
          "loco_groupnorm_silu_fwd_ex")
    return y, stats


def groupnorm_silu_vjp_ex(xp, gy, gamma, beta, eps, silu, addend=None, accumulate=False, gx=None, stats=None,
                          stages=3):
    assert gy.dtype in (torch.float32, torch.float16) and gy.is_contiguous() and xp.dtype == gy.dtype
    K, H, W_, Cc = gy.shape
    gx = torch.empty_like(gy) if gx is None else gx
    stats = torch.empty(64 * (K + 1), dtype=torch.float64, device=gy.device) if stats is None else stats
    check(_lib.load().loco_groupnorm_silu_vjp_ex(ptr(xp), int(gy.dtype == torch.float16), H, W_, Cc, ptr(gy), K,
                                                 ptr(_f32(gamma)), ptr(_f32(beta)), float(eps),
                                                 1 if silu else 0, ptr(addend), 1 if accumulate else 0, ptr(gx),
                                                 ptr(stats), stages, stream_ptr(gy)),
          "loco_groupnorm_silu_vjp_ex")
    return gx, stats


def attention_fwd(qkv, n_primal, head_ch=0):
    """head_ch = 0: one head, channels q|k|v (DDPM); > 0: per-head q|k|v (guided-diffusion legacy)."""
    qkv = _f32(qkv)
    N, T, C3 = qkv.shape
    Cc = C3 // 3
    heads = 1 if head_ch <= 0 else Cc // head_ch
    S = torch.empty(N, heads, T, T, dtype=torch.float32, device=qkv.device)
    o = torch.empty(N, T, Cc, dtype=torch.float32, device=qkv.device)
    check(_lib.load().loco_attention_fwd(ptr(qkv), N, T, Cc, n_primal, head_ch, ptr(S), ptr(o),
                                         stream_ptr(qkv)), "loco_attention_fwd")
    return o, (S if head_ch > 0 else S[:, 0])


def attention_vjp(go, qkv0, P0, head_ch=0):
    go = _f32(go)
    K, T, Cc = go.shape
    heads = 1 if head_ch <= 0 else Cc // head_ch
    gP = torch.empty(K, heads, T, T, dtype=torch.float32, device=go.device)
    gqkv = torch.empty(K, T, 3 * Cc, dtype=torch.float32, device=go.device)
    check(_lib.load().loco_attention_vjp(ptr(go), K, T, Cc, head_ch, ptr(_f32(qkv0)), ptr(_f32(P0)),
                                         ptr(gP), ptr(gqkv), stream_ptr(go)), "loco_attention_vjp")
    return gqkv


def cross_attention_fwd(q, kv, n_primal, heads, tk_valid):
    """q [N, Tq, C] fp32, kv [Tk, 2C] = K_c | V_c (rows >= tk_valid zero).  Returns (o [N, Tq, C], P [N, heads, Tq, Tk])."""
    q, kv = _f32(q), _f32(kv)
    N, Tq, Cc = q.shape
    Tk = kv.shape[0]
    S = torch.empty(N, heads, Tq, Tk, dtype=torch.float32, device=q.device)
    o = torch.empty(N, Tq, Cc, dtype=torch.float32, device=q.device)
    check(_lib.load().loco_cross_attention_fwd(ptr(q), N, Tq, Cc, n_primal, ptr(kv), Tk, tk_valid, heads, ptr(S), ptr(o),
                                               stream_ptr(q)), "loco_cross_attention_fwd")
    return o, S


def cross_attention_vjp(go, kv, P0, heads, tk_valid):
    go, kv, P0 = _f32(go), _f32(kv), _f32(P0)
    K, Tq, Cc = go.shape
    gq = torch.empty_like(go)
    check(_lib.load().loco_cross_attention_vjp(ptr(go), K, Tq, Cc, ptr(kv), kv.shape[0], tk_valid, heads, ptr(P0), ptr(gq),
                                               stream_ptr(go)), "loco_cross_attention_vjp")
    return gq
