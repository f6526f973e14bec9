"""Host-side handle of the B200 U-Net executor.

`B200UNet` is what the Edit classes hold as `self.unet`: calling it evaluates eps_theta(x, t) like
the reference's `self.unet(xt, t)` (src/modules/edit.py:2151, 2375, 2572); `jvp`/`vjp` expose the
fused primal+tangent and transposed passes the power method is built from.  All arithmetic runs in
libloco_b200.so; torch only owns the device memory and the stream.
"""
import collections
import ctypes as C
import os

import torch

from . import _lib
from ._lib import Arch, check, ptr, stream_ptr


def _make_arch(a):
    s = Arch()
    s.ch = a["ch"]
    mult = tuple(a["ch_mult"])
    s.n_levels = len(mult)
    for i, m in enumerate(mult):
        s.ch_mult[i] = m
    s.num_res_blocks = a["num_res_blocks"]
    attn = tuple(a["attn_resolutions"])
    s.n_attn = len(attn)
    for i, r in enumerate(attn):
        s.attn_resolutions[i] = r
    s.resolution = a["resolution"]
    s.in_ch = a.get("in_ch", 3)
    s.out_ch = a.get("out_ch", 3)
    s.gn_eps = a.get("gn_eps", 1e-6)
    s.kind = {"p2": 1, "vae_decoder": 2}.get(a.get("kind"), 0)
    s.head_ch = a.get("head_ch", 0) if s.kind == 1 else 0
    s.ctx_dim = a.get("ctx_dim", 0) if s.kind == 0 else 0
    s.ctx_heads = a.get("ctx_heads", 1)
    return s


class Plan:
    """A static launch program for a fixed (n_primal, n_tangent, n_cotangent) batch."""

    def __init__(self, unet, n_primal, n_tangent, n_cot, half=False):
        self.lib = _lib.load()
        self.unet = unet
        self.shape = (n_primal, n_tangent, n_cot)
        self.half = bool(half)
        h = C.c_void_p()
        check(self.lib.loco_plan_create_ex(unet.handle, n_primal, n_tangent, n_cot, 1 if half else 0,
                                           C.byref(h)), "loco_plan_create_ex")
        self.handle = h
        nbytes = self.lib.loco_plan_workspace_bytes(h)
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=unet.device)
        off = (-self.workspace.data_ptr()) % 256
        self._ws_ptr = C.c_void_p(self.workspace.data_ptr() + off)
        check(self.lib.loco_plan_bind(h, self._ws_ptr), "loco_plan_bind")
        ff, vf = C.c_double(), C.c_double()
        fo, vo = C.c_int(), C.c_int()
        check(self.lib.loco_plan_info(h, C.byref(ff), C.byref(vf), C.byref(fo), C.byref(vo)))
        self.fwd_flops, self.vjp_flops = ff.value, vf.value
        self.fwd_ops, self.vjp_ops = fo.value, vo.value
        self.released = False

    def release(self):
        """Free the plan's workspace and launch programs now (evicted from the U-Net's plan cache)."""
        if not self.released:
            self.released = True
            self.lib.loco_plan_destroy(self.handle)
            self.handle = None
            self.workspace = None

    def forward(self, x, t, out=None):
        n = self.shape[0] + self.shape[1]
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        assert tuple(x.shape) == (n,) + self.unet.in_shape, (tuple(x.shape), n, self.unet.in_shape)
        if out is None:
            out = torch.empty((n,) + self.unet.out_shape, dtype=torch.float32, device=x.device)
        check(self.lib.loco_unet_forward(self.handle, ptr(x), float(t), ptr(out), stream_ptr(x)),
              "loco_unet_forward")
        return out

    def set_condition(self, cond):
        """Conditioning embedding [4*ch] (device tensor, or None) added to the timestep embedding of the
        following forward() calls of this plan."""
        if cond is not None:
            assert cond.is_cuda and cond.dtype == torch.float32 and cond.is_contiguous()
            assert cond.numel() == 4 * self.unet.arch["ch"], (cond.numel(), self.unet.arch["ch"])
        check(self.lib.loco_plan_set_condition(self.handle, ptr(cond), stream_ptr(self.unet.device)),
              "loco_plan_set_condition")

    def set_context(self, ctx):
        """Prompt embedding [n_tok <= 128, ctx_dim] (device tensor) for the cross-attention layers of a
        U-Net built with ctx_dim > 0: the `encoder_hidden_states` of the following forward() calls."""
        assert ctx.is_cuda and ctx.dtype == torch.float32 and ctx.is_contiguous() and ctx.dim() == 2
        assert ctx.shape[1] == self.unet.arch.get("ctx_dim", 0), (tuple(ctx.shape), self.unet.arch.get("ctx_dim"))
        check(self.lib.loco_plan_set_context(self.handle, ptr(ctx), int(ctx.shape[0]), stream_ptr(ctx)),
              "loco_plan_set_context")

    def vjp(self, g_eps, out=None):
        k = self.shape[2]
        assert g_eps.is_cuda and g_eps.dtype == torch.float32 and g_eps.is_contiguous()
        assert tuple(g_eps.shape) == (k,) + self.unet.out_shape, (tuple(g_eps.shape), k, self.unet.out_shape)
        if out is None:
            out = torch.empty((k,) + self.unet.in_shape, dtype=torch.float32, device=g_eps.device)
        check(self.lib.loco_unet_vjp(self.handle, ptr(g_eps), ptr(out), stream_ptr(g_eps)), "loco_unet_vjp")
        return out

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class B200UNet:
    # every plan pins a full activation workspace (1.5 GB for (1,0,0), 16 GB for (1,5,5) at 256^2):
    # keep the most recently used few, release the rest (a run that mixes many ranks / batch sizes
    # must not grow without bound)
    max_cached_plans = 8

    def __init__(self, arch, state_dict, device="cuda:0"):
        self.lib = _lib.load()
        self.arch = dict(arch)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.LocoError("B200UNet needs a CUDA device; there is no CPU fallback")
        self._arch_struct = _make_arch(arch)
        a = self._arch_struct
        # rows of the network input / output ([C, H, W]); the VAE decoder upsamples once per level but the last
        r_out = a.resolution << (a.n_levels - 1) if a.kind == 2 else a.resolution
        self.in_shape = (a.in_ch, a.resolution, a.resolution)
        self.out_shape = (a.out_ch, r_out, r_out)
        h = C.c_void_p()
        check(self.lib.loco_unet_create(C.byref(self._arch_struct), C.byref(h)), "loco_unet_create")
        self.handle = h
        n = self.lib.loco_unet_weight_floats(h)
        self.arena = torch.zeros(n + 64, dtype=torch.float32, device=self.device)
        off = ((-self.arena.data_ptr()) % 256) // 4
        check(self.lib.loco_unet_bind_weights(h, C.c_void_p(self.arena.data_ptr() + 4 * off)))
        self._plans = collections.OrderedDict()
        # arithmetic of the Jacobian-free programs (DDIM loops): fp16 storage + kind::f16 tensor cores
        # (default), or fp32 storage + kind::tf32 (LOCO_FWD_FP16=0).  Same 10-bit operand mantissa either
        # way; parity of both against the unmodified reference: tests/test_gpu_full256.py
        self.fwd_half = os.environ.get("LOCO_FWD_FP16", "1") != "0"
        # arithmetic of the Jacobian programs (fused primal + k-tangent JVP, k-cotangent VJP): same switch;
        # their tangent / cotangent rows are range-scaled by a power of two inside the library
        self.jac_half = os.environ.get("LOCO_JAC_FP16", "1") != "0"
        self.load_state_dict(state_dict)

    def param_shapes(self):
        out = {}
        buf = C.create_string_buffer(256)
        shape = (C.c_int * 4)()
        nd = C.c_int()
        for i in range(self.lib.loco_unet_num_params(self.handle)):
            check(self.lib.loco_unet_param_info(self.handle, i, buf, 256, shape, C.byref(nd)))
            out[buf.value.decode()] = tuple(shape[j] for j in range(nd.value))
        return out

    def load_state_dict(self, sd):
        shapes = self.param_shapes()
        missing = [k for k in shapes if k not in sd]
        if missing:
            raise _lib.LocoError("state_dict lacks %d parameters, e.g. %s" % (len(missing), missing[:3]))
        with torch.cuda.device(self.device):
            for name, shape in shapes.items():
                w = sd[name]
                if tuple(w.shape) != shape:
                    raise _lib.LocoError("parameter %s has shape %s, expected %s" % (name, tuple(w.shape), shape))
                wd = w.detach().to(device=self.device, dtype=torch.float32).contiguous()
                check(self.lib.loco_unet_load_param(self.handle, name.encode(), ptr(wd), wd.numel(),
                                                    stream_ptr(self.device)), "loco_unet_load_param(%s)" % name)
            torch.cuda.current_stream(self.device).synchronize()

    def plan(self, n_primal, n_tangent=0, n_cot=0, half=None, slot=0):
        """`half` (fp16 activations + tcgen05 kind::f16) defaults to `self.fwd_half` for the
        Jacobian-free programs and to `self.jac_half` for the JVP / VJP programs."""
        if half is None:
            half = self.fwd_half if (n_tangent == 0 and n_cot == 0) else self.jac_half
        # `slot` separates otherwise identical plans that must keep their own saved activations (one per
        # conditioning of a classifier-free-guidance Jacobian product)
        key = (n_primal, n_tangent, n_cot, bool(half), int(slot))
        if key in self._plans:
            self._plans.move_to_end(key)
            return self._plans[key]
        while len(self._plans) >= self.max_cached_plans:
            _, old = self._plans.popitem(last=False)
            old.release()
        with torch.cuda.device(self.device):
            self._plans[key] = Plan(self, n_primal, n_tangent, n_cot, half=half)
        return self._plans[key]

    def release_plans(self):
        """Drop every cached plan and its workspace (e.g. before switching to another model)."""
        while self._plans:
            self._plans.popitem()[1].release()
        self.__dict__.pop("_pb_cache", None)

    def __call__(self, x, t):
        """eps = unet(x, t); x [B,C,R,R] fp32 on the device, t scalar (tensor or float)."""
        x = x.contiguous()
        return self.plan(x.shape[0]).forward(x, float(t))

    def jvp(self, x, t, V):
        """(eps, d_eps) for tangents V [k,3,R,R] at x [1,3,R,R] in one fused pass."""
        k = V.shape[0]
        p = self.plan(1, k, k)
        xin = torch.cat([x.reshape(1, *V.shape[1:]), V], 0).contiguous()
        out = p.forward(xin, float(t))
        return out[:1], out[1:]

    def vjp(self, k, g_eps):
        """J_eps^T g for k cotangents at the primal point of the last jvp() call."""
        return self.plan(1, k, k).vjp(g_eps.contiguous())

    def __del__(self):
        try:
            self.release_plans()
            self.lib.loco_unet_destroy(self.handle)
        except Exception:
            pass


class B200VAEDecoder(B200UNet):
    """The `self.vae` of the latent-space twins (src/modules/edit.py:770: `self.vae.decode(z).sample`):
    AutoencoderKL's decode = Decoder(post_quant_conv(z)) on the same CUDA executor (arch kind
    "vae_decoder": latent [in_ch, R, R] -> image [3, R << (levels-1), ...], no timestep input), with the
    fused primal + k-tangent pass and the k-cotangent pass the latent-space power method needs."""

    def __init__(self, arch, state_dict, device="cuda:0"):
        arch = dict(arch)
        arch["kind"] = "vae_decoder"
        arch.setdefault("in_ch", 4)
        arch.setdefault("attn_resolutions", ())
        from .weights import hf_autoencoderkl_to_decoder, is_hf_autoencoderkl_state_dict
        if is_hf_autoencoderkl_state_dict(state_dict):          # a diffusers AutoencoderKL checkpoint
            state_dict = hf_autoencoderkl_to_decoder(state_dict, arch)
        super().__init__(arch, state_dict, device=device)

    def decode(self, z):
        z = z.contiguous()
        return self.plan(z.shape[0]).forward(z, 0.0)

    def __call__(self, z, t=0.0):
        return self.decode(z)

    def jvp(self, z, dZ):
        """(x, dX): decode(z) [1,3,H,W] and the k tangents J_dec dZ in one fused pass."""
        k = dZ.shape[0]
        out = self.plan(1, k, k).forward(torch.cat([z.reshape(1, *self.in_shape), dZ.reshape(k, *self.in_shape)], 0).contiguous(), 0.0)
        return out[:1], out[1:]

    def vjp(self, k, g_x):
        """J_dec^T g for k image-space cotangents at the latent of the last jvp() call."""
        return self.plan(1, k, k).vjp(g_x.reshape(k, *self.out_shape).contiguous())
