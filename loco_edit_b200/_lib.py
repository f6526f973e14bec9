"""ctypes binding of libloco_b200.so (the C ABI declared in include/loco_b200.h).

The library is the only compute path of this package: if it cannot be loaded the import fails
loudly, and every compute entry point fails when no CUDA device is present (no CPU fallback).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libloco_b200.so")


class LocoError(RuntimeError):
    pass


class Arch(C.Structure):
    """Mirror of loco_arch_t."""
    _fields_ = [
        ("ch", C.c_int), ("n_levels", C.c_int), ("ch_mult", C.c_int * 8),
        ("num_res_blocks", C.c_int), ("n_attn", C.c_int), ("attn_resolutions", C.c_int * 4),
        ("resolution", C.c_int), ("in_ch", C.c_int), ("out_ch", C.c_int), ("gn_eps", C.c_float),
        ("kind", C.c_int), ("head_ch", C.c_int), ("ctx_dim", C.c_int), ("ctx_heads", C.c_int),
    ]


_P = C.c_void_p
_I = C.c_int
_LL = C.c_longlong
_F = C.c_float

# name -> (restype, argtypes); every symbol declared in include/loco_b200.h
PROTOTYPES = {
    "loco_abi_version": (_I, []),
    "loco_last_error": (C.c_char_p, []),
    "loco_launch_count": (_LL, []),
    "loco_profile_enable": (_I, [_I]),
    "loco_profile_collect": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_LL), _I]),
    "loco_unet_create": (_I, [C.POINTER(Arch), C.POINTER(_P)]),
    "loco_unet_destroy": (None, [_P]),
    "loco_unet_weight_floats": (_LL, [_P]),
    "loco_unet_bind_weights": (_I, [_P, _P]),
    "loco_unet_num_params": (_I, [_P]),
    "loco_unet_param_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_I), C.POINTER(_I)]),
    "loco_unet_load_param": (_I, [_P, C.c_char_p, _P, _LL, _P]),
    "loco_plan_create": (_I, [_P, _I, _I, _I, C.POINTER(_P)]),
    "loco_plan_create_ex": (_I, [_P, _I, _I, _I, _I, C.POINTER(_P)]),
    "loco_plan_destroy": (None, [_P]),
    "loco_plan_workspace_bytes": (_LL, [_P]),
    "loco_plan_bind": (_I, [_P, _P]),
    "loco_plan_info": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_I), C.POINTER(_I)]),
    "loco_unet_forward": (_I, [_P, _P, _F, _P, _P]),
    "loco_plan_set_condition": (_I, [_P, _P, _P]),
    "loco_plan_set_context": (_I, [_P, _P, _I, _P]),
    "loco_unet_vjp": (_I, [_P, _P, _P, _P]),
    "loco_pullback_scratch_bytes": (_LL, [_I, _LL]),
    "loco_pullback_iteration": (_I, [_P, _P, _F, _F, _P, _I, _P, _I, _LL, _I, _P, _P, _P, _P, _P, _P]),
    "loco_pullback_probe": (_I, [_P, _P, _F, _F, _P, _I, _P, _I, _LL, _P, _P, _P, _P]),
    "loco_pullback_probe_pair": (_I, [_P, _P, _F, _F, _P, _I, _P, _I, _I, _LL, _P, _P, _P, _P]),
    "loco_pullback_pair_iteration": (_I, [_P, _P, _F, _F, _P, _I, _P, _I, _I, _LL, _I, _P, _P, _P, _P, _P, _P]),
    "loco_pmp_forward": (_I, [_P, _P, _F, _LL, _P, _P]),
    "loco_combine3": (_I, [_P, _F, _P, _F, _P, _F, _LL, _P, _P]),
    "loco_pmp_jvp_epilogue": (_I, [_P, _P, _P, _F, _I, _I, _I, _LL, _P, _P, _P, _P]),
    "loco_orthonormalise_scratch_bytes": (_LL, [_I]),
    "loco_orthonormalise": (_I, [_P, _I, _LL, _P, _P, _P, _P, _P]),
    "loco_nullspace_project": (_I, [_P, _I, _P, _I, _LL, _I, _P, _P, _P]),
    "loco_ddim_step": (_I, [_P, _P, _P, _F, _F, _F, _LL, _P, _P, _P]),
    "loco_axpy": (_I, [_P, _P, _F, _LL, _P, _P]),
    "loco_mask_indices": (_I, [_P, _LL, _P, _P, _P]),
    "loco_gather_rows": (_I, [_P, _I, _LL, _P, _I, _P, _P]),
    "loco_scatter_rows": (_I, [_P, _I, _LL, _P, _I, _P, _P]),
    "loco_gram": (_I, [_P, _I, _P, _I, _LL, _P, _P]),
    "loco_conv_halo_eligible": (_I, [_I, _I, _I, _I, _I]),
    "loco_conv2d_fused_nhwc": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _I, _P]),
    "loco_conv2d_nhwc": (_I, [_I, _P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _I, _P, _I, _P, _P, _LL, _P]),
    "loco_conv2d_nhwc_ex": (_I, [_I, _P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _I, _P, _I, _P, _P, _LL, _I, _I, _P]),
    "loco_conv_bench": (_I, [_I, _P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _LL, _I, _I, C.POINTER(_F),
                             C.POINTER(_I), C.POINTER(_I), _P]),
    "loco_conv_bench_ex": (_I, [_I, _P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _LL, _I, _I, _I, _I, _P, _P,
                                C.POINTER(_F), C.POINTER(_I), C.POINTER(_I), _P]),
    "loco_groupnorm_silu_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _F, _I, _P, _P, _P]),
    "loco_groupnorm_silu_vjp": (_I, [_P, _I, _I, _I, _P, _I, _P, _P, _F, _I, _P, _P, _P]),
    "loco_groupnorm_silu_fwd_ex": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _F, _I, _P, _P, _I, _P]),
    "loco_groupnorm_silu_vjp_ex": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P, _F, _I, _P, _I, _P, _P, _I, _P]),
    "loco_attention_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "loco_cross_attention_fwd": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P]),
    "loco_cross_attention_vjp": (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P]),
    "loco_attention_vjp": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (building is a separate, explicit step: loco_edit_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LocoError(
            "libloco_b200.so is missing (%s). Build it with `python -m loco_edit_b200.build`; "
            "this package has no CPU/PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().loco_last_error()
        raise LocoError("%s failed (code %d): %s" % (what or "loco call", rc, (msg or b"").decode()))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """torch's current stream ON `device` (a tensor, a torch.device or None = current device).
    Kernels are enqueued on the device that owns the tensors, not on the thread's current device, so
    `--device cuda:1` needs no prior torch.cuda.set_device (the library switches devices itself)."""
    import torch
    if isinstance(device, torch.Tensor):
        device = device.device
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
