"""Host-side mirror of the reference's `EditUncondDiffusion` hot path (src/modules/edit.py:2034-2625).

Method names, argument lists, file names and return values follow the reference so the class is a
drop-in for that path; the arithmetic runs in libloco_b200.so (no PyTorch fallback):

  reference (PyTorch)                                    here
  -----------------------------------------------------  -----------------------------------------
  torch.func.jacfwd over chunks of v       (:2449-2458)   one fused primal+k-tangent U-Net pass
  fresh forward + k sequential backwards   (:2460-2480)   one k-cotangent VJP pass over the saved
                                                          primal activations of the JVP pass
  torch.linalg.svd(v_)                     (:2482)        Gram + k x k Jacobi eigensolver kernel
  vT_null.T @ (vT_null @ vT_modify.T)      (:2317-2323)   nullspace_project kernels
  scheduler.step / unet per DDIM step      (:2544-2593)   graph-friendly kernels, no host staging

Deliberate deviations (documented in DESIGN.md): no CPU staging of latents (`buffer_device`,
:2562-2584) and of U (:2456-2458); row signs of the basis are aligned with the previous iterate so
the reference's convergence test (:2489-2494) is not defeated by SVD sign flips (`align_sign`);
`v0=` lets a caller inject the initial basis for deterministic parity runs.
"""
import collections
import os

import torch

from . import ops
from .scheduler import YHCustomScheduler


def _pb_workspace(unet, k, slot=0):
    """Power-method buffers for rank k, LRU-bounded (each entry pins the [k,d] iteration buffers; the
    (1,k,k) plan behind it lives in the U-Net's own bounded plan cache).  `slot` separates the buffers (and
    plans) of power methods that run concurrently on different streams."""
    cache = unet.__dict__.setdefault("_pb_cache", collections.OrderedDict())
    key = k if slot == 0 else (k, slot)
    if key in cache:
        cache.move_to_end(key)
        if cache[key].plan.released:          # its plan was evicted meanwhile: rebuild
            del cache[key]
    if key not in cache:
        while len(cache) >= unet.max_cached_plans:
            cache.popitem(last=False)
        cache[key] = ops.PullbackWorkspace(unet, k, slot=slot)
    return cache[key]


def random_basis(d, k, device, generator=None):
    """Algorithm-1 init (src/modules/edit.py:2435-2438: Q of qr(randn(d, k))): an orthonormalised
    Gaussian block, through the library's own orthonormalisation kernel (no cuSOLVER on the path).
    Same distribution (Haar) as the reference's; parity runs inject the reference's draw via v0=."""
    g = torch.randn(k, d, device=device, dtype=torch.float32, generator=generator)
    V, _ = ops.orthonormalise(g)
    return V


def local_basis(unet, scheduler, x, t, pca_rank, v0=None, min_iter=10, max_iter=100,
                convergence_threshold=1e-3, mask=None, noise=False, align_sign=True, verbose=True,
                chunk_size=None):
    """Power-method local basis of the (masked) PMP Jacobian at x_t.

    Restates `local_encoder_decoder_pullback_xt` (src/modules/edit.py:2406-2504): returns
    (u [l_o,k] = J V^T of the last iterate, s [k] = sqrt(svdvals(U^T J)), vT [k,d]).
    Up to `chunk_size` tangents (default 25, the reference's value) run in one fused pass; larger
    ranks are probed chunk by chunk exactly like the reference's `v.chunk(num_chunk)` (:2419, :2448)
    -- the 1+k rows of activations of a k = 64 plan would not fit 180 GB -- and orthonormalised
    together."""
    k = int(pca_rank)
    chunk_size = 25 if chunk_size is None else int(chunk_size)
    if k > chunk_size:
        return _local_basis_chunked(unet, scheduler, x, t, k, v0, min_iter, max_iter,
                                    convergence_threshold, mask, noise, align_sign, verbose, chunk_size)
    ws = _pb_workspace(unet, k)
    d = ws.d
    dev = unet.device
    x = x.to(device=dev, dtype=torch.float32).contiguous().reshape(1, -1)
    assert x.numel() == d
    t_host = float(t)
    at = scheduler.alpha_at(t_host)
    mask_u8 = None
    if mask is not None:
        mask_u8 = mask.to(device=dev).reshape(-1).to(torch.uint8).contiguous()
        assert mask_u8.numel() == d, "mask must cover (c, h, w) of x"
    if v0 is None:
        v0 = random_basis(d, k, dev)
    cur, nxt = ws.V
    cur.copy_(v0.reshape(k, d))
    for i in range(max_iter):
        ws.iterate(x, t_host, at, mask_u8, noise, cur, nxt, align_sign=align_sign)
        need_check = i > min_iter
        if verbose or need_check:
            convergence = torch.dist(cur, nxt).item()                       # :2489
            if verbose:
                print(f'power method : {i}-th step convergence : ', convergence)
        cur, nxt = nxt, cur
        if need_check and torch.allclose(nxt, cur, atol=convergence_threshold):   # :2492
            if verbose:
                print('reach convergence threshold : ', convergence)
            break
    ws.V = [cur, nxt]
    vT = cur.clone()
    s = ws.s.clone()
    if mask_u8 is None:
        u = ws.u_full.clone()
    else:
        idx = ops.mask_indices(mask_u8)
        u = ops.gather_rows(ws.u_full, idx)
    return u.T, s, vT


def _local_basis_chunked(unet, scheduler, x, t, k, v0, min_iter, max_iter, convergence_threshold, mask,
                         noise, align_sign, verbose, chunk_size):
    """Rank k > chunk_size: probe in chunks (torch.chunk sizes, src/modules/edit.py:2419, 2448)."""
    dev = unet.device
    R = unet.arch["resolution"]
    d = 3 * R * R
    x = x.to(device=dev, dtype=torch.float32).contiguous().reshape(1, -1)
    t_host = float(t)
    at = scheduler.alpha_at(t_host)
    mask_u8 = None if mask is None else mask.to(device=dev).reshape(-1).to(torch.uint8).contiguous()
    if v0 is None:
        v0 = random_basis(d, k, dev)
    V = v0.reshape(k, d).contiguous().clone()
    num_chunk = k // chunk_size if k % chunk_size == 0 else k // chunk_size + 1
    sizes = [c.shape[0] for c in torch.empty(k, 1).chunk(num_chunk)]
    W = torch.empty(k, d, dtype=torch.float32, device=dev)
    U = torch.empty(k, d, dtype=torch.float32, device=dev)
    s = None
    for i in range(max_iter):
        lo = 0
        for kc in sizes:
            ws = _pb_workspace(unet, kc)
            u_c, w_c = ws.probe(x, t_host, at, mask_u8, noise, V[lo:lo + kc].contiguous())
            U[lo:lo + kc].copy_(u_c)
            W[lo:lo + kc].copy_(w_c)
            lo += kc
        V_new, s = ops.orthonormalise(W, v_prev=V if align_sign else None)
        need_check = i > min_iter
        if verbose or need_check:
            convergence = torch.dist(V, V_new).item()
            if verbose:
                print(f'power method : {i}-th step convergence : ', convergence)
        done = need_check and torch.allclose(V, V_new, atol=convergence_threshold)
        V = V_new
        if done:
            break
    u = U if mask_u8 is None else ops.gather_rows(U, ops.mask_indices(mask_u8))
    return u.T, s, V


def local_basis_pair(unet, scheduler, x, t, k, k_null, mask, v0=None, v0_null=None, n_iter=12,
                     noise=False, align_sign=True, slot=0):
    """Edit basis (mask) and null basis (~mask) of run_edit_null_space_projection
    (src/modules/edit.py:2294-2310) computed together: both probe the Jacobian at the same x_t, so
    their k + k_null tangents share one fused JVP pass, one VJP pass and one set of primal
    activations per iteration.  Fixed iteration count (the reference's loop with min_iter >=
    max_iter).  Returns (vT_modify [k,d], s_modify, vT_null [k_null,d], s_null)."""
    kt = k + k_null
    ws = _pb_workspace(unet, kt, slot)
    d = ws.d
    dev = unet.device
    x = x.to(device=dev, dtype=torch.float32).contiguous().reshape(1, -1)
    t_host = float(t)
    at = scheduler.alpha_at(t_host)
    mask_u8 = mask.to(device=dev).reshape(-1).to(torch.uint8).contiguous()
    cur, nxt = ws.V
    for v, lo, kk in ((v0, 0, k), (v0_null, k, k_null)):
        if v is None:
            v = random_basis(d, kk, dev)
        cur[lo:lo + kk].copy_(v.reshape(kk, d))
    for _ in range(n_iter):
        ws.iterate_pair(x, t_host, at, mask_u8, noise, cur, nxt, k, k_null, align_sign=align_sign)
        cur, nxt = nxt, cur
    ws.V = [cur, nxt]
    s = ws.s.clone()
    return cur[:k].clone(), s[:k], cur[k:].clone(), s[k:]


def local_basis_pair_converging(unet, scheduler, x, t, k, k_null, mask, v0=None, v0_null=None, min_iter=10,
                                max_iter=50, convergence_threshold=1e-4, noise=False, align_sign=True,
                                verbose=False):
    """The two power methods of `run_edit_null_space_projection` (src/modules/edit.py:2294-2310: the edit basis
    through `mask`, then the null basis through `~mask`, each with the loop and the stopping rule of :2440-2497)
    advanced TOGETHER while both still iterate: their tangents probe the same Jacobian at the same x_t, so one
    fused (1, k + k_null) pass per iteration serves both (rows are independent: each basis sees exactly the
    iterates of its own loop).  Convergence is tested per basis with the reference's rule; when one basis stops,
    the other finishes alone from its current iterate with the remaining iteration budget.
    Returns (vT_modify [k,d], s_modify, vT_null [k_null,d], s_null)."""
    kt = k + k_null
    ws = _pb_workspace(unet, kt)
    d = ws.d
    dev = unet.device
    xr = x.to(device=dev, dtype=torch.float32).contiguous().reshape(1, -1)
    t_host = float(t)
    at = scheduler.alpha_at(t_host)
    mask_u8 = mask.to(device=dev).reshape(-1).to(torch.uint8).contiguous()
    cur, nxt = ws.V
    for v, lo, kk in ((v0, 0, k), (v0_null, k, k_null)):       # same order of draws as the two separate calls
        if v is None:
            v = random_basis(d, kk, dev)
        cur[lo:lo + kk].copy_(v.reshape(kk, d))
    stop = [False, False]
    it = 0
    for it in range(max_iter):
        ws.iterate_pair(xr, t_host, at, mask_u8, noise, cur, nxt, k, k_null, align_sign=align_sign)
        need_check = it > min_iter
        if verbose:
            for name, sl in (("modify", slice(0, k)), ("null", slice(k, kt))):
                print(f'power method ({name}) : {it}-th step convergence : ', torch.dist(cur[sl], nxt[sl]).item())
        cur, nxt = nxt, cur
        if need_check:
            stop = [bool(torch.allclose(nxt[:k], cur[:k], atol=convergence_threshold)),      # :2492
                    bool(torch.allclose(nxt[k:], cur[k:], atol=convergence_threshold))]
            if stop[0] or stop[1]:
                break
    ws.V = [cur, nxt]
    vT_mod, vT_null = cur[:k].clone(), cur[k:].clone()
    s = ws.s.clone()
    s_mod, s_null = s[:k], s[k:]
    left = max_iter - (it + 1)
    if (stop[0] != stop[1]) and left > 0:
        # one loop ended: the other one goes on by itself (its check is already active: min_iter = -1)
        if stop[0]:
            _, s_null, vT_null = local_basis(unet, scheduler, x, t, k_null, v0=vT_null, min_iter=-1, max_iter=left,
                                             convergence_threshold=convergence_threshold, mask=~mask.to(dev),
                                             noise=noise, align_sign=align_sign, verbose=verbose)
        else:
            _, s_mod, vT_mod = local_basis(unet, scheduler, x, t, k, v0=vT_mod, min_iter=-1, max_iter=left,
                                           convergence_threshold=convergence_threshold, mask=mask.to(dev),
                                           noise=noise, align_sign=align_sign, verbose=verbose)
    return vT_mod, s_mod, vT_null, s_null


class SyntheticDataset(object):
    """Seeded stand-in for the reference's datasets (src/utils/utils.py:472-672,
    src/dataset/celeba_hq_dataloader.py): `ds[idx] -> [1,3,R,R]` in [-1,1], `getmask -> bool[3,R,R]`.
    (image clamp(0.5*randn,-1,1) seeded by idx; rectangle mask.)"""

    def __init__(self, resolution=256):
        self.R = resolution

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(int(idx))
        return (0.5 * torch.randn(1, 3, self.R, self.R, generator=g)).clamp(-1, 1)

    def getmask(self, idx=0, choose_sem=None):
        R = self.R
        m = torch.zeros(3, R, R, dtype=torch.bool)
        m[:, (3 * R) // 8:(5 * R) // 8, R // 4:(3 * R) // 4] = True
        return m


class EditUncondDiffusion(object):
    """Drop-in for the reference class of the same name (src/modules/edit.py:2034-2625).

    `unet` / `dataset` may be injected (random-init synthetic weights otherwise: there is no
    network for checkpoints); everything else is read from `args` exactly like the reference."""

    def __init__(self, args, unet=None, dataset=None):
        self.pca_device = getattr(args, "pca_device", "cpu")
        self.buffer_device = getattr(args, "buffer_device", "cpu")     # accepted, unused (no staging)
        self.memory_bound = getattr(args, "memory_bound", 50)
        self.device = torch.device(args.device)
        self.dtype = args.dtype
        self.seed = getattr(args, "seed", 0)
        self.save_result_as = getattr(args, "save_result_as", "image")
        self.model_name = getattr(args, "model_name", "")
        self.image_size = getattr(args, "image_size", 256)
        self.c_in = 3
        if self.dtype != torch.float32:
            raise NotImplementedError("the uncond hot path runs in fp32/TF32 (all reference scripts use --dtype fp32)")
        if unet is None:
            from .unet import B200UNet
            from .weights import (DDPM256, P2_256, hf_unet2d_to_ddpm, is_hf_unet2d_state_dict,
                                  random_state_dict)
            # the reference picks the network family by model name (utils/utils.py:95-131):
            # "*_P2" -> guided-diffusion UNetModel(P2_DICT); "*_HF" -> the DDPM U-Net architecture
            if self.model_name == "FFHQ_HF":
                # google/ncsnpp-ffhq-256 (utils/utils.py:99-100) is an NCSN++ score network, not this U-Net
                raise NotImplementedError("FFHQ_HF (google/ncsnpp-ffhq-256, NCSN++) is not an architecture of "
                                          "this hot path; the DDPM (*_HF, CelebA_HQ) and P2 (*_P2) U-Nets are")
            base = P2_256 if self.model_name.endswith("_P2") else DDPM256
            arch = dict(base, resolution=self.image_size)
            wp = getattr(args, "weights_path", "")
            sd = torch.load(wp, map_location="cpu") if wp else random_state_dict(arch, seed=1234)
            if base is DDPM256 and is_hf_unet2d_state_dict(sd):
                sd = hf_unet2d_to_ddpm(sd, arch)     # a diffusers UNet2DModel checkpoint (*_HF models)
            unet = B200UNet(arch, sd, device=self.device)
        self.unet = unet
        self.scheduler = YHCustomScheduler(args, device=self.device)
        self.dataset = dataset if dataset is not None else SyntheticDataset(self.image_size)
        self.dataset_name = getattr(args, "dataset_name", "CelebA_HQ_mask")
        self.for_steps = args.for_steps
        self.inv_steps = args.inv_steps
        self.use_yh_custom_scheduler = getattr(args, "use_yh_custom_scheduler", True)
        self.edit_t = args.edit_t
        self.scheduler.set_timesteps(self.for_steps, device=self.device)
        # src/modules/edit.py:2072-2073
        self.edit_t_idx = int((self.scheduler.timesteps - self.edit_t * 1000).abs().argmin())
        pbt = getattr(args, "performance_boosting_t", 0.0)
        self.performance_boosting_t_idx = int((self.scheduler.timesteps - pbt * 1000).abs().argmin()) if pbt > 0 else 1000
        self.use_x_space_guidance = getattr(args, "use_x_space_guidance", False)
        self.x_space_guidance_edit_step = getattr(args, "x_space_guidance_edit_step", 1)
        self.x_space_guidance_scale = getattr(args, "x_space_guidance_scale", 0)
        self.x_space_guidance_num_step = getattr(args, "x_space_guidance_num_step", 0)
        rf = getattr(args, "result_folder", "./runs/")
        if self.dataset_name == "Random":
            self.result_folder = os.path.join(rf, f"sample_seed{self.seed}")
        else:
            self.result_folder = os.path.join(rf, f"sample_idx{getattr(args, 'sample_idx', 0)}")
        os.makedirs(self.result_folder, exist_ok=True)
        self.obs_folder = getattr(args, "obs_folder", self.result_folder)
        self.vT_path = getattr(args, "vT_path", "")
        self.vT1_path = getattr(args, "vT1_path", "")
        self.mask_type = getattr(args, "mask_type", "SAM")
        self.args = args
        self.verbose = getattr(args, "verbose", True)
        self.save_images = getattr(args, "save_images", True)
        self.align_sign = getattr(args, "align_sign", True)
        self.v0 = None            # optional injected initial basis (parity runs)
        # run_edit_null_space_projection: advance the edit and the null power method in one fused pass per iteration
        # while both still iterate (same iterates and stopping rule per basis as two separate calls)
        self.fuse_bases = bool(getattr(args, "fuse_bases", True))
        self.noise_fn = None      # optional callable(i, xt) -> eta=1 noise (parity runs)
        self.last_images = []     # outputs of the performance-boosted DDIM passes

    # ------------------------------------------------------------------ simple experiments
    @torch.no_grad()
    def run_DDIMforward(self, num_samples=5):
        self.EXP_NAME = 'DDIMforward'
        xT = torch.randn(num_samples, self.c_in, self.image_size, self.image_size, device=self.device, dtype=self.dtype)
        return self.DDIMforwardsteps(xT, t_start_idx=0, t_end_idx=-1, vis_psd=False)

    @torch.no_grad()
    def run_DDIMinversion(self, idx):
        """src/modules/edit.py:2117-2167."""
        EXP_NAME = f'DDIMinversion-{self.dataset_name}_{idx}'
        self.scheduler.set_timesteps(self.inv_steps, device=self.device, is_inversion=True)
        n = len(self.scheduler._ts_host)
        x0 = self.dataset[idx]
        self._save_image(x0, 'original.png')
        xt = x0.to(self.device, dtype=self.dtype).contiguous()
        for i in range(n):
            if i == n - 1:
                break
            t = self.scheduler._ts_host[i]
            et = self.unet(xt, t)
            xt = self.scheduler.step(et, t, xt, eta=0, t_idx=i).prev_sample
        self._save_image(xt, f'xT-{EXP_NAME}.png')
        return xt

    # ------------------------------------------------------------------ editing drivers
    @torch.no_grad()
    def group_edit_null_space_projection(self, idx, **kwargs):
        """src/modules/edit.py:2171-2212 (two directions, cumulative; `vT_paths=` generalises to n)."""
        if self.dataset_name == 'Random':
            xT = torch.randn(1, 3, self.image_size, self.image_size, dtype=self.dtype, device=self.device)
        else:
            xT = self.run_DDIMinversion(idx=idx)
        xt, t, t_idx = self.DDIMforwardsteps(xT, t_start_idx=0, t_end_idx=self.edit_t_idx)
        assert t_idx == self.edit_t_idx
        paths = kwargs.get("vT_paths") or [self.vT_path, self.vT1_path]
        vT_list = [torch.load(p, map_location=self.device) for p in paths]
        BASIS_NAME = f"load-basis-{len(vT_list)}"
        xt_temp = xt.detach().clone()
        xt_vis_list = [xt_temp]
        for vT in vT_list:
            vk = vT[0, :].view(-1, *xt.shape[1:]).to(self.dtype).contiguous()
            xt_temp = ops.axpy(xt_temp, vk, self.x_space_guidance_scale * self.x_space_guidance_num_step)
            xt_vis_list.append(xt_temp)
        self.EXP_NAME = f'{idx}-Edit_xt-noise-{BASIS_NAME}'
        xt_vis = torch.cat(xt_vis_list, dim=0)
        self.DDIMforwardsteps(xt_vis, t_start_idx=self.edit_t_idx, t_end_idx=-1, performance_boosting=True)
        return xt

    def _get_masks(self, idx, use_mask):
        """Mask sources of src/modules/edit.py:2234-2267.
          CelebA_HQ_mask : the dataset's ground-truth semantic mask (:2248-2251), always used;
          Random / FFHQ / AFHQ / ... : `mask/mask.pt` (bool [n,1,res,res] written by
            MaskSegmentation.mask_segmentation, mask_segmentation.py:18-26), row `--mask_index`,
            repeated over the 3 channels (:2247, :2262-2265); `use_mask=False` -> None (unmasked
            pull-back, :2266-2267).
        The SAM network that writes mask.pt in the reference is out of scope: a
        missing file is an error here, never a silent substitute."""
        if self.dataset_name == "CelebA_HQ_mask":
            return self.dataset.getmask(idx=getattr(self.args, "sample_idx", idx),
                                        choose_sem=getattr(self.args, "choose_sem", None))
        if not use_mask and self.dataset_name != "Random":
            return None
        mpath = os.path.join(self.result_folder, "mask/mask.pt")
        if not os.path.exists(mpath):
            if hasattr(self.dataset, "getmask") and getattr(self.args, "allow_dataset_mask", False):
                return self.dataset.getmask(idx=idx, choose_sem=None)
            raise FileNotFoundError(
                f"{mpath} not found: dataset '{self.dataset_name}' takes its masks from a cached mask.pt "
                "(the reference generates it with SAM, which this path does not ship). Write a bool "
                "[n,1,res,res] tensor there (loco_edit_b200.masks.save_masks), or pass "
                "--allow_dataset_mask True to use the synthetic dataset's rectangle on purpose.")
        from .masks import load_mask
        return load_mask(self.result_folder, getattr(self.args, "mask_index", 0))

    @torch.no_grad()
    def run_edit_null_space_projection(
            self, idx, vis_num, vis_num_pc=5, pca_rank=50, pca_rank_null=10, op='mid', block_idx=0,
            null_space_projection=True, encoder_decoder_by_et=False, use_mask=True, random_edit=False,
            **kwargs):
        """src/modules/edit.py:2216-2366."""
        if self.dataset_name == 'Random':
            xT = torch.randn(1, self.c_in, self.image_size, self.image_size, dtype=self.dtype, device=self.device)
        else:
            xT = self.run_DDIMinversion(idx=idx)
        mask = self._get_masks(idx, use_mask)
        if getattr(self.args, "sampling_mode", False):
            return None
        if mask is not None:
            mask = mask.to(self.device)

        xt, t, t_idx = self.DDIMforwardsteps(xT, t_start_idx=0, t_end_idx=self.edit_t_idx)
        assert t_idx == self.edit_t_idx

        sem = getattr(self.args, "choose_sem", None) if self.dataset_name == "CelebA_HQ_mask" else getattr(self.args, "mask_index", 0)
        if not os.path.exists(self.vT_path):
            save_dir = os.path.join(self.result_folder, "basis", f'local_basis-{self.edit_t}T-select-mask-{sem}')
            os.makedirs(save_dir, exist_ok=True)
            vT_modify_path = os.path.join(save_dir, f'vT-modify-pca-rank-{pca_rank}.pt')
            vT_null_path = os.path.join(save_dir, f'vT-null-{pca_rank_null}.pt')
            vT_null_fused = None
            if (self.fuse_bases and mask is not None and null_space_projection and not os.path.exists(vT_modify_path)
                    and not os.path.exists(vT_null_path) and pca_rank + pca_rank_null <= 25):
                # neither basis is cached: both power methods advance in one fused pass per iteration
                vT_modify, vT_null_fused = self.local_encoder_decoder_pullback_xt_pair(
                    x=xt, t=t, pca_rank=pca_rank, pca_rank_null=pca_rank_null, min_iter=10, max_iter=50,
                    convergence_threshold=1e-4, mask=mask, noise=encoder_decoder_by_et)
                torch.save(vT_modify, vT_modify_path)
            elif os.path.exists(vT_modify_path):
                vT_modify = torch.load(vT_modify_path, map_location=self.device).type(self.dtype)
            else:
                u_modify, s_modify, vT_modify = self.local_encoder_decoder_pullback_xt(
                    x=xt, t=t, op=op, block_idx=block_idx, pca_rank=pca_rank,
                    min_iter=10, max_iter=50, convergence_threshold=1e-4, mask=mask, noise=encoder_decoder_by_et)
                torch.save(vT_modify, vT_modify_path)
            if vT_null_fused is not None:
                vT_null = vT_null_fused
                torch.save(vT_null, vT_null_path)
            elif null_space_projection and os.path.exists(vT_null_path):
                vT_null = torch.load(vT_null_path, map_location=self.device).type(self.dtype)
            elif not null_space_projection:
                vT_null = None
            else:
                u_null, s_null, vT_null = self.local_encoder_decoder_pullback_xt(
                    x=xt, t=t, op=op, block_idx=block_idx, pca_rank=pca_rank_null,
                    min_iter=10, max_iter=50, convergence_threshold=1e-4,
                    mask=(~mask if mask is not None else None), noise=encoder_decoder_by_et)
                torch.save(vT_null, vT_null_path)
            if random_edit:
                vT_modify = torch.randn_like(vT_modify)
            # :2317-2323
            if not null_space_projection:
                vT = ops.nullspace_project(vT_modify.contiguous(), None, project=False)
            else:
                vT = ops.nullspace_project(vT_modify.contiguous(), vT_null[:pca_rank_null, :].contiguous(), project=True)
            BASIS_NAME = f"{encoder_decoder_by_et}_{sem}-edit_{self.edit_t}T_null_proj_{null_space_projection}_rank{pca_rank_null}_scale_{self.x_space_guidance_scale}"
            for pc_idx in range(max(vis_num_pc, vT.shape[0])):        # :2329 (IndexError if > k, as in the reference)
                self.EXP_NAME = f'{idx}-Edit_xt-noise-{BASIS_NAME}-pc_{pc_idx:0=3d}'
                torch.save(vT[[pc_idx], :], os.path.join(save_dir, f'{self.EXP_NAME}-vT.pt'))
        else:
            vT = torch.load(self.vT_path, map_location=self.device).type(self.dtype)
            BASIS_NAME = f"edit_{self.edit_t}T-load-basis-'{os.path.basename(self.vT_path)}'"

        # edit (:2339-2364)
        original_xt = xt.detach()
        self.last_images = []
        for pc_idx in range(min(vis_num_pc, vT.shape[0])):
            self.EXP_NAME = f'{idx}-Edit-random{random_edit}_xt-noise-{BASIS_NAME}-pc_{pc_idx:0=3d}'
            xt = self.build_edit_batch(original_xt, vT[pc_idx, :], vis_num)
            self.DDIMforwardsteps(xt, t_start_idx=self.edit_t_idx, t_end_idx=-1, performance_boosting=True)
        return xt

    def build_edit_batch(self, original_xt, v_row, vis_num):
        """src/modules/edit.py:2341-2363: +/- direction, num_step cumulative edits, subsample."""
        xts = {}
        for direction in [1, -1]:
            vk = (direction * v_row).view(-1, *original_xt.shape[1:]).contiguous()
            xt_list = [original_xt.clone()]
            for _ in range(self.x_space_guidance_num_step):
                xt_list.append(self.x_space_guidance_direct(
                    xt_list[-1], t_idx=self.edit_t_idx, vk=vk, single_edit_step=self.x_space_guidance_edit_step))
            xt = torch.cat(xt_list, dim=0)
            xt = xt[[0, -1], :] if vis_num == 1 else xt[::(xt.size(0) // vis_num)]
            xts[direction] = xt
        return torch.cat([(xts[-1].flip(dims=[0]))[:-1], xts[1]], dim=0).contiguous()

    # ------------------------------------------------------------------ hot-path pieces
    def get_x0(self, t, x, mask=None):
        """src/modules/edit.py:2369-2391."""
        et = self.unet(x.contiguous(), float(t))
        P_xt = ops.pmp_forward(x.contiguous(), et, self.scheduler.alpha_at(float(t)))
        if mask is not None:
            idx = ops.mask_indices(mask.to(self.device))
            P_xt = ops.gather_rows(P_xt.reshape(P_xt.shape[0], -1), idx)
        return P_xt

    def get_et(self, t, x, mask=None):
        """src/modules/edit.py:2394-2403."""
        et = self.unet(x.contiguous(), float(t))
        if mask is not None:
            idx = ops.mask_indices(mask.to(self.device))
            et = ops.gather_rows(et.reshape(et.shape[0], -1), idx)
        return et

    def local_encoder_decoder_pullback_xt_pair(self, x, t, pca_rank, pca_rank_null, mask, min_iter=10, max_iter=50,
                                               convergence_threshold=1e-4, noise=False):
        """Edit basis (mask) and null basis (~mask) of src/modules/edit.py:2294-2310 in one fused loop
        (`local_basis_pair_converging`); returns (vT_modify, vT_null)."""
        v0a = v0b = None
        if isinstance(self.v0, dict):               # injected initial bases (parity runs), keyed by rank
            v0a, v0b = self.v0.get(pca_rank), self.v0.get(pca_rank_null)
        vm, _, vn, _ = local_basis_pair_converging(
            self.unet, self.scheduler, x, t, pca_rank, pca_rank_null, mask, v0=v0a, v0_null=v0b, min_iter=min_iter,
            max_iter=max_iter, convergence_threshold=convergence_threshold, noise=noise,
            align_sign=self.align_sign, verbose=self.verbose)
        return vm, vn

    def local_encoder_decoder_pullback_xt(
            self, x, t, op=None, block_idx=None,
            pca_rank=50, chunk_size=25, min_iter=10, max_iter=100, convergence_threshold=1e-3,
            mask=None, noise=False, v0=None):
        """src/modules/edit.py:2406-2504; `op`/`block_idx` are accepted and ignored like there."""
        if v0 is None and self.v0 is not None:      # injected initial basis (dict keyed by rank, or tensor)
            v0 = self.v0.get(pca_rank) if isinstance(self.v0, dict) else self.v0
        return local_basis(self.unet, self.scheduler, x, t, pca_rank, v0=v0,
                           min_iter=min_iter, max_iter=max_iter, convergence_threshold=convergence_threshold,
                           mask=mask, noise=noise, align_sign=self.align_sign, verbose=self.verbose,
                           chunk_size=chunk_size)

    @torch.no_grad()
    def DDIMforwardsteps(self, xt, t_start_idx, t_end_idx, vis_psd=False, save_image=True,
                         return_xt=True, performance_boosting=False):
        """src/modules/edit.py:2508-2614 (CPU staging of the latents dropped; `memory_bound` chunking kept)."""
        assert (t_start_idx < self.for_steps) & (t_end_idx <= self.for_steps)
        self.scheduler.set_timesteps(self.for_steps, device=self.device)
        ts = self.scheduler._ts_host
        n = len(ts)
        xt = xt.to(device=self.device, dtype=self.dtype).contiguous()
        for i in range(n):
            if t_end_idx == i:
                return xt, self.scheduler.timesteps[i], i
            elif i < t_start_idx:
                continue
            boost = performance_boosting and (self.performance_boosting_t_idx <= i) and \
                (self.performance_boosting_t_idx != n - 1)
            eta = 1 if boost else 0
            if xt.size(0) <= self.memory_bound:
                et = self.unet(xt, ts[i])
            else:   # src/modules/edit.py:2564-2576: the batch goes through the U-Net in chunks
                et = torch.cat([self.unet(c.contiguous(), ts[i]) for c in xt.split(self.memory_bound)], 0)
            noise = self.noise_fn(i, xt) if (eta and self.noise_fn is not None) else None
            xt = self.scheduler.step(et, ts[i], xt, eta=eta, t_idx=i, noise=noise).prev_sample
        if performance_boosting:
            self.last_images.append(xt)
        if save_image:
            self._save_image(xt, f'{getattr(self, "EXP_NAME", "DDIMforward")}.png', nrow=xt.size(0))
        if return_xt:
            return xt
        return

    @torch.no_grad()
    def x_space_guidance_direct(self, xt, t_idx, vk, single_edit_step):
        """src/modules/edit.py:2618-2625."""
        return ops.axpy(xt.contiguous(), vk.expand_as(xt).contiguous(), self.x_space_guidance_scale * single_edit_step)

    def _save_image(self, x, name, nrow=8):
        if not self.save_images:
            return
        try:
            import torchvision.utils as tvu
            tvu.save_image((x / 2 + 0.5).clamp(0, 1), os.path.join(self.result_folder, name), nrow=nrow)
        except Exception as e:   # image writing is I/O polish, never fatal for the hot path
            print("image not written:", e)
