"""T-LOCO Edit in the latent space of a latent-diffusion model: the `EditStableDiffusion` twin of the
editing-direction hot path (reference: src/modules/edit.py:483-1196).

The Jacobian of this twin is  J = d(mask o VAE.decode(z0_hat(z_t) / 0.18215)) / d z_t  : the noise
prediction of a U-Net over 4-channel latents under classifier-free guidance, the posterior-mean
formula, and the VAE DECODER (`get_x0`, :757-781).  Every Jacobian product therefore runs through two
networks on the CUDA executor:

    J V^T  : one fused primal + k-tangent pass of the U-Net per guidance conditioning (combined
             linearly, e = sum_i w_i eps(z, t, c_i)), the PMP formula on all 1 + k rows, then ONE fused
             primal + k-tangent pass of the decoder (arch kind "vae_decoder", csrc/unet.cu);
    J^T U  : the decoder's k-cotangent pass, then the U-Net's per conditioning.

Built: `_classifer_free_guidance` (:636-674, four modes), `get_x0` (:757-781), `DDIMforwardsteps`
(:677-755), `local_encoder_decoder_pullback_zt` (:830-915), `get_delta_zt_via_grad` (:784-828),
`run_edit_null_space_projection_zt` (:918-1043) and `run_edit_null_space_projection_zt_semantic` (:1045-1175)
with the reference's basis file names, and
`x_space_guidance_direct` (:1177-1185).

What is NOT the reference's: the two networks.  Stable Diffusion's U-Net and VAE are diffusers models
(`diffusers==0.11.0`, not under /root/reference, no checkpoints offline; SURVEY 8c "parity unpinned at the
network level").  The U-Net stand-in is this package's text-conditioned U-Net over 4-channel latents
(`weights.latent_unet_arch`), the decoder is the published AutoencoderKL decoder architecture
(`weights.SD_VAE_DECODER`, `unet.B200VAEDecoder`).  The Edit-class arithmetic around them is pinned
against the UNMODIFIED reference class run on the same stand-ins (tests/golden/make_golden_sd.py).
The CLIP text encoder and SAM are out of scope: prompt embeddings and masks are supplied by the caller.
"""
import os

import torch

from . import ops
from .edit import random_basis
from .masks import load_mask
from .t2i import EditDeepFloydIF

VAE_SCALE = 0.18215                       # src/modules/edit.py:769
SD_MODES = ["null+(for-null)+(edit-null)", "null+(for-null)", "null+(edit-null)", "(for-edit)"]


class EditStableDiffusion(EditDeepFloydIF):
    """Drop-in for the hot-path methods of the reference class of the same name.  `unet`: a
    `TextB200UNet` / `CondB200UNet` over latents [4, r, r]; `vae`: a `B200VAEDecoder` whose input shape is
    that latent shape."""
    default_t_max = 999                              # get_stable_diffusion_scheduler, src/utils/utils.py:150
    default_noise_schedule = "scaled_linear"

    def __init__(self, args, unet, vae, for_prompt_emb, edit_prompt_emb, null_prompt_emb, dataset=None):
        if not hasattr(args, "image_size"):
            args.image_size = vae.out_shape[-1]
        super().__init__(args, unet, for_prompt_emb, edit_prompt_emb, null_prompt_emb, dataset=dataset)
        assert tuple(vae.in_shape) == tuple(unet.base.in_shape) == tuple(unet.base.out_shape), \
            (vae.in_shape, unet.base.in_shape, unet.base.out_shape)
        self.vae = vae
        self.c_in = unet.base.in_shape[0]
        self.latent_shape = tuple(unet.base.in_shape)             # (4, 64, 64) for SD 1.x (:938)
        # :491 (no model-size suffix in this class's folder name)
        rf = getattr(args, "result_folder", "./runs/")
        self.result_folder = os.path.join(rf, f"for_prompt_{getattr(args, 'for_prompt', '')}_cfg{self.guidance_scale}_seed{self.seed}")
        os.makedirs(self.result_folder, exist_ok=True)
        self.zT = None            # optional injected z_T (parity runs; the reference draws randn, :938)

    # ------------------------------------------------------------------ guidance / x0_hat
    def _classifer_free_guidance(self, latents, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mode,
                                 do_classifier_free_guidance):
        """src/modules/edit.py:636-674."""
        assert mode in SD_MODES
        return super()._classifer_free_guidance(latents, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mode,
                                                do_classifier_free_guidance)

    def _pmp_coefficients(self, t):
        """z0_hat / 0.18215 = c1 z_t + c2 eps   (:763-769)."""
        at = self.scheduler.alpha_at(float(t))
        c1 = 1.0 / (at ** 0.5 * VAE_SCALE)
        return c1, -((1.0 - at) ** 0.5) * c1

    def decode(self, z0_hat):
        """`self.vae.decode(1 / 0.18215 * z0_hat).sample` (:769-770)."""
        return self.vae.decode(ops.combine3(z0_hat.contiguous(), 1.0 / VAE_SCALE))

    @torch.no_grad()
    def get_x0(self, zt, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None,
               mode="null+(for-null)+(edit-null)", flatten=False):
        """src/modules/edit.py:757-781: x0_hat in PIXEL space, [B, l_o] under a mask."""
        assert mode in SD_MODES
        do_cfg = self.guidance_scale > 1.0
        zt = zt.to(self.device, torch.float32).contiguous()
        noise_pred = self._classifer_free_guidance(zt, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mode, do_cfg)
        c1, c2 = self._pmp_coefficients(t)
        x0_hat = self.vae.decode(ops.combine3(zt, c1, noise_pred, c2))
        if mask is not None:
            return ops.gather_rows(x0_hat.reshape(x0_hat.shape[0], -1), ops.mask_indices(mask.to(self.device)))
        if flatten:
            return x0_hat.reshape(-1, x0_hat[0].numel())
        return x0_hat

    @torch.no_grad()
    def DDIMforwardsteps(self, zt, t_start_idx, t_end_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb,
                         mode="null+(for-null)", **kwargs):
        """src/modules/edit.py:677-755: guided DDIM (eta = 0) over latents; returns (latents, t, t_idx) at
        t_end_idx, else (latents / 0.18215, uint8 images [B, H, W, 3]) after the VAE decode."""
        assert mode in ["null+(for-null)+(edit-null)", "null+(for-null)", "null+(edit-null)"]
        do_cfg = self.guidance_scale > 1.0
        self.scheduler.set_timesteps(self.for_steps, device=self.device)
        ts = self.scheduler._ts_host
        latents = zt.to(self.device, torch.float32).contiguous()
        for t_idx, t in enumerate(ts):
            if t_idx < t_start_idx:
                continue
            elif t_idx == t_end_idx and t_idx != t_start_idx:
                return latents, self.scheduler.timesteps[t_idx], t_idx
            noise_pred = self._classifer_free_guidance(latents, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mode, do_cfg)
            latents = self.scheduler.step(noise_pred, t, latents, eta=0, t_idx=t_idx).prev_sample
        latents = ops.combine3(latents, 1.0 / VAE_SCALE)                       # :747
        x0 = self.vae.decode(latents)
        self.last_images.append(x0)
        img = (x0 / 2 + 0.5).clamp(0, 1)
        return latents, (img * 255).to(torch.uint8).permute(0, 2, 3, 1)

    # ------------------------------------------------------------------ Jacobian products through both networks
    def _jvp_decoded(self, z_row, t, V, slots):
        """(x0_hat [1, d_x], dX [k, d_x]): tangents of decode(z0_hat(z) / 0.18215) along the rows of V."""
        k = V.shape[0]
        zin = torch.cat([z_row.reshape(1, -1), V], 0).contiguous()                      # [1+k, d_z]
        xin = zin.reshape(1 + k, *self.latent_shape)
        outs = []
        for slot, wi, emb in slots:
            plan = self.unet.base.plan(1, k, k, slot=slot)
            self.unet.apply_condition(plan, emb)
            outs.append((plan.forward(xin, float(t)).reshape(1 + k, -1), wi))
        eps_all = self._combine(outs)                                                   # guided eps and its tangents
        c1, c2 = self._pmp_coefficients(t)
        z0_all = ops.combine3(zin, c1, eps_all, c2)                                     # linear: same map on every row
        X = self.vae.plan(1, k, k).forward(z0_all.reshape(1 + k, *self.latent_shape).contiguous(), 0.0)
        X = X.reshape(1 + k, -1)
        return X[:1], X[1:]

    def _vjp_decoded(self, G, t, slots):
        """J^T G for k pixel-space cotangents G [k, d_x] at the primal point of the last `_jvp_decoded`."""
        k = G.shape[0]
        gz0 = self.vae.plan(1, k, k).vjp(G.reshape(k, *self.vae.out_shape).contiguous()).reshape(k, -1)
        c1, c2 = self._pmp_coefficients(t)
        g_eps = ops.combine3(gz0, c2)
        return ops.combine3(gz0, c1, self._vjp_cfg(g_eps, slots), 1.0)

    def local_encoder_decoder_pullback_zt(self, zt, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb,
                                          op=None, block_idx=None, pca_rank=50, chunk_size=25, min_iter=10,
                                          max_iter=100, convergence_threshold=1e-3, mask=None,
                                          mode="null+(for-null)+(edit-null)", v0=None):
        """src/modules/edit.py:830-915: subspace iteration on J^T J with J through the U-Net, the PMP and
        the VAE decoder.  `mask`: bool [3, H, W] over the DECODED image.  Returns (u [l_o, k],
        s [k] = sqrt(svdvals), vT [k, d_z]) like the reference."""
        assert mode in SD_MODES
        slots = self._cfg_slots(mode, (for_prompt_emb, edit_prompt_emb, null_prompt_emb))
        return self._power_method(zt, t, slots, pca_rank, min_iter, max_iter, convergence_threshold, mask, v0)

    def _power_method(self, zt, t, slots, pca_rank, min_iter, max_iter, convergence_threshold, mask, v0):
        """The subspace iteration shared by the latent-space classes: `slots` = the conditionings whose noise
        predictions combine linearly into the one the posterior mean uses."""
        k = int(pca_rank)
        d = zt[0].numel()
        z = zt.to(self.device, torch.float32).contiguous().reshape(1, -1)
        idx = None if mask is None else ops.mask_indices(mask.to(self.device))
        if v0 is None and self.v0 is not None:
            v0 = self.v0.get(k) if isinstance(self.v0, dict) else self.v0
        V = (random_basis(d, k, self.device) if v0 is None else v0.to(self.device, torch.float32).reshape(k, d)).contiguous()
        u = s = None
        for i in range(max_iter):
            _, dX = self._jvp_decoded(z, t, V, slots)                                   # :872-880 (jacfwd)
            if idx is None:
                u, G = dX, dX
            else:
                u = ops.gather_rows(dX, idx)                                            # x0_hat[:, mask]
                G = ops.scatter_rows(u, idx, dX.shape[1])                               # its transpose
            w = self._vjp_decoded(G, t, slots)                                          # :882-894 (jacobian)
            V_new, s = ops.orthonormalise(w, v_prev=V if self.align_sign else None)     # :896
            need_check = i > min_iter
            if self.verbose or need_check:
                convergence = torch.dist(V, V_new).item()
                if self.verbose:
                    print(f'power method : {i}-th step convergence : ', convergence)
            done = need_check and torch.allclose(V, V_new, atol=convergence_threshold)
            V = V_new
            if done:
                break
        return u.T, s, V

    @torch.no_grad()
    def get_delta_zt_via_grad(self, zt, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None,
                              mode="null+(for-null)+(edit-null)"):
        """src/modules/edit.py:784-828: text-supervised latent direction
        v = normalise(J_mode^T (x0_hat_after - x0_hat)) with both x0_hat in pixel space -- one transposed pass
        through the decoder and the U-Net.  (With mask = None the reference reshapes the pixel-space
        difference with the LATENT extents, :804-811, and fails; here all pixels are used.)"""
        z = zt.to(self.device, torch.float32).contiguous()
        x0_hat = self.get_x0(z, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None, mode="null+(for-null)")
        x0_hat_after = self.get_x0(z, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None, mode=mode)
        dx = x0_hat[0].numel()
        delta = ops.combine3(x0_hat_after.reshape(1, dx), 1.0, x0_hat.reshape(1, dx), -1.0)
        if mask is not None:
            idx = ops.mask_indices(mask.to(self.device))
            delta = ops.scatter_rows(ops.gather_rows(delta, idx), idx, dx)
        slots = self._cfg_slots(mode, (for_prompt_emb, edit_prompt_emb, null_prompt_emb))
        # primal activations of both networks at z_t under every conditioning of `mode` (dummy tangent row)
        self._jvp_decoded(z.reshape(1, -1), t, torch.zeros(1, z[0].numel(), device=self.device), slots)
        v_ = self._vjp_decoded(delta, t, slots)
        return ops.nullspace_project(v_, None, project=False)                          # v_ / ||v_||   (:816)

    @torch.no_grad()
    def x_space_guidance_direct(self, zt, t_idx, vk, single_edit_step):
        """src/modules/edit.py:1177-1185."""
        return ops.axpy(zt.contiguous(), vk.expand_as(zt).contiguous(), self.x_space_guidance_scale * single_edit_step)

    # ------------------------------------------------------------------ drivers
    def _start_zt(self, mask_index):
        """z_T, the mask and z_t at the edit timestep (:933-966 without SAM; masks come from `mask/mask.pt`
        under the result folder, a bool [H, W] over the DECODED image repeated over its 3 channels, :959)."""
        self.scheduler.set_timesteps(self.for_steps, device=self.device)
        zT = self.zT if self.zT is not None else torch.randn(1, *self.latent_shape, dtype=torch.float32, device=self.device)
        zT = zT.to(self.device)
        mask = load_mask(self.result_folder, mask_index).to(self.device)
        kw = dict(for_prompt_emb=self.for_prompt_emb, edit_prompt_emb=self.edit_prompt_emb,
                  null_prompt_emb=self.null_prompt_emb, mode="null+(for-null)")
        zt, t, t_idx = self.DDIMforwardsteps(zT, t_start_idx=0, t_end_idx=self.edit_t_idx, **kw)
        assert t_idx == self.edit_t_idx
        return mask, zt, t, t_idx, kw

    def _basis_paths(self, save_dir, pca_rank_null):
        os.makedirs(save_dir, exist_ok=True)
        return dict(u_m=os.path.join(save_dir, 'u-modify.pt'), v_m=os.path.join(save_dir, 'vT-modify.pt'),
                    u_n=os.path.join(save_dir, f'u-null-null_space_rank_{pca_rank_null}.pt'),
                    v_n=os.path.join(save_dir, f'vT-null-null_space_rank_{pca_rank_null}.pt'))

    def _null_basis(self, zt, t, t_idx, mask, op, block_idx, pca_rank_null, paths):
        """:996-1003 / :1120-1127: the basis of the complement region, `~mask`."""
        u_null, _, vT_null = self.local_encoder_decoder_pullback_zt(
            zt, t, t_idx, self.for_prompt_emb, self.edit_prompt_emb, self.null_prompt_emb, op=op, block_idx=block_idx,
            pca_rank=pca_rank_null, chunk_size=5, min_iter=10, max_iter=50, convergence_threshold=1e-3, mask=~mask,
            mode="null+(for-null)")
        torch.save(u_null, paths["u_n"])
        torch.save(vT_null, paths["v_n"])
        return vT_null

    def _project_and_edit(self, zt, vT_modify, vT_null, null_space_projection, pca_rank_null, vis_num, vis_num_pc, name, kw):
        """:1005-1043: normalise / project the directions, build the edited latents, denoise and decode."""
        if not null_space_projection:
            vT = ops.nullspace_project(vT_modify.contiguous(), None, project=False)
        else:
            vT = ops.nullspace_project(vT_modify.contiguous(), vT_null[:pca_rank_null, :].contiguous(), project=True)
        self.last_images = []
        imgs = None
        for pc_idx in range(min(vis_num_pc, vT.shape[0])):
            self.EXP_NAME = name(pc_idx)
            batch = self._edit_batch(zt, vT[pc_idx, :], vis_num)
            _, imgs = self.DDIMforwardsteps(batch, t_start_idx=self.edit_t_idx, t_end_idx=-1, **kw)
        return dict(vT=vT, vT_modify=vT_modify, vT_null=vT_null, zt=zt, images=imgs)

    @torch.no_grad()
    def run_edit_null_space_projection_zt(self, op, block_idx, vis_num, mask_index=0, vis_num_pc=1, vis_vT=False,
                                          pca_rank=50, edit_prompt=None, null_space_projection=False, pca_rank_null=50,
                                          non_semantic=False):
        """src/modules/edit.py:918-1043 (unsupervised latent-space edit)."""
        if self.sampling_mode:
            return None
        mask, zt, t, t_idx, kw = self._start_zt(mask_index)
        paths = self._basis_paths(os.path.join(self.result_folder, "basis",
                                               f'local_basis-{self.edit_t}T-pca-rank-{pca_rank}-select-mask{mask_index}'), pca_rank_null)
        if all(os.path.exists(p) for p in paths.values()):
            vT_modify = torch.load(paths["v_m"], map_location=self.device).type(torch.float32)
            vT_null = torch.load(paths["v_n"], map_location=self.device).type(torch.float32)
        else:
            u_modify, _, vT_modify = self.local_encoder_decoder_pullback_zt(
                zt, t, t_idx, self.for_prompt_emb, self.edit_prompt_emb, self.null_prompt_emb, op=op, block_idx=block_idx,
                pca_rank=pca_rank, chunk_size=5, min_iter=10, max_iter=50, convergence_threshold=1e-3, mask=mask,
                mode="null+(for-null)")
            torch.save(u_modify, paths["u_m"])
            torch.save(vT_modify, paths["v_m"])
            vT_null = self._null_basis(zt, t, t_idx, mask, op, block_idx, pca_rank_null, paths) if null_space_projection else None
        name = lambda pc: (f'Edit_zt-edit_{self.edit_t}T-pc_{pc}_select_mask{mask_index}_null_space_projection_'
                           f'{null_space_projection}_null_space_rank_{pca_rank_null}')
        return self._project_and_edit(zt, vT_modify, vT_null, null_space_projection, pca_rank_null, vis_num, vis_num_pc, name, kw)

    @torch.no_grad()
    def run_edit_null_space_projection_zt_semantic(self, op, block_idx, vis_num, mask_index=0, vis_num_pc=1, vis_vT=False,
                                                   pca_rank=50, edit_prompt=None, null_space_projection=False,
                                                   pca_rank_null=50):
        """src/modules/edit.py:1045-1175 (text-supervised latent-space edit: the direction is
        `get_delta_zt_via_grad`, projected off the null basis of `~mask`; the SEGA branch, :1164-1175, is a
        plain guided denoising with the three-term mode)."""
        if self.sampling_mode:
            return None
        mask, zt, t, t_idx, kw = self._start_zt(mask_index)
        if getattr(self, "use_sega", False):
            self.EXP_NAME = f'sega-edit_prompt-{self.edit_prompt}'
            _, imgs = self.DDIMforwardsteps(zt, t_start_idx=self.edit_t_idx, t_end_idx=-1,
                                            **dict(kw, mode="null+(for-null)+(edit-null)"))
            return dict(zt=zt, images=imgs)
        paths = self._basis_paths(os.path.join(
            self.result_folder, "basis", f'local_basis-{self.edit_t}T-"{self.edit_prompt}"-pca-rank-{pca_rank}-select-mask{mask_index}'),
            pca_rank_null)
        if all(os.path.exists(p) for p in paths.values()):
            vT_modify = torch.load(paths["v_m"], map_location=self.device).type(torch.float32)
            vT_null = torch.load(paths["v_n"], map_location=self.device).type(torch.float32)
        else:
            vT_modify = self.get_delta_zt_via_grad(zt, t, t_idx, self.for_prompt_emb, self.edit_prompt_emb, self.null_prompt_emb,
                                                   mask=mask, mode=self.tilda_v_score_type)
            torch.save(vT_modify, paths["v_m"])
            vT_null = self._null_basis(zt, t, t_idx, mask, op, block_idx, pca_rank_null, paths) if null_space_projection else None
        name = lambda pc: (f'Edit_zt-edit_{self.edit_t}T-{op}-block_{block_idx}-pc_{pc:0=3d}_pos-edit_prompt-{self.edit_prompt}'
                           f'_select_mask{mask_index}_null_space_projection_{null_space_projection}_null_space_rank_{pca_rank_null}'
                           f'_{self.tilda_v_score_type}')
        return self._project_and_edit(zt, vT_modify, vT_null, null_space_projection, pca_rank_null, vis_num, vis_num_pc, name, kw)


# ================================================================================================
# Latent consistency twin
# ================================================================================================
def guidance_scale_embedding(w, embedding_dim=256):
    """`pipe.get_guidance_scale_embedding(w, embedding_dim)` of diffusers' LatentConsistencyModelPipeline (called at
    src/modules/edit.py:118, 161, 218): sinusoidal embedding of 1000 w, [sin | cos], divisor half - 1."""
    import math
    w = torch.as_tensor(w, dtype=torch.float32).reshape(-1) * 1000.0
    half = embedding_dim // 2
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))
    emb = w[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


class LCMB200UNet(object):
    """Stand-in for the distilled LCM U-Net `unet(z, t, timestep_cond=w_embedding, encoder_hidden_states=...)`
    (src/modules/edit.py:182-188): this package's text-conditioned latent U-Net (cross-attention to the prompt tokens)
    whose timestep embedding additionally receives the guidance-scale embedding through a fixed seeded projection
    (`loco_plan_set_condition`; diffusers adds `cond_proj(timestep_cond)` to the sinusoid before the time MLP)."""

    def __init__(self, base, w_dim=256, seed=0):
        from .t2i import TextB200UNet, cond_projection
        self.text = TextB200UNet(base)
        self.base, self.device, self.arch = base, base.device, base.arch
        self.P = cond_projection(w_dim, 4 * base.arch["ch"], seed)
        self.cond = None

    def set_guidance(self, w_embedding):
        """w_embedding [1, w_dim] -> the conditioning vector added to the timestep embedding of every pass."""
        self.cond = (w_embedding.detach().to("cpu", torch.float32).reshape(-1) @ self.P).contiguous().to(self.device)

    def apply_condition(self, plan, prompt_emb):
        self.text.apply_condition(plan, prompt_emb)
        plan.set_condition(self.cond)

    def eps(self, x, t, prompt_emb):
        x = x.contiguous()
        plan = self.base.plan(x.shape[0])
        self.apply_condition(plan, prompt_emb)
        return plan.forward(x, float(t))

    def __call__(self, x, t, timestep_cond=None, encoder_hidden_states=None, return_dict=False):
        if timestep_cond is not None:
            self.set_guidance(timestep_cond[:1])
        return (self.eps(x, t, encoder_hidden_states[:1]),)


class EditLatentConsistency(EditStableDiffusion):
    """Drop-in for the hot-path methods of the reference class of the same name (src/modules/edit.py:42-480): the
    latent-space pull-back with the consistency model's boundary-condition x0 estimate instead of the DDIM posterior
    mean, ONE conditioning (guidance is distilled into the network through the w embedding, no CFG batch), prompts
    passed as strings and encoded by `encode_prompt` (the CLIP text encoder is out of scope: the stand-in is
    `synthetic_prompt_embedding`).  The scheduler is `scheduler.LCMScheduler` (parity unpinned, see there)."""

    def __init__(self, args, unet, vae, encode_prompt=None, dataset=None):
        from .scheduler import LCMScheduler
        from .t2i import synthetic_prompt_embedding
        self.encode_prompt = encode_prompt or (lambda p: synthetic_prompt_embedding(p, 77, unet.arch["ctx_dim"]))
        self.for_prompt = getattr(args, "for_prompt", "")
        self.edit_prompt = getattr(args, "edit_prompt", "")
        embs = [self.encode_prompt(p) for p in (self.for_prompt, self.edit_prompt, "")]
        if not hasattr(args, "edit_t"):
            args.edit_t = 0.0
        super().__init__(args, unet, vae, *embs, dataset=dataset)
        self.num_inference_steps = getattr(args, "num_inference_steps", 4)
        self.scheduler = LCMScheduler(self.device)
        self.scheduler.set_timesteps(self.num_inference_steps, device=self.device)
        self.edit_t_idx = int(getattr(args, "edit_t_idx", 0))                                      # :97
        self.use_sega = getattr(args, "use_sega", False)
        w = torch.tensor(self.guidance_scale - 1).repeat(1)                                         # :117
        self.unet.set_guidance(guidance_scale_embedding(w, getattr(args, "time_cond_proj_dim", 256)))
        self.step_noise = None        # optional injected noise list (parity runs; the stock scheduler draws randn)

    def _emb(self, prompt):
        return prompt if torch.is_tensor(prompt) else self.encode_prompt(prompt)

    def _pmp_coefficients(self, t):
        """denoised / scaling_factor = c1 z_t + c2 eps   (:237-238)."""
        c1, c2 = self.scheduler.denoised_coefficients(t)
        return c1 / VAE_SCALE, c2 / VAE_SCALE

    @torch.no_grad()
    def get_x0(self, zt, prompt, t, t_idx, mask=None, flatten=False):
        """src/modules/edit.py:206-248."""
        zt = zt.to(self.device, torch.float32).contiguous()
        model_pred = self.unet.eps(zt, t, self._emb(prompt))
        c1, c2 = self._pmp_coefficients(t)
        x0_hat = self.vae.decode(ops.combine3(zt, c1, model_pred, c2))
        if mask is not None:
            return ops.gather_rows(x0_hat.reshape(x0_hat.shape[0], -1), ops.mask_indices(mask.to(self.device)))
        if flatten:
            return x0_hat.reshape(-1, x0_hat[0].numel())
        return x0_hat

    @torch.no_grad()
    def LCMforwardsteps(self, zt, prompt, t_start_idx=0, t_end_idx=-1):
        """src/modules/edit.py:148-203: multistep consistency sampling; returns (latents, t, t_idx) at t_end_idx, else
        (latents, uint8 images) decoded from the last `denoised`."""
        self.scheduler.set_timesteps(self.num_inference_steps, device=self.device)
        emb = self._emb(prompt)
        latents = zt.to(self.device, torch.float32).contiguous()
        denoised = latents
        for t_idx, t in enumerate(self.scheduler._ts_host):
            if t_idx < t_start_idx:
                continue
            elif t_idx == t_end_idx and t_idx != t_start_idx:
                return latents, self.scheduler.timesteps[t_idx], t_idx
            model_pred = self.unet.eps(latents, t, emb)
            noise = None if self.step_noise is None else self.step_noise[t_idx].to(self.device)[:latents.shape[0]]
            latents, denoised = self.scheduler.step(model_pred, t, latents, t_idx=t_idx, noise=noise)
        image = self.vae.decode(ops.combine3(denoised, 1.0 / VAE_SCALE))
        self.last_images.append(image)
        x0 = (image / 2 + 0.5).clamp(0, 1)
        return latents, (x0 * 255).to(torch.uint8).permute(0, 2, 3, 1)

    def run_LCMforward(self, zT, prompt, num_samples=1):
        """src/modules/edit.py:103-144."""
        return self.LCMforwardsteps(zT, prompt, t_start_idx=0, t_end_idx=-1)

    def local_encoder_decoder_pullback_zt(self, zt, t, t_idx, for_prompt, op=None, block_idx=None, pca_rank=50, chunk_size=25,
                                          min_iter=10, max_iter=100, convergence_threshold=1e-3, mask=None, v0=None):
        """src/modules/edit.py:283-369."""
        return self._power_method(zt, t, [(0, 1.0, self._emb(for_prompt))], pca_rank, min_iter, max_iter,
                                  convergence_threshold, mask, v0)

    @torch.no_grad()
    def get_delta_zt_via_grad(self, zt, t, t_idx, for_prompt, edit_prompt, mask=None):
        """src/modules/edit.py:251-280: v = normalise(J_edit^T (x0_hat(edit) - x0_hat(for))), the Jacobian taken under
        the EDIT prompt (:271)."""
        z = zt.to(self.device, torch.float32).contiguous()
        x0_hat = self.get_x0(z, for_prompt, t, t_idx)
        x0_hat_after = self.get_x0(z, edit_prompt, t, t_idx)
        dx = x0_hat[0].numel()
        delta = ops.combine3(x0_hat_after.reshape(1, dx), 1.0, x0_hat.reshape(1, dx), -1.0)
        if mask is not None:
            idx = ops.mask_indices(mask.to(self.device))
            delta = ops.scatter_rows(ops.gather_rows(delta, idx), idx, dx)
        slots = [(0, 1.0, self._emb(edit_prompt))]
        self._jvp_decoded(z.reshape(1, -1), t, torch.zeros(1, z[0].numel(), device=self.device), slots)
        v_ = self._vjp_decoded(delta, t, slots)
        return ops.nullspace_project(v_, None, project=False)

    @torch.no_grad()
    def run_edit_null_space_projection_zt(self, op, block_idx, vis_num, mask_index=0, vis_num_pc=1, vis_vT=False, pca_rank=50,
                                          edit_prompt=None, null_space_projection=False, pca_rank_null=50, non_semantic=False):
        """src/modules/edit.py:374-473: text-supervised (or, `non_semantic`, unsupervised) direction at step
        `edit_t_idx` of the consistency sampler, projected off the null basis of ~mask; masks from `mask/mask.pt`."""
        if edit_prompt is not None:
            self.edit_prompt = edit_prompt
        if self.sampling_mode:
            return None
        self.scheduler.set_timesteps(self.num_inference_steps, device=self.device)
        zT = self.zT if self.zT is not None else torch.randn(1, *self.latent_shape, dtype=torch.float32, device=self.device)
        mask = load_mask(self.result_folder, mask_index).to(self.device)
        if self.edit_t_idx > 0:
            zt, t, t_idx = self.LCMforwardsteps(zT.to(self.device), self.for_prompt, t_start_idx=0, t_end_idx=self.edit_t_idx)
        else:
            zt, t_idx = zT.to(self.device), 0
        t = int(self.scheduler._ts_host[t_idx])
        if self.use_sega:
            self.EXP_NAME = f'sega_{self.edit_t_idx}T-{op}-block_{block_idx}_pos-edit_prompt-{self.edit_prompt}'
            _, imgs = self.LCMforwardsteps(zt, self.edit_prompt, t_start_idx=self.edit_t_idx, t_end_idx=-1)
            return dict(zt=zt, images=imgs)
        pb = dict(op=op, block_idx=block_idx, chunk_size=5, min_iter=10, max_iter=50, convergence_threshold=1e-3)
        if non_semantic:
            _, _, vT_modify = self.local_encoder_decoder_pullback_zt(zt, t, t_idx, self.for_prompt, pca_rank=pca_rank, mask=mask, **pb)
        else:
            vT_modify = self.get_delta_zt_via_grad(zt.clone(), t, t_idx, self.for_prompt, self.edit_prompt, mask=mask)
        vT_null = None
        if null_space_projection:
            _, _, vT_null = self.local_encoder_decoder_pullback_zt(zt, t, t_idx, self.for_prompt, pca_rank=pca_rank_null,
                                                                   mask=~mask, **pb)
            vT = ops.nullspace_project(vT_modify.contiguous(), vT_null[:pca_rank_null, :].contiguous(), project=True)
        else:
            vT = ops.nullspace_project(vT_modify.contiguous(), None, project=False)
        self.EXP_NAME = (f'Edit_zt-edit_{self.edit_t_idx}T-{op}-block_{block_idx}_pos-edit_prompt-{self.edit_prompt}_select_mask'
                         f'{mask_index}_null_space_projection_{null_space_projection}_null_space_rank_{pca_rank_null}')
        batch = self._edit_batch(zt, vT[0, :], vis_num)      # the reference edits along ALL rows at once (:445); row 0 here
        self.last_images = []
        _, imgs = self.LCMforwardsteps(batch, self.for_prompt, t_start_idx=self.edit_t_idx, t_end_idx=-1)
        return dict(vT=vT, vT_modify=vT_modify, vT_null=vT_null, zt=zt, images=imgs)
