"""T-LOCO Edit: the text-conditioned twins of the editing-direction hot path
(reference: `EditDeepFloydIF`, src/modules/edit.py:1198-2031; the same logic sits in
`EditStableDiffusion`, :483-1196, around a VAE that is out of scope here).

What is built (SURVEY section 8 rows a11, f1 (CFG folding), f3, f4):
  * `_classifer_free_guidance` with the reference's eight modes (:1286-1373);
  * `get_x0` (:1566-1587), `DDPMforwardsteps` (:1410-1483), `mask_diffedit` (:1395-1407);
  * `local_encoder_decoder_pullback_xt` under classifier-free guidance (:1589-1676): the guided noise
    prediction is a LINEAR combination  e = sum_i w_i eps(x, t, c_i)  of the same network under 2-3
    conditionings, so J = sum_i w_i J_i: one fused primal + k-tangent pass and one k-cotangent pass of
    the CUDA executor per conditioning, combined by `loco_combine3`;
  * the text-supervised direction `get_delta_xt_via_grad` (:1680-1720: ONE VJP with the data-dependent
    cotangent x0_hat_after - x0_hat) and `get_v_modify` (:1723-1741);
  * the two drivers `run_edit_null_space_projection_xt` (:1745-1871, non-semantic) and
    `run_edit_null_space_projection_xt_semantic` (:1874-2018, ablation "null-space-proj"), with the
    reference's basis file names, and an n-direction `group_edit_null_space_projection` (the reference
    composes two directions, :2171-2212; BASELINE config 5 asks for three).

What is NOT the reference's: the network.  DeepFloyd-IF stage I / Stable Diffusion are diffusers models
that are not under /root/reference and cannot be downloaded here ("parity unpinned at the
network level").  `CondB200UNet` is the stand-in the survey prescribes: this package's DDPM U-Net with
a conditioning embedding added to the timestep embedding, c = mean_tokens(prompt_emb) @ P (P a fixed
seeded projection).  It has no cross-attention; every Edit-class formula around it is pinned against
the UNMODIFIED reference class run on the same stand-in (tests/golden/make_golden_t2i.py).  Prompt
embeddings are tensors [1, n_tok, D] supplied by the caller (the T5 / CLIP text encoders are out of
scope); `synthetic_prompt_embedding` makes seeded ones from a prompt string.  Super-resolution
(stages II / III, :1376-1393) is skipped.
"""
import os
import types
import zlib

import torch

from . import ops
from .edit import _pb_workspace, random_basis
from .masks import diffedit_mask, load_mask
from .scheduler import YHCustomScheduler

CFG_MODES = ["null+(for-null)+(edit-null)", "null+(for-null)", "null+(edit-null)", "(for-edit)", "(for-null)",
             "(edit-null)", "edit-proj[for](edit)", "null+for+edit-proj[for](edit)"]


def synthetic_prompt_embedding(prompt, n_tok=77, dim=64, seed=0):
    """Seeded stand-in for `stage_1.encode_prompt` (src/modules/edit.py:1274-1284): [1, n_tok, dim]."""
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(prompt.encode())) % (2 ** 63))
    return torch.randn(1, n_tok, dim, generator=g)


def cond_projection(dim, temb_ch, seed=0):
    """The fixed map pooled prompt embedding [dim] -> conditioning embedding [temb_ch]."""
    g = torch.Generator().manual_seed(7919 + seed)
    return torch.randn(dim, temb_ch, generator=g) / dim ** 0.5


def linear_cfg_weights(mode, g, g_edit, do_cfg=True):
    """Weights (w_for, w_edit, w_null) with which the guided prediction of `mode` combines the three
    conditional predictions (src/modules/edit.py:1326-1357); None for the two projection modes, which
    are not linear in the predictions."""
    if not do_cfg:
        return (1.0, 0.0, 0.0)                                     # :1318-1320
    return {
        "null+(for-null)+(edit-null)": (g, g_edit, 1.0 - g - g_edit),
        "null+(for-null)": (g, 0.0, 1.0 - g),
        "null+(edit-null)": (0.0, g, 1.0 - g),
        "(for-edit)": (g, -g, 0.0),
        "(for-null)": (g, 0.0, -g),
        "(edit-null)": (0.0, g, -g),
    }.get(mode)


class CondB200UNet(object):
    """eps(x, t, c) on the CUDA executor: `base` (a B200UNet) with the conditioning embedding
    c = mean_tokens(encoder_hidden_states) @ P added to the timestep embedding (loco_plan_set_condition).
    Callable like the reference's `self.unet(x, t, encoder_hidden_states=...)` -> `.sample`."""

    def __init__(self, base, dim, seed=0):
        self.base = base
        self.device = base.device
        self.arch = base.arch
        self.P = cond_projection(dim, 4 * base.arch["ch"], seed)
        self._cache = {}

    def cond_vector(self, prompt_emb):
        """Device vector [4*ch] of one prompt embedding [1, n_tok, D] (host arithmetic, cached)."""
        key = (prompt_emb.data_ptr(), tuple(prompt_emb.shape))
        if key not in self._cache:
            c = prompt_emb.detach().to("cpu", torch.float32).reshape(-1, prompt_emb.shape[-1]).mean(0) @ self.P
            self._cache[key] = (prompt_emb, c.contiguous().to(self.device))
        return self._cache[key][1]

    def apply_condition(self, plan, prompt_emb):
        """Make `plan` evaluate eps(., t, prompt_emb) from now on."""
        plan.set_condition(self.cond_vector(prompt_emb))

    def eps(self, x, t, prompt_emb):
        """Noise prediction of all rows of x under ONE prompt."""
        x = x.contiguous()
        plan = self.base.plan(x.shape[0])
        self.apply_condition(plan, prompt_emb)
        return plan.forward(x, float(t))

    def __call__(self, x, t, encoder_hidden_states):
        """Reference call pattern (:1319-1322): row b of x under encoder_hidden_states[b]; rows that
        share an embedding (CFG batches are [latents] * 2 or * 3) go through the network together."""
        ehs = encoder_hidden_states
        out = torch.empty_like(x)
        done = [False] * x.shape[0]
        for b in range(x.shape[0]):
            if done[b]:
                continue
            rows = [r for r in range(b, x.shape[0]) if not done[r] and torch.equal(ehs[r], ehs[b])]
            out[rows] = self.eps(x[rows], t, ehs[b:b + 1])
            for r in rows:
                done[r] = True
        return types.SimpleNamespace(sample=out)


class TextB200UNet(CondB200UNet):
    """eps(x, t, prompt) with REAL cross-attention: `base` is a B200UNet built with ctx_dim > 0, every
    AttnBlock of which attends to the prompt tokens (loco_plan_set_context; fused tcgen05
    cross-attention with JVP / VJP, csrc/attention_tc.cu).  `encoder_hidden_states` [1, n_tok <= 128,
    ctx_dim] is used as is, like diffusers' UNet2DConditionModel does (src/modules/edit.py:655-658,
    1319-1322).  Same call protocol as CondB200UNet, so the Edit classes take either."""

    def __init__(self, base):
        assert base.arch.get("ctx_dim", 0) > 0, "TextB200UNet needs a U-Net with cross-attention layers (ctx_dim > 0)"
        self.base = base
        self.device = base.device
        self.arch = base.arch
        self._cache = {}

    def context(self, prompt_emb):
        key = (prompt_emb.data_ptr(), tuple(prompt_emb.shape))
        if key not in self._cache:
            c = prompt_emb.detach().reshape(-1, prompt_emb.shape[-1]).to(device=self.device, dtype=torch.float32)
            self._cache[key] = (prompt_emb, c.contiguous())
        return self._cache[key][1]

    def apply_condition(self, plan, prompt_emb):
        plan.set_context(self.context(prompt_emb))


class EditDeepFloydIF(object):
    """Drop-in for the hot-path methods of the reference class of the same name."""
    default_t_max = 990                              # get_deepfloyd_if_scheduler, src/utils/utils.py:162
    default_noise_schedule = "squaredcos_cap_v2"

    def __init__(self, args, unet, for_prompt_emb, edit_prompt_emb, null_prompt_emb, dataset=None):
        self.seed = getattr(args, "seed", 0)
        self.device = torch.device(args.device)
        self.dtype = getattr(args, "dtype", torch.float32)
        self.unet = unet                                    # CondB200UNet or TextB200UNet
        # get_deepfloyd_if_scheduler (src/utils/utils.py:159-170): the stage-I scheduler's own alpha_bar
        # table with the custom timestep grid on t_max = 990
        sargs = types.SimpleNamespace(device=self.device, dtype=torch.float32, t_max=getattr(args, "t_max", self.default_t_max),
                                      noise_schedule=getattr(args, "noise_schedule", self.default_noise_schedule))
        self.scheduler = YHCustomScheduler(sargs, device=self.device)
        self.for_steps = args.for_steps
        self.use_yh_custom_scheduler = True
        self.c_in = 3
        self.image_size = getattr(args, "image_size", unet.arch["resolution"])
        self.dataset_name = getattr(args, "dataset_name", "Random")
        self.for_prompt_emb, self.edit_prompt_emb, self.null_prompt_emb = for_prompt_emb, edit_prompt_emb, null_prompt_emb
        self.edit_prompt = getattr(args, "edit_prompt", "")
        self.guidance_scale = args.guidance_scale
        self.guidance_scale_edit = getattr(args, "guidance_scale_edit", 4.0)
        self.x_space_guidance_edit_step = getattr(args, "x_space_guidance_edit_step", 1)
        self.x_space_guidance_scale = getattr(args, "x_space_guidance_scale", 0)
        self.x_space_guidance_num_step = getattr(args, "x_space_guidance_num_step", 0)
        self.memory_bound = getattr(args, "memory_bound", 50)
        self.scheduler.set_timesteps(self.for_steps, device=self.device)
        self.edit_t = args.edit_t
        self.edit_t_idx = int((self.scheduler.timesteps - self.edit_t * 1000).abs().argmin())      # :1264
        self.sampling_mode = getattr(args, "sampling_mode", False)
        self.tilda_v_score_type = getattr(args, "tilda_v_score_type", "null+(for-null)+(edit-null)")
        self.ablation_method = getattr(args, "ablation_method", "null-space-proj")
        self.mask_type = getattr(args, "mask_type", "SAM")
        self.vT_path = getattr(args, "vT_path", "")
        rf = getattr(args, "result_folder", "./runs/")
        # :1206 (the model-size suffix comes from the diffusers model name; "stand-in" here)
        self.result_folder = os.path.join(rf, f"for_prompt_{getattr(args, 'for_prompt', '')}_cfg{self.guidance_scale}_seed{self.seed}_standin")
        os.makedirs(self.result_folder, exist_ok=True)
        self.verbose = getattr(args, "verbose", False)
        self.align_sign = getattr(args, "align_sign", True)
        self.v0 = None            # optional injected initial basis (dict keyed by rank, or tensor)
        self.xT = None            # optional injected x_T (parity runs; the reference draws randn, :1769)
        self.last_images = []

    # ------------------------------------------------------------------ guidance
    def _eps3(self, x, t, need, f, e, n):
        """Conditional predictions the mode needs: (eps_for, eps_edit, eps_null), None where unused."""
        return tuple(self.unet.eps(x, t, emb) if flag else None for flag, emb in zip(need, (f, e, n)))

    def _classifer_free_guidance(self, latents, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mode,
                                 do_classifier_free_guidance):
        """src/modules/edit.py:1286-1373 (the network returns eps only: no variance split)."""
        assert mode in CFG_MODES
        x = latents.to(self.device, torch.float32).contiguous()
        w = linear_cfg_weights(mode, self.guidance_scale, self.guidance_scale_edit, do_classifier_free_guidance)
        if w is not None:
            ef, ee, en = self._eps3(x, t, [wi != 0.0 for wi in w], for_prompt_emb, edit_prompt_emb, null_prompt_emb)
            terms = [(a, wi) for a, wi in zip((ef, ee, en), w) if a is not None]
            (a, wa), rest = terms[0], terms[1:]
            (b, wb) = rest[0] if len(rest) > 0 else (None, 0.0)
            (c, wc) = rest[1] if len(rest) > 1 else (None, 0.0)
            if b is None and wa == 1.0:
                return a
            return ops.combine3(a, wa, b, wb, c, wc)
        if mode == "edit-proj[for](edit)":
            # the reference reads `noise_pred_uncond` without computing it in this branch (:1358-1364) and
            # raises NameError; there is nothing to be compatible with
            raise NotImplementedError("mode 'edit-proj[for](edit)' is broken in the reference (src/modules/edit.py:1361)")
        # "null+for+edit-proj[for](edit)" (:1365-1372): projection of (edit-null) off (for-null) in noise space
        ef, ee, en = self._eps3(x, t, [True, True, True], for_prompt_emb, edit_prompt_emb, null_prompt_emb)
        nf = ops.combine3(ef, 1.0, en, -1.0)
        ne = ops.combine3(ee, 1.0, en, -1.0)
        coef = float(ops.gram(ne.reshape(1, -1), nf.reshape(1, -1))[0, 0] / ops.gram(nf.reshape(1, -1), nf.reshape(1, -1))[0, 0])
        # e_null + g * nf + g_edit * (ne - coef * nf)
        return ops.combine3(en, 1.0, nf, self.guidance_scale - self.guidance_scale_edit * coef, ne, self.guidance_scale_edit)

    def get_x0(self, xt, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None,
               mode="null+(for-null)+(edit-null)", flatten=False):
        """src/modules/edit.py:1566-1587."""
        do_cfg = self.guidance_scale > 1.0
        xt = xt.to(self.device, torch.float32).contiguous()
        noise_pred = self._classifer_free_guidance(xt, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mode, do_cfg)
        x0_hat = ops.pmp_forward(xt, noise_pred, self.scheduler.alpha_at(float(t)))
        if mask is not None:
            return ops.gather_rows(x0_hat.reshape(x0_hat.shape[0], -1), ops.mask_indices(mask.to(self.device)))
        if flatten:
            return x0_hat.reshape(-1, xt[0].numel())
        return x0_hat

    @torch.no_grad()
    def mask_diffedit(self, x0, for_prompt_emb, edit_prompt_emb, null_prompt_emb, noise=None):
        """src/modules/edit.py:1395-1407 (t = 500, ten noised copies; `noise` injects the draw)."""
        t = 500.0
        at = self.scheduler.alpha_at(t)
        if noise is None:
            noise = torch.randn(10, self.c_in, self.image_size, self.image_size, dtype=torch.float32, device=self.device)
        xt = (at ** 0.5) * x0.to(self.device, torch.float32) + ((1 - at) ** 0.5) * noise.to(self.device)
        eps_1 = self._classifer_free_guidance(xt, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, "null+(for-null)", True)
        eps_2 = self._classifer_free_guidance(xt, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, "null+(edit-null)", True)
        return diffedit_mask(eps_1, eps_2)

    @torch.no_grad()
    def DDPMforwardsteps(self, xt, t_start_idx, t_end_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb,
                         mode="null+(for-null)", **kwargs):
        """src/modules/edit.py:1410-1483: DDIM (eta = 0) denoising under classifier-free guidance; returns
        (xt, t, t_idx) at t_end_idx, else the uint8 images [B, H, W, 3]."""
        assert mode in ["null+(for-null)+(edit-null)", "null+(for-null)", "null+(edit-null)"]
        do_cfg = self.guidance_scale > 1.0
        self.scheduler.set_timesteps(self.for_steps, device=self.device)
        ts = self.scheduler._ts_host
        xt = xt.to(self.device, torch.float32).contiguous()
        for t_idx, t in enumerate(ts):
            if t_idx < t_start_idx:
                continue
            elif t_idx == t_end_idx and t_idx != t_start_idx:
                return xt, self.scheduler.timesteps[t_idx], t_idx
            noise_pred = self._classifer_free_guidance(xt, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mode, do_cfg)
            xt = self.scheduler.step(noise_pred, t, xt, eta=0, t_idx=t_idx).prev_sample
        self.last_images.append(xt)
        img = (xt / 2 + 0.5).clamp(0, 1)
        return (img * 255).to(torch.uint8).permute(0, 2, 3, 1)

    # ------------------------------------------------------------------ Jacobian products under guidance
    def _cfg_slots(self, mode, embs):
        """[(plan slot, weight, prompt embedding)] of the conditionings the guided prediction of `mode`
        combines; embs = (for, edit, null)."""
        do_cfg = self.guidance_scale > 1.0
        w = linear_cfg_weights(mode, self.guidance_scale, self.guidance_scale_edit, do_cfg)
        if w is None:
            raise NotImplementedError(f"Jacobian products need a linear guidance mode, not '{mode}'")
        return [(i, wi, embs[i]) for i, wi in enumerate(w) if wi != 0.0]

    def _combine(self, terms):
        """sum_i w_i a_i for up to three (tensor, weight) terms."""
        (a, wa), rest = terms[0], terms[1:]
        (b, wb) = rest[0] if len(rest) > 0 else (None, 0.0)
        (c, wc) = rest[1] if len(rest) > 1 else (None, 0.0)
        return ops.combine3(a, wa, b, wb, c, wc)

    def _jvp_cfg(self, x_row, t, V, slots):
        """Guided tangents sum_i w_i J_i V^T: one fused primal + k-tangent pass per conditioning; the
        primal activations of conditioning i stay in plan slot i for the transposed pass."""
        k = V.shape[0]
        xin = torch.cat([x_row.reshape(1, -1), V], 0).reshape(1 + k, *self.unet.base.in_shape).contiguous()
        outs = []
        for slot, wi, emb in slots:
            plan = self.unet.base.plan(1, k, k, slot=slot)
            self.unet.apply_condition(plan, emb)
            outs.append((plan.forward(xin, float(t))[1:].reshape(k, -1), wi))
        return self._combine(outs)

    def _vjp_cfg(self, g_eps, slots):
        k = g_eps.shape[0]
        g = g_eps.reshape(k, *self.unet.base.out_shape).contiguous()
        outs = [(self.unet.base.plan(1, k, k, slot=slot).vjp(g).reshape(k, -1), wi) for slot, wi, _ in slots]
        return self._combine(outs)

    def local_encoder_decoder_pullback_xt(self, xt, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb,
                                          op=None, block_idx=None, pca_rank=50, chunk_size=25, min_iter=10,
                                          max_iter=100, convergence_threshold=1e-3, mask=None,
                                          mode="null+(for-null)+(edit-null)", v0=None):
        """src/modules/edit.py:1589-1676: subspace iteration on J^T J, J = d(mask o x0_hat)/d x_t with the
        GUIDED noise prediction.  Returns (u [l_o, k], s [k] = sqrt(svdvals), vT [k, d]) like the reference.
        (`chunk_size`: every chunk of tangents is a separate fused pass in the reference; here all k run
        in one pass per conditioning.)"""
        assert mode in ["null+(for-null)+(edit-null)", "null+(for-null)", "null+(edit-null)", "(for-edit)"]
        slots = self._cfg_slots(mode, (for_prompt_emb, edit_prompt_emb, null_prompt_emb))
        k = int(pca_rank)
        d = xt[0].numel()
        x = xt.to(self.device, torch.float32).contiguous().reshape(1, -1)
        at = self.scheduler.alpha_at(float(t))
        m8 = None if mask is None else mask.to(self.device).reshape(-1).to(torch.uint8).contiguous()
        if v0 is None and self.v0 is not None:
            v0 = self.v0.get(k) if isinstance(self.v0, dict) else self.v0
        V = (random_basis(d, k, self.device) if v0 is None else v0.to(self.device, torch.float32).reshape(k, d)).contiguous()
        u = s = None
        for i in range(max_iter):
            deps = self._jvp_cfg(x, t, V, slots)                              # :1637-1641 (jacfwd)
            u, g_eps, gx_direct = ops.pmp_jvp_epilogue(V, deps, m8, at)
            w = ops.combine3(gx_direct, 1.0, self._vjp_cfg(g_eps, slots), 1.0)    # :1645-1654 (jacobian)
            V_new, s = ops.orthonormalise(w, v_prev=V if self.align_sign else None)   # :1656
            need_check = i > min_iter
            if self.verbose or need_check:
                convergence = torch.dist(V, V_new).item()
                if self.verbose:
                    print(f'power method : {i}-th step convergence : ', convergence)
            done = need_check and torch.allclose(V, V_new, atol=convergence_threshold)
            V = V_new
            if done:
                break
        u_out = u if m8 is None else ops.gather_rows(u, ops.mask_indices(m8))
        return u_out.T, s, V

    @torch.no_grad()
    def get_delta_xt_via_grad(self, xt, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None,
                              mode="null+(for-null)+(edit-null)"):
        """src/modules/edit.py:1680-1720: text-supervised direction v = normalise(J_mode^T (x0_hat_after - x0_hat)),
        ONE transposed pass with a data-dependent cotangent."""
        x = xt.to(self.device, torch.float32).contiguous()
        d = x[0].numel()
        at = self.scheduler.alpha_at(float(t))
        x0_hat = self.get_x0(x, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None, mode="null+(for-null)")
        x0_hat_after = self.get_x0(x, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None, mode=mode)
        delta = ops.combine3(x0_hat_after.reshape(1, d), 1.0, x0_hat.reshape(1, d), -1.0)
        m8 = None if mask is None else mask.to(self.device).reshape(-1).to(torch.uint8).contiguous()
        slots = self._cfg_slots(mode, (for_prompt_emb, edit_prompt_emb, null_prompt_emb))
        # the primal activations of every conditioning at x_t (the tangent row of the pass is a dummy)
        self._jvp_cfg(x.reshape(1, -1), t, torch.zeros(1, d, device=self.device), slots)
        # <delta_masked, P(x)>: seeds of the transposed pass = the PMP epilogue with V := delta and no eps tangent,
        # i.e. u = mask o delta / sqrt(at) ... the epilogue's seeds are linear in u, so feed u = mask o delta
        # through it by scaling V by sqrt(at)
        zero = torch.zeros_like(delta)
        _, g_eps, gx_direct = ops.pmp_jvp_epilogue(ops.axpy(zero, delta, at ** 0.5), zero, m8, at)
        v_ = ops.combine3(gx_direct, 1.0, self._vjp_cfg(g_eps, slots), 1.0)
        return ops.nullspace_project(v_, None, project=False)              # v_ / ||v_||   (:1710)

    @torch.no_grad()
    def get_v_modify(self, xt, t, t_idx, for_prompt_emb, edit_prompt_emb, null_prompt_emb, mask=None,
                     mode="(for-edit)-direct", jacobian=False):
        """src/modules/edit.py:1723-1741."""
        if jacobian:
            return self.get_delta_xt_via_grad(xt, t, t_idx, self.for_prompt_emb, self.edit_prompt_emb,
                                              self.null_prompt_emb, mask=mask, mode=self.tilda_v_score_type)
        cfg = lambda m: self._classifer_free_guidance(xt, t, for_prompt_emb, edit_prompt_emb, null_prompt_emb, m, True)
        if mode == "(for-edit)-direct":
            return cfg("(for-edit)").reshape(1, -1)
        if mode == "(edit-null)-direct":
            e = cfg("(edit-null)").reshape(1, -1)
            return ops.combine3(e, -1.0)
        if mode == "proj_null[for-null](edit-null)-direct":
            eps_1 = cfg("(for-null)").reshape(1, -1)
            eps_2 = cfg("(edit-null)").reshape(1, -1)
            coef = float(ops.gram(eps_2, eps_1)[0, 0] / ops.gram(eps_1, eps_1)[0, 0])
            return ops.combine3(eps_2, -1.0, eps_1, coef)
        raise ValueError(mode)

    @torch.no_grad()
    def x_space_guidance_direct(self, xt, t_idx, vk, single_edit_step):
        """src/modules/edit.py:2021-2029."""
        return ops.axpy(xt.contiguous(), vk.expand_as(xt).contiguous(), self.x_space_guidance_scale * single_edit_step)

    # ------------------------------------------------------------------ drivers
    def _edit_batch(self, original_xt, v_row, vis_num):
        """:1840-1858 (same list building as the unconditional driver)."""
        xts = {}
        for direction in [1, -1]:
            vk = (direction * v_row).view(-1, *original_xt.shape[1:]).contiguous()
            xt_list = [original_xt.clone()]
            for _ in range(self.x_space_guidance_num_step):
                xt_list.append(self.x_space_guidance_direct(xt_list[-1], t_idx=self.edit_t_idx, vk=vk,
                                                            single_edit_step=self.x_space_guidance_edit_step))
            xt = torch.cat(xt_list, dim=0)
            xts[direction] = xt[[0, -1], :] if vis_num == 1 else xt[::(xt.size(0) // vis_num)]
        return torch.cat([(xts[-1].flip(dims=[0]))[:-1], xts[1]], dim=0).contiguous()

    def _start(self, mask_index):
        """x_T, the mask and x_t at the edit timestep (:1766-1790 without SAM / super-resolution)."""
        self.scheduler.set_timesteps(self.for_steps, device=self.device)
        xT = self.xT if self.xT is not None else torch.randn(1, self.c_in, self.image_size, self.image_size,
                                                             dtype=torch.float32, device=self.device)
        xT = xT.to(self.device)
        mask = load_mask(self.result_folder, mask_index).to(self.device)            # :1777-1780
        kw = dict(for_prompt_emb=self.for_prompt_emb, edit_prompt_emb=self.edit_prompt_emb,
                  null_prompt_emb=self.null_prompt_emb, mode="null+(for-null)")
        xt, t, t_idx = self.DDPMforwardsteps(xT, t_start_idx=0, t_end_idx=self.edit_t_idx, **kw)
        assert t_idx == self.edit_t_idx
        return xT, mask, xt, t, t_idx, kw

    @torch.no_grad()
    def run_edit_null_space_projection_xt(self, op, block_idx, vis_num, mask_index=0, vis_num_pc=1, vis_vT=False,
                                          pca_rank=50, edit_prompt=None, null_space_projection=False, pca_rank_null=50):
        """src/modules/edit.py:1745-1871 (non-semantic T-LOCO edit)."""
        if self.sampling_mode:
            return None
        xT, mask, xt, t, t_idx, kw = self._start(mask_index)
        save_dir = os.path.join(self.result_folder, "basis", f'local_basis-{self.edit_t}T-pca-rank-{pca_rank}-select-mask{mask_index}')
        os.makedirs(save_dir, exist_ok=True)
        paths = dict(u_m=os.path.join(save_dir, 'u-modify.pt'), v_m=os.path.join(save_dir, 'vT-modify.pt'),
                     u_n=os.path.join(save_dir, f'u-null-null_space_rank_{pca_rank_null}.pt'),
                     v_n=os.path.join(save_dir, f'vT-null-null_space_rank_{pca_rank_null}.pt'))
        if all(os.path.exists(p) for p in paths.values()):
            vT_modify = torch.load(paths["v_m"], map_location=self.device).type(torch.float32)
            vT_null = torch.load(paths["v_n"], map_location=self.device).type(torch.float32)
        else:
            pb = dict(op=op, block_idx=block_idx, chunk_size=5, min_iter=10, max_iter=50, convergence_threshold=1e-3,
                      mode="null+(for-null)")
            u_modify, _, vT_modify = self.local_encoder_decoder_pullback_xt(
                xt, t, t_idx, self.for_prompt_emb, self.edit_prompt_emb, self.null_prompt_emb, pca_rank=pca_rank, mask=mask, **pb)
            torch.save(u_modify, paths["u_m"])
            torch.save(vT_modify, paths["v_m"])
            vT_null = None
            if null_space_projection:
                u_null, _, vT_null = self.local_encoder_decoder_pullback_xt(
                    xt, t, t_idx, self.for_prompt_emb, self.edit_prompt_emb, self.null_prompt_emb, pca_rank=pca_rank_null,
                    mask=~mask, **pb)
                torch.save(u_null, paths["u_n"])
                torch.save(vT_null, paths["v_n"])
        if not null_space_projection:
            vT = ops.nullspace_project(vT_modify.contiguous(), None, project=False)
        else:
            vT = ops.nullspace_project(vT_modify.contiguous(), vT_null[:pca_rank_null, :].contiguous(), project=True)
        self.last_images = []
        imgs = None
        for pc_idx in range(vis_num_pc):
            self.EXP_NAME = (f'Non-semantic_Edit_xt-edit_{self.edit_t}T-select_mask{mask_index}-edit_space_rank-{pc_idx}-'
                             f'null_space_projection_{null_space_projection}-null_space_rank_{pca_rank_null}_{self.tilda_v_score_type}')
            batch = self._edit_batch(xt, vT[pc_idx, :], vis_num)
            imgs = self.DDPMforwardsteps(batch, t_start_idx=self.edit_t_idx, t_end_idx=-1, **kw)
        return dict(vT=vT, vT_modify=vT_modify, vT_null=vT_null, xt=xt, images=imgs)

    @torch.no_grad()
    def run_edit_null_space_projection_xt_semantic(self, op, block_idx, vis_num, mask_index=0, vis_num_pc=1,
                                                   vis_vT=False, pca_rank=50, edit_prompt=None, null_space_projection=False,
                                                   pca_rank_null=50, jacobian=False):
        """src/modules/edit.py:1874-2018, ablation_method "null-space-proj" (the SEGA / DiffEdit branches
        are baselines of the paper, not this path)."""
        if self.ablation_method != "null-space-proj":
            raise NotImplementedError("only the 'null-space-proj' method is part of the LOCO path")
        if self.sampling_mode:
            return None
        xT, mask, xt, t, t_idx, kw = self._start(mask_index)
        save_dir = os.path.join(self.result_folder, "basis")
        os.makedirs(save_dir, exist_ok=True)
        if not os.path.exists(self.vT_path):
            vT_modify = self.get_v_modify(xt, t, t_idx, self.for_prompt_emb, self.edit_prompt_emb, self.null_prompt_emb,
                                          mask=mask, mode=self.tilda_v_score_type, jacobian=jacobian)
            if null_space_projection:
                _, _, vT_null = self.local_encoder_decoder_pullback_xt(
                    xt, t, t_idx, self.for_prompt_emb, self.edit_prompt_emb, self.null_prompt_emb, op=op, block_idx=block_idx,
                    pca_rank=pca_rank_null, chunk_size=5, min_iter=10, max_iter=50, convergence_threshold=1e-3,
                    mask=~mask, mode="null+(for-null)")
                vT = ops.nullspace_project(vT_modify.contiguous(), vT_null[:pca_rank_null, :].contiguous(), project=True)
            else:
                vT = ops.nullspace_project(vT_modify.contiguous(), None, project=False)
            BASIS_NAME = (f"edit-{self.edit_t}T-edit_prompt-{self.edit_prompt}-select_mask{mask_index}-null_space_projection_"
                          f"{null_space_projection}_null_space_rank_{pca_rank_null}_{self.tilda_v_score_type}")
            for pc_idx in range(min(vT.shape[0], vis_num_pc)):
                self.EXP_NAME = f'Semantic_Edit_xt-{BASIS_NAME}-pc_{pc_idx:0=3d}'
                torch.save(vT[[pc_idx], :], os.path.join(save_dir, f'{self.EXP_NAME}-vT.pt'))
        else:
            vT = torch.load(self.vT_path, map_location=self.device)
            BASIS_NAME = f"load-basis-'{os.path.basename(self.vT_path)}'"
        self.last_images = []
        batch = None
        for pc_idx in range(vis_num_pc):
            self.EXP_NAME = f'Semantic_Edit_xt-{BASIS_NAME}_scale_{self.x_space_guidance_scale}'
            batch = self._edit_batch(xt, vT[pc_idx, :], vis_num)
        imgs = self.DDPMforwardsteps(batch, t_start_idx=self.edit_t_idx, t_end_idx=-1, **kw)
        return dict(vT=vT, xt=xt, images=imgs)

    @torch.no_grad()
    def group_edit_null_space_projection(self, vT_paths, mask_index=0):
        """Composition of n saved directions, cumulatively with step scale * num_step, like the
        unconditional `group_edit_null_space_projection` (src/modules/edit.py:2171-2212, two directions
        there; BASELINE config 5 composes three)."""
        xT, mask, xt, t, t_idx, kw = self._start(mask_index)
        xt_temp = xt.detach().clone()
        vis = [xt_temp]
        for p in vT_paths:
            vT = torch.load(p, map_location=self.device)
            vk = vT[0, :].view(-1, *xt.shape[1:]).to(torch.float32).contiguous()
            xt_temp = ops.axpy(xt_temp, vk, self.x_space_guidance_scale * self.x_space_guidance_num_step)
            vis.append(xt_temp)
        self.EXP_NAME = f'Group_Edit_xt-load-basis-{len(vT_paths)}'
        imgs = self.DDPMforwardsteps(torch.cat(vis, dim=0), t_start_idx=self.edit_t_idx, t_end_idx=-1, **kw)
        return dict(xt=xt, latents=torch.cat(vis, dim=0), images=imgs)
