"""Golden fixtures for the latent-space (Stable-Diffusion) twin of the hot path from the UNMODIFIED reference
class `EditStableDiffusion` (src/modules/edit.py:483-1196), run on CPU against stand-in networks (build
container only; see make_golden.py for the import stubs):

    python tests/golden/make_golden_sd.py        # -> tests/golden/sd_tiny.pt

Stable Diffusion's U-Net and VAE are diffusers models that are not under /root/reference, so
both networks are stand-ins plugged into the reference object as `self.unet` / `self.vae`:
  * U-Net: the oracle's text-conditioned DDPM U-Net over 4-channel latents (cross-attention to the prompt
    tokens in every AttnBlock), `unet(z, t, encoder_hidden_states=...) -> .sample`;
  * VAE:   the oracle's restatement of the AutoencoderKL decoder (oracle/vae_ref.py), `vae.decode(z).sample`.
The scheduler is a stub holding Stable Diffusion's "scaled_linear" alpha_bar table, monkey-patched by the
reference's own `get_stable_diffusion_scheduler` (src/utils/utils.py:147-157).  Everything else is the
reference's code: _classifer_free_guidance (4 modes), get_x0 (with the decode), the latent-space power method
local_encoder_decoder_pullback_zt (N = 1, 2 and 6; two guidance modes; mask and ~mask over the DECODED image),
get_delta_zt_via_grad, DDIMforwardsteps (last 9 steps + decode).
"""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

DIM, NTOK, RZ = 64, 8, 16
G, G_EDIT = 7.5, 4.0


def main():
    torch.set_num_threads(os.cpu_count())
    ddpm, uu, edit = mg.import_reference()
    from loco_edit_b200.scheduler import scaled_linear_betas
    from loco_edit_b200.t2i import synthetic_prompt_embedding
    from loco_edit_b200.weights import latent_unet_arch, random_state_dict, tiny_vae_decoder_arch
    from oracle import ddpm_ref, vae_ref
    arch = latent_unet_arch(resolution=RZ, ch_mult=(1, 2), attn_resolutions=(8,), num_res_blocks=1, ctx_dim=DIM, ctx_heads=2)
    varch = tiny_vae_decoder_arch(resolution=RZ, ch_mult=(1, 2), num_res_blocks=1)
    sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
    vsd = random_state_dict(varch, seed=4321, perturb_norm=0.1)
    RX = RZ << (len(varch["ch_mult"]) - 1)

    class StandIn(torch.nn.Module):
        def forward(self, x, t, encoder_hidden_states=None):
            outs = [ddpm_ref.unet_forward(sd, arch, x[b:b + 1], t, ctx=encoder_hidden_states[b]) for b in range(x.shape[0])]
            return types.SimpleNamespace(sample=torch.cat(outs, 0))

    args = mg.ns(use_yh_custom_scheduler=True, device=torch.device("cpu"), dtype=torch.float32)
    betas = scaled_linear_betas(1000)
    stub = types.SimpleNamespace(alphas_cumprod=torch.cumprod(1.0 - betas, 0), betas=betas)
    sched = uu.get_stable_diffusion_scheduler(args, stub)
    e = object.__new__(edit.EditStableDiffusion)
    e.unet = StandIn()
    e.vae = vae_ref.RefVAE(varch, vsd)
    e.scheduler = sched
    e.device, e.dtype = torch.device("cpu"), torch.float32
    e.guidance_scale, e.guidance_scale_edit = G, G_EDIT
    e.for_steps, e.use_yh_custom_scheduler = 100, True
    e.buffer_device, e.memory_bound = "cpu", 2
    e.c_in, e.image_size = 4, RX
    e.tilda_v_score_type = "null+(for-null)+(edit-null)"
    e.result_folder = tempfile.mkdtemp()
    e.EXP_NAME = "golden"
    prompts = ["a photo of a dog", "a dog with glasses", ""]
    embs = [synthetic_prompt_embedding(p, NTOK, DIM) for p in prompts]
    e.for_prompt_emb, e.edit_prompt_emb, e.null_prompt_emb = embs
    sched.set_timesteps(100)
    t_idx = 60
    t = sched.timesteps[t_idx]
    g = torch.Generator().manual_seed(5)
    zt = torch.randn(1, 4, RZ, RZ, generator=g)
    z2 = torch.randn(2, 4, RZ, RZ, generator=g)
    mask = torch.zeros(3, RX, RX, dtype=torch.bool)
    mask[:, 12:20, 8:24] = True
    out = {"arch": arch, "vae_arch": varch, "dim": DIM, "ntok": NTOK, "prompts": prompts, "g": G, "g_edit": G_EDIT,
           "t_idx": t_idx, "t": t.clone(), "zt": zt, "z2": z2, "mask": mask,
           "timesteps": sched.timesteps.clone(), "alphas_cumprod": sched.alphas_cumprod.clone()}
    with torch.no_grad():
        out["decode"] = e.vae.decode(z2).sample
        out["cfg"] = {}
        for mode in ["null+(for-null)+(edit-null)", "null+(for-null)", "null+(edit-null)", "(for-edit)"]:
            out["cfg"][mode] = e._classifer_free_guidance(z2, t, *embs, mode=mode, do_classifier_free_guidance=True)
        out["cfg_off"] = e._classifer_free_guidance(z2, t, *embs, mode="null+(for-null)", do_classifier_free_guidance=False)
        out["x0_masked"] = e.get_x0(zt, t, t_idx, *embs, mask=mask, mode="null+(for-null)")
        out["x0_flat"] = e.get_x0(zt, t, t_idx, *embs, mask=None, mode="null+(for-null)+(edit-null)", flatten=True)
    out["pullback"] = {}
    for mode, m, mname in [("null+(for-null)", mask, "mask"), ("null+(for-null)", ~mask, "~mask"),
                           ("null+(for-null)+(edit-null)", mask, "mask")]:
        for n_iter in (1, 2):
            torch.manual_seed(7)
            u, s, vT = e.local_encoder_decoder_pullback_zt(zt, t, t_idx, *embs, pca_rank=2, chunk_size=5, min_iter=10 ** 6,
                                                           max_iter=n_iter, convergence_threshold=1e-3, mask=m, mode=mode)
            out["pullback"][(mode, mname, n_iter)] = {"u": u.clone(), "s": s.clone(), "vT": vT.clone()}
            print(mode, mname, n_iter, s.tolist())
    # depth: six iterations of the masked power method (the error of the CUDA path adds up over iterations)
    torch.manual_seed(7)
    u, s, vT = e.local_encoder_decoder_pullback_zt(zt, t, t_idx, *embs, pca_rank=2, chunk_size=5, min_iter=10 ** 6,
                                                   max_iter=6, convergence_threshold=1e-3, mask=mask, mode="null+(for-null)")
    out["pullback"][("null+(for-null)", "mask", 6)] = {"u": u.clone(), "s": s.clone(), "vT": vT.clone()}
    print("N=6", s.tolist())
    out["delta_masked"] = e.get_delta_zt_via_grad(zt, t, t_idx, *embs, mask=mask, mode="null+(for-null)+(edit-null)").clone()
    edit.tvu.save_image = lambda *a, **k: None
    with torch.no_grad():
        lat, u8 = e.DDIMforwardsteps(z2.clone(), t_start_idx=90, t_end_idx=-1, for_prompt_emb=embs[0],
                                     edit_prompt_emb=embs[1], null_prompt_emb=embs[2], mode="null+(for-null)")
        out["ddim_final"] = {"latents": lat, "u8": u8}
        z_mid, t_mid, i_mid = e.DDIMforwardsteps(z2.clone(), t_start_idx=88, t_end_idx=92, for_prompt_emb=embs[0],
                                                 edit_prompt_emb=embs[1], null_prompt_emb=embs[2],
                                                 mode="null+(for-null)+(edit-null)")
        out["ddim_mid"] = {"zt": z_mid, "t": t_mid, "idx": i_mid}
    torch.save(out, os.path.join(HERE, "sd_tiny.pt"))
    print("written", {k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in out.items()})


if __name__ == "__main__":
    main()
