"""Golden fixtures for the latent-consistency twin of the hot path from the UNMODIFIED reference class
`EditLatentConsistency` (src/modules/edit.py:42-480), run on CPU against stand-ins (build container only):

    python tests/golden/make_golden_lcm.py        # -> tests/golden/lcm_tiny.pt

Everything this class calls outside /root/reference is diffusers code (the LCM pipeline, its U-Net, VAE and
`LCMScheduler`), so all of it is stood in for ("parity unpinned at the network level" -- here also at
the scheduler level, see loco_edit_b200/scheduler.py:LCMScheduler):
  * `pipe.encode_prompt(prompt, ...)` -> seeded prompt embedding; `pipe.get_guidance_scale_embedding` -> the
    published sinusoidal w embedding (loco_edit_b200/sd.py:guidance_scale_embedding);
  * `unet(z, t, timestep_cond=w_emb, encoder_hidden_states=..., return_dict=False)[0]` -> the oracle's
    text-conditioned latent U-Net with w_emb @ P added to its timestep embedding;
  * `vae.decode(z, return_dict=False)[0]`, `vae.config.scaling_factor` -> oracle/vae_ref.py;
  * `scheduler.step(eps, t, z, return_dict=False) -> (prev_sample, denoised)` -> a CPU restatement of diffusers'
    LCMScheduler (below; noise injected from a seeded list so that both sides see the same draws).
What IS the reference's code and is pinned by this fixture: get_x0 (:206-248), get_delta_zt_via_grad (:251-280),
local_encoder_decoder_pullback_zt (:283-369, N = 1, 2; mask and ~mask), LCMforwardsteps (:148-203).
"""
import math
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

DIM, NTOK, RZ, WDIM = 64, 8, 16, 256
GUIDANCE = 8.0
STEPS = 4


class RefLCMScheduler(object):
    """diffusers `LCMScheduler` (epsilon prediction, original_inference_steps 50, timestep_scaling 10), CPU torch."""

    def __init__(self, alphas_cumprod, noise_list):
        self.alphas_cumprod = alphas_cumprod
        self.noise_list = noise_list
        self.sigma_data, self.timestep_scaling = 0.5, 10.0

    def set_timesteps(self, n, device=None):
        origin = torch.arange(1, 51) * 20 - 1
        self.timesteps = origin.flip(0)[::50 // n][:n].clone()
        self.n = n

    def step(self, model_output, timestep, sample, return_dict=False):
        i = self.timesteps.tolist().index(int(timestep))
        a = self.alphas_cumprod[int(timestep)]
        s = self.timestep_scaling * float(int(timestep))
        c_skip = self.sigma_data ** 2 / (s ** 2 + self.sigma_data ** 2)
        c_out = s / (s ** 2 + self.sigma_data ** 2) ** 0.5
        pred_x0 = (sample - (1 - a).sqrt() * model_output) / a.sqrt()
        denoised = c_out * pred_x0 + c_skip * sample
        if i == self.n - 1:
            return denoised, denoised
        a_prev = self.alphas_cumprod[int(self.timesteps[i + 1])]
        noise = self.noise_list[i][:sample.shape[0]]
        return a_prev.sqrt() * denoised + (1 - a_prev).sqrt() * noise, denoised


def main():
    torch.set_num_threads(os.cpu_count())
    ddpm, uu, edit = mg.import_reference()
    from loco_edit_b200.scheduler import scaled_linear_betas
    from loco_edit_b200.sd import guidance_scale_embedding
    from loco_edit_b200.t2i import cond_projection, synthetic_prompt_embedding
    from loco_edit_b200.weights import latent_unet_arch, random_state_dict, tiny_vae_decoder_arch
    from oracle import ddpm_ref, vae_ref
    arch = latent_unet_arch(resolution=RZ, ch_mult=(1, 2), attn_resolutions=(8,), num_res_blocks=1, ctx_dim=DIM, ctx_heads=2)
    varch = tiny_vae_decoder_arch(resolution=RZ, ch_mult=(1, 2), num_res_blocks=1)
    sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
    vsd = random_state_dict(varch, seed=4321, perturb_norm=0.1)
    RX = RZ << (len(varch["ch_mult"]) - 1)
    P = cond_projection(WDIM, 4 * arch["ch"])

    class StandInUNet(torch.nn.Module):
        config = types.SimpleNamespace(time_cond_proj_dim=WDIM)

        def forward(self, x, t, timestep_cond=None, encoder_hidden_states=None, return_dict=False):
            outs = [ddpm_ref.unet_forward(sd, arch, x[b:b + 1], t, cond=timestep_cond[b] @ P, ctx=encoder_hidden_states[b])
                    for b in range(x.shape[0])]
            return (torch.cat(outs, 0),)

    ref_vae = vae_ref.RefVAE(varch, vsd)

    class StandInVAE(object):
        config = types.SimpleNamespace(scaling_factor=0.18215)

        def decode(self, z, return_dict=False):
            return (ref_vae.decode(z).sample,)

    pipe = types.SimpleNamespace(
        encode_prompt=lambda prompt, device, num_images_per_prompt=1, do_classifier_free_guidance=False:
        (synthetic_prompt_embedding(prompt, NTOK, DIM), None),
        get_guidance_scale_embedding=lambda w, embedding_dim: guidance_scale_embedding(w, embedding_dim))
    g = torch.Generator().manual_seed(5)
    noise_list = [torch.randn(5, 4, RZ, RZ, generator=g) for _ in range(STEPS)]
    betas = scaled_linear_betas(1000)
    sched = RefLCMScheduler(torch.cumprod(1.0 - betas, 0), noise_list)

    e = object.__new__(edit.EditLatentConsistency)
    e.pipe, e.unet, e.vae, e.scheduler = pipe, StandInUNet(), StandInVAE(), sched
    e.device, e.dtype = torch.device("cpu"), torch.float32
    e.guidance_scale, e.guidance_scale_edit = GUIDANCE, 4.0
    e.num_inference_steps = STEPS
    e.result_folder = tempfile.mkdtemp()
    e.EXP_NAME = "golden"
    for_prompt, edit_prompt = "a photo of a dog", "a dog with glasses"
    e.for_prompt, e.edit_prompt = for_prompt, edit_prompt
    sched.set_timesteps(STEPS)
    t_idx = 1
    t = sched.timesteps[t_idx]
    zt = torch.randn(1, 4, RZ, RZ, generator=g)
    z2 = torch.randn(2, 4, RZ, RZ, generator=g)
    mask = torch.zeros(3, RX, RX, dtype=torch.bool)
    mask[:, 12:20, 8:24] = True
    out = {"arch": arch, "vae_arch": varch, "dim": DIM, "ntok": NTOK, "wdim": WDIM, "guidance": GUIDANCE, "steps": STEPS,
           "prompts": [for_prompt, edit_prompt], "t_idx": t_idx, "t": t.clone(), "zt": zt, "z2": z2, "mask": mask,
           "timesteps": sched.timesteps.clone(), "noise_list": noise_list}
    with torch.no_grad():
        out["x0_masked"] = e.get_x0(zt, for_prompt, t, t_idx, mask=mask)
        out["x0_flat"] = e.get_x0(z2, edit_prompt, t, t_idx, flatten=True)
    out["pullback"] = {}
    for m, mname in [(mask, "mask"), (~mask, "~mask")]:
        for n_iter in (1, 2):
            torch.manual_seed(7)
            u, s, vT = e.local_encoder_decoder_pullback_zt(zt, t, t_idx, for_prompt, pca_rank=2, chunk_size=5, min_iter=10 ** 6,
                                                           max_iter=n_iter, convergence_threshold=1e-3, mask=m)
            out["pullback"][(mname, n_iter)] = {"u": u.clone(), "s": s.clone(), "vT": vT.clone()}
            print(mname, n_iter, s.tolist())
    out["delta_masked"] = e.get_delta_zt_via_grad(zt, t, t_idx, for_prompt, edit_prompt, mask=mask).clone()
    edit.tvu.save_image = lambda *a, **k: None
    with torch.no_grad():
        lat, u8 = e.LCMforwardsteps(z2.clone(), for_prompt, t_start_idx=0, t_end_idx=-1)
        out["lcm_final"] = {"latents": lat, "u8": u8}
        z_mid, t_mid, i_mid = e.LCMforwardsteps(z2.clone(), for_prompt, t_start_idx=0, t_end_idx=2)
        out["lcm_mid"] = {"zt": z_mid, "t": t_mid, "idx": i_mid}
    torch.save(out, os.path.join(HERE, "lcm_tiny.pt"))
    print("written", {k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in out.items()})


if __name__ == "__main__":
    main()
