"""Golden fixtures for the T-LOCO (text-conditioned) Edit-class logic from the UNMODIFIED reference
class `EditDeepFloydIF` (src/modules/edit.py:1198-2031), run on CPU against a stand-in conditional
U-Net (build container only; see make_golden.py for the import stubs):

    python tests/golden/make_golden_t2i.py        # -> tests/golden/t2i_tiny.pt
    python tests/golden/make_golden_t2i.py cross  # -> tests/golden/t2i_cross_tiny.pt (stand-in with real
                                                  #    cross-attention to the prompt embedding)

The DeepFloyd-IF / Stable-Diffusion networks are diffusers models that are not under /root/reference, so
the network is a stand-in: the oracle's DDPM U-Net with the conditioning embedding
c = mean_tokens(encoder_hidden_states) @ P added to the timestep embedding, returning 6 channels
(eps | zeros) like the learned-variance IF U-Net so that the reference's `.split(3, dim=1)` works.
The scheduler is a stub holding a cosine alpha_bar table, monkey-patched by the reference's own
`get_deepfloyd_if_scheduler` (src/utils/utils.py:159-213).  Everything else is the reference's code:
  _classifer_free_guidance (7 of its 8 modes; 'edit-proj[for](edit)' raises NameError there),
  get_x0, local_encoder_decoder_pullback_xt (N = 1, 2; two guidance modes), get_delta_xt_via_grad,
  get_v_modify (three direct modes), DDPMforwardsteps (last 9 steps).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

DIM, NTOK, R = 64, 8, 32
G, G_EDIT = 7.5, 4.0


def main(cross=False):
    torch.set_num_threads(os.cpu_count())
    ddpm, uu, edit = mg.import_reference()
    from loco_edit_b200.scheduler import cosine_betas
    from loco_edit_b200.t2i import cond_projection, synthetic_prompt_embedding
    from loco_edit_b200.weights import random_state_dict, tiny_arch
    from oracle import ddpm_ref
    arch = tiny_arch(resolution=R, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1,
                     ctx_dim=DIM if cross else 0, ctx_heads=2)
    sd = random_state_dict(arch, seed=1234, perturb_norm=0.1)
    P = cond_projection(DIM, 4 * arch["ch"])

    class StandIn(torch.nn.Module):
        def forward(self, x, t, encoder_hidden_states=None):
            outs = []
            for b in range(x.shape[0]):
                if cross:      # every AttnBlock attends to the prompt tokens themselves
                    outs.append(ddpm_ref.unet_forward(sd, arch, x[b:b + 1], t, ctx=encoder_hidden_states[b]))
                    continue
                c = encoder_hidden_states[b].mean(0) @ P
                outs.append(ddpm_ref.unet_forward(sd, arch, x[b:b + 1], t, cond=c))
            eps = torch.cat(outs, 0)
            return types.SimpleNamespace(sample=torch.cat([eps, torch.zeros_like(eps)], dim=1))

    args = mg.ns(use_yh_custom_scheduler=True, device=torch.device("cpu"), dtype=torch.float32)
    stub = types.SimpleNamespace(alphas_cumprod=torch.cumprod(1.0 - cosine_betas(1000), 0).float(),
                                 betas=cosine_betas(1000).float(), scale_model_input=lambda x, t: x)
    sched = uu.get_deepfloyd_if_scheduler(args, stub)
    e = object.__new__(edit.EditDeepFloydIF)
    e.unet = StandIn()
    e.scheduler = sched
    e.device, e.dtype = torch.device("cpu"), torch.float32
    e.guidance_scale, e.guidance_scale_edit = G, G_EDIT
    e.for_steps, e.use_yh_custom_scheduler = 100, True
    e.buffer_device, e.memory_bound = "cpu", 2      # batches of 2..memory_bound-1 hit chunk(0) in the reference (:1455)
    e.c_in, e.image_size = 3, R
    e.tilda_v_score_type = "null+(for-null)+(edit-null)"
    import tempfile
    e.result_folder = tempfile.mkdtemp()
    e.EXP_NAME = "golden"
    embs = [synthetic_prompt_embedding(p, NTOK, DIM) for p in ("a photo of a dog", "a dog with glasses", "")]
    e.for_prompt_emb, e.edit_prompt_emb, e.null_prompt_emb = embs
    sched.set_timesteps(100)
    t_idx = 60
    t = sched.timesteps[t_idx]
    g = torch.Generator().manual_seed(5)
    xt = torch.randn(1, 3, R, R, generator=g)
    x2 = torch.randn(2, 3, R, R, generator=g)
    mask = torch.zeros(3, R, R, dtype=torch.bool)
    mask[:, 12:20, 8:24] = True
    out = {"arch": arch, "dim": DIM, "ntok": NTOK, "prompts": ["a photo of a dog", "a dog with glasses", ""],
           "g": G, "g_edit": G_EDIT, "t_idx": t_idx, "t": t.clone(), "xt": xt, "x2": x2, "mask": mask,
           "timesteps": sched.timesteps.clone(), "alphas_cumprod": sched.alphas_cumprod.clone()}
    with torch.no_grad():
        out["cfg"] = {}
        for mode in ["null+(for-null)+(edit-null)", "null+(for-null)", "null+(edit-null)", "(for-edit)", "(for-null)",
                     "(edit-null)", "null+for+edit-proj[for](edit)"]:
            out["cfg"][mode] = e._classifer_free_guidance(x2, t, *embs, mode=mode, do_classifier_free_guidance=True)
        out["cfg_off"] = e._classifer_free_guidance(x2, t, *embs, mode="null+(for-null)", do_classifier_free_guidance=False)
        out["x0_masked"] = e.get_x0(xt, t, t_idx, *embs, mask=mask, mode="null+(for-null)")
        out["x0_flat"] = e.get_x0(xt, t, t_idx, *embs, mask=None, mode="null+(for-null)+(edit-null)", flatten=True)
    out["pullback"] = {}
    for mode in ["null+(for-null)", "null+(for-null)+(edit-null)"]:
        for n_iter in (1, 2):
            torch.manual_seed(7)
            u, s, vT = e.local_encoder_decoder_pullback_xt(xt, t, t_idx, *embs, pca_rank=2, chunk_size=5, min_iter=10 ** 6,
                                                           max_iter=n_iter, convergence_threshold=1e-3, mask=mask, mode=mode)
            out["pullback"][(mode, n_iter)] = {"u": u.clone(), "s": s.clone(), "vT": vT.clone()}
            print(mode, n_iter, s.tolist())
    out["delta_masked"] = e.get_delta_xt_via_grad(xt, t, t_idx, *embs, mask=mask, mode="null+(for-null)+(edit-null)").clone()
    out["delta_nomask"] = e.get_delta_xt_via_grad(xt, t, t_idx, *embs, mask=None, mode="null+(for-null)+(edit-null)").clone()
    out["v_modify"] = {}
    for mode in ["(for-edit)-direct", "(edit-null)-direct", "proj_null[for-null](edit-null)-direct"]:
        out["v_modify"][mode] = e.get_v_modify(xt, t, t_idx, *embs, mask=mask, mode=mode, jacobian=False).clone()
    out["v_modify_jacobian"] = e.get_v_modify(xt, t, t_idx, *embs, mask=mask, jacobian=True).clone()
    # last 9 DDIM steps under guidance (the reference writes a PNG: give torchvision's writer a no-op)
    edit.tvu.save_image = lambda *a, **k: None
    with torch.no_grad():
        out["ddpm_final_u8"] = e.DDPMforwardsteps(x2.clone(), t_start_idx=90, t_end_idx=-1, for_prompt_emb=embs[0],
                                                  edit_prompt_emb=embs[1], null_prompt_emb=embs[2], mode="null+(for-null)")
        xt_mid, t_mid, i_mid = e.DDPMforwardsteps(x2.clone(), t_start_idx=88, t_end_idx=92, for_prompt_emb=embs[0],
                                                  edit_prompt_emb=embs[1], null_prompt_emb=embs[2],
                                                  mode="null+(for-null)+(edit-null)")
        out["ddpm_mid"] = {"xt": xt_mid, "t": t_mid, "idx": i_mid}
    torch.save(out, os.path.join(HERE, "t2i_cross_tiny.pt" if cross else "t2i_tiny.pt"))
    print("written", {k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in out.items()})


if __name__ == "__main__" and "cross" in sys.argv[1:]:
    main(cross=True)
elif __name__ == "__main__":
    main()
