"""Entry point mirroring the reference's src/main.py (:12-103): same flags, same dispatch on the
`--run_*` booleans.  The unconditional path runs `EditUncondDiffusion`; DeepFloyd-IF model names run the
T-LOCO class `EditDeepFloydIF` of loco_edit_b200/t2i.py on a stand-in text-conditioned U-Net (cross-attention
to seeded prompt embeddings (the IF network and its T5 encoder are diffusers / transformers models that cannot
be obtained here); Stable Diffusion model names run the latent-space class
`EditStableDiffusion` of loco_edit_b200/sd.py on a stand-in latent U-Net and the SD-shaped VAE decoder (the
decoder Jacobian sits inside every Jacobian product); LCM model names run `EditLatentConsistency` (same file) the same way.

    python -m loco_edit_b200.main --model_name LSUN_church_HF --dataset_name LSUN_church --dtype fp32 \
        --edit_t 0.6 --performance_boosting_t 0.2 --pca_rank 5 --pca_rank_null 5 \
        --x_space_guidance_scale 0.5 --x_space_guidance_num_step 16 --vis_num 2 \
        --null_space_projection True --run_edit_null_space_projection True
"""
from .define_argparser import parse_args, preset
from .edit import EditUncondDiffusion


def main_deepfloyd(args):
    """src/main.py:28-32, 72-85 for `--model_name DeepFloyd/IF-I-*`: pixel-space T-LOCO on 64 x 64."""
    import torch
    from .t2i import EditDeepFloydIF, TextB200UNet, synthetic_prompt_embedding
    from .unet import B200UNet
    from .weights import if_standin_arch, random_state_dict
    print("DeepFloyd-IF: running the stand-in text-conditioned U-Net (cross-attention to seeded prompt "
          "embeddings; the IF checkpoint / T5 encoder are not available offline)")
    arch = if_standin_arch(args.image_size)
    net = TextB200UNet(B200UNet(arch, random_state_dict(arch, seed=1234), device=torch.device(args.device)))
    embs = [synthetic_prompt_embedding(p or "", 77, 768) for p in (args.for_prompt, args.edit_prompt, "")]
    edit = EditDeepFloydIF(args, net, *embs)
    common = dict(op='mid', block_idx=0, mask_index=args.mask_index, vis_num=args.vis_num, vis_num_pc=args.pca_rank,
                  pca_rank=args.pca_rank, edit_prompt=args.edit_prompt, null_space_projection=args.null_space_projection,
                  pca_rank_null=args.pca_rank_null)
    if args.run_edit_null_space_projection_xt:
        edit.run_edit_null_space_projection_xt(**common)
    if args.run_edit_null_space_projection_xt_semantic:
        edit.run_edit_null_space_projection_xt_semantic(jacobian=args.jacobian, **common)
    return edit


def main_stable_diffusion(args):
    """src/main.py:22-27, 52-71 for `--model_name *stable-diffusion*`: latent-space T-LOCO, z_t [4, 64, 64],
    x0_hat [3, 512, 512] through the VAE decoder."""
    import torch
    from .sd import EditStableDiffusion
    from .t2i import TextB200UNet, synthetic_prompt_embedding
    from .unet import B200UNet, B200VAEDecoder
    from .weights import SD_VAE_DECODER, random_state_dict, sd_standin_unet_arch
    print("Stable Diffusion: running the stand-in latent U-Net and a random-init VAE decoder of the SD 1.x shape "
          "(the checkpoints / CLIP text encoder are not available offline)")
    dev = torch.device(args.device)
    arch = sd_standin_unet_arch(args.image_size // 8)
    varch = dict(SD_VAE_DECODER, resolution=args.image_size // 8)
    net = TextB200UNet(B200UNet(arch, random_state_dict(arch, seed=1234), device=dev))
    vae = B200VAEDecoder(varch, random_state_dict(varch, seed=4321), device=dev)
    embs = [synthetic_prompt_embedding(p or "", 77, 768) for p in (args.for_prompt, args.edit_prompt, "")]
    edit = EditStableDiffusion(args, net, vae, *embs)
    common = dict(op='mid', block_idx=0, mask_index=args.mask_index, vis_num=args.vis_num, vis_num_pc=args.pca_rank,
                  pca_rank=args.pca_rank, edit_prompt=args.edit_prompt, null_space_projection=args.null_space_projection,
                  pca_rank_null=args.pca_rank_null)
    if args.run_edit_null_space_projection_zt:
        edit.run_edit_null_space_projection_zt(non_semantic=getattr(args, "non_semantic", False), **common)
    if args.run_edit_null_space_projection_zt_semantic:
        edit.run_edit_null_space_projection_zt_semantic(**common)
    return edit


def main_lcm(args):
    """src/main.py:30-32, 52-62 for `--model_name *LCM*`: the latent-consistency class (z_t [4, 64, 64], 512^2 images)."""
    import torch
    from .sd import EditLatentConsistency, LCMB200UNet
    from .unet import B200UNet, B200VAEDecoder
    from .weights import SD_VAE_DECODER, random_state_dict, sd_standin_unet_arch
    print("LCM: running the stand-in latent U-Net (w-conditioned) and a random-init VAE decoder of the SD 1.x shape "
          "(the checkpoints / CLIP text encoder are not available offline)")
    dev = torch.device(args.device)
    arch = sd_standin_unet_arch(args.image_size // 8)
    varch = dict(SD_VAE_DECODER, resolution=args.image_size // 8)
    net = LCMB200UNet(B200UNet(arch, random_state_dict(arch, seed=1234), device=dev))
    vae = B200VAEDecoder(varch, random_state_dict(varch, seed=4321), device=dev)
    edit = EditLatentConsistency(args, net, vae)
    if args.run_edit_null_space_projection_zt:
        edit.run_edit_null_space_projection_zt(
            op='mid', block_idx=0, mask_index=args.mask_index, vis_num=args.vis_num, vis_num_pc=args.pca_rank,
            pca_rank=args.pca_rank, edit_prompt=args.edit_prompt, null_space_projection=args.null_space_projection,
            pca_rank_null=args.pca_rank_null, non_semantic=getattr(args, "non_semantic", False))
    return edit


def main(argv=None):
    args = preset(parse_args(argv))
    if args.is_LCM:
        return main_lcm(args)
    if args.is_stable_diffusion:
        return main_stable_diffusion(args)
    if args.is_DeepFloyd_IF_diffusion:
        return main_deepfloyd(args)
    print('is custmized diffusion model')
    edit = EditUncondDiffusion(args)
    if args.run_edit_null_space_projection:
        edit.run_edit_null_space_projection(
            idx=args.sample_idx, op='mid', block_idx=0, vis_num=args.vis_num, vis_num_pc=args.pca_rank,
            pca_rank=args.pca_rank, edit_prompt=args.edit_prompt,
            null_space_projection=args.null_space_projection, pca_rank_null=args.pca_rank_null,
            encoder_decoder_by_et=args.encoder_decoder_by_et, use_mask=args.use_mask,
            random_edit=args.random_edit)
    for flag in ("run_edit_null_space_projection_zt", "run_edit_null_space_projection_zt_semantic",
                 "run_edit_null_space_projection_xt", "run_edit_null_space_projection_xt_semantic"):
        if getattr(args, flag):
            raise NotImplementedError(flag + " belongs to the T2I classes")
    if args.group_edit_null_space_projection:
        edit.group_edit_null_space_projection(
            idx=args.sample_idx, op='mid', block_idx=0, vis_num_pc=1, pca_rank=1, edit_prompt=args.edit_prompt,
            null_space_projection=args.null_space_projection, pca_rank_null=args.pca_rank_null,
            encoder_decoder_by_et=args.encoder_decoder_by_et)
    if args.run_ddim_forward:
        edit.run_DDIMforward(num_samples=5)
    if args.run_ddim_inversion:
        edit.run_DDIMinversion(idx=args.sample_idx)
    return edit


if __name__ == "__main__":
    main()
