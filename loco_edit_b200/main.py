"""Entry point mirroring the reference's src/main.py (:12-103): same flags, same dispatch on the
`--run_*` booleans.  Only the unconditional-diffusion path (EditUncondDiffusion) is implemented in
this round; the text-to-image classes (SD / DeepFloyd-IF / LCM) need U-Nets whose arithmetic lives
in un-vendored third-party code and raise NotImplementedError.

    python -m loco_edit_b200.main --model_name LSUN_church_HF --dataset_name LSUN_church --dtype fp32 \
        --edit_t 0.6 --performance_boosting_t 0.2 --pca_rank 5 --pca_rank_null 5 \
        --x_space_guidance_scale 0.5 --x_space_guidance_num_step 16 --vis_num 2 \
        --null_space_projection True --run_edit_null_space_projection True
"""
from .define_argparser import parse_args, preset
from .edit import EditUncondDiffusion


def main(argv=None):
    args = preset(parse_args(argv))
    if args.is_stable_diffusion or args.is_DeepFloyd_IF_diffusion or args.is_LCM:
        raise NotImplementedError("T2I editing classes are listed under 'next' in DESIGN.md (SURVEY 8f)")
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
