"""Flag surface of the reference CLI (src/utils/define_argparser.py:14-258), kept flag-for-flag so
the reference's launch scripts (src/scripts/*.sh) drive this package unchanged.

Differences: `preset` does not copy the launching script into the run folder when it does not exist
(the reference requires cwd = src/ and --sh_file_name), and three optional flags are added:
--weights_path (state_dict with the reference's DDPM names; random-init synthetic weights
otherwise), --verbose, --align_sign.
"""
import argparse
import os
import random
import shutil

import numpy as np
import torch


def str2bool(v):
    """src/utils/define_argparser.py:128-136 (accepts any substring of 'true' / 'false')."""
    if isinstance(v, bool):
        return v
    if v.lower() in ('true'):
        return True
    elif v.lower() in ('false'):
        return False
    else:
        raise argparse.ArgumentTypeError('Boolean value expected.')


_STR = [("sh_file_name", ""), ("device", ""), ("dtype", "fp16"), ("result_folder", "./runs/"),
        ("cache_folder", ""), ("dataset_root", ""), ("model_name", ""), ("dataset_name", ""),
        ("for_prompt", ""), ("inv_prompt", ""), ("neg_prompt", ""), ("edit_prompt", ""),
        ("original_prompt", ""), ("edit_xt", "default"), ("pca_device", "cpu"), ("buffer_device", "cpu"),
        ("save_result_as", "image"), ("note", None), ("choose_sem", "hair"),
        ("mask_model_name", "facebook/sam-vit-large"), ("vT_path", ""), ("vT1_path", ""),
        ("weights_path", "")]
_INT = [("seed", 0), ("num_imgs", 100), ("image_size", 256), ("c_in", 3), ("sample_idx", 0),
        ("for_steps", 100), ("inv_steps", 100), ("x_space_guidance_num_step", 0), ("pca_rank_null", 5),
        ("pca_rank", 5), ("vis_num", 4), ("filter_mask", 100), ("mask_index", 0), ("edit_t_idx", 1),
        ("num_inference_steps", 3)]
_FLOAT = [("performance_boosting_t", 0.0), ("guidance_scale", 0), ("guidance_scale_edit", 4.0),
          ("x_space_guidance_edit_step", 1), ("x_space_guidance_scale", 0), ("h_t", 0.8), ("edit_t", 1.0),
          ("no_edit_t", 0.5), ("h_edit_step_size", 0), ("x_edit_step_size", 0)]
_BOOL = [("use_yh_custom_scheduler", True), ("use_x_space_guidance", False), ("x_space_guidance_direct", False),
         ("x_space_guidance_use_edit_prompt", True), ("run_cfg_forward", False), ("run_mcg_forward", False),
         ("run_pfg_forward", False), ("run_ddim_forward", False), ("run_ddim_inversion", False),
         ("run_edit_local_encoder_pullback_zt", False), ("run_edit_local_decoder_pullback_zt", False),
         ("run_edit_local_encoder_decoder_pullback_zt", False), ("encoder_decoder_by_et", False),
         ("use_mask", True), ("run_edit_local_x0_decoder_pullback_zt", False), ("run_edit_local_pca_zt", False),
         ("run_edit_null_space_projection", False), ("run_edit_null_space_projection_zt", False),
         ("run_edit_null_space_projection_zt_semantic", False), ("run_edit_null_space_projection_xt", False),
         ("run_edit_null_space_projection_xt_semantic", False), ("group_edit_null_space_projection", False),
         ("null_space_projection", False), ("debug_mode", False), ("sampling_mode", False),
         ("non_semantic", False), ("jacobian", False), ("use_sega", False), ("random_edit", False),
         ("verbose", True), ("align_sign", True)]


def build_parser():
    p = argparse.ArgumentParser()
    for n, d in _STR:
        p.add_argument("--" + n, type=str, default=d, required=False)
    for n, d in _INT:
        p.add_argument("--" + n, type=int, default=d, required=False)
    for n, d in _FLOAT:
        p.add_argument("--" + n, type=float, default=d, required=False)
    for n, d in _BOOL:
        p.add_argument("--" + n, type=str2bool, default=d, required=False)
    p.add_argument("--mask_type", type=str, default="SAM", choices=["SAM", "diffedit"])
    p.add_argument("--ablation_method", type=str, required=False, choices=["null-space-proj", "sega", "diffedit"])
    p.add_argument("--tilda_v_score_type", type=str, required=False)
    return p


def parse_args(argv=None):
    return build_parser().parse_args(argv)


def seed_everything(seed):
    """src/utils/define_argparser.py:251-258."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed % (2 ** 32))
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def preset(args):
    """src/utils/define_argparser.py:138-249: model-family detection, folders, derived fields, asserts."""
    if args.seed == 0:
        args.seed = int(torch.randint(2 ** 32, ()))
    seed_everything(args.seed)
    args.is_stable_diffusion = 'stable-diffusion' in args.model_name
    args.is_DeepFloyd_IF_diffusion = 'DeepFloyd' in args.model_name
    args.is_LCM = 'LCM' in args.model_name
    if args.is_stable_diffusion:
        args.exp = f'Stable_Diffusion-{args.dataset_name}-{args.note}'
    elif args.is_DeepFloyd_IF_diffusion:
        args.exp = f'DeepFloyd-IF-{args.dataset_name}-{args.note}'
    elif args.is_LCM:
        args.exp = f'LCM-{args.dataset_name}-{args.note}'
    else:
        if args.model_name == 'CelebA_HQ':
            raise NotImplementedError('Model weight deprecated...')
        elif args.model_name in ["FFHQ_P2", "AFHQ_P2", "Flower_P2", "Cub_P2", "Metface_P2",
                                 'CelebA_HQ_HF', 'LSUN_church_HF', 'LSUN_bedroom_HF', 'FFHQ_HF']:
            pass
        else:
            raise ValueError('model_name choice: [CelebA_HQ_HF, LSUN_church_HF, FFHQ_HF]')
        args.exp = f'{args.model_name}-{args.dataset_name}'
    args.exp_folder = os.path.join(args.result_folder, args.exp)
    os.makedirs(args.exp_folder, exist_ok=True)
    sh = os.path.join('scripts', args.sh_file_name)
    if args.sh_file_name and os.path.exists(sh):
        shutil.copy(sh, os.path.join(args.exp_folder, args.sh_file_name))
    args.obs_folder = os.path.join(args.exp_folder, 'obs')
    args.result_folder = os.path.join(args.exp_folder, 'results')
    os.makedirs(args.obs_folder, exist_ok=True)
    os.makedirs(args.result_folder, exist_ok=True)
    args.device = torch.device(args.device if args.device else "cuda:0")
    args.dtype = torch.float32 if args.dtype == 'fp32' else torch.float16
    if args.is_stable_diffusion:
        args.c_in, args.image_size, args.memory_bound = 4, 64, 5
    elif args.is_DeepFloyd_IF_diffusion:
        args.c_in, args.image_size, args.memory_bound = 3, 64, 5
    else:
        args.c_in, args.image_size, args.memory_bound = 3, 256, 50
        args.noise_schedule = 'linear'
    if args.is_stable_diffusion or args.is_DeepFloyd_IF_diffusion:
        assert args.use_yh_custom_scheduler
        assert args.performance_boosting_t <= 0
    elif args.is_LCM:
        pass
    else:
        assert args.use_yh_custom_scheduler
        assert args.for_steps == 100
        assert args.performance_boosting_t == 0.2
    return args
