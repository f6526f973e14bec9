"""CPU: host-side logic of the drop-in layer (scheduler tables, timestep bookkeeping, edit-batch
construction, CLI surface) against the reference-generated fixtures."""
import os

import pytest
import torch

from loco_edit_b200.scheduler import YHCustomScheduler


def test_scheduler_tables_match_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "scheduler.pt"), weights_only=False)
    s = YHCustomScheduler(device="cpu")
    assert torch.equal(s.alphas_cumprod, g["alphas_cumprod"])
    s.set_timesteps(100)
    assert torch.equal(s.timesteps, g["timesteps"]) and torch.equal(s.timesteps_next, g["timesteps_next"])
    # edit_t = 0.6 -> idx 40, fractional t, alpha index floor(t)     
    idx = int((s.timesteps - 0.6 * 1000).abs().argmin())
    assert idx == 40 and abs(s._ts_host[40] - 595.3636) < 1e-3
    assert s.index_of(s.timesteps[40]) == 40
    assert s.alpha_at(s._ts_host[40]) == float(g["alphas_cumprod"][595])
    assert int((s.timesteps - 0.2 * 1000).abs().argmin()) == 79
    assert torch.equal(s.get_timesteps(s.timesteps[40]).reshape(()), s.timesteps_next[40])
    s.set_timesteps(100, is_inversion=True)
    assert torch.equal(s.timesteps, g["inv_timesteps"])
    assert torch.equal(s.timesteps_next, g["inv_timesteps_next"])
    assert s.alpha_at(s._ts_host[0]) == float(g["alphas_cumprod"][0])


def test_cli_surface_matches_reference_flags():
    """Every flag of the reference's parse_args (src/utils/define_argparser.py:18-124) is accepted."""
    from loco_edit_b200.define_argparser import build_parser, str2bool
    p = build_parser()
    flags = {a.dest for a in p._actions}
    expected = """sh_file_name device dtype seed result_folder cache_folder dataset_root model_name dataset_name
    num_imgs image_size c_in sample_idx for_prompt inv_prompt neg_prompt for_steps inv_steps
    performance_boosting_t use_yh_custom_scheduler guidance_scale guidance_scale_edit edit_prompt
    original_prompt edit_xt use_x_space_guidance x_space_guidance_direct x_space_guidance_edit_step
    x_space_guidance_scale x_space_guidance_num_step x_space_guidance_use_edit_prompt pca_rank_null pca_rank
    h_t edit_t no_edit_t h_edit_step_size x_edit_step_size pca_device buffer_device save_result_as note
    run_cfg_forward run_mcg_forward run_pfg_forward run_ddim_forward run_ddim_inversion
    run_edit_local_encoder_pullback_zt run_edit_local_decoder_pullback_zt
    run_edit_local_encoder_decoder_pullback_zt encoder_decoder_by_et use_mask
    run_edit_local_x0_decoder_pullback_zt run_edit_local_pca_zt run_edit_null_space_projection
    run_edit_null_space_projection_zt run_edit_null_space_projection_zt_semantic
    run_edit_null_space_projection_xt run_edit_null_space_projection_xt_semantic
    group_edit_null_space_projection vis_num choose_sem null_space_projection debug_mode sampling_mode
    non_semantic mask_model_name filter_mask mask_index mask_type ablation_method tilda_v_score_type vT_path
    vT1_path jacobian use_sega edit_t_idx num_inference_steps random_edit""".split()
    missing = [f for f in expected if f not in flags]
    assert not missing, missing
    # str2bool accepts any substring of 'true'/'false' like the reference (:128-136)
    assert str2bool("True") is True and str2bool("t") is True and str2bool("fal") is False
    a = p.parse_args(["--model_name", "LSUN_church_HF", "--dtype", "fp32", "--edit_t", "0.6",
                      "--performance_boosting_t", "0.2", "--pca_rank", "5", "--run_edit_null_space_projection", "True"])
    assert a.run_edit_null_space_projection is True and a.pca_rank == 5 and a.for_steps == 100


def test_edit_batch_layout_matches_reference_lines():
    """modules/edit.py:2341-2363 restated in the oracle == list/flip/concat logic used by the driver."""
    from oracle import pullback_ref
    xt = torch.arange(12.0).reshape(1, 3, 2, 2)
    v = torch.ones(12)
    for num_step, vis_num, n_out in [(16, 2, 5), (1, 2, 3), (3, 3, 7), (16, 1, 3)]:
        b = pullback_ref.edit_batch(xt, v, 0.5, num_step, vis_num)
        assert b.shape[0] == n_out
        mid = n_out // 2
        assert torch.equal(b[mid], xt[0])
        assert float((b[-1] - xt[0]).mean()) > 0 and float((b[0] - xt[0]).mean()) < 0


def test_preset_derives_the_reference_fields(tmp_path, monkeypatch):
    """preset() (src/utils/define_argparser.py:138-249): experiment folder names, derived fields and
    the asserts of the unconditional models; model-name dispatch of the two U-Net families."""
    from loco_edit_b200.define_argparser import parse_args, preset
    monkeypatch.chdir(tmp_path)
    base = ["--dtype", "fp32", "--device", "cpu", "--result_folder", str(tmp_path / "runs"), "--seed", "3",
            "--use_yh_custom_scheduler", "True", "--performance_boosting_t", "0.2", "--for_steps", "100"]
    a = preset(parse_args(base + ["--model_name", "FFHQ_P2", "--dataset_name", "FFHQ"]))
    assert a.exp == "FFHQ_P2-FFHQ" and a.result_folder.endswith(os.path.join("FFHQ_P2-FFHQ", "results"))
    assert os.path.isdir(a.result_folder) and os.path.isdir(a.obs_folder)
    assert (a.c_in, a.image_size, a.memory_bound) == (3, 256, 50) and a.dtype == torch.float32
    assert not (a.is_stable_diffusion or a.is_DeepFloyd_IF_diffusion or a.is_LCM)
    b = preset(parse_args(base + ["--model_name", "LSUN_church_HF", "--dataset_name", "LSUN_church"]))
    assert b.exp == "LSUN_church_HF-LSUN_church"
    # seed 0 means "draw a seed" (:140-141)
    c = preset(parse_args([x if x != "3" else "0" for x in base] + ["--model_name", "CelebA_HQ_HF",
                                                                    "--dataset_name", "CelebA_HQ_mask"]))
    assert c.seed != 0
    # unconditional models assert for_steps == 100 and performance_boosting_t == 0.2 (:244-247)
    bad = [x for x in base]
    bad[bad.index("--for_steps") + 1] = "50"
    with pytest.raises(AssertionError):
        preset(parse_args(bad + ["--model_name", "FFHQ_P2", "--dataset_name", "FFHQ"]))
    with pytest.raises(ValueError):
        preset(parse_args(base + ["--model_name", "NotAModel", "--dataset_name", "x"]))
    # the Edit class picks the network family by model name (reference: utils/utils.py:101-131)
    from loco_edit_b200.weights import P2_256, DDPM256, param_shapes
    assert "input_blocks.0.0.weight" in param_shapes(P2_256) and "conv_in.weight" in param_shapes(DDPM256)


def _ddpm_to_hf_names(arch, legacy_attn):
    """Inverse name table (test-side): DDPM.state_dict() name -> diffusers UNet2DModel name."""
    from loco_edit_b200.weights import ddpm_param_shapes
    L = len(arch["ch_mult"])
    res = {"norm1": "norm1", "conv1": "conv1", "temb_proj": "time_emb_proj", "norm2": "norm2", "conv2": "conv2",
           "nin_shortcut": "conv_shortcut"}
    att = ({"norm": "group_norm", "q": "query", "k": "key", "v": "value", "proj_out": "proj_attn"} if legacy_attn
           else {"norm": "group_norm", "q": "to_q", "k": "to_k", "v": "to_v", "proj_out": "to_out.0"})
    out = {}
    for name, shp in ddpm_param_shapes(arch).items():
        p = name.split(".")
        leaf = p[-1]
        if p[0] == "temb":
            hf = "time_embedding.linear_%d.%s" % (int(p[2]) + 1, leaf)
        elif p[0] in ("conv_in", "conv_out"):
            hf = name
        elif p[0] == "norm_out":
            hf = "conv_norm_out." + leaf
        elif p[0] == "mid":
            hf = ("mid_block.resnets.%d.%s.%s" % (int(p[1][-1]) - 1, res[p[2]], leaf) if p[1].startswith("block")
                  else "mid_block.attentions.0.%s.%s" % (att[p[2]], leaf))
        else:
            blk = "down_blocks.%s" % p[1] if p[0] == "down" else "up_blocks.%d" % (L - 1 - int(p[1]))
            if p[2] == "block":
                hf = "%s.resnets.%s.%s.%s" % (blk, p[3], res[p[4]], leaf)
            elif p[2] == "attn":
                hf = "%s.attentions.%s.%s.%s" % (blk, p[3], att[p[4]], leaf)
            else:
                hf = "%s.%ssamplers.0.conv.%s" % (blk, "down" if p[2] == "downsample" else "up", leaf)
        is_attn_proj = ".attentions." in hf and "group_norm" not in hf and leaf == "weight"
        out[name] = (hf, tuple(shp[:2]) if is_attn_proj else tuple(shp))
    return out


@pytest.mark.parametrize("legacy_attn", [True, False])
def test_hf_unet2d_checkpoint_names_map_onto_the_ddpm_state_dict(legacy_attn):
    """`*_HF` models are diffusers UNet2DModel checkpoints (src/utils/utils.py:93-98, 122-125); the
    remapper must produce exactly DDPM.state_dict() (names, shapes; Linear q/k/v -> 1x1 conv) for
    both the diffusers-0.11 and the current attention naming, and refuse anything else."""
    from loco_edit_b200.weights import DDPM256, hf_unet2d_to_ddpm, is_hf_unet2d_state_dict, random_state_dict, tiny_arch
    for arch in (tiny_arch(resolution=32, ch_mult=(1, 2), attn_resolutions=(16,), num_res_blocks=1), DDPM256):
        names = _ddpm_to_hf_names(arch, legacy_attn)
        if arch is DDPM256:      # names/shapes only for the 113 M-parameter model
            with torch.device("meta"):
                hf = {h: torch.empty(shp) for h, shp in names.values()}
            got = hf_unet2d_to_ddpm(hf, arch)
            assert len(got) == len(names)
            continue
        sd = random_state_dict(arch, seed=5)
        hf = {h: sd[n].reshape(shp).clone() for n, (h, shp) in names.items()}
        assert is_hf_unet2d_state_dict(hf) and not is_hf_unet2d_state_dict(sd)
        got = hf_unet2d_to_ddpm(hf, arch)
        assert list(sorted(got)) == list(sorted(sd))
        for n in sd:
            assert torch.equal(got[n], sd[n]), n
        bad = dict(hf)
        bad["down_blocks.0.resnets.0.mystery.weight"] = torch.zeros(1)
        with pytest.raises(KeyError):
            hf_unet2d_to_ddpm(bad, arch)
        short = dict(hf)
        short.pop("conv_out.bias")
        with pytest.raises(KeyError):
            hf_unet2d_to_ddpm(short, arch)


@pytest.mark.parametrize("legacy_attn", [True, False])
def test_hf_autoencoderkl_checkpoint_names_map_onto_the_decoder_state_dict(legacy_attn):
    """The `vae` of a Stable Diffusion pipeline is a diffusers AutoencoderKL (src/utils/utils.py:217); its decoder
    half must map onto exactly the names / shapes B200VAEDecoder takes (Linear q/k/v -> 1x1 conv, up_blocks in
    execution order -> levels), the encoder half is ignored."""
    import re
    from loco_edit_b200.weights import (SD_VAE_DECODER, hf_autoencoderkl_to_decoder, is_hf_autoencoderkl_state_dict,
                                        random_state_dict, tiny_vae_decoder_arch, vae_decoder_param_shapes)
    res = {"norm1": "norm1", "conv1": "conv1", "norm2": "norm2", "conv2": "conv2", "nin_shortcut": "conv_shortcut"}
    attn = ({"norm": "group_norm", "q": "query", "k": "key", "v": "value", "proj_out": "proj_attn"} if legacy_attn else
            {"norm": "group_norm", "q": "to_q", "k": "to_k", "v": "to_v", "proj_out": "to_out.0"})

    def to_hf(arch):
        L = len(arch["ch_mult"])
        names = {}
        for n, shp in vae_decoder_param_shapes(arch).items():
            leaf = n.rsplit(".", 1)[1]
            m = re.match(r"decoder\.mid\.block_(\d)\.(\w+)\.", n)
            if m:
                h = "decoder.mid_block.resnets.%d.%s.%s" % (int(m.group(1)) - 1, res[m.group(2)], leaf)
            elif n.startswith("decoder.mid.attn_1."):
                sub = n.split(".")[3]
                h = "decoder.mid_block.attentions.0.%s.%s" % (attn[sub], leaf)
                if sub != "norm" and leaf == "weight":
                    shp = shp[:2]
            elif re.match(r"decoder\.up\.(\d+)\.block\.", n):
                m = re.match(r"decoder\.up\.(\d+)\.block\.(\d+)\.(\w+)\.", n)
                h = "decoder.up_blocks.%d.resnets.%s.%s.%s" % (L - 1 - int(m.group(1)), m.group(2), res[m.group(3)], leaf)
            elif ".upsample.conv." in n:
                h = "decoder.up_blocks.%d.upsamplers.0.conv.%s" % (L - 1 - int(n.split(".")[2]), leaf)
            elif n.startswith("decoder.norm_out."):
                h = "decoder.conv_norm_out." + leaf
            else:
                h = n
            names[n] = (h, shp)
        return names

    for arch in (tiny_vae_decoder_arch(), SD_VAE_DECODER):
        names = to_hf(arch)
        if arch is SD_VAE_DECODER:       # names / shapes only
            with torch.device("meta"):
                hf = {h: torch.empty(shp) for h, shp in names.values()}
                hf["encoder.conv_in.weight"] = torch.empty(128, 3, 3, 3)
                hf["quant_conv.weight"] = torch.empty(8, 8, 1, 1)
            assert len(hf_autoencoderkl_to_decoder(hf, arch)) == len(names)
            continue
        sd = random_state_dict(arch, seed=5)
        hf = {h: sd[n].reshape(shp).clone() for n, (h, shp) in names.items()}
        assert is_hf_autoencoderkl_state_dict(hf) and not is_hf_autoencoderkl_state_dict(sd)
        got = hf_autoencoderkl_to_decoder(hf, arch)
        assert sorted(got) == sorted(sd)
        for n in sd:
            assert torch.equal(got[n], sd[n]), n
        bad = dict(hf)
        bad["decoder.up_blocks.0.mystery.weight"] = torch.zeros(1)
        with pytest.raises(KeyError):
            hf_autoencoderkl_to_decoder(bad, arch)
        short = dict(hf)
        short.pop("decoder.conv_out.bias")
        with pytest.raises(KeyError):
            hf_autoencoderkl_to_decoder(short, arch)


def test_mask_files_and_postprocessing(tmp_path):
    """mask/mask.pt format of the SAM wrapper (src/modules/mask_segmentation.py:18-26), the row
    selection of the drivers (src/modules/edit.py:2247) and the DiffEdit formula (:1401-1402)."""
    import torch.nn.functional as F
    from loco_edit_b200 import masks as M
    g = torch.Generator().manual_seed(2)
    raw = torch.rand(3, 37, 53, generator=g) > 0.5
    want = torch.round(F.interpolate(raw.unsqueeze(dim=1).to(torch.float32), [16, 16]).squeeze(dim=1)).to(torch.bool)
    got = M.resize_masks(raw, 16)
    assert got.dtype == torch.bool and torch.equal(got, want)
    path = M.save_masks(str(tmp_path), got)
    assert path.endswith(os.path.join("mask", "mask.pt")) and torch.load(path).shape == (3, 16, 16)
    m1 = M.load_mask(str(tmp_path), 1)
    assert m1.shape == (3, 16, 16) and torch.equal(m1[0], got[1]) and torch.equal(m1[2], got[1])
    e1, e2 = torch.randn(10, 3, 8, 8, generator=g), torch.randn(10, 3, 8, 8, generator=g)
    mask = (e1 - e2).mean(dim=0, keepdim=True).mean(dim=1)
    ref = torch.round((mask - mask.min() / (mask.max() - mask.min()))).to(torch.bool)      # reference line, verbatim
    assert torch.equal(M.diffedit_mask(e1, e2), ref)
    assert int(M.rectangle_mask(256).sum()) == 3 * 64 * 128


def test_driver_refuses_to_invent_a_mask(tmp_path):
    """A non-CelebA dataset without mask/mask.pt is an error, `use_mask=False` means no
    mask (src/modules/edit.py:2266-2267) -- checked on the host logic only (no GPU needed)."""
    import types
    from loco_edit_b200.edit import EditUncondDiffusion
    e = object.__new__(EditUncondDiffusion)
    e.dataset_name, e.result_folder = "FFHQ", str(tmp_path)
    e.args = types.SimpleNamespace(mask_index=0)
    e.dataset = None
    assert e._get_masks(0, use_mask=False) is None
    with pytest.raises(FileNotFoundError):
        e._get_masks(0, use_mask=True)
    from loco_edit_b200.masks import save_masks
    save_masks(str(tmp_path), torch.ones(2, 8, 8, dtype=torch.bool))
    assert e._get_masks(0, use_mask=True).shape == (3, 8, 8)


def test_lcm_scheduler_restatement(golden_dir):
    """`scheduler.LCMScheduler`: the multistep grid of diffusers' LCMScheduler (50 original steps) and the
    boundary-condition coefficients denoised = c1 x + c2 eps, against the CPU restatement the LCM golden was
    generated with (tests/golden/make_golden_lcm.py)."""
    import os
    from loco_edit_b200.scheduler import LCMScheduler, scaled_linear_betas
    g = torch.load(os.path.join(golden_dir, "lcm_tiny.pt"), weights_only=False)
    s = LCMScheduler("cpu")
    s.set_timesteps(g["steps"], device="cpu")
    assert torch.equal(s.timesteps, g["timesteps"]) and s._ts_host == [999, 759, 519, 279]
    s.set_timesteps(8, device="cpu")
    assert s._ts_host == [999, 879, 759, 639, 519, 399, 279, 159]
    acp = torch.cumprod(1.0 - scaled_linear_betas(1000), 0)
    for t in (999, 759, 279):
        a = float(acp[t])
        st = 10.0 * t
        c_skip, c_out = 0.25 / (st * st + 0.25), st / (st * st + 0.25) ** 0.5
        c1, c2 = s.denoised_coefficients(t)
        assert abs(c1 - (c_out / a ** 0.5 + c_skip)) < 1e-12 and abs(c2 + c_out * (1 - a) ** 0.5 / a ** 0.5) < 1e-12


def test_committed_bench_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/ (one B200 run each of `python bench.py` and
    `python bench.py --impl reference`) carry every key of the measurement contract."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ours = json.loads(open(os.path.join(root, "profiles", "r2_bench_final4.json")).read().strip().splitlines()[-1])
    ref = json.loads(open(os.path.join(root, "profiles", "r2_bench_ref_final4.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in ours, k
    assert ours["warmup"] >= 3 and ours["gpu_launches"] > 0 and ours["vs_baseline"] is None
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(ours["roofline"])
    assert abs(ours["roofline"]["frac"] - ours["roofline"]["achieved"] / ours["roofline"]["peak"]) < 1e-9
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(ours["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(ours["e2e"])
    assert ours["e2e"]["h2d_bytes_per_step"] > 0 and ours["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(ours["clocks"])
    assert "workload" in ours["config"] and "model" not in ours["config"]
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["config"]["workload"] == ours["config"]["workload"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0
    assert ref["cpu_baseline"]["kind"] in ("port", "reference") and ref["cpu_baseline"]["value"] == ref["value"]
