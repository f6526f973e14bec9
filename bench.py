#!/usr/bin/env python
"""bench.py -- LOCO-Edit editing-direction hot path on B200 (contract: see the task statement).

One "edit" = the BASELINE config-1 pipeline for one synthetic 256x256 image on the DDPM-256 U-Net
(random-init, seed 1234): DDIM inversion (98 U-Net calls) + forward to t=0.6T (40) + rank-5 local
basis of the masked PMP Jacobian (12 power iterations) + rank-5 null basis (12) + null-space
projection + 59 DDIM steps on the 5-latent edit batch (= 697 U-Net-forward equivalents, 346 TFLOP).
One step = one batch of --batch image/mask pairs edited together on each GPU (BASELINE config 2:
64 pairs over 8 GPUs = 8 per GPU; --batch 1 = config 1).  Each rank edits its own images (weak
scaling, no data-path collective).

  python bench.py --gpus 1 --steps 3 --warmup 3          # our CUDA path
  python bench.py --impl reference ...                    # reference algorithm on the host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_DDPM256 = 0.4970e12          # algorithmic FLOPs of one U-Net forward, B=1, 256^2
K_RANK, K_NULL, N_ITER = 5, 5, 12
EDIT_FWD_EQUIV = 138 + 2 * N_ITER * (1 + 2 * K_RANK) + 59 * 5    # 697: BASELINE.md's definition of an edit
# executed here: both bases share one primal row per iteration -> 138 + 12*(1 + 2*10) + 295 = 685
EDIT_FWD_EXECUTED = 138 + N_ITER * (1 + 2 * (K_RANK + K_NULL)) + 59 * 5


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_inputs(R, idx):
    import torch
    g = torch.Generator().manual_seed(int(idx))
    x0 = (0.5 * torch.randn(1, 3, R, R, generator=g)).clamp(-1, 1)
    mask = torch.zeros(3, R, R, dtype=torch.bool)
    mask[:, (3 * R) // 8:(5 * R) // 8, R // 4:(3 * R) // 4] = True
    return x0, mask


# ------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (CPU oracle port of its PyTorch path) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(threads):
    """Bounded sample of the config-1 edit on the host cores (CPU oracle port of the reference's
    PyTorch path, fp32): ONE full rank-5 power iteration at 256x256 (5 forward-mode products + a fresh
    forward + 5 reverse passes = the reference's (1+3k) F, src/modules/edit.py:2443-2494), one U-Net
    forward at B = 1 and one at B = 5 (the batch of the final DDIM stage).  The edit is assembled from
    those measured pieces with the reference's own counts: 138 forwards (B=1) + 2 bases x 12
    iterations + 59 steps at B=5."""
    import torch
    from loco_edit_b200.weights import DDPM256, random_state_dict
    from oracle import ddpm_ref, pullback_ref
    torch.set_num_threads(threads)
    sd = random_state_dict(DDPM256, seed=1234)
    unet = ddpm_ref.RefUNet(DDPM256, sd)
    sched = pullback_ref.RefScheduler()
    sched.set_timesteps(100)
    x0, mask = synthetic_inputs(256, 0)
    t = sched.timesteps[40]
    g = torch.Generator().manual_seed(7)
    v0, _ = torch.linalg.qr(torch.randn(x0.numel(), K_RANK, generator=g))
    with torch.no_grad():
        unet(x0, t)                                   # warm-up (thread pool, allocator)
        t0 = time.perf_counter()
        unet(x0, t)
        f1 = time.perf_counter() - t0
        x5 = x0.repeat(5, 1, 1, 1)
        t0 = time.perf_counter()
        unet(x5, t)
        f5 = time.perf_counter() - t0
    t0 = time.perf_counter()
    pullback_ref.power_iteration(unet, sched, x0, t, v0.T.contiguous(), mask=mask)
    it5 = time.perf_counter() - t0
    edit_s = 138 * f1 + 2 * N_ITER * it5 + 59 * f5
    return {"forward_b1_s": f1, "forward_b5_s": f5, "rank5_iter_s": it5, "edit_s_assembled": edit_s,
            "jvp_probes_per_s": K_RANK / it5}


CPU_SAMPLE_TEXT = ("measured on the host cores, fp32 torch CPU, 256^2: one FULL rank-5 power iteration (5 JVP + "
                   "forward + 5 VJP), one U-Net forward at B=1, one at B=5; edit = 138 x fwd(B=1) + 24 x "
                   "iteration + 59 x fwd(B=5)")


def workload_config(batch):
    """The `config` keys both arms share (the reference arm runs the same workload on the host cores)."""
    return {"workload": "batch edit (BASELINE config 2 shape): DDPM-256 (ddpm-ema-celebahq-256 arch, "
                        "random init), %d image/mask pairs of 256x256 per step per GPU, each a full "
                        "config-1 edit: rank 5 + null rank 5, N=12 power iterations, t=0.6T, "
                        "98+40+59 DDIM steps, 5 edited latents per image" % batch,
            "images_per_step_per_gpu": batch,
            "fwd_equivalents_per_edit": EDIT_FWD_EQUIV,
            "tflop_per_edit": EDIT_FWD_EQUIV * F_DDPM256 / 1e12}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(args.warmup if args.warmup < 1 else 0):
        pass
    vals = []
    t_all = time.perf_counter()
    steps = max(1, min(args.steps, 2))               # each step is ~tens of seconds of CPU work
    for _ in range(steps):
        vals.append(cpu_sample(threads))
    wall = time.perf_counter() - t_all
    edit_s = statistics.mean(v["edit_s_assembled"] for v in vals)
    value = 1.0 / edit_s
    sample = "per step: " + CPU_SAMPLE_TEXT
    line = {
        "impl": "reference", "metric": "edits/sec (rank-5 @256^2, t=0.6T)", "value": value,
        "unit": "edits/s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
        "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.batch),
                       arithmetic="fp32 torch CPU (the reference's --dtype fp32), all host threads",
                       timing="host wall clock; edits are independent and the host runs them one after the other, so "
                              "edits/s of the batch = 1 / (seconds per edit); the edit is assembled from the measured "
                              "pieces of a bounded sample (cpu_baseline.sample)"),
        "cpu_baseline": {"value": value, "unit": "edits/s", "cores": threads, "kind": "port",
                         "sample": sample,
                         "forward_b1_s": statistics.mean(v["forward_b1_s"] for v in vals),
                         "forward_b5_s": statistics.mean(v["forward_b5_s"] for v in vals),
                         "rank5_iter_s": statistics.mean(v["rank5_iter_s"] for v in vals),
                         "jvp_probes_per_s": statistics.mean(v["jvp_probes_per_s"] for v in vals)},
        "e2e": {"value": value, "unit": "edits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm: extra measurements (all CUDA events on torch's current stream = the launching stream)
# ------------------------------------------------------------------------------------------------
def _timed(fn, reps, sync):
    import torch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for _ in range(reps):
        sync()
        a.record()
        fn()
        b.record()
        sync()
        out.append(a.elapsed_time(b))
    return out


def measure_latency_b1(pipe, gen, R):
    """north_star target: ONE config-1 edit (rank 5 + null 5, N = 12, 98+40+59 DDIM steps, 5 edited
    latents) through the public call `EditPipeline.edit` with pinned host buffers in and images back
    on the host: median of 3 after one warm-up call."""
    import torch
    x0, m = synthetic_inputs(R, 777)
    x0, m = x0.pin_memory(), m.pin_memory()
    gen.manual_seed(6000)
    pipe.edit(x0, m, gen=gen)
    ts = _timed(lambda: pipe.edit(x0, m, gen=gen), 3, torch.cuda.synchronize)
    return statistics.median(ts)


def measure_dropin(unet, dev, R):
    """The reference-facing driver itself: `EditUncondDiffusion.run_edit_null_space_projection` with the
    reference's own settings (the two power methods with min_iter=10 / max_iter=50: with
    random-init weights neither converges, so 50 iterations each; basis files written with
    torch.save), vis_num 2, one direction.  Second call of two, fresh result folder each."""
    import shutil
    import tempfile
    import types
    import torch
    from loco_edit_b200.edit import EditUncondDiffusion
    ms = []
    for rep in range(2):
        tmp = tempfile.mkdtemp(prefix="loco_dropin_")
        a = types.SimpleNamespace(
            device=dev, dtype=torch.float32, seed=1, model_name="CelebA_HQ_HF", dataset_name="CelebA_HQ_mask",
            image_size=R, for_steps=100, inv_steps=100, edit_t=0.6, performance_boosting_t=0.2,
            x_space_guidance_edit_step=1.0, x_space_guidance_scale=0.5, x_space_guidance_num_step=16,
            result_folder=tmp, sample_idx=0, choose_sem="hair", mask_index=0, sampling_mode=False,
            vT_path="", vT1_path="", verbose=False, save_images=False, noise_schedule=None)
        e = EditUncondDiffusion(a, unet=unet)
        ms.append(_timed(lambda: e.run_edit_null_space_projection(idx=0, vis_num=2, vis_num_pc=1, pca_rank=5,
                                                                  pca_rank_null=5), 1, torch.cuda.synchronize)[0])
        shutil.rmtree(tmp, ignore_errors=True)
    return {"ms": ms[-1], "power_iterations": "50 iterations of the edit and the null power method, advanced in one fused "
                                              "pass per iteration while both still iterate (min_iter=10, max_iter=50; "
                                              "neither converges on random weights)",
            "fwd_equivalents_executed": 138 + 50 * (1 + 2 * (K_RANK + K_NULL)) + 59 * 5,
            "entry_point": "EditUncondDiffusion.run_edit_null_space_projection (files written)"}


def measure_bandwidth_kernels(dev, hbm_peak):
    """north_star (2): achieved HBM GB/s of the bandwidth-bound pieces around the U-Net, each on its
    config-1 shapes (k = k_null = 5, d = 196608, l_o = 24576) and on the batch-edit shape of the DDIM
    update.  Working sets rotate through > 256 MB of distinct buffers so that no call is served by the
    126 MB L2; bytes are the algorithmic ones."""
    import torch
    from loco_edit_b200 import ops
    d, k = 3 * 256 * 256, 5
    nset = 48                                      # 48 x [5, d] fp32 = 189 MB per operand family
    g = torch.Generator(device=dev).manual_seed(9)
    W = [torch.randn(k, d, device=dev, generator=g) for _ in range(nset)]
    Vn = [ops.orthonormalise(torch.randn(k, d, device=dev, generator=g))[0] for _ in range(nset)]
    mask = torch.zeros(3, 256, 256, dtype=torch.bool, device=dev)
    mask[:, 96:160, 64:192] = True
    idx = ops.mask_indices(mask)
    xb = [torch.randn(40, 3, 256, 256, device=dev, generator=g) for _ in range(4)]      # 4 x 31 MB x 3 operands
    eb = [torch.randn(40, 3, 256, 256, device=dev, generator=g) for _ in range(4)]
    nb = [torch.randn(40, 3, 256, 256, device=dev, generator=g) for _ in range(4)]
    out = {}

    def run(name, fn, nbytes, reps):
        for i in range(3):
            fn(i)
        t = _timed(lambda: [fn(i) for i in range(reps)], 3, torch.cuda.synchronize)
        us = 1e3 * min(t) / reps
        out[name] = {"us": us, "bytes": nbytes, "gbs": nbytes / (us * 1e-6) / 1e9,
                     "frac_of_hbm_peak": nbytes / (us * 1e-6) / 1e9 / hbm_peak}

    # algorithmic bytes as in SURVEY 8(d): orthonormalise 3*k*d*4 (read W for the Gram matrix, read W and
    # write V for the transform; the implementation's second re-orthonormalisation pass moves twice
    # that), projection (2k + k_null)*d*4, mask gather (d + l_o)*4 per row
    run("orthonormalise_k5", lambda i: ops.orthonormalise(W[i % nset]), 3 * k * d * 4, nset)
    run("nullspace_project_k5", lambda i: ops.nullspace_project(W[i % nset], Vn[i % nset]), (2 * k + k) * d * 4, nset)
    run("gather_rows_mask", lambda i: ops.gather_rows(W[i % nset], idx), k * (d + idx.numel()) * 4, nset)
    run("ddim_step_b40_eta0", lambda i: ops.ddim_step(xb[i % 4], eb[i % 4], 0.5, 0.6), 3 * xb[0].numel() * 4, 8)
    run("ddim_step_b40_eta1", lambda i: ops.ddim_step(xb[i % 4], eb[i % 4], 0.5, 0.6, eta=1.0, noise=nb[i % 4]),
        4 * xb[0].numel() * 4, 8)
    run("axpy_b40", lambda i: ops.axpy(xb[i % 4], eb[i % 4], 0.5), 3 * xb[0].numel() * 4, 8)
    # BASELINE config 3 (rank 64): the replicated step of the probe-sharded power method
    del W, Vn
    W64 = [torch.randn(64, d, device=dev, generator=g) for _ in range(4)]                # 4 x 50 MB
    run("orthonormalise_k64", lambda i: ops.orthonormalise(W64[i % 4]), 3 * 64 * d * 4, 4)
    return out


def measure_text_unet(dev, R):
    """SURVEY 8(f1), BASELINE configs 4-5 stand-in: the DDPM-256 U-Net with a cross-attention sub-block
    (8 heads, 77 x 768 prompt embedding) in each of its six AttnBlocks.  Device time of the fused rank-5
    primal + tangent pass, of the rank-5 cotangent pass and of a B = 1 forward, next to which the
    unconditional numbers of the main line show what the cross-attention layers cost."""
    import torch
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import DDPM256_TEXT, random_state_dict
    net = B200UNet(DDPM256_TEXT, random_state_dict(DDPM256_TEXT, seed=1234), device=dev)
    g = torch.Generator(device=dev).manual_seed(21)
    ctx = torch.randn(77, 768, device=dev, generator=g)
    pj, p1 = net.plan(1, K_RANK, K_RANK), net.plan(1)
    pj.set_context(ctx); p1.set_context(ctx)
    xin = torch.randn(1 + K_RANK, 3, R, R, device=dev, generator=g)
    gin = torch.randn(K_RANK, 3, R, R, device=dev, generator=g)
    for _ in range(2):
        pj.forward(xin, 595.3636); pj.vjp(gin); p1.forward(xin[:1].contiguous(), 595.3636)
    torch.cuda.synchronize()
    a, b, c, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    reps = 5
    a.record()
    for _ in range(reps):
        pj.forward(xin, 595.3636)
    b.record()
    for _ in range(reps):
        pj.vjp(gin)
    c.record()
    for _ in range(reps):
        p1.forward(xin[:1].contiguous(), 595.3636)
    d.record()
    torch.cuda.synchronize()
    out = {"arch": "DDPM-256 + cross-attention to a 77 x 768 prompt embedding in 6 AttnBlocks (8 heads), random init",
           "jvp_pass_ms": a.elapsed_time(b) / reps, "vjp_pass_ms": b.elapsed_time(c) / reps,
           "fwd_b1_ms": c.elapsed_time(d) / reps,
           "jvp_probes_per_s": K_RANK * reps / (a.elapsed_time(b) * 1e-3)}
    net.release_plans()
    del net
    torch.cuda.empty_cache()
    return out


def measure_sd_latent(dev):
    """SURVEY 8(f2), BASELINE config 4 shape: the latent-space twin.  The VAE decoder of the Stable Diffusion
    1.x shape (latent [4, 64, 64] -> image [3, 512, 512], 49.5 M parameters, random init) inside every Jacobian
    product: device time of its fused rank-5 primal + tangent pass, of the rank-5 cotangent pass and of a
    B = 1 decode; then one power iteration of `EditStableDiffusion.local_encoder_decoder_pullback_zt`
    (stand-in latent U-Net under two-term guidance + PMP + decoder) per probe."""
    import types
    import torch
    from loco_edit_b200.masks import rectangle_mask
    from loco_edit_b200.sd import EditStableDiffusion
    from loco_edit_b200.t2i import TextB200UNet, synthetic_prompt_embedding
    from loco_edit_b200.unet import B200UNet, B200VAEDecoder
    from loco_edit_b200.weights import SD_VAE_DECODER, random_state_dict, sd_standin_unet_arch
    vae = B200VAEDecoder(SD_VAE_DECODER, random_state_dict(SD_VAE_DECODER, seed=4321), device=dev)
    g = torch.Generator(device=dev).manual_seed(22)
    pj, p1 = vae.plan(1, K_RANK, K_RANK), vae.plan(1)
    zin = torch.randn(1 + K_RANK, 4, 64, 64, device=dev, generator=g)
    gin = torch.randn(K_RANK, 3, 512, 512, device=dev, generator=g)
    for _ in range(2):
        pj.forward(zin, 0.0); pj.vjp(gin); p1.forward(zin[:1].contiguous(), 0.0)
    torch.cuda.synchronize()
    a, b, c, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    reps = 5
    a.record()
    for _ in range(reps):
        pj.forward(zin, 0.0)
    b.record()
    for _ in range(reps):
        pj.vjp(gin)
    c.record()
    for _ in range(reps):
        p1.forward(zin[:1].contiguous(), 0.0)
    d.record()
    torch.cuda.synchronize()
    jvp_ms, vjp_ms = a.elapsed_time(b) / reps, b.elapsed_time(c) / reps
    out = {"decoder": "SD 1.x VAE decoder shape (ch 128, mult 1-2-4-4, 3 ResnetBlocks per level, mid attention over "
                      "4096 tokens), latent 4x64x64 -> image 3x512x512, random init",
           "decoder_tflop_per_row": pj.fwd_flops / (1 + K_RANK) / 1e12,
           "decoder_jvp_pass_ms": jvp_ms, "decoder_vjp_pass_ms": vjp_ms, "decode_b1_ms": c.elapsed_time(d) / reps,
           "decoder_jvp_tflops": pj.fwd_flops / (jvp_ms * 1e-3) / 1e12,
           "decoder_vjp_tflops": pj.vjp_flops / (vjp_ms * 1e-3) / 1e12}
    # one latent-space power iteration through both networks (rank 5, mask over the decoded image)
    arch = sd_standin_unet_arch(64)
    net = TextB200UNet(B200UNet(arch, random_state_dict(arch, seed=1234), device=dev))
    embs = [synthetic_prompt_embedding(p, 77, 768) for p in ("a photo of a dog", "a dog with glasses", "")]
    import tempfile
    args = types.SimpleNamespace(device=dev, dtype=torch.float32, seed=3, for_steps=100, edit_t=0.6, guidance_scale=7.5,
                                 guidance_scale_edit=4.0, result_folder=tempfile.mkdtemp(), for_prompt="bench")
    e = EditStableDiffusion(args, net, vae, *embs)
    zt = torch.randn(1, 4, 64, 64, device=dev, generator=g)
    mask = rectangle_mask(512).to(dev)
    t = float(e.scheduler.timesteps[e.edit_t_idx])
    run = lambda n: e.local_encoder_decoder_pullback_zt(zt, t, e.edit_t_idx, *embs, pca_rank=K_RANK, min_iter=10 ** 6,
                                                        max_iter=n, mask=mask, mode="null+(for-null)")
    run(2)
    torch.cuda.synchronize()
    a.record()
    n_it = 4
    _, s, _ = run(n_it)
    b.record()
    torch.cuda.synchronize()
    it_ms = a.elapsed_time(b) / n_it
    out.update({"unet": "stand-in latent U-Net (ch 128, mult 1-2-4-4, self + cross attention at 16^2 / 8^2, 77 x 768 prompt), "
                        "guidance mode null+(for-null): 2 conditionings per Jacobian product",
                "power_iteration_ms": it_ms, "probes_per_s": K_RANK / (it_ms * 1e-3),
                "singular_values_finite": bool(torch.isfinite(s).all())})
    net.base.release_plans(); vae.release_plans()
    del net, vae, e
    torch.cuda.empty_cache()
    return out


def measure_probe_shard(unet, sched, dev, world, rank, R, barrier, max_over_ranks, n_it=3, k=64):
    """BASELINE config 3 under torchrun: rank-64 subspace iteration (mask = None, t idx 40), probe
    tangents sharded over the ranks (64 / world rows each, probed in chunks of <= 25), one all-gather
    of the [64, d] W rows per iteration, replicated orthonormalisation.  Rank 0 then repeats the same
    iterations alone (3 chunks of 22/22/20) for the strong-scaling reference and the parity figures."""
    import torch
    import torch.distributed as dist
    from loco_edit_b200 import dist as ld
    from loco_edit_b200.edit import local_basis, random_basis
    d = 3 * R * R
    sched.set_timesteps(100, device=dev)
    t = sched._ts_host[40]
    g = torch.Generator(device=dev).manual_seed(11)
    xt = torch.randn(1, 3, R, R, device=dev, generator=g)
    v0 = random_basis(d, k, dev, generator=g)
    dist.broadcast(xt, src=0)
    dist.broadcast(v0, src=0)
    unet.release_plans()
    torch.cuda.empty_cache()
    ld.sharded_local_basis_cuda(unet, sched, xt, t, k, v0, 1, mask=None)          # warm-up: plans, graphs, NCCL
    barrier()
    timer = ld.CommTimer()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _, s, V = ld.sharded_local_basis_cuda(unet, sched, xt, t, k, v0, n_it, mask=None, timer=timer)
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    ag = max_over_ranks(timer.total_ms())
    res = None
    unet.release_plans()
    torch.cuda.empty_cache()
    if rank == 0:
        local_basis(unet, sched, xt, t, k, v0=v0, min_iter=10 ** 6, max_iter=1, mask=None, verbose=False)
        torch.cuda.synchronize()
        a.record()
        _, s1, V1 = local_basis(unet, sched, xt, t, k, v0=v0, min_iter=10 ** 6, max_iter=n_it, mask=None,
                                verbose=False)
        b.record()
        torch.cuda.synchronize()
        ms1 = a.elapsed_time(b)
        qa, _ = torch.linalg.qr(V.double().T.cpu())
        qb, _ = torch.linalg.qr(V1.double().T.cpu())
        ang = torch.rad2deg(torch.acos(torch.linalg.svdvals(qa.T @ qb).clamp(max=1.0)))
        res = {"rank": k, "iterations": n_it, "probes_per_rank": [hi - lo for lo, hi in
                                                                  (ld.shard_range(k, world, r) for r in range(world))],
               "ms_per_iter": ms / n_it, "probes_per_s": k * n_it / (ms * 1e-3),
               "allgather_ms_per_iter": ag / n_it, "allgather_share": ag / ms,
               "allgather_bytes_per_iter": k * d * 4,
               "single_gpu_ms_per_iter": ms1 / n_it, "single_gpu_probes_per_s": k * n_it / (ms1 * 1e-3),
               "speedup_vs_1gpu": ms1 / ms,
               "parity_s_rel": float(((s - s1).abs() / s1).max()), "parity_deg": float(ang.max()),
               "s_top3": [float(x) for x in s[:3]],
               "timing": "CUDA events, max over ranks; parity = sharded vs rank 0 alone, same V0, same iterations"}
        unet.release_plans()
        torch.cuda.empty_cache()
    barrier()
    return res


def measure_sharded_latency(pipe, gen, dev, R, world, barrier, max_over_ranks):
    """ONE config-1 edit by all ranks together (`EditPipeline.edit_sharded`): replicated serial chain,
    the 10 probes of {edit, null} bases sharded jointly, the 5 final latents sharded; host buffers."""
    import torch
    from loco_edit_b200 import dist as ld
    x0, m = synthetic_inputs(R, 777)
    x0, m = x0.pin_memory(), m.pin_memory()
    gen.manual_seed(6000)
    pipe.edit_sharded(x0, m, gen=gen)
    ts, ags = [], []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        timer = ld.CommTimer()
        barrier()
        a.record()
        pipe.edit_sharded(x0, m, gen=gen, timer=timer)
        b.record()
        barrier()
        ts.append(max_over_ranks(a.elapsed_time(b)))
        ags.append(max_over_ranks(timer.total_ms()))
    i = ts.index(statistics.median(ts))
    return {"ms": ts[i], "allgather_ms": ags[i], "n_gpus": world,
            "what": "one config-1 edit, 138-step chain replicated, 10 probes and 5 final latents sharded"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from loco_edit_b200 import _lib
    from loco_edit_b200.pipeline import EditPipeline
    from loco_edit_b200.unet import B200UNet
    from loco_edit_b200.weights import DDPM256, random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the single JSON line
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    sd = random_state_dict(DDPM256, seed=1234)
    unet = B200UNet(DDPM256, sd, device=dev)
    pipe = EditPipeline(unet, k=K_RANK, k_null=K_NULL, n_iter=N_ITER)
    pipe_streams = pipe.basis_streams     # concurrent per-image power methods in the batch edit (LOCO_BASIS_STREAMS)
    gen = torch.Generator(device=dev)
    R = 256

    BATCH = args.batch

    def host_inputs(i):
        xs, ms = zip(*[synthetic_inputs(R, i * BATCH + j) for j in range(BATCH)])
        return torch.cat(xs, 0).pin_memory(), torch.stack(ms, 0).pin_memory()

    def run_dev(x0, m):
        if BATCH == 1:
            return pipe.edit_device(x0, m[0], gen=gen)["images"]
        return pipe.edit_batch_device(x0, m, gen=gen)["images"]

    def run_host(x0, m):
        if BATCH == 1:
            return pipe.edit(x0, m[0], gen=gen)
        return pipe.edit_batch(x0, m, gen=gen)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also builds plans / workspaces) ----
    for w in range(args.warmup):
        gen.manual_seed(1000 + w)
        x0, m = host_inputs(rank * 1000 + w)
        run_host(x0, m)
    barrier()

    # ---- device-resident timed region: `value` ----
    dev_inputs = []
    for s in range(args.steps):
        x0, m = host_inputs(rank * 1000 + 100 + s)
        dev_inputs.append((x0.to(dev), m.to(dev)))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = lib.loco_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        gen.manual_seed(2000 + s)
        run_dev(dev_inputs[s][0], dev_inputs[s][1])
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.loco_launch_count() - l0

    # ---- end-to-end through the public API with pinned host buffers: `e2e` ----
    host = [host_inputs(rank * 1000 + 200 + s) for s in range(args.steps)]
    barrier()
    e0.record()
    d2h = 0
    for s in range(args.steps):
        gen.manual_seed(3000 + s)
        imgs = run_host(host[s][0], host[s][1])
        d2h = imgs.numel() * 4
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    h2d = host[0][0].numel() * 4 + host[0][1].numel()

    # ---- roofline leg: per-kernel-family event timing of one more edit (rank 0) ----
    roof = None
    probes = None
    pk, pk_kind = peaks()
    if rank == 0:
        lib.loco_profile_enable(1)
        gen.manual_seed(4000)
        # one stream for this leg: the event pair around a launch must not include kernels of a concurrent stream
        streams_timed, pipe.basis_streams = pipe.basis_streams, 1
        run_dev(dev_inputs[0][0], dev_inputs[0][1])
        pipe.basis_streams = streams_timed
        import ctypes as C
        ms = (C.c_double * 3)(); work = (C.c_double * 3)(); nl = (C.c_longlong * 3)()
        lib.loco_profile_collect(ms, work, nl, 3)
        lib.loco_profile_enable(0)
        conv_tflops = work[0] / (ms[0] * 1e-3) / 1e12 if ms[0] > 0 else 0.0
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        roof = {"bound": "tensor",
                "kernel": "tcgen05 implicit-GEMM conv family, kind::f16 (fp16 operands, fp32 accumulation in TMEM): "
                          "conv_gemm_tf32_pair_kernel<1,1> = cta_group::2 halo variant on the >=64^2 layers, "
                          "conv_gemm_tf32_kernel<.,1,1> on the small ones (the kernel names keep their round-1 "
                          "'tf32' stem; the template arguments select the fp16 instantiation)",
                "achieved": conv_tflops,
                "peak": peak, "unit": "TFLOP/s", "frac": conv_tflops / peak,
                # dram__bytes_read + dram__bytes_write of the dominant launch from the committed ncu --set full
                # capture profiles/r2u_conv_pair_f16_jvp6_ncu_full_raw.csv (CTA-pair halo variant, 3x3 128->128
                # at 256^2, 6 rows = primal + 5 tangents, fp16 in / out: algorithmic 100.7 MB in + 100.7 MB out;
                # part of the output is still in L2 when the kernel ends)
                "traffic": 161.6e6, "traffic_launch": "conv 3x3 128->128, 6 x 256x256 fp16 (116 GFLOP, 83.0 us under ncu)",
                # same capture: sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active / _elapsed
                "ncu_tensor_pipe_active_pct": 86.8, "ncu_tensor_pipe_elapsed_pct": 75.2,
                "peak_source": pk_kind + " bf16 dense sustained (kind::f16 runs at the bf16 rate)",
                "launches": int(nl[0]), "avg_launch_ms": ms[0] / max(1, nl[0]),
                "flops_per_launch": work[0] / max(1, nl[0]),
                "conv_ms_per_step": ms[0], "groupnorm_ms_per_step": ms[1],
                "groupnorm_gbs": work[1] / (ms[1] * 1e-3) / 1e9 if ms[1] > 0 else 0.0,
                "hbm_peak_gbs": pk["hbm_gbs"]}
        # JVP / VJP probe throughput (fused rank-5 passes)
        plan = unet.plan(1, K_RANK, K_RANK)
        xin = torch.randn(1 + K_RANK, 3, R, R, device=dev)
        gin = torch.randn(K_RANK, 3, R, R, device=dev)
        for _ in range(2):
            plan.forward(xin, 595.3636); plan.vjp(gin)
        torch.cuda.synchronize()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 5
        a.record()
        for _ in range(reps):
            plan.forward(xin, 595.3636)
        b.record()
        for _ in range(reps):
            plan.vjp(gin)
        c.record()
        torch.cuda.synchronize()
        probes = {"jvp_probes_per_s": K_RANK * reps / (a.elapsed_time(b) * 1e-3),
                  "vjp_probes_per_s": K_RANK * reps / (b.elapsed_time(c) * 1e-3),
                  "jvp_pass_ms": a.elapsed_time(b) / reps, "vjp_pass_ms": b.elapsed_time(c) / reps}
        for bsz in sorted({1, BATCH, 5 * BATCH}):
            pl = unet.plan(bsz)
            xb = torch.randn(bsz, 3, R, R, device=dev)
            pl.forward(xb, 595.3636)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                pl.forward(xb, 595.3636)
            b.record()
            torch.cuda.synchronize()
            probes["fwd_b%d_ms" % bsz] = a.elapsed_time(b) / reps

    # ---- single-edit latency (north_star: < 1 s), the drop-in driver, the small HBM-bound kernels ----
    latency_b1 = dropin = bw = text = sd_latent = None
    if rank == 0 and world == 1 and not args.no_extras:
        latency_b1 = measure_latency_b1(pipe, gen, R)
        dropin = measure_dropin(unet, dev, R)
        bw = measure_bandwidth_kernels(dev, pk["hbm_gbs"])
        text = measure_text_unet(dev, R)
        sd_latent = measure_sd_latent(dev)
    # ---- multi-GPU data path with a collective: one edit by all ranks, and BASELINE config 3 ----
    sharded_latency = probe_shard = None
    if world > 1 and not args.no_extras:
        unet.release_plans()
        torch.cuda.empty_cache()
        sharded_latency = measure_sharded_latency(pipe, gen, dev, R, world, barrier, max_over_ranks)
        probe_shard = measure_probe_shard(unet, pipe.driver.scheduler, dev, world, rank, R, barrier,
                                          max_over_ranks)

    # ---- BASELINE config 2 proper: the P2 / guided-diffusion U-Net with the FFHQ_P2 script settings
    # (edit_t 0.2, rank 3 + null 5, scale 12, 1 step; scripts/main_hf_null_space_projection_FFHQ_P2.sh),
    # one warm-up and one timed batch of BATCH image/mask pairs (single-GPU runs only; extra key) ----
    p2 = None
    if rank == 0 and world == 1 and not args.no_p2:
        from loco_edit_b200.weights import P2_256
        del pipe
        unet.release_plans()
        torch.cuda.empty_cache()
        net2 = B200UNet(P2_256, random_state_dict(P2_256, seed=1234), device=dev)
        pipe2 = EditPipeline(net2, k=3, k_null=5, edit_t=0.2, n_iter=N_ITER, scale=12.0, num_step=1, vis_num=2)
        xs, ms_ = host_inputs(rank * 1000 + 300)
        xs, ms_ = xs.to(dev), ms_.to(dev)
        gen.manual_seed(5000)
        pipe2.edit_batch_device(xs, ms_, gen=gen)
        torch.cuda.synchronize()
        e0.record()
        out2 = pipe2.edit_batch_device(xs, ms_, gen=gen)["images"]
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1)
        f_p2 = 0.3879e12
        fwd_eq = 98 + 79 + N_ITER * (1 + 2 * (3 + 5)) + 20 * 3
        p2 = {"value": BATCH / (ms2 * 1e-3), "unit": "edits/s", "ms_per_step": ms2,
              "workload": "P2 U-Net (P2_DICT, 93.6 M params, random init), %d pairs per step, FFHQ_P2 script "
                          "settings: t=0.2T, rank 3 + null 5, N=12, 98+79+20 DDIM steps, 3 edited latents" % BATCH,
              "fwd_equivalents_executed": fwd_eq,
              "achieved_tflops": BATCH * fwd_eq * f_p2 / (ms2 * 1e-3) / 1e12,
              "finite": bool(torch.isfinite(out2).all())}
        del pipe2, net2

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        smp = cpu_sample(threads)
        cpu = {"value": 1.0 / smp["edit_s_assembled"], "unit": "edits/s", "cores": threads, "kind": "port",
               "sample": CPU_SAMPLE_TEXT,
               "forward_b1_s": smp["forward_b1_s"], "forward_b5_s": smp["forward_b5_s"],
               "rank5_iter_s": smp["rank5_iter_s"], "jvp_probes_per_s": smp["jvp_probes_per_s"]}

    if rank == 0:
        n_edits = args.steps * world * BATCH
        value = n_edits / (ms_dev * 1e-3)
        line = {
            "metric": "edits/sec (rank-5 @256^2, t=0.6T)", "value": value, "unit": "edits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {**workload_config(BATCH),
                       "fwd_equivalents_executed": EDIT_FWD_EXECUTED,
                       "arithmetic": "fp16 storage + tcgen05 kind::f16 products with fp32 accumulation for the convs "
                                     "(tangent / cotangent rows range-scaled by powers of two), fp32 GroupNorm / "
                                     "softmax / PMP / DDIM math, tf32 tcgen05 attention products; LOCO_FWD_FP16=0 "
                                     "LOCO_JAC_FP16=0 select the round-1 tf32 programs",
                       "l2": "per-edit working set (>8 GB of activations) exceeds the 126 MB L2; no flush needed",
                       "basis_streams": pipe_streams,
                       "parallelism": "dp%d (independent images per rank, no collective)" % world},
            "achieved_tflops": value * EDIT_FWD_EXECUTED * F_DDPM256 / 1e12,
            "e2e": {"value": n_edits / (ms_e2e * 1e-3), "unit": "edits/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "p2_ffhq": p2,
            "latency_b1_ms": latency_b1, "latency_b1_target_ms": 1000.0,
            "latency_b1_sharded": sharded_latency, "probe_shard": probe_shard,
            "dropin_driver": dropin, "bandwidth_kernels": bw, "text_conditioned": text,
            "sd_latent": sd_latent,
        }
        if probes:
            line.update(probes)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="image/mask pairs edited together per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-p2", action="store_true", help="skip the extra P2 / FFHQ_P2 batch-edit measurement")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the single-edit latency / drop-in driver / small-kernel / probe-sharding legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
