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
    """Bounded sample of the config-1 edit on the host: one U-Net forward (B=1) and one rank-1
    power iteration (jacfwd-style JVP + autograd VJP) at 256x256.  The edit is extrapolated with the
    reference's own op counts: 138 F + 24 iterations of (1+3k)/(1+3) x the
    rank-1 iteration + 59*5 F."""
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
    v0, _ = torch.linalg.qr(torch.randn(x0.numel(), 1, generator=g))
    with torch.no_grad():
        unet(x0, t)                                   # warm-up (thread pool, allocator)
        t0 = time.perf_counter()
        unet(x0, t)
        f1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    pullback_ref.power_iteration(unet, sched, x0, t, v0.T, mask=mask)
    it1 = time.perf_counter() - t0
    it5 = it1 * (1 + 3 * K_RANK) / 4.0
    edit_s = 138 * f1 + 2 * N_ITER * it5 + 59 * 5 * f1
    return {"forward_s": f1, "rank1_iter_s": it1, "edit_s_extrapolated": edit_s,
            "jvp_probes_per_s": K_RANK / it5}


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
    edit_s = statistics.mean(v["edit_s_extrapolated"] for v in vals)
    value = 1.0 / edit_s
    sample = ("per step: 1 U-Net forward (B=1) + 1 rank-1 power iteration (JVP+VJP) at 256^2, fp32, "
              "torch CPU; edit extrapolated as 138 F + 24 x (16/4) x iter + 295 F (BASELINE.md s2)")
    line = {
        "impl": "reference", "metric": "edits/sec (rank-5 @256^2, t=0.6T)", "value": value,
        "unit": "edits/s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
        "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config-1 edit: DDPM-256 random-init, 1 image, rank 5 + null 5, N=12, t=0.6T",
                   "timing": "host wall clock, extrapolated from a bounded sample"},
        "cpu_baseline": {"value": value, "unit": "edits/s", "cores": threads, "kind": "port",
                         "sample": sample,
                         "jvp_probes_per_s": statistics.mean(v["jvp_probes_per_s"] for v in vals)},
        "e2e": {"value": value, "unit": "edits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
        run_dev(dev_inputs[0][0], dev_inputs[0][1])
        import ctypes as C
        ms = (C.c_double * 3)(); work = (C.c_double * 3)(); nl = (C.c_longlong * 3)()
        lib.loco_profile_collect(ms, work, nl, 3)
        lib.loco_profile_enable(0)
        conv_tflops = work[0] / (ms[0] * 1e-3) / 1e12 if ms[0] > 0 else 0.0
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        roof = {"bound": "tensor",
                "kernel": "tcgen05 implicit-GEMM conv family (conv_gemm_tf32_pair_kernel = cta_group::2 halo "
                          "variant on the >=64^2 layers, conv_gemm_tf32_kernel on the small ones)",
                "achieved": conv_tflops,
                "peak": peak, "unit": "TFLOP/s", "frac": conv_tflops / peak,
                # dram__bytes_read+write of the dominant launch (CTA-pair halo variant, 3x3 128->128 at
                # 256^2, 6 rows; algorithmic 403 MB) from the committed ncu --set full capture,
                # profiles/r1_conv_pair_ncu_full_raw.csv
                "traffic": 352.8e6, "traffic_launch": "conv 3x3 128->128, 6 x 256x256 (116 GFLOP)",
                # same capture: sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active / _elapsed
                "ncu_tensor_pipe_active_pct": 80.9, "ncu_tensor_pipe_elapsed_pct": 74.0,
                "peak_source": pk_kind + " bf16 dense sustained (kernel runs kind::tf32: half the bf16 rate)",
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
        cpu = {"value": 1.0 / smp["edit_s_extrapolated"], "unit": "edits/s", "cores": threads, "kind": "port",
               "sample": "1 U-Net forward (B=1) + 1 rank-1 power iteration at 256^2 on the host cores; edit "
                         "extrapolated with the reference's op counts (138 F + 24 x 4 x iter + 295 F)",
               "forward_s": smp["forward_s"], "rank1_iter_s": smp["rank1_iter_s"],
               "jvp_probes_per_s": smp["jvp_probes_per_s"]}

    if rank == 0:
        n_edits = args.steps * world * BATCH
        value = n_edits / (ms_dev * 1e-3)
        line = {
            "metric": "edits/sec (rank-5 @256^2, t=0.6T)", "value": value, "unit": "edits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "batch edit (BASELINE config 2 shape): DDPM-256 (ddpm-ema-celebahq-256 arch, "
                                   "random init), %d image/mask pairs of 256x256 per step per GPU, each a full "
                                   "config-1 edit: rank 5 + null rank 5, N=12 power iterations, t=0.6T, "
                                   "98+40+59 DDIM steps, 5 edited latents per image" % BATCH,
                       "images_per_step_per_gpu": BATCH,
                       "fwd_equivalents_per_edit": EDIT_FWD_EQUIV,
                       "tflop_per_edit": EDIT_FWD_EQUIV * F_DDPM256 / 1e12,
                       "fwd_equivalents_executed": EDIT_FWD_EXECUTED,
                       "l2": "per-edit working set (>8 GB of activations) exceeds the 126 MB L2; no flush needed",
                       "parallelism": "dp%d (independent images per rank, no collective)" % world},
            "achieved_tflops": value * EDIT_FWD_EXECUTED * F_DDPM256 / 1e12,
            "e2e": {"value": n_edits / (ms_e2e * 1e-3), "unit": "edits/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "p2_ffhq": p2,
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
