/*
 * loco_b200.h -- C ABI of the B200-native LOCO-Edit editing-direction hot path.
 *
 * The reference (ChicyChen/LOCO-Edit) is pure Python/PyTorch and has no FFI layer; its boundary for
 * this path is the set of tensor-valued Python methods of `EditUncondDiffusion`
 * (src/modules/edit.py:2034-2625) and `YHCustomScheduler` (src/utils/utils.py:305-423).  Each entry
 * point below names the reference call it replaces.  The Python mirror of those methods
 * (loco_edit_b200/edit.py, scheduler.py) binds this library with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors in practice); nothing is
 *     allocated or freed on the device by this library; host handles are created/destroyed here;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync;
 *   - return value 0 = ok, non-zero = error (message via loco_last_error()); nothing throws;
 *   - images / tangents / cotangents cross the ABI in the reference's NCHW fp32 layout;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef LOCO_B200_H_
#define LOCO_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct loco_unet loco_unet_t;
typedef struct loco_plan loco_plan_t;

/* Architecture of the U-Net.
 * kind 0: DDPM.__init__, src/models/ddpm/diffusion.py:24-126 (hyper-parameters of
 *         src/configs/custom_celeba_ddpm.yml:21-30; == google/ddpm-ema-celebahq-256);
 * kind 1: guided-diffusion UNetModel as built by create_model(**P2_DICT),
 *         src/models/guided_diffusion/script_util.py:166-190,379-435 and unet.py:398-684
 *         (scale-shift norm, ResBlock up/down-sampling, learn_sigma: eps = first 3 of 6 output
 *         channels, multi-head "legacy" attention with head_ch channels per head);
 * kind 2: decoder half of a latent-diffusion VAE -- the `self.vae.decode(z).sample` inside the Stable
 *         Diffusion twin's get_x0, src/modules/edit.py:764-774 (diffusers AutoencoderKL.decode =
 *         Decoder(post_quant_conv(z)); module tree conv_in, mid.{block_1, attn_1, block_2},
 *         up.{l}.block.{0..num_res_blocks} [+ up.{l}.upsample.conv], norm_out, conv_out with the
 *         ResnetBlock / AttnBlock / Upsample of src/models/ddpm/diffusion.py:816-966, no timestep input).
 *         `resolution` is the LATENT resolution; rows of the input are [in_ch, R, R], rows of the output
 *         [out_ch, R << (n_levels-1), R << (n_levels-1)]; `t` is ignored.  Same plan / forward / vjp
 *         entry points as the U-Nets (fused primal + k-tangent pass, k-cotangent pass).
 * in_ch / out_ch other than 3 (the 4-channel latents of the Stable-Diffusion-shaped U-Net, kind 0, and
 * of the decoder): that end of the network runs on the tcgen05 conv kernels over a thin side zero-padded
 * to 64 channels instead of the 3-channel edge kernels. */
typedef struct loco_arch {
  int ch;                  /* base channels (multiple of 128) */
  int n_levels;            /* len(ch_mult) <= 8 */
  int ch_mult[8];
  int num_res_blocks;
  int n_attn;              /* len(attn_resolutions) <= 4 */
  int attn_resolutions[4];
  int resolution;          /* input H = W */
  int in_ch, out_ch;       /* 3, 3 (image-space U-Nets); 4, 4 (latent U-Net, kind 0); 4, 3 (kind 2); each 1..4 */
  float gn_eps;            /* 1e-6 (DDPM) / 1e-5 (guided diffusion) */
  int kind;                /* 0 = DDPM, 1 = P2 / guided diffusion, 2 = VAE decoder */
  int head_ch;             /* kind 1: channels per attention head (num_head_channels); else 0 */
  int ctx_dim;             /* kind 0: > 0 adds a cross-attention sub-block to every AttnBlock, keys / values =
                              Linear(ctx_dim -> 2C) of a prompt embedding (text-conditioned twins); 0 = none */
  int ctx_heads;           /* heads of those cross-attention layers (C / heads a multiple of 64, <= 512) */
} loco_arch_t;

int loco_abi_version(void);
const char* loco_last_error(void);
/* number of CUDA kernels this library has launched so far in this process */
long long loco_launch_count(void);
/* per-kernel-family CUDA-event timing (family 0 = tcgen05 conv GEMM [work = FLOPs], 1 = GroupNorm
 * [work = algorithmic bytes], 2 = other): enable, run the workload, then collect (device sync). */
int loco_profile_enable(int on);
int loco_profile_collect(double* ms, double* work, long long* launches, int nfam);

/* ---------------- model: parameters by reference state_dict name ---------------- */
int loco_unet_create(const loco_arch_t* arch, loco_unet_t** out);
void loco_unet_destroy(loco_unet_t* m);
/* size (in floats) of the packed-weight arena the caller must provide */
long long loco_unet_weight_floats(const loco_unet_t* m);
int loco_unet_bind_weights(loco_unet_t* m, float* arena);
int loco_unet_num_params(const loco_unet_t* m);
/* name/shape of parameter i (same names as DDPM.state_dict()); shape has up to 4 entries */
int loco_unet_param_info(const loco_unet_t* m, int i, char* name, int name_cap, int* shape, int* ndim);
/* pack one parameter (torch layout, fp32, device memory) into the arena */
int loco_unet_load_param(loco_unet_t* m, const char* name, const float* src, long long numel,
                         void* stream);

/* ---------------- execution plans ---------------- */
/* n_primal samples, n_tangent probe tangents (JVP rows, need n_primal == 1), n_cotangent rows for
 * the transposed pass.  (B,0,0) = plain eps_theta(x,t) for the DDIM loops; (1,k,k) = pull-back. */
int loco_plan_create(const loco_unet_t* m, int n_primal, int n_tangent, int n_cotangent,
                     loco_plan_t** out);
/* Same with storage / arithmetic flags.  LOCO_PLAN_FP16: every activation of the program is stored
 * in fp16 and the GEMM-shaped layers run on tcgen05.mma kind::f16 (fp32 accumulation) instead of
 * fp32 storage + kind::tf32: same 10-bit operand mantissa as the TF32 path (the reference's own GPU
 * numerics, torch's cudnn.allow_tf32 default), half the HBM bytes, twice the tensor rate.  Meant for
 * the Jacobian-free programs (n_tangent = n_cotangent = 0): the DDIM inversion / denoising loops
 * `self.unet(xt, t)` of src/modules/edit.py:2151 and :2572. */
#define LOCO_PLAN_FP16 1
int loco_plan_create_ex(const loco_unet_t* m, int n_primal, int n_tangent, int n_cotangent, int flags,
                        loco_plan_t** out);
void loco_plan_destroy(loco_plan_t* p);
long long loco_plan_workspace_bytes(const loco_plan_t* p);
int loco_plan_bind(loco_plan_t* p, void* workspace);
int loco_plan_info(const loco_plan_t* p, double* fwd_flops, double* vjp_flops, int* fwd_ops,
                   int* vjp_ops);

/* eps = unet(x, t): replaces `self.unet(xt, t)` (src/modules/edit.py:2151, 2375, 2572).
 * x: [n_primal + n_tangent, in_ch, R, R], eps: the same rows of the network output ([.., out_ch, R, R];
 * kind 2: the decoded images [.., out_ch, R << (n_levels-1), ...], replacing `self.vae.decode(z).sample`,
 * src/modules/edit.py:770); tangent rows hold dx on input and d eps on output, i.e. the forward-mode
 * product torch.func.jacfwd computes at src/modules/edit.py:2455 (:880 for the latent-space twin). */
int loco_unet_forward(loco_plan_t* p, const float* x, float t, float* eps, void* stream);
/* Conditional U-Net eps(x, t, c): `cond` [4*ch] (device; NULL = unconditional) is added to the
 * timestep embedding of every following loco_unet_forward of this plan -- the class / pooled-text
 * conditioning of guided-diffusion style U-Nets.  It is the stand-in for the text-conditioned U-Nets
 * of the T-LOCO twins (`self.unet(x, t, encoder_hidden_states=...)`, src/modules/edit.py:655-658,
 * 1319-1322): classifier-free guidance evaluates the same network under 2-3 conditionings and
 * combines eps linearly, so its Jacobian products are the same linear combination of this plan's
 * JVP / VJP passes. */
int loco_plan_set_condition(loco_plan_t* p, const float* cond, void* stream);
/* Prompt embedding ctx [n_tokens <= 128, ctx_dim] (device memory) of a U-Net created with ctx_dim > 0:
 * the `encoder_hidden_states` of `self.unet(x, t, encoder_hidden_states=...)` (src/modules/edit.py:655-658,
 * 1319-1322).  Computes K_c | V_c of every cross-attention layer once; must precede loco_unet_forward.
 * The context is a constant of x -> eps(x, t, ctx), so JVP / VJP differentiate the query side only. */
int loco_plan_set_context(loco_plan_t* p, const float* ctx, int n_tokens, void* stream);
/* gx[j] = (d eps / d x)^T g_eps[j] at the primal point of the last loco_unet_forward: replaces
 * the k backward passes of torch.autograd.functional.jacobian (src/modules/edit.py:2479; :893 for the
 * latent-space twin, where g_eps rows have the shape of the network OUTPUT and gx rows of its input). */
int loco_unet_vjp(loco_plan_t* p, const float* g_eps, float* gx, void* stream);

/* ---------------- pull-back (power method) ---------------- */
long long loco_pullback_scratch_bytes(int k, long long d);
/* One subspace iteration of local_encoder_decoder_pullback_xt (src/modules/edit.py:2443-2483):
 *   U = mask o J V^T (JVP), W = J^T U (VJP), (s, V_out) = orthonormalise(W).
 * xt [d], V [k,d] in; u_full [k,d] (masked, zeros outside the mask), w_out [k,d] (un-orthonormalised
 * W, may be NULL), V_out [k,d], s_out [k] = sqrt(singular values of W) out.  at = alphas_cumprod[floor t].
 * align_sign != 0 picks row signs with <V_out_i, V_i> >= 0. */
int loco_pullback_iteration(loco_plan_t* p, const float* xt, float t, float at,
                            const unsigned char* mask, int noise, const float* V, int k, long long d,
                            int align_sign, float* u_full, float* w_out, float* V_out, float* s_out,
                            void* scratch, void* stream);

/* The two Jacobian products of one iteration without the orthonormalisation: rows of
 * U = mask o J V^T and W = J^T U for the k given rows of V.  This is the unit a rank runs on its
 * shard of the probe tangents before the all-gather of W (multi-GPU power method). */
int loco_pullback_probe(loco_plan_t* p, const float* xt, float t, float at, const unsigned char* mask,
                        int noise, const float* V, int k, long long d, float* u_full, float* w_out,
                        void* scratch, void* stream);

/* Same for a mixed shard of the joint {edit basis, null basis} probe set: rows [0,k_invert) of V are
 * probed through the masked Jacobian, rows [k_invert,k) through the complement-mask Jacobian
 * (src/modules/edit.py:2294 and :2307 share x_t, t and the primal activations). */
int loco_pullback_probe_pair(loco_plan_t* p, const float* xt, float t, float at,
                             const unsigned char* mask, int noise, const float* V, int k, int k_invert,
                             long long d, float* u_full, float* w_out, void* scratch, void* stream);

/* One iteration of BOTH local bases of run_edit_null_space_projection (src/modules/edit.py:2294 and
 * :2307) in a single fused pass: rows [0,k1) of V are probes of the masked Jacobian (edit basis),
 * rows [k1,k1+k2) probes of the complement-mask Jacobian (null basis).  They share x_t, t and the
 * primal activations, so one (1, k1+k2, k1+k2) plan serves both; each group is orthonormalised on
 * its own.  Results equal two loco_pullback_iteration calls. */
int loco_pullback_pair_iteration(loco_plan_t* p, const float* xt, float t, float at,
                                 const unsigned char* mask, int noise, const float* V, int k1, int k2,
                                 long long d, int align_sign, float* u_full, float* w_out, float* V_out,
                                 float* s_out, void* scratch, void* stream);

/* ---------------- bandwidth-bound pieces ---------------- */
/* P = (x - eps*sqrt(1-at))/sqrt(at): get_x0 without the mask (src/modules/edit.py:2386) */
int loco_pmp_forward(const float* x, const float* eps, float at, long long n, float* out, void* stream);
/* out = wa*a + wb*b + wc*c (b, c may be NULL): the classifier-free-guidance combination
 * `e_null + g (e_for - e_null) + g_edit (e_edit - e_null)` of src/modules/edit.py:660-673 / 1326-1373,
 * also applied to the tangents / cotangent products of the conditionings (the combination is linear). */
int loco_combine3(const float* a, float wa, const float* b, float wb, const float* c, float wc, long long n,
                  float* out, void* stream);
/* PMP differentiated along k directions (get_x0, src/modules/edit.py:1566-1587, 2369-2391):
 *   u = mask o (V - eps_dot sqrt(1-at)) / sqrt(at)   (noise != 0: u = mask o eps_dot)
 * and the seeds of the transposed pass g_eps = d<u,P>/d eps, gx_direct = d<u,P>/d x.  Rows >= k_invert
 * use the complement mask.  All [k, d]. */
int loco_pmp_jvp_epilogue(const float* V, const float* eps_dot, const unsigned char* mask, float at, int noise,
                          int k, int k_invert, long long d, float* u, float* g_eps, float* gx_direct,
                          void* stream);
/* Vh and sqrt(singular values) of W [k,d]: torch.linalg.svd at src/modules/edit.py:2482.
 * scratch >= loco_orthonormalise_scratch_bytes(k). */
long long loco_orthonormalise_scratch_bytes(int k);
int loco_orthonormalise(const float* W, int k, long long d, const float* v_prev, float* V,
                        float* s_out, void* scratch, void* stream);
/* vT = normalise_rows(vT_mod - (Vn^T (Vn vT_mod^T))^T): src/modules/edit.py:2317-2323.
 * scratch >= 8*(k_null*k + k) bytes. */
int loco_nullspace_project(const float* vT_mod, int k, const float* Vn, int k_null, long long d,
                           int project, float* out, void* scratch, void* stream);
/* YHCustomScheduler.step (src/utils/utils.py:342-374); x0_pred may be NULL */
int loco_ddim_step(const float* xt, const float* et, const float* noise, float at, float at_next,
                   float eta, long long n, float* xt_next, float* x0_pred, void* stream);
/* x + scale*v: x_space_guidance_direct (src/modules/edit.py:2618-2625) */
int loco_axpy(const float* x, const float* v, float scale, long long n, float* out, void* stream);
/* row-major indices of mask != 0 (selection order of `P_xt[:, mask]`, src/modules/edit.py:2390) */
int loco_mask_indices(const unsigned char* mask, long long d, int* idx, int* count, void* stream);
int loco_gather_rows(const float* src, int rows, long long d, const int* idx, int count, float* out,
                     void* stream);
int loco_scatter_rows(const float* src, int rows, long long d, const int* idx, int count, float* out,
                      void* stream);
int loco_gram(const float* A, int ka, const float* B, int kb, long long d, double* G, void* stream);

/* ---------------- single layers (used by the parity tests) ---------------- */
/* Channels-last convolution on the tcgen05 path.  kind: 0 = 3x3 s1 p1, 1 = 1x1, 2 = 3x3 s2 pad
 * (0,1,0,1), 3 = data gradient of kind 0, 4 = data gradient of kind 2.  w is the torch-layout
 * weight [Cout,Cin,k,k] of the FORWARD conv; wpack is scratch of the same size.  x: [N,H,W,Cx],
 * y: [N,Ho,Wo,Cy] contiguous.  bias/addend may be NULL.  splitk_scratch (optional, >= 1 MiB, first
 * 16 KiB zeroed) enables the split-K path for layers with fewer tiles than SMs. */
int loco_conv2d_nhwc(int kind, const float* x, int N, int H, int W, int Cx, const float* w, int Cout,
                     int Cin, float* wpack, const float* bias, int bias_rows, const float* addend,
                     int accumulate, float* y, void* splitk_scratch, long long splitk_bytes,
                     void* stream);
/* Same with typed tensors: in16 != 0: x (and the packed weights in wpack) are fp16 -> tcgen05.mma
 * kind::f16; out16 != 0: y and addend are fp16.  Accumulation is fp32 either way. */
int loco_conv2d_nhwc_ex(int kind, const void* x, int N, int H, int W, int Cx, const float* w, int Cout,
                        int Cin, void* wpack, const float* bias, int bias_rows, const void* addend,
                        int accumulate, void* y, void* splitk_scratch, long long splitk_bytes,
                        int in16, int out16, void* stream);
/* Pure host logic: 1 if a stride-1 3x3 convolution (kind 0) or its data gradient (kind 3) of this
 * shape is served by the halo / CTA-pair tcgen05 kernels (wave-quantisation cost model). */
int loco_conv_halo_eligible(int kind, int N, int H, int W, int Cout);
/* y = conv3x3(x; w) + conv1x1(x2; w2) + bias in one launch: `ResnetBlock.conv2` + `nin_shortcut`
 * (src/models/ddpm/diffusion.py:905-912) / `out_layers` conv + `skip_connection`
 * (src/models/guided_diffusion/unet.py:252-258) as the U-Net programs run them, with the optional
 * fused GroupNorm statistics of the result (stats: 64*N doubles = [row][group](sum, sum sq), groups of
 * stat_cg channels).  x2 may be null.  Only shapes served by the halo conv variants (H, W multiples of
 * 16, Cout multiple of 128, >= ~1 tile per SM). wpack/wpack2: scratch of w/w2's size. */
int loco_conv2d_fused_nhwc(const float* x, int N, int H, int W, int Cin, const float* w, int Cout,
                           const float* x2, int C2, const float* w2, float* wpack, float* wpack2,
                           const float* bias, int bias_rows, float* y, double* stats, int stat_cg,
                           void* stream);
/* micro-benchmark of one prepared conv launch (wpack must already hold packed weights):
 * average device ms over `reps` back-to-back launches; reports the split-K factor / grid used. */
int loco_conv_bench(int kind, float* x, int N, int H, int W, int Cx, float* wpack, int Cout, int Cin,
                    float* y, void* splitk_scratch, long long splitk_bytes, int max_ksplit, int reps,
                    float* ms_out, int* ksplit_out, int* grid_out, void* stream);
/* Same with typed tensors (in16 / out16 as in loco_conv2d_nhwc_ex; wpack must hold fp16 data when
 * in16), an optional residual addend of the output's type and optional fused GroupNorm statistics
 * (stats: 64*N doubles, zeroed by the caller; they keep accumulating over the repetitions). */
int loco_conv_bench_ex(int kind, void* x, int N, int H, int W, int Cx, void* wpack, int Cout, int Cin,
                       void* y, void* splitk_scratch, long long splitk_bytes, int max_ksplit, int reps,
                       int in16, int out16, const void* addend, double* stats, float* ms_out,
                       int* ksplit_out, int* grid_out, void* stream);
/* GroupNorm(32)+optional SiLU on [N,H,W,C]; rows >= n_primal are tangents of row 0.
 * stats: 8*N*64 bytes scratch. */
int loco_groupnorm_silu_fwd(const float* x, int N, int H, int W, int C, int n_primal,
                            const float* gamma, const float* beta, float eps, int silu, float* y,
                            void* stats, void* stream);
/* VJP of the above at primal xp [1,H,W,C] for K cotangent rows gy -> gx. stats: 8*(K+1)*64 bytes. */
int loco_groupnorm_silu_vjp(const float* xp, int H, int W, int C, const float* gy, int K,
                            const float* gamma, const float* beta, float eps, int silu, float* gx,
                            void* stats, void* stream);
/* Typed variants of the two entry points above: half = 1 takes fp16 tensors (the storage type of
 * the default programs; x / y / gy / gx / addend are then __half*), and `stages` selects the passes
 * (bit 0: statistics incl. clearing `stats`, bit 1: apply) so that each bandwidth-bound kernel can
 * be verified and timed alone; bit 2: the one-launch kernel the programs use on their small sites (a
 * (row, group) slice of <= 8 K elements: statistics + apply by one block per slice, no atomics).  VJP: gx = J^T gy (+ addend) (+ gx if accumulate). */
int loco_groupnorm_silu_fwd_ex(const void* x, int half, int N, int H, int W, int C, int n_primal,
                               const float* gamma, const float* beta, float eps, int silu, void* y,
                               void* stats, int stages, void* stream);
int loco_groupnorm_silu_vjp_ex(const void* xp, int half, int H, int W, int C, const void* gy, int K,
                               const float* gamma, const float* beta, float eps, int silu,
                               const void* addend, int accumulate, void* gx, void* stats, int stages,
                               void* stream);
/* attention core on qkv [N,T,3C]; S scratch [N,heads,T,T]; o [N,T,C].  head_ch = 0: one head,
 * channels q|k|v (ddpm/diffusion.py:941-966); head_ch > 0: C/head_ch heads, per head q|k|v
 * (QKVAttentionLegacy, guided_diffusion/unet.py:339-356). */
int loco_attention_fwd(const float* qkv, int N, int T, int C, int n_primal, int head_ch, float* S,
                       float* o, void* stream);
int loco_attention_vjp(const float* go, int K, int T, int C, int head_ch, const float* qkv0,
                       const float* P0, float* gP, float* gqkv, void* stream);

/* Cross-attention of image tokens to a fixed context (text-conditioned U-Nets; SD / IF twins of the
 * path, src/modules/edit.py:636-674, 1286-1373 call diffusers' UNet2DConditionModel, whose attention
 * processor computes softmax(q k^T / sqrt(d)) v with k, v = linear maps of encoder_hidden_states):
 * q [N,Tq,C] in `heads` heads of C/heads channels; kv [Tk,2C] = K_c | V_c, Tk a multiple of 64 (<= 256)
 * of which the first Tk_valid rows exist (the rest must be zero); S [N,heads,Tq,Tk] receives the
 * probabilities; o [N,Tq,C].  Rows >= n_primal are tangents of row 0 (the context is a constant).
 * VJP at the primal probabilities P0 [heads,Tq,Tk]: go [K,Tq,C] -> gq [K,Tq,C].  Fused tcgen05 kernels. */
int loco_cross_attention_fwd(const float* q, int N, int Tq, int C, int n_primal, const float* kv, int Tk,
                             int Tk_valid, int heads, float* S, float* o, void* stream);
int loco_cross_attention_vjp(const float* go, int K, int Tq, int C, const float* kv, int Tk, int Tk_valid,
                             int heads, const float* P0, float* gq, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LOCO_B200_H_ */
