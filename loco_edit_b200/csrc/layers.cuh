// Bandwidth-bound layer kernels of the U-Net hot path (GroupNorm/SiLU and their JVP/VJP, edge
// convolutions with 3 channels on one side, resampling, timestep embedding, weight packing).
#pragma once
#include "common.cuh"

namespace loco {

int layers_init();   // one-time kernel attribute setup

// ---- weight packing (done once when parameters are loaded) ----
// torch conv weight [Cout][Cin][kh][kw] -> GEMM operand [Cout][(r*kw+s)*Cin + ci], tf32-rounded.
int pack_conv_fprop(const float* w, float* dst, int Cout, int Cin, int kh, int kw, cudaStream_t s);
// torch conv weight [Cout][Cin][kh][kw] -> data-gradient operand [Cin][(r*kw+s)*Cout + co].
// cout_total/co_off place the block inside a wider fused operand (q|k|v): row length is
// kh*kw*cout_total and this weight fills columns [co_off, co_off+Cout) of every tap.
int pack_conv_dgrad(const float* w, float* dst, int Cout, int Cin, int kh, int kw, int cout_total,
                    int co_off, cudaStream_t s);
// fp16 variants of the two GEMM packs (operands of tcgen05.mma kind::f16; dst holds __half)
int pack_conv_fprop16(const float* w, void* dst, int Cout, int Cin, int kh, int kw, cudaStream_t s);
int pack_conv_dgrad16(const float* w, void* dst, int Cout, int Cin, int kh, int kw, int cout_total,
                      int co_off, cudaStream_t s);
// 3-channel edge convolutions: -> [9][3][C].  in_is_3: weight is [C][3][3][3] (conv_in), else
// [3][C][3][3] (conv_out).  Kept in full fp32 (these run on CUDA cores).
int pack_conv_edge(const float* w, float* dst, int C, int in_is_3, cudaStream_t s);

// ---- 3-channel edge convolutions (NCHW <-> channels-last boundary of the network) ----
// out[n,y,x,c] = bias[c]*(n<bias_rows) + sum_{tap,j} We[tap][j][c] * in3[n,j,y+dy,x+dx];
// flip = 0: (dy,dx) = (r-1,s-1) (conv_in forward); flip = 1: (1-r,1-s) (conv_out data gradient).
// scale_dev (optional, device float[2] = {s, 1/s}): rows n >= scale_from of the expand output are
// multiplied by s, of the reduce output by 1/s (power-of-two range scaling of the fp16 tangent /
// cotangent rows, see pow2_scale).
int edge_conv_expand(const float* in3_nchw, const float* We, const float* bias, int bias_rows,
                     View out, int flip, int round_out, cudaStream_t s, const float* scale_dev = nullptr,
                     int scale_from = 0);
// out3[n,j,y,x] = bias[j]*(n<bias_rows) + sum_{tap,c} Wr[tap][j][c] * in[n,y+dy,x+dx,c];
// flip = 0: conv_out forward; flip = 1: conv_in data gradient.
int edge_conv_reduce(View in, const float* Wr, const float* bias, int bias_rows, float* out3_nchw,
                     int flip, cudaStream_t s, const float* scale_dev = nullptr, int scale_from = 0);
// scale2[0] = 2^floor(log2(target / max|v|)), scale2[1] = its inverse; tmp: one device word of scratch
int pow2_scale(const float* v, long long n, float target, float* scale2, unsigned* tmp, cudaStream_t s);

// ---- thin edges (<= 4 channels) as zero-padded 64-channel tensors for the tcgen05 conv kernels ----
// (4-channel latents of the Stable-Diffusion-shaped U-Net and the VAE decoder)
constexpr int kThinPad = 64;
// out[n,y,x,0:c] = scale(n) * (mix ? W in + b*(n<bias_rows) : in)[n,:,y,x], out[...,c:64] = 0; in is NCHW fp32.
// mix = [c*c + c] device floats (1x1 conv weight [o][i], then bias) or null; scale(n) = scale_dev[0] for
// rows n >= scale_from (1 if scale_dev is null).
int thin_pad(const float* in_nchw, int c, View out, const float* mix, int bias_rows, const float* scale_dev,
             int scale_from, int round_out, cudaStream_t s);
// out[n,j,y,x] = scale'(n) * (mix ? W^T in : in)[n,y,x,j] for j < c; scale'(n) = scale_dev[1] for rows >= scale_from.
int thin_extract(View in, int c, float* out_nchw, const float* mix, const float* scale_dev, int scale_from, cudaStream_t s);

// ---- GroupNorm(32 groups) + optional SiLU: forward, JVP and VJP ----
// stats layout: double [rows][32][2].
// Forward/JVP: rows < n_primal accumulate (sum x, sum x^2); tangent rows (sum dx, sum x0*dx) with
// x0 = primal row 0.  (reference: Normalize/nonlinearity, ddpm/diffusion.py:806-811)
int gn_stats_fwd(View x, int n_primal, double* stats, cudaStream_t s);
// Statistics of a tensor with cg channels per group, regrouped for a consumer whose groups are PAIRS of them (a decoder
// concat buffer with as many channels again): dst[row][group0 + g] += src[row][2 g] + src[row][2 g + 1], g < 16.
int gn_stats_fold_pairs(const double* src, double* dst, int rows, int group0, cudaStream_t s);
int gn_apply_fwd(View x, int n_primal, const double* stats, const float* gamma, const float* beta,
                 float eps, int silu, int round_out, View y, cudaStream_t s);
// VJP: xp = saved primal input (1 row), pstats = its (sum x, sum x^2); gy = k cotangent rows of the
// layer output.  a = gamma * act'(u) * gy;  stats rows accumulate (sum a, sum x*a).
int gn_stats_vjp(View xp, const double* pstats, View gy, const float* gamma, const float* beta,
                 float eps, int silu, double* stats, cudaStream_t s);
// gx (+)= rstd * (a - mean(a) - xhat * mean(xhat a)) + addend
int gn_apply_vjp(View xp, const double* pstats, View gy, const double* stats, const float* gamma,
                 const float* beta, float eps, int silu, const View* addend, int accumulate,
                 int round_out, View gx, cudaStream_t s);

// Small sites (a (row, group) slice of <= 8 K elements, LOCO_GN_SMALL_MAX: the <= 32^2 layers): statistics + apply in ONE launch, a
// block per (row, group), no atomics.  gn_small_fwd also stores the statistics (the VJP pass reads the primal row's).
bool gn_small_eligible(const View& x);
int gn_small_fwd(View x, int n_primal, double* stats, const float* gamma, const float* beta, float eps, int silu,
                 int round_out, View y, cudaStream_t s);
int gn_small_vjp(View xp, const double* pstats, View gy, const float* gamma, const float* beta, float eps, int silu,
                 const View* addend, int accumulate, int round_out, View gx, cudaStream_t s);

// ---- resampling ----
// out (+)= scale * nearest_upsample(in), out = 2H x 2W.  scale 1: the DDPM Upsample / P2 up ResBlock
// (ddpm/diffusion.py:816-832, guided_diffusion/unet.py:95-124); scale 1/4: VJP of the 2x2 avg-pool.
int upsample2x(View in, View out, float scale, int accumulate, int round_out, cudaStream_t s);
// tensor-core (mma.sync) versions of the two edge convolutions for fp16 tensors with 128 channels whose
// image is a multiple of 8 x 32 pixels (edge_mma.cu); LOCO_EDGE_MMA=0 keeps the CUDA-core kernels
int edge_mma_init();
bool edge_mma_eligible(const View& v);
int edge_conv_reduce_mma(View in, const float* Wr, const float* bias, int bias_rows, float* out3, int flip, cudaStream_t s,
                         const float* scale_dev, int scale_from);
int edge_conv_expand_mma(const float* in3, const float* We, const float* bias, int bias_rows, View out, int flip,
                         cudaStream_t s, const float* scale_dev, int scale_from);
// K_c | V_c = ctx W^T + b for the first n_tok of `rows` context rows (the rest zero), tf32-rounded
int context_kv(const float* ctx, int n_tok, int dim, const float* w, const float* b, int cout, int rows, float* out,
               cudaStream_t s);
// out (+)= scale * 2x2 sum-pool(in).  scale 1: VJP of upsample2x; scale 1/4: the avg-pool of the P2
// down ResBlock (guided_diffusion/unet.py:127-158 with use_conv = False).
int sumpool2x(View in, View out, float scale, int accumulate, int round_out, cudaStream_t s);
// out (+)= in   (cotangent fan-in where no producing kernel can fuse it)
int add_views(View in, View out, int accumulate, cudaStream_t s);

// ---- timestep embedding (reference: get_timestep_embedding + temb.dense, ddpm/diffusion.py:
// 154-157, 783-804): temb_act = silu(dense1(silu(dense0([sin(t w), cos(t w)])))), then every
// ResnetBlock's temb_proj Linear(temb_ch -> Cout) evaluated into one packed vector.
// t is read from device memory so a captured CUDA graph can be replayed for any timestep.
// style 0 = DDPM sinusoid ([sin, cos], divisor half-1), 1 = guided-diffusion ([cos, sin], half).
// cond: optional [4ch] conditioning embedding (device) added to the timestep embedding.
int temb_forward(const float* t_dev, int ch, const float* w0, const float* b0, const float* w1,
                 const float* b1, float* scratch /*2*4ch*/, int style, const float* cond, cudaStream_t s);
// P2 scale-shift norm folded into the GroupNorm affine, all sites of a program in one launch:
// out[out_off + c] = gamma[c] (1 + scale[c]), out[out_off + C + c] = beta[c] (1 + scale[c]) + shift[c]
// with (scale | shift) = tproj[tproj_off .. tproj_off + 2C).
struct AffineSite { long long gamma_off, beta_off; int tproj_off, C; long long out_off; };
int scale_shift_affine(const AffineSite* sites_dev, int n_sites, const float* weights,
                       const float* tproj, float* out, cudaStream_t s);
int set_scalar(float* dst, float v, cudaStream_t s);
int temb_project(const float* temb_act, int temb_ch, const float* w, const float* b, int cout,
                 float* out, cudaStream_t s);

}  // namespace loco
