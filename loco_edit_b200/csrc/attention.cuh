// Single-head self-attention core of the DDPM AttnBlock (reference: ddpm/diffusion.py:941-966) and
// its JVP / VJP.  The token counts are tiny (256 or 64 tokens x 512 channels, 0.14 % of the U-Net
// FLOPs) and the reference computes them with fp32 bmm, so this stays on CUDA cores in exact fp32.
#pragma once
#include "common.cuh"

namespace loco {

int attention_init();   // one-time kernel attribute setup

// C[b](m,n) = alpha * sum_k A[b](m,k) * B[b](k,n) + beta * C[b](m,n), arbitrary element strides.
struct GemmOperand {
  const float* ptr;
  long long s0, s1, sb;   // strides of (first index, second index, batch)
  long long sh = 0;       // stride of the attention head (second batch dimension)
};
int batched_gemm(GemmOperand A, GemmOperand B, float* C, long long sCm, long long sCn, long long sCb,
                 long long sCh, int M, int N, int K, int batch, int heads, float alpha, float beta,
                 int round_out, cudaStream_t s);

// P[b][i][:] = softmax(scale * S[b][i][:]) in place.
int softmax_rows(float* S, int T, int batch, float scale, cudaStream_t s);
// X[b][i][:] = scale * P0[i][:] * (X[b][i][:] - sum_j P0[i][j] X[b][i][j])   (softmax JVP and VJP)
int softmax_lin_rows(const float* P0, float* X, int T, int batch, int heads, float scale,
                     cudaStream_t s);

// qkv: [N, T, 3C]; S: [N, heads, T, T] scratch that keeps P afterwards; o: [N, T, C].
// head_ch == 0: one head, channels q | k | v (DDPM AttnBlock); head_ch > 0: C / head_ch heads in
// the guided-diffusion "legacy" order (per head q | k | v of head_ch channels each,
// guided_diffusion/unet.py:339-356).  Rows < n_primal are ordinary forward passes; the remaining
// rows are tangents of primal row 0.
int attention_forward(View qkv, int n_primal, int head_ch, float* S, View o, cudaStream_t s);
// go: [k, T, C] cotangent of o; qkv0/P0: saved primal row; gP: [k, heads, T, T] scratch;
// gqkv: [k, T, 3C] result.
int attention_vjp(View go, View qkv0, int head_ch, const float* P0, float* gP, View gqkv,
                  cudaStream_t s);

// Fused tcgen05 path (attention_tc.cu): taken by attention_forward / attention_vjp whenever the
// shape is eligible (T <= 256 tokens, T % 32 == 0, head width D % 32 == 0, D <= 512, fp32 tensors);
// LOCO_ATTN_TC=0 keeps the CUDA-core path above (profiling aid).
bool attention_tc_enabled();
bool attention_tc_eligible(int T, int C, int head_ch);
int attention_tc_init();
int attention_forward_tc(View qkv, int n_primal, int head_ch, float* S, View o, cudaStream_t s);
int attention_vjp_tc(View go, View qkv0, int head_ch, const float* P0, float* gP, View gqkv, cudaStream_t s);
// Cross-attention to a fixed context (attention_tc.cu): q [N, Tq, C], kv [1, Tk, 2C] = K_c | V_c with
// Tk a multiple of 64 (<= 256) of which Tk_valid rows exist, S [N, heads, Tq, Tk], o [N, Tq, C];
// rows >= n_primal are tangents of row 0 (d/dq only).  VJP: go [K, Tq, C] -> gq [K, Tq, C].
bool attention_cross_eligible(int Tk, int C, int heads);
int attention_cross_forward_tc(View q, View kv, int n_primal, int heads, int Tk_valid, float* S, View o, cudaStream_t s);
int attention_cross_vjp_tc(View go, View kv, int heads, int Tk_valid, const float* P0, View gq, cudaStream_t s);

}  // namespace loco
