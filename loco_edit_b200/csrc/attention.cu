// Attention core (scores, softmax, value mixing) with JVP and VJP: see attention.cuh.
#include "attention.cuh"
#include <math.h>

namespace loco {

namespace {

constexpr int BM = 32, BN = 64, BK = 16;
constexpr int kGemmGroup = 128;                 // threads that own one K slab at a time
constexpr int kGemmThreads = 2 * kGemmGroup;

// 32x64 output tile per block.  The attention operands are tiny (256 or 64 tokens), so a launch has
// only about one block per SM and is bound by the latency of its K loop, not by FMA throughput:
// the block therefore runs TWO 128-thread groups (8x16 threads, 4x4 outputs each) that walk
// alternate K slabs with their own shared-memory tiles and named barriers, which halves the
// length of the dependent chain; the next slab of a group is prefetched into registers while the
// current one is multiplied, and group 1 hands its partial sums to group 0 through shared memory
// (fixed order: bit-reproducible).
__global__ void __launch_bounds__(kGemmThreads)
batched_gemm_kernel(GemmOperand A, GemmOperand B, float* __restrict__ C, long long sCm,
                    long long sCn, long long sCb, long long sCh, int heads, int M, int N, int K,
                    float alpha, float beta, int round_out, int a_kcontig, int b_ncontig) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  __shared__ float red[kGemmGroup][17];
  const int b = blockIdx.z / heads;       // batch row
  const int hd = blockIdx.z % heads;      // attention head
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* Ab = A.ptr + b * A.sb + hd * A.sh;
  const float* Bb = B.ptr + b * B.sb + hd * B.sh;
  const int grp = threadIdx.x / kGemmGroup;
  const int tid = threadIdx.x % kGemmGroup;
  const int tx = tid % 16, ty = tid / 16;   // 8x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + kGemmGroup * r;
      int m, k;
      if (a_kcontig) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      ra[r] = (m0 + m < M && k0 + k < K) ? Ab[(long long)(m0 + m) * A.s0 + (long long)(k0 + k) * A.s1] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = tid + kGemmGroup * r;
      int n, k;
      if (b_ncontig) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      rb[r] = (n0 + n < N && k0 + k < K) ? Bb[(long long)(k0 + k) * B.s0 + (long long)(n0 + n) * B.s1] : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + kGemmGroup * r;
      int m, k;
      if (a_kcontig) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      As[grp][k][m] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = tid + kGemmGroup * r;
      int n, k;
      if (b_ncontig) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      Bs[grp][k][n] = rb[r];
    }
  };
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kGemmGroup) : "memory"); };
  const int kstep = 2 * BK;
  if (grp * BK < K) fetch(grp * BK);
  for (int k0 = grp * BK; k0 < K; k0 += kstep) {
    stash();
    group_sync();
    if (k0 + kstep < K) fetch(k0 + kstep);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[grp][k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[grp][k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    group_sync();
  }
  if (grp == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[tid][i * 4 + j] = acc[i][j];
  }
  __syncthreads();
  if (grp == 1) return;
  float* Cb = C + b * sCb + hd * sCh;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* cp = Cb + (long long)m * sCm + (long long)n * sCn;
      float v = alpha * (acc[i][j] + red[tid][i * 4 + j]);
      if (beta != 0.f) v += beta * (*cp);
      if (round_out) v = round_tf32(v);
      *cp = v;
    }
  }
}

// One warp per row.
__global__ void softmax_rows_kernel(float* __restrict__ S, int T, long long rows, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* p = S + row * T;
  float mx = -INFINITY;
  for (int j = lane; j < T; j += 32) mx = fmaxf(mx, p[j] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float e = expf(p[j] * scale - mx);
    p[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < T; j += 32) p[j] *= inv;
}

// rows of X are [batch][heads * T]; P0 holds the primal probabilities [heads * T][T]
__global__ void softmax_lin_rows_kernel(const float* __restrict__ P0, float* __restrict__ X, int T,
                                        long long rows, long long prow_mod, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = P0 + (row % prow_mod) * T;
  float* x = X + row * T;
  float dot = 0.f;
  for (int j = lane; j < T; j += 32) dot += p[j] * x[j];
  dot = warp_sum(dot);
  for (int j = lane; j < T; j += 32) x[j] = scale * p[j] * (x[j] - dot);
}

int check_tokens(const View& v, const char* what) {
  LOCO_REQUIRE(v.sH == (long long)v.W * v.sW, "%s: tokens must be uniformly strided", what);
  return 0;
}

}  // namespace

int attention_init() {
  static bool done[kMaxDevices] = {false};
  if (!first_time_on_device(done)) return 0;
  const int co = cudaSharedmemCarveoutMaxShared;
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(batched_gemm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(softmax_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(softmax_lin_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
  return attention_tc_init();
}

int batched_gemm(GemmOperand A, GemmOperand B, float* C, long long sCm, long long sCn, long long sCb,
                 long long sCh, int M, int N, int K, int batch, int heads, float alpha, float beta,
                 int round_out, cudaStream_t s) {
  if (batch <= 0) return 0;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch * heads);
  batched_gemm_kernel<<<grid, kGemmThreads, 0, s>>>(A, B, C, sCm, sCn, sCb, sCh, heads, M, N, K, alpha,
                                                    beta, round_out, A.s1 == 1 ? 1 : 0, B.s1 == 1 ? 1 : 0);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int softmax_rows(float* S, int T, int batch, float scale, cudaStream_t s) {
  if (batch <= 0) return 0;
  const long long rows = (long long)batch * T;
  softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(S, T, rows, scale);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int softmax_lin_rows(const float* P0, float* X, int T, int batch, int heads, float scale,
                     cudaStream_t s) {
  if (batch <= 0) return 0;
  const long long rows = (long long)batch * heads * T;
  softmax_lin_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(P0, X, T, rows,
                                                                     (long long)heads * T, scale);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace {
// Channel geometry of the fused q|k|v tensor.
//   head_ch == 0: DDPM AttnBlock, one head over all C channels, q | k | v blocks of C channels,
//                 scores scaled by C^-1/2 (ddpm/diffusion.py:941-966);
//   head_ch  > 0: guided-diffusion QKVAttentionLegacy, heads split BEFORE q/k/v: head h owns the
//                 channels [3 h dh, 3 (h+1) dh) = q | k | v of dh each, scores scaled by dh^-1/2
//                 (q and k each by dh^-1/4; guided_diffusion/unet.py:339-356).
struct HeadGeom { int heads, dh; long long qo, ko, vo, hs; float scale; };
HeadGeom head_geom(int C, int head_ch) {
  HeadGeom g;
  if (head_ch <= 0) { g.heads = 1; g.dh = C; g.qo = 0; g.ko = C; g.vo = 2 * C; g.hs = 0; }
  else { g.heads = C / head_ch; g.dh = head_ch; g.qo = 0; g.ko = head_ch; g.vo = 2 * head_ch; g.hs = 3 * head_ch; }
  g.scale = 1.0f / sqrtf((float)g.dh);
  return g;
}
}  // namespace

int attention_forward(View qkv, int n_primal, int head_ch, float* S, View o, cudaStream_t s) {
  LOCO_TRY(check_tokens(qkv, "attention_forward(qkv)"));
  LOCO_TRY(check_tokens(o, "attention_forward(o)"));
  const int C = qkv.C / 3;
  const int T = qkv.H * qkv.W;
  const int N = qkv.N;
  const int nt = N - n_primal;
  LOCO_REQUIRE(o.C == C && o.N == N, "attention_forward: shape mismatch");
  LOCO_REQUIRE(nt == 0 || n_primal == 1, "attention_forward: tangents need exactly one primal row");
  LOCO_REQUIRE(head_ch <= 0 || C % head_ch == 0, "attention_forward: %d channels, heads of %d", C, head_ch);
  if (attention_tc_eligible(T, C, head_ch) && !qkv.half && !o.half)
    return attention_forward_tc(qkv, n_primal, head_ch, S, o, s);
  const HeadGeom G = head_geom(C, head_ch);
  const int Hh = G.heads, D = G.dh;
  const float scale = G.scale;
  const long long ts = qkv.sW;                  // token stride
  const long long TT = (long long)T * T;        // one head's score matrix
  const long long HTT = TT * Hh;                // one batch row's score matrices
  const float* q = qkv.ptr + G.qo;
  const float* k = qkv.ptr + G.ko;
  const float* v = qkv.ptr + G.vo;
  // primal rows: S = q k^T ; P = softmax(scale S) ; o = P v
  LOCO_TRY(batched_gemm({q, ts, 1, qkv.sN, G.hs}, {k, 1, ts, qkv.sN, G.hs}, S, T, 1, HTT, TT, T, T, D,
                        n_primal, Hh, 1.f, 0.f, 0, s));
  LOCO_TRY(softmax_rows(S, T, n_primal * Hh, scale, s));
  LOCO_TRY(batched_gemm({S, T, 1, HTT, TT}, {v, ts, 1, qkv.sN, G.hs}, o.ptr, o.sW, 1, o.sN, D, T, D, T,
                        n_primal, Hh, 1.f, 0.f, 1, s));
  if (nt > 0) {
    const float* qd = q + qkv.sN;
    const float* kd = k + qkv.sN;
    const float* vd = v + qkv.sN;
    float* Sd = S + HTT;
    float* od = o.ptr + o.sN;
    // dS = dq k0^T + q0 dk^T
    LOCO_TRY(batched_gemm({qd, ts, 1, qkv.sN, G.hs}, {k, 1, ts, 0, G.hs}, Sd, T, 1, HTT, TT, T, T, D, nt,
                          Hh, 1.f, 0.f, 0, s));
    LOCO_TRY(batched_gemm({q, ts, 1, 0, G.hs}, {kd, 1, ts, qkv.sN, G.hs}, Sd, T, 1, HTT, TT, T, T, D, nt,
                          Hh, 1.f, 1.f, 0, s));
    // dP = scale * P0 o (dS - rowsum(P0 o dS))
    LOCO_TRY(softmax_lin_rows(S, Sd, T, nt, Hh, scale, s));
    // do = P0 dv + dP v0
    LOCO_TRY(batched_gemm({S, T, 1, 0, TT}, {vd, ts, 1, qkv.sN, G.hs}, od, o.sW, 1, o.sN, D, T, D, T, nt,
                          Hh, 1.f, 0.f, 0, s));
    LOCO_TRY(batched_gemm({Sd, T, 1, HTT, TT}, {v, ts, 1, 0, G.hs}, od, o.sW, 1, o.sN, D, T, D, T, nt, Hh,
                          1.f, 1.f, 1, s));
  }
  return 0;
}

int attention_vjp(View go, View qkv0, int head_ch, const float* P0, float* gP, View gqkv,
                  cudaStream_t s) {
  LOCO_TRY(check_tokens(go, "attention_vjp(go)"));
  LOCO_TRY(check_tokens(qkv0, "attention_vjp(qkv0)"));
  LOCO_TRY(check_tokens(gqkv, "attention_vjp(gqkv)"));
  const int C = qkv0.C / 3;
  const int T = qkv0.H * qkv0.W;
  const int K = go.N;
  LOCO_REQUIRE(go.C == C && gqkv.C == 3 * C && gqkv.N == K, "attention_vjp: shape mismatch");
  if (attention_tc_eligible(T, C, head_ch) && !go.half && !qkv0.half && !gqkv.half)
    return attention_vjp_tc(go, qkv0, head_ch, P0, gP, gqkv, s);
  const HeadGeom G = head_geom(C, head_ch);
  const int Hh = G.heads, D = G.dh;
  const float scale = G.scale;
  const long long ts = qkv0.sW, gts = gqkv.sW;
  const long long TT = (long long)T * T;
  const long long HTT = TT * Hh;
  const float* q0 = qkv0.ptr + G.qo;
  const float* k0 = qkv0.ptr + G.ko;
  const float* v0 = qkv0.ptr + G.vo;
  float* gq = gqkv.ptr + G.qo;
  float* gk = gqkv.ptr + G.ko;
  float* gv = gqkv.ptr + G.vo;
  // gv = P0^T go      : (j, c) = sum_i P0[i][j] go[i][c]
  LOCO_TRY(batched_gemm({P0, 1, T, 0, TT}, {go.ptr, go.sW, 1, go.sN, D}, gv, gts, 1, gqkv.sN, G.hs, T, D,
                        T, K, Hh, 1.f, 0.f, 1, s));
  // gP = go v0^T      : (i, j) = sum_c go[i][c] v0[j][c]
  LOCO_TRY(batched_gemm({go.ptr, go.sW, 1, go.sN, D}, {v0, 1, ts, 0, G.hs}, gP, T, 1, HTT, TT, T, T, D, K,
                        Hh, 1.f, 0.f, 0, s));
  // gS = scale * P0 o (gP - rowsum(P0 o gP))   (gradient w.r.t. the unscaled scores)
  LOCO_TRY(softmax_lin_rows(P0, gP, T, K, Hh, scale, s));
  // gq = gS k0        : (i, c) = sum_j gS[i][j] k0[j][c]
  LOCO_TRY(batched_gemm({gP, T, 1, HTT, TT}, {k0, ts, 1, 0, G.hs}, gq, gts, 1, gqkv.sN, G.hs, T, D, T, K,
                        Hh, 1.f, 0.f, 1, s));
  // gk = gS^T q0      : (j, c) = sum_i gS[i][j] q0[i][c]
  LOCO_TRY(batched_gemm({gP, 1, T, HTT, TT}, {q0, ts, 1, 0, G.hs}, gk, gts, 1, gqkv.sN, G.hs, T, D, T, K,
                        Hh, 1.f, 0.f, 1, s));
  return 0;
}

}  // namespace loco
