// Bandwidth-bound layer kernels: see layers.cuh.  All activations are channels-last fp32; every
// kernel moves 16-byte vectors with consecutive threads on consecutive channels (coalesced), does
// its group reductions with shared-memory + one double atomic per (row, group, block), and sizes
// its grid from the tensor so large layers launch many waves over the 148 SMs.
#include "layers.cuh"
#include <math.h>
#include <algorithm>
#include <type_traits>

namespace loco {

namespace {

__device__ __forceinline__ float sigmoidf_(float u) { return 1.0f / (1.0f + expf(-u)); }
__device__ __forceinline__ float silu_f(float u) { return u * sigmoidf_(u); }
// d/du [u * sigmoid(u)]
__device__ __forceinline__ float silu_grad(float u) {
  const float s = sigmoidf_(u);
  return s * (1.0f + u * (1.0f - s));
}

// ------------------------------------------------------------------------------------------------
// Weight packing
// ------------------------------------------------------------------------------------------------
// destination element: tf32-rounded fp32 (tensor core kind::tf32) or fp16 (kind::f16)
__device__ __forceinline__ void put_w(float* dst, long long i, float v) { dst[i] = round_tf32(v); }
__device__ __forceinline__ void put_w(__half* dst, long long i, float v) { dst[i] = __float2half_rn(v); }

template <typename T>
__global__ void pack_fprop_kernel(const float* __restrict__ w, T* __restrict__ dst, int Cout,
                                  int Cin, int taps) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int t = (int)((i / Cin) % taps);
    const int co = (int)(i / ((long long)Cin * taps));
    put_w(dst, i, w[((long long)co * Cin + ci) * taps + t]);
  }
}
template <typename T>
__global__ void pack_dgrad_kernel(const float* __restrict__ w, T* __restrict__ dst, int Cout,
                                  int Cin, int taps, int cout_total, int co_off) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int t = (int)((i / Cout) % taps);
    const int ci = (int)(i / ((long long)Cout * taps));
    put_w(dst, ((long long)ci * taps + t) * cout_total + co_off + co, w[((long long)co * Cin + ci) * taps + t]);
  }
}
__global__ void pack_edge_kernel(const float* __restrict__ w, float* __restrict__ dst, int C,
                                 int in_is_3) {
  const int total = 9 * 3 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C;
    const int j = (i / C) % 3;
    const int t = i / (3 * C);
    // conv_in: w[c][j][t]; conv_out: w[j][c][t]
    dst[i] = in_is_3 ? w[(c * 3 + j) * 9 + t] : w[(j * C + c) * 9 + t];
  }
}

// ------------------------------------------------------------------------------------------------
// Edge convolutions
// ------------------------------------------------------------------------------------------------
constexpr int kEdgeIters = 16;
template <bool F16>
__global__ void edge_expand_kernel(const float* __restrict__ in3, const float* __restrict__ We,
                                   const float* __restrict__ bias, int bias_rows, View out,
                                   int flip, int round_out, const float* __restrict__ scale_dev, int scale_from) {
  extern __shared__ float sw[];   // [27][C]
  const int C = out.C, H = out.H, W = out.W;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = We[i];
  __syncthreads();
  const int cvn = C >> 2;
  const int ppb = blockDim.x / cvn;
  const int cv = threadIdx.x % cvn;
  const int n = blockIdx.y;
  // the 27 x C weight slab is staged once per block and amortised over kEdgeIters pixel groups
  for (int it = 0; it < kEdgeIters; ++it) {
  const long long pix = ((long long)blockIdx.x * kEdgeIters + it) * ppb + threadIdx.x / cvn;
  if (pix >= (long long)H * W) return;
  const int y = (int)(pix / W), x = (int)(pix % W);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias && n < bias_rows) acc = *reinterpret_cast<const float4*>(bias + cv * 4);
  const float* src = in3 + (long long)n * 3 * H * W;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = y + (flip ? 1 - r : r - 1);
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int xx = x + (flip ? 1 - s : s - 1);
      if (xx < 0 || xx >= W) continue;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float v = __ldg(src + ((long long)j * H + yy) * W + xx);
        const float4 w = *reinterpret_cast<const float4*>(&sw[((r * 3 + s) * 3 + j) * C + cv * 4]);
        acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y);
        acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
      }
    }
  }
  if (scale_dev != nullptr && n >= scale_from) {
    const float sc = scale_dev[0];
    acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
  }
  if (round_out) {
    acc.x = round_tf32(acc.x); acc.y = round_tf32(acc.y);
    acc.z = round_tf32(acc.z); acc.w = round_tf32(acc.w);
  }
  st4t<F16>(out.ptr, n * out.sN + y * out.sH + x * out.sW + cv * 4, acc);
  }
}

// One warp per output pixel; lanes stride over channels, three warp reductions per pixel.
template <bool F16>
__global__ void edge_reduce_kernel(View in, const float* __restrict__ Wr,
                                   const float* __restrict__ bias, int bias_rows,
                                   float* __restrict__ out3, int flip, const float* __restrict__ scale_dev,
                                   int scale_from) {
  extern __shared__ float sw[];   // [27][C]
  const int C = in.C, H = in.H, W = in.W;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = Wr[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int n = blockIdx.y;
  for (int it = 0; it < kEdgeIters; ++it) {
  const long long pix = ((long long)blockIdx.x * kEdgeIters + it) * wpb + (threadIdx.x >> 5);
  if (pix >= (long long)H * W) return;
  const int y = (int)(pix / W), x = (int)(pix % W);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int r = 0; r < 3; ++r) {
    const int yy = y + (flip ? 1 - r : r - 1);
    if (yy < 0 || yy >= H) continue;
    for (int s = 0; s < 3; ++s) {
      const int xx = x + (flip ? 1 - s : s - 1);
      if (xx < 0 || xx >= W) continue;
      const long long soff = n * in.sN + yy * in.sH + xx * in.sW;
      const float* w = &sw[(r * 3 + s) * 3 * C];
      for (int c = lane * 4; c < C; c += 128) {
        const float4 v = ld4t<F16>(in.ptr, soff + c);
        const float4 w0 = *reinterpret_cast<const float4*>(w + c);
        const float4 w1 = *reinterpret_cast<const float4*>(w + C + c);
        const float4 w2 = *reinterpret_cast<const float4*>(w + 2 * C + c);
        a0 += v.x * w0.x + v.y * w0.y + v.z * w0.z + v.w * w0.w;
        a1 += v.x * w1.x + v.y * w1.y + v.z * w1.z + v.w * w1.w;
        a2 += v.x * w2.x + v.y * w2.y + v.z * w2.z + v.w * w2.w;
      }
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if (lane == 0) {
    const bool ub = bias && n < bias_rows;
    const float sc = (scale_dev != nullptr && n >= scale_from) ? scale_dev[1] : 1.f;
    float* dst = out3 + (long long)n * 3 * H * W + (long long)y * W + x;
    dst[0] = a0 * sc + (ub ? bias[0] : 0.f);
    dst[(long long)H * W] = a1 * sc + (ub ? bias[1] : 0.f);
    dst[2LL * H * W] = a2 * sc + (ub ? bias[2] : 0.f);
  }
  }
}

// ---- fast paths for C == 128 (every shipped architecture: conv_in / conv_out have `ch` = 128
// channels on the wide side).  A warp owns a strip of 4 output pixels of one image row; lane l owns
// channels [4l, 4l+4).  The 3 x 6 input window of the strip is loaded once and the 27 weight vectors
// are read from shared memory once per strip (not once per pixel), which turns the old
// shared-memory-bound kernels into HBM-streaming ones.
constexpr int kStrip = 4;
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <bool F16>
__global__ void __launch_bounds__(256, 2)
edge_reduce128_kernel(View in, const float* __restrict__ Wr, const float* __restrict__ bias,
                      int bias_rows, float* __restrict__ out3, int flip, const float* __restrict__ scale_dev,
                      int scale_from) {
  __shared__ float4 sw[27 * 32];   // [tap][j][channel quad]
  const int H = in.H, W = in.W;
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(Wr)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int sx = W / kStrip;
  const long long nstrips = (long long)in.N * H * sx;
  for (long long st = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); st < nstrips;
       st += (long long)gridDim.x * wpb) {
    const int x0 = (int)(st % sx) * kStrip;
    const int y = (int)((st / sx) % H);
    const int n = (int)(st / ((long long)sx * H));
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    // input window rows y-1..y+1, columns x0-1..x0+4 (zero outside the image), one row at a time
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = y + r - 1;
      float4 win[kStrip + 2];
#pragma unroll
      for (int c = 0; c < kStrip + 2; ++c) {
        const int xx = x0 + c - 1;
        win[c] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                     ? ld4t<F16>(in.ptr, (long long)n * in.sN + (long long)yy * in.sH + (long long)xx * in.sW + lane * 4)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // window offset (r, c) <-> filter tap: forward (dy,dx) = (r-1,c-1) is tap (r,c); the flipped
        // (data-gradient) form reads in(y - (r'-1), x - (c'-1)), i.e. tap (2-r, 2-c)
        const int t = flip ? (2 - r) * 3 + (2 - c) : r * 3 + c;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 w = sw[(t * 3 + j) * 32 + lane];
#pragma unroll
          for (int q = 0; q < kStrip; ++q) {
            const float4 v = win[q + c];
            acc[q * 3 + j] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[q * 3 + j]))));
          }
        }
      }
    }
    const float tot = butterfly16(acc, lane);
    const int idx = lane >> 1;           // value index q * 3 + j
    if ((lane & 1) == 0 && idx < kStrip * 3) {
      const int q = idx / 3, j = idx % 3;
      const bool ub = bias && n < bias_rows;
      const float sc = (scale_dev != nullptr && n >= scale_from) ? scale_dev[1] : 1.f;
      out3[((long long)n * 3 + j) * H * W + (long long)y * W + x0 + q] = tot * sc + (ub ? bias[j] : 0.f);
    }
  }
}

template <bool F16>
__global__ void __launch_bounds__(256)
edge_expand128_kernel(const float* __restrict__ in3, const float* __restrict__ We,
                      const float* __restrict__ bias, int bias_rows, View out, int flip,
                      int round_out, const float* __restrict__ scale_dev, int scale_from) {
  __shared__ float4 sw[27 * 32];   // [tap][j][channel quad]
  const int H = out.H, W = out.W;
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(We)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int sx = W / kStrip;
  const long long nstrips = (long long)out.N * H * sx;
  const float4 b4 = bias ? ld4(bias + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long st = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); st < nstrips;
       st += (long long)gridDim.x * wpb) {
    const int x0 = (int)(st % sx) * kStrip;
    const int y = (int)((st / sx) % H);
    const int n = (int)(st / ((long long)sx * H));
    const float* src = in3 + (long long)n * 3 * H * W;
    // lanes 0..17 each fetch one (row, column) of the 3 x 6 window of the three image planes
    float wv[3] = {0.f, 0.f, 0.f};
    if (lane < 3 * (kStrip + 2)) {
      const int r = lane / (kStrip + 2), c = lane % (kStrip + 2);
      const int yy = y + r - 1, xx = x0 + c - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
#pragma unroll
        for (int j = 0; j < 3; ++j) wv[j] = __ldg(src + ((long long)j * H + yy) * W + xx);
      }
    }
    float4 acc[kStrip];
    const bool ub = bias && n < bias_rows;
#pragma unroll
    for (int q = 0; q < kStrip; ++q) acc[q] = ub ? b4 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int t = flip ? (2 - r) * 3 + (2 - c) : r * 3 + c;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 w = sw[(t * 3 + j) * 32 + lane];
#pragma unroll
          for (int q = 0; q < kStrip; ++q) {
            const float v = __shfl_sync(0xffffffffu, wv[j], r * (kStrip + 2) + q + c);
            acc[q].x = fmaf(v, w.x, acc[q].x); acc[q].y = fmaf(v, w.y, acc[q].y);
            acc[q].z = fmaf(v, w.z, acc[q].z); acc[q].w = fmaf(v, w.w, acc[q].w);
          }
        }
      }
#pragma unroll
    const float sc = (scale_dev != nullptr && n >= scale_from) ? scale_dev[0] : 1.f;
#pragma unroll
    for (int q = 0; q < kStrip; ++q) {
      float4 o = acc[q];
      o.x *= sc; o.y *= sc; o.z *= sc; o.w *= sc;
      if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
      st4t<F16>(out.ptr, (long long)n * out.sN + (long long)y * out.sH + (long long)(x0 + q) * out.sW + lane * 4, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm
// ------------------------------------------------------------------------------------------------
constexpr int kGroups = 32;
constexpr int kRC = 6;          // batch rows processed per register chunk (1 primal + 5 tangents)
template <int MODE, bool F16> struct GnRows { static constexpr int value = kRC; };
constexpr int kGnMaxRows = 96;  // rows whose per-group scalars fit the shared table

// A thread owns VEC consecutive channels = one 16-byte vector (4 fp32 / 8 fp16 channels) of a pixel.
// With 8-byte vectors the fp16 kernels were latency-bound at the speed of the fp32 ones (2.9 / 3.6
// TB/s on the 6 x 256^2 x 128 JVP site); 16 bytes per load keep the bytes in flight per thread equal.
template <bool F16> struct GnVec { static constexpr int value = F16 ? 8 : 4; };
__host__ __device__ inline int gn_block_dim(int C, int vec) { return (256 % (C / vec) == 0) ? 256 : 192; }

struct GnGeom {
  int block, pstep, ppb, nblk;
};
// A thread owns one channel vector of a pixel and walks ALL batch rows of that pixel: the primal
// value is loaded (and its sigmoid evaluated) once and shared by the k tangent / cotangent rows, and
// the loads of a row chunk are issued together.
// `resident` = blocks of the kernel that fit the GPU at once (occupancy x SMs): the grid is one
// full wave of persistent blocks that stride over the pixel chunks.
inline GnGeom gn_geom(int C, int vec, long long HW, int resident) {
  GnGeom g;
  g.block = gn_block_dim(C, vec);
  g.pstep = g.block / (C / vec);
  g.ppb = g.pstep;
  const long long nchunks = (HW + g.pstep - 1) / g.pstep;
  g.nblk = (int)(nchunks < resident ? nchunks : resident);
  return g;
}

__device__ __forceinline__ float fast_sigmoid(float u) { return __fdividef(1.0f, 1.0f + __expf(-u)); }

// (mean, rstd) of a primal row's group from its (sum x, sum x^2)
__device__ __forceinline__ float2 mean_rstd(const double* st, double cnt, float eps) {
  const double mu = st[0] / cnt;
  const double var = st[1] / cnt - mu * mu;
  return make_float2((float)mu, (float)(1.0 / sqrt((var > 0 ? var : 0) + (double)eps)));
}

// VEC consecutive channels of an activation tensor <-> registers
template <bool F16>
__device__ __forceinline__ void ldvec(const float* base, long long off, float (&v)[GnVec<F16>::value]) {
  if (F16) {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(base) + off);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      v[2 * k] = f.x; v[2 * k + 1] = f.y;
    }
  } else {
    const float4 f = *reinterpret_cast<const float4*>(base + off);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
}
template <bool F16>
__device__ __forceinline__ void stvec(float* base, long long off, const float (&v)[GnVec<F16>::value]) {
  if (F16) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
      w[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(base) + off) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    *reinterpret_cast<float4*>(base + off) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// the same 16 bytes kept packed (4 registers) until they are used: the kernels below issue the loads
// of a whole row chunk first and unpack row by row, so that the bytes in flight per thread are not
// bounded by the register cost of the unpacked fp32 values
template <bool F16>
__device__ __forceinline__ uint4 ldraw(const float* base, long long off) {
  return F16 ? *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(base) + off)
             : *reinterpret_cast<const uint4*>(base + off);
}
template <bool F16>
__device__ __forceinline__ void unpack(const uint4& u, float (&v)[GnVec<F16>::value]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  if (F16) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      v[2 * k] = f.x; v[2 * k + 1] = f.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __uint_as_float(w[k]);
  }
}

// ---- statistics ---------------------------------------------------------------------------------
// mode 0 (forward/JVP): rows < n_primal: (sum x, sum x^2); tangent rows: (sum dx, sum x0 dx).
// mode 1 (VJP): a = gamma * act'(u) * gy;  rows: (sum a, sum x a), x = primal input (x has 1 row).
// A thread's VEC channels are VEC / 4 quads; a quad never straddles a group (cg is a multiple of 4).
template <int MODE, bool F16>
__global__ void __launch_bounds__(256)
gn_stats_kernel(View x, int n_primal, View gy, const double* __restrict__ pstats,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                double* __restrict__ stats, int lin) {
  constexpr int RC = GnRows<MODE, F16>::value;
  constexpr int VEC = GnVec<F16>::value;
  constexpr int NQ = VEC / 4;                      // quads per thread
  __shared__ float part[2 * RC][256 * NQ];         // [row, which][pixel row of the block][quad of the pixel]
  const int C = x.C;
  const int cvn = C / VEC;
  const int qn = C >> 2;                           // quads per pixel
  const int cg = C / kGroups;
  const int cv = threadIdx.x % cvn;
  const int prow = threadIdx.x / cvn;
  const int pstep = blockDim.x / cvn;
  const int HW = x.H * x.W;
  // persistent blocks: chunk c covers pixels [c * pstep, (c + 1) * pstep); a block walks the chunks
  // c = blockIdx.x, blockIdx.x + gridDim.x, ... so the grid is one full wave and has no tail
  // lin: every view has sH == W * sW, so a pixel's offset is p * sW (no division in the loop)
  const View& rows = (MODE == 0) ? x : gy;
  const int N = rows.N;
  const bool jvp = (MODE == 0) && (n_primal < N);   // row 0 primal, rows 1.. tangents

  // per-element factors of the primal point (VJP: a = fac * gy)
  float fac[VEC];
  float mu[NQ], rstd[NQ];
  float gs[VEC], bs[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { fac[i] = 1.f; gs[i] = 1.f; bs[i] = 0.f; }
#pragma unroll
  for (int q = 0; q < NQ; ++q) { mu[q] = 0.f; rstd[q] = 1.f; }
  if (MODE == 1) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int g = (cv * VEC + 4 * q) / cg;
      const float2 mr = mean_rstd(pstats + g * 2, (double)HW * cg, eps);
      mu[q] = mr.x; rstd[q] = mr.y;
      const float4 ga = ld4(gamma + cv * VEC + 4 * q), be = ld4(beta + cv * VEC + 4 * q);
      gs[4 * q] = ga.x; gs[4 * q + 1] = ga.y; gs[4 * q + 2] = ga.z; gs[4 * q + 3] = ga.w;
      bs[4 * q] = be.x; bs[4 * q + 1] = be.y; bs[4 * q + 2] = be.z; bs[4 * q + 3] = be.w;
    }
  }

  using Elem = typename std::conditional<F16, __half, float>::type;
  const Elem* xbase = reinterpret_cast<const Elem*>(x.ptr) + cv * VEC;
  for (int n0 = 0; n0 < N; n0 += RC) {
    float s1[RC][NQ], s2[RC][NQ];
#pragma unroll
    for (int r = 0; r < RC; ++r)
#pragma unroll
      for (int q = 0; q < NQ; ++q) { s1[r][q] = 0.f; s2[r][q] = 0.f; }
    // row pointers; a chunk's rows past N re-read row N - 1 (an L1 hit) and their sums are never
    // published, so the loop body carries no per-row predicates
    const Elem* rbase[RC];
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      const int nr = (n0 + r < N) ? n0 + r : N - 1;
      rbase[r] = reinterpret_cast<const Elem*>(rows.ptr) + (long long)nr * rows.sN + cv * VEC;
    }
    const bool x_is_row0 = (MODE == 0 && n0 == 0);
    const bool need_x = (jvp || MODE == 1) && !x_is_row0;
    for (int p = blockIdx.x * pstep + prow; p < HW; p += gridDim.x * pstep) {
      long long xoff, roff;
      if (lin) {
        xoff = (long long)p * x.sW; roff = (long long)p * rows.sW;
      } else {
        const int y = p / x.W, xx = p - y * x.W;
        xoff = (long long)y * x.sH + (long long)xx * x.sW;
        roff = (long long)y * rows.sH + (long long)xx * rows.sW;
      }
      // all loads of the chunk are issued before the first use
      uint4 rawx = make_uint4(0, 0, 0, 0);
      if (need_x) rawx = *reinterpret_cast<const uint4*>(xbase + xoff);
      uint4 raw[RC];
#pragma unroll
      for (int r = 0; r < RC; ++r) raw[r] = *reinterpret_cast<const uint4*>(rbase[r] + roff);
      float xs[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) xs[i] = 0.f;
      if (jvp || MODE == 1) unpack<F16>(x_is_row0 ? raw[0] : rawx, xs);
      if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float u = gs[i] * ((xs[i] - mu[i >> 2]) * rstd[i >> 2]) + bs[i];
          float d = 1.0f;
          if (silu) { const float sg = fast_sigmoid(u); d = sg * (1.0f + u * (1.0f - sg)); }
          fac[i] = gs[i] * d;
        }
      }
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        const bool tangent = jvp && (n0 + r >= n_primal);
        float v[VEC];
        unpack<F16>(raw[r], v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          if (MODE == 0) {
            s1[r][i >> 2] += v[i];
            s2[r][i >> 2] += v[i] * (tangent ? xs[i] : v[i]);
          } else {
            const float av = fac[i] * v[i];
            s1[r][i >> 2] += av;
            s2[r][i >> 2] += av * xs[i];
          }
        }
      }
    }
    // fixed-order block reduction (bit-reproducible per block); fp64 atomics across blocks
#pragma unroll
    for (int r = 0; r < RC; ++r)
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        part[2 * r][prow * qn + cv * NQ + q] = s1[r][q];
        part[2 * r + 1][prow * qn + cv * NQ + q] = s2[r][q];
      }
    __syncthreads();
    for (int e = threadIdx.x; e < RC * kGroups * 2; e += blockDim.x) {
      const int r = e / (kGroups * 2), gw = e % (kGroups * 2);
      if (n0 + r >= N) continue;
      const int gg = gw >> 1, which = gw & 1;
      const int q0 = gg * (cg >> 2), q1 = q0 + (cg >> 2);
      float acc = 0.f;
      for (int pr2 = 0; pr2 < pstep; ++pr2)
        for (int c = q0; c < q1; ++c) acc += part[2 * r + which][pr2 * qn + c];
      atomicAdd(&stats[(long long)(n0 + r) * kGroups * 2 + gw], (double)acc);
    }
    __syncthreads();
  }
}

// ---- apply ----------------------------------------------------------------------------------------
// mode 0: y = act(gn(x)) for primal rows, JVP rule for tangent rows.  mode 1: VJP rule.
// RC rows of a pixel are loaded together (packed, 4 registers per 16-byte vector) and then unpacked,
// transformed and stored one at a time.  The VJP kernel has two instantiations: RC = 8 for the plain
// rule, RC = 4 when an addend and / or the previous gx ride along (up to three loads per row).
template <int MODE, bool F16, int RC>
__global__ void __launch_bounds__(256, 2)
gn_apply_kernel(View x, int n_primal, View gy, const double* __restrict__ pstats,
                const double* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int silu, int round_out,
                const float* __restrict__ addend, long long add_sN, long long add_sH,
                long long add_sW, int accumulate, View out, int lin) {
  pdl_trigger();   // a dependent conv launch may start its prologue now (it waits for this grid's completion)
  constexpr int VEC = GnVec<F16>::value;
  constexpr int NQ = VEC / 4;
  constexpr bool EXTRA = (MODE == 1 && RC == 4);
  // per (row, group) scalars: primal rows (mean, rstd); tangent / cotangent rows (m1, m2)
  __shared__ float2 tab[kGnMaxRows][kGroups];
  const int C = x.C;
  const int cvn = C / VEC;
  const int cg = C / kGroups;
  const int cv = threadIdx.x % cvn;
  const int prow = threadIdx.x / cvn;
  const int pstep = blockDim.x / cvn;
  int g[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) g[q] = (cv * VEC + 4 * q) / cg;
  const int HW = x.H * x.W;
  const double cnt = (double)HW * cg;
  const View& rows = (MODE == 0) ? x : gy;
  const int N = rows.N;
  const bool jvp = (MODE == 0) && (n_primal < N);
  const double* prim = (MODE == 0) ? stats : pstats;    // statistics of primal row 0
  for (int e = threadIdx.x; e < N * kGroups; e += blockDim.x) {
    const int n = e / kGroups, gg = e % kGroups;
    const bool is_primal = (MODE == 0) && (n < n_primal);
    if (is_primal) {
      tab[n][gg] = mean_rstd(stats + ((long long)n * kGroups + gg) * 2, cnt, eps);
    } else {
      const float2 mr = mean_rstd(prim + gg * 2, cnt, eps);
      const double sa = stats[((long long)n * kGroups + gg) * 2];
      const double sxa = stats[((long long)n * kGroups + gg) * 2 + 1];
      tab[n][gg] = make_float2((float)(sa / cnt),
                               (float)((sxa - (double)mr.x * sa) * (double)mr.y / cnt));
    }
  }
  __syncthreads();
  float gs[VEC], bs[VEC];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 ga4 = ld4(gamma + cv * VEC + 4 * q), be4 = ld4(beta + cv * VEC + 4 * q);
    gs[4 * q] = ga4.x; gs[4 * q + 1] = ga4.y; gs[4 * q + 2] = ga4.z; gs[4 * q + 3] = ga4.w;
    bs[4 * q] = be4.x; bs[4 * q + 1] = be4.y; bs[4 * q + 2] = be4.z; bs[4 * q + 3] = be4.w;
  }
  float2 mr0[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    mr0[q] = make_float2(0.f, 1.f);
    if (jvp || MODE == 1) mr0[q] = (MODE == 0) ? tab[0][g[q]] : mean_rstd(pstats + g[q] * 2, cnt, eps);
  }

  const int nchunks = (HW + pstep - 1) / pstep;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int p = ch * pstep + prow;
    if (p >= HW) break;
    int y = 0, xx = p;
    if (!lin) { y = p / x.W; xx = p - y * x.W; }
    const long long xoff = (long long)y * x.sH + (long long)xx * x.sW + cv * VEC;
    const long long roff = (long long)y * rows.sH + (long long)xx * rows.sW + cv * VEC;
    const long long ooff = (long long)y * out.sH + (long long)xx * out.sW + cv * VEC;
    const long long aoff = (long long)y * add_sH + (long long)xx * add_sW + cv * VEC;
    // primal-point quantities shared by every tangent / cotangent row of this pixel
    float xh[VEC], coef[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { xh[i] = 0.f; coef[i] = 1.f; }
    uint4 rawx = make_uint4(0, 0, 0, 0);
    if (jvp || MODE == 1) rawx = ldraw<F16>(x.ptr, xoff);
    for (int n0 = 0; n0 < N; n0 += RC) {
      uint4 rv[RC], re[EXTRA ? RC : 1], rc[EXTRA ? RC : 1];
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        rv[r] = make_uint4(0, 0, 0, 0);
        if (EXTRA) { re[r] = rv[r]; rc[r] = rv[r]; }
        if (n0 + r < N) {
          rv[r] = ldraw<F16>(rows.ptr, (long long)(n0 + r) * rows.sN + roff);
          if (EXTRA && addend) re[r] = ldraw<F16>(addend, (long long)(n0 + r) * add_sN + aoff);
          if (EXTRA && accumulate) rc[r] = ldraw<F16>(out.ptr, (long long)(n0 + r) * out.sN + ooff);
        }
      }
      if (n0 == 0 && (jvp || MODE == 1)) {
        float xs[VEC];
        unpack<F16>(rawx, xs);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          xh[i] = (xs[i] - mr0[i >> 2].x) * mr0[i >> 2].y;
          const float u = gs[i] * xh[i] + bs[i];
          float d = 1.0f;
          if (silu) { const float sg = fast_sigmoid(u); d = sg * (1.0f + u * (1.0f - sg)); }
          coef[i] = d * gs[i];          // act'(u) * gamma
        }
      }
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        const int n = n0 + r;
        if (n >= N) break;
        float2 t2[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) t2[q] = tab[n][g[q]];
        float v[VEC], o[VEC];
        unpack<F16>(rv[r], v);
        if (MODE == 0 && n < n_primal) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float u = gs[i] * ((v[i] - t2[i >> 2].x) * t2[i >> 2].y) + bs[i];
            o[i] = silu ? u * fast_sigmoid(u) : u;
          }
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float2 tq = t2[i >> 2];
            const float rs0 = mr0[i >> 2].y;
            if (MODE == 0) o[i] = coef[i] * rs0 * (v[i] - tq.x - xh[i] * tq.y);
            else o[i] = rs0 * (coef[i] * v[i] - tq.x - xh[i] * tq.y);
          }
          if (EXTRA) {
            float e[VEC], c[VEC];
            unpack<F16>(re[r], e);
            unpack<F16>(rc[r], c);
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] += e[i] + c[i];
          }
        }
        if (round_out && !F16) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) o[i] = round_tf32(o[i]);
        }
        stvec<F16>(out.ptr, (long long)n * out.sN + ooff, o);
      }
    }
  }
}

// ---- fp16 forward-only apply (the Jacobian-free programs): y = act(gn(x)), 8 channels = 16 bytes
// per thread and four pixels in flight per thread.  blockIdx.y is the batch row, so the per-channel
// affine u = a * x + b (a = gamma * rstd, b = beta - mean * a) is loop-invariant, and the pixel
// offset is p * sW when the views are pixel-contiguous (`lin`): the loop body has no division (the
// first version of this kernel spent ~120 of its ~230 instructions per vector on 64-bit divisions
// and was issue-bound at 3.5 TB/s on the 40 x 256^2 x 128 site).
constexpr int kGn16Unroll = 4;
__global__ void __launch_bounds__(256)
gn_apply_fwd16_kernel(View x, const double* __restrict__ stats, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, int silu, View y, int lin) {
  pdl_trigger();   // a dependent conv launch may start its prologue now (it waits for this grid's completion)
  const int C = x.C;
  const int c8n = C >> 3;
  const int cg = C / kGroups;
  const int HW = x.H * x.W;
  const double cnt = (double)HW * cg;
  const int n = blockIdx.y;
  const int cv = threadIdx.x % c8n;
  const int prow = threadIdx.x / c8n;
  const int pstep = blockDim.x / c8n;
  float a[8], b[8];
  {
    const float2 m0 = mean_rstd(stats + ((long long)n * kGroups + (cv * 8) / cg) * 2, cnt, eps);
    const float2 m1 = mean_rstd(stats + ((long long)n * kGroups + (cv * 8 + 4) / cg) * 2, cnt, eps);
    const float4 g0 = ld4(gamma + cv * 8), g1 = ld4(gamma + cv * 8 + 4);
    const float4 b0 = ld4(beta + cv * 8), b1 = ld4(beta + cv * 8 + 4);
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 mr = i < 4 ? m0 : m1;
      a[i] = ga[i] * mr.y;
      b[i] = be[i] - mr.x * a[i];
    }
  }
  const int stride = gridDim.x * pstep;
  const __half* xp = reinterpret_cast<const __half*>(x.ptr) + (long long)n * x.sN + cv * 8;
  __half* yp = reinterpret_cast<__half*>(y.ptr) + (long long)n * y.sN + cv * 8;
  for (int p0 = blockIdx.x * pstep + prow; p0 < HW; p0 += stride * kGn16Unroll) {
    uint4 v[kGn16Unroll];
    long long yo[kGn16Unroll];
#pragma unroll
    for (int u = 0; u < kGn16Unroll; ++u) {
      const int p = p0 + u * stride;
      yo[u] = -1;
      if (p < HW) {
        int yy = 0, xx = p;
        if (!lin) { yy = p / x.W; xx = p - yy * x.W; }
        v[u] = *reinterpret_cast<const uint4*>(xp + (long long)yy * x.sH + (long long)xx * x.sW);
        yo[u] = (long long)yy * y.sH + (long long)xx * y.sW;
      }
    }
#pragma unroll
    for (int u = 0; u < kGn16Unroll; ++u) {
      if (yo[u] < 0) continue;
      const uint32_t* w = &v[u].x;
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        float u0 = fmaf(a[2 * k], f.x, b[2 * k]);
        float u1 = fmaf(a[2 * k + 1], f.y, b[2 * k + 1]);
        if (silu) { u0 = __fdividef(u0, 1.0f + __expf(-u0)); u1 = __fdividef(u1, 1.0f + __expf(-u1)); }
        const __half2 h = __floats2half2_rn(u0, u1);
        o[k] = *reinterpret_cast<const uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(yp + yo[u]) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---- small sites: statistics + apply in ONE launch ------------------------------------------------
// The <= 64^2 GroupNorm sites of the Jacobian programs are launch-latency bound (two ~6 us kernels
// for a tensor that sits in L2).  Here a block owns one (batch row, group): it reads the group's
// slice of the row (and of the primal row) twice from L2 -- once for the sums, once to apply -- so a
// site is one launch, needs no cleared statistics buffer and no atomics (bit-reproducible).
//   MODE 0: primal rows y = act(gn(x)), tangent rows the JVP rule; the sums are also stored in the
//           statistics buffer (the VJP pass reads the primal row's).
//   MODE 1: VJP rule for cotangent row blockIdx.y at the primal point (xp, pstats).
// A thread moves VW = 8 (or 4 when a group has 4 channels) consecutive channels of a pixel.
constexpr int kGnSmallMax = 8192;             // elements of one (row, group) slice (measured: larger slices lose to the two-launch path)

__device__ __forceinline__ double block_sum2(double a, double b, double* sh, double& out_b) {
  a = warp_sum(a); b = warp_sum(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { sh[2 * w] = a; sh[2 * w + 1] = b; }
  __syncthreads();
  double ra = 0.0, rb = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { ra += sh[2 * i]; rb += sh[2 * i + 1]; }   // fixed order
  out_b = rb;
  return ra;
}

template <int MODE, bool F16, int VW>
__global__ void __launch_bounds__(512)
gn_small_kernel(View x, int n_primal, View gy, const double* __restrict__ pstats, double* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu, int round_out,
                const float* __restrict__ addend, long long add_sN, long long add_sW, int accumulate, View out) {
  pdl_trigger();   // a dependent conv launch may start its prologue now (it waits for this grid's completion)
  __shared__ double sh[32];
  using Elem = typename std::conditional<F16, __half, float>::type;
  const int g = blockIdx.x, n = blockIdx.y;
  const int C = x.C, cg = C / kGroups, vpp = cg / VW;
  const int HW = x.H * x.W, nvec = HW * vpp;
  const double cnt = (double)HW * cg;
  const View& rows = (MODE == 0) ? x : gy;
  const bool tangent = (MODE == 0) && n >= n_primal;
  const Elem* __restrict__ xrow = reinterpret_cast<const Elem*>(x.ptr) + ((MODE == 0 && !tangent) ? (long long)n * x.sN : 0) + g * cg;
  const Elem* __restrict__ rrow = reinterpret_cast<const Elem*>(rows.ptr) + (long long)n * rows.sN + g * cg;
  auto ldv = [&](const Elem* base, long long off, float (&v)[VW]) {
    if (F16) {
      if (VW == 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(base + off);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k])); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(base + off);
        const uint32_t w[2] = {u.x, u.y};
#pragma unroll
        for (int k = 0; k < 2; ++k) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k])); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
      }
    } else {
#pragma unroll
      for (int q = 0; q < VW / 4; ++q) {
        const float4 f = *reinterpret_cast<const float4*>(base + off + 4 * q);
        v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
      }
    }
  };
  auto stv = [&](Elem* base, long long off, const float (&v)[VW]) {
    if (F16) {
      uint32_t w[VW / 2];
#pragma unroll
      for (int k = 0; k < VW / 2; ++k) { const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]); w[k] = *reinterpret_cast<const uint32_t*>(&h); }
      if (VW == 8) *reinterpret_cast<uint4*>(base + off) = make_uint4(w[0], w[1], w[VW / 2 - 2], w[VW / 2 - 1]);
      else *reinterpret_cast<uint2*>(base + off) = make_uint2(w[0], w[1]);
    } else {
#pragma unroll
      for (int q = 0; q < VW / 4; ++q)
        *reinterpret_cast<float4*>(base + off + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  };
  // ---- statistics of the primal slice (mean, rstd) ----
  float mu, rstd;
  if (MODE == 1) {
    const float2 mr = mean_rstd(pstats + g * 2, cnt, eps);
    mu = mr.x; rstd = mr.y;
  } else {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll 4
    for (int e = threadIdx.x; e < nvec; e += blockDim.x) {
      float v[VW];
      ldv(xrow, (long long)(e / vpp) * x.sW + (e % vpp) * VW, v);
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int i = 0; i < VW; ++i) { a += v[i]; b += v[i] * v[i]; }
      s1 += a; s2 += b;
    }
    double t2;
    const double t1 = block_sum2(s1, s2, sh, t2);
    if (!tangent && threadIdx.x == 0) { stats[((long long)n * kGroups + g) * 2] = t1; stats[((long long)n * kGroups + g) * 2 + 1] = t2; }
    const double m = t1 / cnt, var = t2 / cnt - m * m;
    mu = (float)m; rstd = (float)(1.0 / sqrt((var > 0 ? var : 0) + (double)eps));
  }
  float gs[VW], bs[VW];
  // channel-dependent affine: a thread may see any of the vpp vectors of a pixel, so it is loaded per vector below
  // ---- primal rows of the forward program: apply and leave ----
  Elem* orow = reinterpret_cast<Elem*>(out.ptr) + (long long)n * out.sN + g * cg;
  if (MODE == 0 && !tangent) {
#pragma unroll 4
    for (int e = threadIdx.x; e < nvec; e += blockDim.x) {
      const int p = e / vpp, c0 = (e % vpp) * VW;
      float v[VW], o[VW];
      ldv(xrow, (long long)p * x.sW + c0, v);
#pragma unroll
      for (int i = 0; i < VW; ++i) { gs[i] = gamma[g * cg + c0 + i]; bs[i] = beta[g * cg + c0 + i]; }
#pragma unroll
      for (int i = 0; i < VW; ++i) {
        const float u = gs[i] * ((v[i] - mu) * rstd) + bs[i];
        o[i] = silu ? u * fast_sigmoid(u) : u;
        if (round_out && !F16) o[i] = round_tf32(o[i]);
      }
      stv(orow, (long long)p * out.sW + c0, o);
    }
    return;
  }
  // ---- tangent / cotangent rows: sums against the primal slice ----
  double s1 = 0.0, s2 = 0.0;
#pragma unroll 4
  for (int e = threadIdx.x; e < nvec; e += blockDim.x) {
    const int p = e / vpp, c0 = (e % vpp) * VW;
    float xv[VW], rv[VW];
    ldv(xrow, (long long)p * x.sW + c0, xv);
    ldv(rrow, (long long)p * rows.sW + c0, rv);
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < VW; ++i) {
      float w = rv[i];
      if (MODE == 1) {
        const float gm = gamma[g * cg + c0 + i];
        const float u = gm * ((xv[i] - mu) * rstd) + beta[g * cg + c0 + i];
        float d = 1.0f;
        if (silu) { const float sg = fast_sigmoid(u); d = sg * (1.0f + u * (1.0f - sg)); }
        w = gm * d * rv[i];
      }
      a += w; b += w * xv[i];
    }
    s1 += a; s2 += b;
  }
  double t2;
  const double t1 = block_sum2(s1, s2, sh, t2);
  if (MODE == 0 && threadIdx.x == 0) { stats[((long long)n * kGroups + g) * 2] = t1; stats[((long long)n * kGroups + g) * 2 + 1] = t2; }
  const float m1 = (float)(t1 / cnt);
  const float m2 = (float)((t2 - (double)mu * t1) * (double)rstd / cnt);
  const Elem* arow = addend ? reinterpret_cast<const Elem*>(addend) + (long long)n * add_sN + g * cg : nullptr;
#pragma unroll 2
  for (int e = threadIdx.x; e < nvec; e += blockDim.x) {
    const int p = e / vpp, c0 = (e % vpp) * VW;
    float xv[VW], rv[VW], o[VW];
    ldv(xrow, (long long)p * x.sW + c0, xv);
    ldv(rrow, (long long)p * rows.sW + c0, rv);
#pragma unroll
    for (int i = 0; i < VW; ++i) {
      const float gm = gamma[g * cg + c0 + i];
      const float xh = (xv[i] - mu) * rstd;
      const float u = gm * xh + beta[g * cg + c0 + i];
      float d = 1.0f;
      if (silu) { const float sg = fast_sigmoid(u); d = sg * (1.0f + u * (1.0f - sg)); }
      const float coef = d * gm;
      if (MODE == 0) o[i] = coef * rstd * (rv[i] - m1 - xh * m2);
      else o[i] = rstd * (coef * rv[i] - m1 - xh * m2);
    }
    if (MODE == 1 && arow) {
      float av[VW];
      ldv(arow, (long long)p * add_sW + c0, av);
#pragma unroll
      for (int i = 0; i < VW; ++i) o[i] += av[i];
    }
    if (MODE == 1 && accumulate) {
      float cv[VW];
      ldv(orow, (long long)p * out.sW + c0, cv);
#pragma unroll
      for (int i = 0; i < VW; ++i) o[i] += cv[i];
    }
    if (round_out && !F16) {
#pragma unroll
      for (int i = 0; i < VW; ++i) o[i] = round_tf32(o[i]);
    }
    stv(orow, (long long)p * out.sW + c0, o);
  }
}

// ------------------------------------------------------------------------------------------------
// Thin edges: <= 4-channel NCHW fp32 tensors <-> zero-padded 64-channel NHWC activations
// ------------------------------------------------------------------------------------------------
// The 4-channel ends of the latent-space networks (Stable-Diffusion-shaped U-Net: 4 -> C and C -> 4;
// VAE decoder: 4 -> C) run on the ordinary tcgen05 conv kernels over a thin side padded with zero
// channels to one K block (64): the padding costs < 3 % of the network's FLOPs and needs no new GEMM
// kernel.  These two kernels are the NCHW <-> padded-NHWC boundary, with the power-of-two range scale
// of the tangent / cotangent rows and the optional c x c channel mix of `post_quant_conv`
// (diffusers AutoencoderKL.decode = decoder(post_quant_conv(z)), src/modules/edit.py:770).
// mix: [c*c + c] = 1x1 weight [o][i] then bias; transposed = 1 applies the transpose (VJP), no bias.
template <bool F16>
__global__ void __launch_bounds__(256)
thin_pad_kernel(const float* __restrict__ in, int c, View out, const float* __restrict__ mix, int bias_rows,
                const float* __restrict__ scale_dev, int scale_from, int round_out) {
  pdl_trigger();   // a dependent conv launch may start its prologue now (it waits for this grid's completion)
  constexpr int VPP = F16 ? kThinPad / 8 : kThinPad / 4;      // 16-byte vectors per pixel
  const long long HW = (long long)out.H * out.W;
  const long long total = (long long)out.N * HW * VPP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % VPP);
    const long long p = i / VPP;
    const long long pix = p % HW;
    const int n = (int)(p / HW);
    float val[4] = {0.f, 0.f, 0.f, 0.f};
    if (v == 0) {
      float raw[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < c; ++j) raw[j] = in[((long long)n * c + j) * HW + pix];
      if (mix != nullptr) {
        for (int o = 0; o < c; ++o) {
          float a = n < bias_rows ? mix[c * c + o] : 0.f;
          for (int j = 0; j < c; ++j) a += mix[o * c + j] * raw[j];
          val[o] = a;
        }
      } else {
        for (int j = 0; j < c; ++j) val[j] = raw[j];
      }
      const float sc = (scale_dev != nullptr && n >= scale_from) ? scale_dev[0] : 1.f;
      for (int j = 0; j < 4; ++j) { val[j] *= sc; if (round_out && !F16) val[j] = round_tf32(val[j]); }
    }
    const long long y = pix / out.W, x = pix % out.W;
    const long long o = n * out.sN + y * out.sH + x * out.sW;
    if (F16) {
      const __half2 a = __floats2half2_rn(val[0], val[1]), b = __floats2half2_rn(val[2], val[3]);
      uint4 u = make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b), 0u, 0u);
      *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out.ptr) + o + v * 8) = u;
    } else {
      *reinterpret_cast<float4*>(out.ptr + o + v * 4) = make_float4(val[0], val[1], val[2], val[3]);
    }
  }
}
template <bool F16>
__global__ void __launch_bounds__(256)
thin_extract_kernel(View in, int c, float* __restrict__ out, const float* __restrict__ mix,
                    const float* __restrict__ scale_dev, int scale_from) {
  const long long HW = (long long)in.H * in.W;
  const long long total = (long long)in.N * HW;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const long long pix = p % HW;
    const int n = (int)(p / HW);
    const long long y = pix / in.W, x = pix % in.W;
    const long long o = n * in.sN + y * in.sH + x * in.sW;
    float val[4];
    if (F16) {
      const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(in.ptr) + o);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      val[0] = a.x; val[1] = a.y; val[2] = b.x; val[3] = b.y;
    } else {
      const float4 f = *reinterpret_cast<const float4*>(in.ptr + o);
      val[0] = f.x; val[1] = f.y; val[2] = f.z; val[3] = f.w;
    }
    const float sc = (scale_dev != nullptr && n >= scale_from) ? scale_dev[1] : 1.f;
    for (int j = 0; j < c; ++j) {
      float a = val[j];
      if (mix != nullptr) {                       // transpose of the channel mix: g_in[j] = sum_o W[o][j] g[o]
        a = 0.f;
        for (int o2 = 0; o2 < c; ++o2) a += mix[o2 * c + j] * val[o2];
      }
      out[((long long)n * c + j) * HW + pix] = a * sc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Resampling / add
// ------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void upsample2x_kernel(View in, View out, float scale, int accumulate, int round_out) {
  pdl_trigger();   // a dependent conv launch may start its prologue now (it waits for this grid's completion)
  const int cvn = out.C >> 2;
  const long long total = (long long)out.N * out.H * out.W * cvn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvn);
    long long r = i / cvn;
    const int x = (int)(r % out.W); r /= out.W;
    const int y = (int)(r % out.H);
    const int n = (int)(r / out.H);
    float4 v = ld4t<F16>(in.ptr, n * in.sN + (y >> 1) * in.sH + (x >> 1) * in.sW + cv * 4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    const long long o = n * out.sN + y * out.sH + x * out.sW + cv * 4;
    if (accumulate) {
      const float4 p = ld4t<F16>(out.ptr, o);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (round_out && !F16) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    st4t<F16>(out.ptr, o, v);
  }
}
// Non-accumulating fast path: one thread per INPUT vector of 16 bytes (4 fp32 / 8 fp16 channels),
// one load and four 16-byte stores (the output-indexed kernel above issues one 8/16-byte load per
// output vector and was latency-bound at 1.4 TB/s on the 128^2 -> 256^2 site).
template <bool F16>
__global__ void __launch_bounds__(256)
upsample2x_fast_kernel(View in, View out, float scale, int round_out) {
  pdl_trigger();   // a dependent conv launch may start its prologue now (it waits for this grid's completion)
  constexpr int VEC = F16 ? 8 : 4;
  const int cvn = in.C / VEC;
  const long long total = (long long)in.N * in.H * in.W * cvn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvn);
    long long r = i / cvn;
    const int x = (int)(r % in.W); r /= in.W;
    const int y = (int)(r % in.H);
    const int n = (int)(r / in.H);
    const long long ioff = n * in.sN + y * in.sH + x * in.sW + cv * VEC;
    const long long ooff = n * out.sN + (2 * y) * out.sH + (2 * x) * out.sW + cv * VEC;
    uint4 o;
    if (F16) {
      uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(in.ptr) + ioff);
      if (scale != 1.f) {
        uint32_t* w = &u.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
          const __half2 h = __floats2half2_rn(f.x * scale, f.y * scale);
          w[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      o = u;
      __half* op = reinterpret_cast<__half*>(out.ptr) + ooff;
      *reinterpret_cast<uint4*>(op) = o;
      *reinterpret_cast<uint4*>(op + out.sW) = o;
      *reinterpret_cast<uint4*>(op + out.sH) = o;
      *reinterpret_cast<uint4*>(op + out.sH + out.sW) = o;
    } else {
      float4 v = *reinterpret_cast<const float4*>(in.ptr + ioff);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
      float* op = out.ptr + ooff;
      *reinterpret_cast<float4*>(op) = v;
      *reinterpret_cast<float4*>(op + out.sW) = v;
      *reinterpret_cast<float4*>(op + out.sH) = v;
      *reinterpret_cast<float4*>(op + out.sH + out.sW) = v;
    }
  }
}
template <bool F16>
__global__ void sumpool2x_kernel(View in, View out, float scale, int accumulate, int round_out) {
  pdl_trigger();   // a dependent conv launch may start its prologue now (it waits for this grid's completion)
  const int cvn = out.C >> 2;
  const long long total = (long long)out.N * out.H * out.W * cvn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvn);
    long long r = i / cvn;
    const int x = (int)(r % out.W); r /= out.W;
    const int y = (int)(r % out.H);
    const int n = (int)(r / out.H);
    const long long b = n * in.sN + (2 * y) * in.sH + (2 * x) * in.sW + cv * 4;
    const float4 a = ld4t<F16>(in.ptr, b);
    const float4 c = ld4t<F16>(in.ptr, b + in.sW);
    const float4 d = ld4t<F16>(in.ptr, b + in.sH);
    const float4 e = ld4t<F16>(in.ptr, b + in.sH + in.sW);
    float4 v = make_float4(scale * ((a.x + c.x) + (d.x + e.x)), scale * ((a.y + c.y) + (d.y + e.y)),
                           scale * ((a.z + c.z) + (d.z + e.z)), scale * ((a.w + c.w) + (d.w + e.w)));
    const long long o = n * out.sN + y * out.sH + x * out.sW + cv * 4;
    if (accumulate) {
      const float4 p = ld4t<F16>(out.ptr, o);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (round_out && !F16) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    st4t<F16>(out.ptr, o, v);
  }
}
template <bool F16>
__global__ void add_views_kernel(View in, View out, int accumulate) {
  pdl_trigger();   // a dependent conv launch may start its prologue now (it waits for this grid's completion)
  const int cvn = out.C >> 2;
  const long long total = (long long)out.N * out.H * out.W * cvn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvn);
    long long r = i / cvn;
    const int x = (int)(r % out.W); r /= out.W;
    const int y = (int)(r % out.H);
    const int n = (int)(r / out.H);
    float4 v = ld4t<F16>(in.ptr, n * in.sN + y * in.sH + x * in.sW + cv * 4);
    const long long o = n * out.sN + y * out.sH + x * out.sW + cv * 4;
    if (accumulate) {
      const float4 p = ld4t<F16>(out.ptr, o);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    st4t<F16>(out.ptr, o, v);
  }
}

// ------------------------------------------------------------------------------------------------
// Timestep embedding
// ------------------------------------------------------------------------------------------------
__global__ void set_scalar_kernel(float* dst, float v) { *dst = v; }
// One dense layer of the timestep MLP: one warp per output, 8 outputs per block, the input vector in shared
// memory.  first = 1: the input is the sinusoidal embedding of *t_dev (computed per block, `ch` values);
// else `in` [n_in].  out[o] = silu(b[o] + W[o,:] . in (+ cond[o])).  (The two layers used to run in ONE block,
// 32 outputs per warp one after the other: 93 us of dependent L2 latency at the head of every forward program.)
__global__ void __launch_bounds__(256)
temb_dense_kernel(const float* __restrict__ t_dev, int style, int first, const float* __restrict__ in, int n_in,
                  const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ cond,
                  int n_out, float* __restrict__ out) {
  extern __shared__ float vec[];   // n_in
  if (first) {
    const float t = *t_dev;
    const int half = n_in / 2;
    // style 0: DDPM [sin, cos], w_i = exp(-ln(1e4) i / (half-1))      (ddpm/diffusion.py:783-804)
    // style 1: guided-diffusion [cos, sin], w_i = exp(-ln(1e4) i / half)  (guided_diffusion/nn.py:103-121)
    const float coef = -(float)(log(10000.0) / (double)(style == 0 ? half - 1 : half));
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const float wi = expf((float)i * coef);
      const float a = t * wi;
      vec[style == 0 ? i : half + i] = sinf(a);
      vec[style == 0 ? half + i : i] = cosf(a);
    }
  } else {
    for (int i = threadIdx.x; i < n_in; i += blockDim.x) vec[i] = in[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + warp;
  if (o >= n_out) return;
  float acc = 0.f;
#pragma unroll 8
  for (int i = lane; i < n_in; i += 32) acc += w[(long long)o * n_in + i] * vec[i];
  acc = warp_sum(acc);
  // cond (optional): conditioning embedding added to the timestep embedding before the blocks'
  // SiLU + projection (emb = time_embed(t) + cond, the class / pooled-text conditioning of
  // guided-diffusion style U-Nets)
  if (lane == 0) out[o] = silu_f(acc + b[o] + (cond ? cond[o] : 0.f));
}
// P2 scale-shift normalisation (guided_diffusion/unet.py:247-252): GN(x) * (1 + scale) + shift with
// (scale | shift) = emb_layers(emb) folds into an effective affine of the GroupNorm:
//   gamma' = gamma (1 + scale),  beta' = beta (1 + scale) + shift.   One block per site.
__global__ void scale_shift_affine_kernel(const AffineSite* __restrict__ sites,
                                          const float* __restrict__ weights,
                                          const float* __restrict__ tproj, float* __restrict__ out) {
  const AffineSite st = sites[blockIdx.x];
  for (int c = threadIdx.x; c < st.C; c += blockDim.x) {
    const float sc = 1.0f + tproj[st.tproj_off + c];
    const float sh = tproj[st.tproj_off + st.C + c];
    out[st.out_off + c] = weights[st.gamma_off + c] * sc;
    out[st.out_off + st.C + c] = weights[st.beta_off + c] * sc + sh;
  }
}
// out[c] = b[c] + W[c,:] . temb_act   (one warp per output)
__global__ void temb_project_kernel(const float* __restrict__ temb_act, int temb_ch,
                                    const float* __restrict__ w, const float* __restrict__ b,
                                    int cout, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= cout) return;
  float acc = 0.f;
  for (int i = lane; i < temb_ch; i += 32) acc += w[(long long)o * temb_ch + i] * temb_act[i];
  acc = warp_sum(acc);
  if (lane == 0) out[o] = acc + b[o];
}

// Key | value projection of a prompt embedding for the cross-attention layers:
// out[t][j] = round_tf32(b[j] + W[j,:] . ctx[t,:]) for the n_tok tokens that exist, 0 for the padding
// rows (one warp per output; the result is an operand of the tcgen05 cross-attention kernels).
__global__ void context_kv_kernel(const float* __restrict__ ctx, int n_tok, int dim, const float* __restrict__ w,
                                  const float* __restrict__ b, int cout, int rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= (long long)rows * cout) return;
  const int t = (int)(o / cout), j = (int)(o % cout);
  float acc = 0.f;
  if (t < n_tok)
    for (int i = lane; i < dim; i += 32) acc += w[(long long)j * dim + i] * ctx[(long long)t * dim + i];
  acc = warp_sum(acc);
  if (lane == 0) out[o] = t < n_tok ? round_tf32(acc + b[j]) : 0.f;
}

// Power-of-two scale that brings max |v| to about `target`: the fp16 JVP / VJP programs carry their
// tangent / cotangent rows scaled by it (both passes are linear in those rows, a power of two is exact)
// so that the rows sit in the middle of fp16's exponent range whatever their natural magnitude
// (tangents of an orthonormal basis of R^196608 are ~2e-3, with activations down to 1e-6).
__global__ void amax_kernel(const float* __restrict__ v, long long n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(v[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));     // non-negative floats order like uints
}
__global__ void pow2_scale_kernel(const unsigned* __restrict__ amax_bits, float target, float* __restrict__ scale2) {
  const float a = __uint_as_float(*amax_bits);
  float sc = 1.f;
  if (a > 0.f && isfinite(a)) sc = exp2f(floorf(log2f(target / a)));
  sc = fminf(fmaxf(sc, 1.0f / 16777216.f), 16777216.f);
  scale2[0] = sc;
  scale2[1] = 1.0f / sc;
}

inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// pixel-contiguous view: the offset of pixel p = y * W + x is p * sW
inline int gn_lin(const View& v) { return v.sH == (long long)v.W * v.sW; }
int check_gn_view(const View& v, const char* what) {
  LOCO_REQUIRE(v.C % 128 == 0, "%s: channels %d must be a multiple of 128", what, v.C);
  LOCO_REQUIRE(v.C <= 1024, "%s: channels %d > 1024", what, v.C);
  LOCO_REQUIRE((v.sW % 4) == 0 && (v.sH % 4) == 0 && (v.sN % 4) == 0 &&
                   (((uintptr_t)v.ptr) & 15) == 0,
               "%s: view not float4-aligned", what);
  LOCO_REQUIRE(!v.half || ((v.sW % 8) == 0 && (v.sH % 8) == 0 && (v.sN % 8) == 0),
               "%s: fp16 view not aligned to 16-byte vectors", what);
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host launchers
// ------------------------------------------------------------------------------------------------
int pack_conv_fprop(const float* w, float* dst, int Cout, int Cin, int kh, int kw, cudaStream_t s) {
  const long long total = (long long)Cout * Cin * kh * kw;
  pack_fprop_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(w, dst, Cout, Cin, kh * kw);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int pack_conv_dgrad(const float* w, float* dst, int Cout, int Cin, int kh, int kw, int cout_total,
                    int co_off, cudaStream_t s) {
  const long long total = (long long)Cout * Cin * kh * kw;
  pack_dgrad_kernel<float><<<grid_for(total, 256), 256, 0, s>>>(w, dst, Cout, Cin, kh * kw, cout_total,
                                                               co_off);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int pack_conv_fprop16(const float* w, void* dst, int Cout, int Cin, int kh, int kw, cudaStream_t s) {
  const long long total = (long long)Cout * Cin * kh * kw;
  pack_fprop_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(w, reinterpret_cast<__half*>(dst), Cout, Cin,
                                                                kh * kw);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int pack_conv_dgrad16(const float* w, void* dst, int Cout, int Cin, int kh, int kw, int cout_total,
                      int co_off, cudaStream_t s) {
  const long long total = (long long)Cout * Cin * kh * kw;
  pack_dgrad_kernel<__half><<<grid_for(total, 256), 256, 0, s>>>(w, reinterpret_cast<__half*>(dst), Cout, Cin,
                                                                kh * kw, cout_total, co_off);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int pack_conv_edge(const float* w, float* dst, int C, int in_is_3, cudaStream_t s) {
  pack_edge_kernel<<<grid_for(27 * C, 256), 256, 0, s>>>(w, dst, C, in_is_3);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int pow2_scale(const float* v, long long n, float target, float* scale2, unsigned* tmp, cudaStream_t s) {
  LOCO_CHECK_CUDA(cudaMemsetAsync(tmp, 0, sizeof(unsigned), s));
  amax_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, s>>>(v, n, tmp);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  pow2_scale_kernel<<<1, 1, 0, s>>>(tmp, target, scale2);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int edge_conv_expand(const float* in3, const float* We, const float* bias, int bias_rows, View out,
                     int flip, int round_out, cudaStream_t s, const float* scale_dev, int scale_from) {
  if (edge_mma_eligible(out)) return edge_conv_expand_mma(in3, We, bias, bias_rows, out, flip, s, scale_dev, scale_from);
  if (out.C == 128 && out.W % kStrip == 0) {
    const long long nstrips = (long long)out.N * out.H * (out.W / kStrip);
    ProfScope prof(2, 0, s);
    const int grid = grid_for((nstrips + 7) / 8, 1, num_sms() * 4);
    if (out.half) edge_expand128_kernel<true><<<grid, 256, 0, s>>>(in3, We, bias, bias_rows, out, flip, round_out, scale_dev, scale_from);
    else edge_expand128_kernel<false><<<grid, 256, 0, s>>>(in3, We, bias, bias_rows, out, flip, round_out, scale_dev, scale_from);
    count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  LOCO_REQUIRE(out.C % 4 == 0 && 256 % (out.C / 4) == 0, "edge_conv_expand: C=%d unsupported", out.C);
  const int ppb = 256 / (out.C / 4);
  const long long HW = (long long)out.H * out.W;
  const long long ppb_all = (long long)ppb * kEdgeIters;
  dim3 grid((unsigned)((HW + ppb_all - 1) / ppb_all), out.N);
  const size_t smem = 27 * out.C * sizeof(float);
  LOCO_REQUIRE(smem <= 48 * 1024, "edge_conv_expand: C=%d too large", out.C);
  if (out.half) edge_expand_kernel<true><<<grid, 256, smem, s>>>(in3, We, bias, bias_rows, out, flip, round_out, scale_dev, scale_from);
  else edge_expand_kernel<false><<<grid, 256, smem, s>>>(in3, We, bias, bias_rows, out, flip, round_out, scale_dev, scale_from);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int edge_conv_reduce(View in, const float* Wr, const float* bias, int bias_rows, float* out3,
                     int flip, cudaStream_t s, const float* scale_dev, int scale_from) {
  if (edge_mma_eligible(in)) return edge_conv_reduce_mma(in, Wr, bias, bias_rows, out3, flip, s, scale_dev, scale_from);
  if (in.C == 128 && in.W % kStrip == 0) {
    const long long nstrips = (long long)in.N * in.H * (in.W / kStrip);
    ProfScope prof(2, 0, s);
    const int grid = grid_for((nstrips + 7) / 8, 1, num_sms() * 4);
    if (in.half) edge_reduce128_kernel<true><<<grid, 256, 0, s>>>(in, Wr, bias, bias_rows, out3, flip, scale_dev, scale_from);
    else edge_reduce128_kernel<false><<<grid, 256, 0, s>>>(in, Wr, bias, bias_rows, out3, flip, scale_dev, scale_from);
    count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  LOCO_REQUIRE(in.C % 4 == 0, "edge_conv_reduce: C=%d unsupported", in.C);
  const long long HW = (long long)in.H * in.W;
  dim3 grid((unsigned)((HW + 8 * kEdgeIters - 1) / (8 * kEdgeIters)), in.N);
  const size_t smem = 27 * in.C * sizeof(float);
  LOCO_REQUIRE(smem <= 48 * 1024, "edge_conv_reduce: C=%d too large", in.C);
  if (in.half) edge_reduce_kernel<true><<<grid, 256, smem, s>>>(in, Wr, bias, bias_rows, out3, flip, scale_dev, scale_from);
  else edge_reduce_kernel<false><<<grid, 256, smem, s>>>(in, Wr, bias, bias_rows, out3, flip, scale_dev, scale_from);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// resident blocks of each GroupNorm kernel at its block size (queried once)
template <typename K>
static int gn_resident(K kernel, int block, int* cache) {
  if (*cache == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    *cache = per_sm * num_sms();
  }
  return *cache;
}
// per device: [storage type][kernel: stats fwd, stats vjp, apply fwd, apply vjp][block == 192]
static int g_res[kMaxDevices][2][5][2];   // kernel 4: the VJP apply with addend / accumulate
static int g_res16[kMaxDevices][2];       // gn_apply_fwd16_kernel, [block == 192]
static int* res_slot(int h, int kernel, int bd192) {
  const int d = current_device();
  return &g_res[d < kMaxDevices ? d : 0][h][kernel][bd192];
}

// All kernels of the U-Net programs ask for the same (maximum shared memory) L1/smem split as the
// tcgen05 conv kernel, so the SMs never have to re-partition between consecutive launches.
int layers_init() {
  static bool done[kMaxDevices] = {false};
  if (!first_time_on_device(done)) return 0;
  const int co = cudaSharedmemCarveoutMaxShared;
#define LOCO_CARVE(k) LOCO_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, co))
#define LOCO_CARVE2(k) LOCO_CARVE(k<false>); LOCO_CARVE(k<true>)
  LOCO_CARVE((gn_stats_kernel<0, false>)); LOCO_CARVE((gn_stats_kernel<1, false>));
  LOCO_CARVE((gn_apply_kernel<0, false, kRC>)); LOCO_CARVE((gn_apply_kernel<1, false, 8>)); LOCO_CARVE((gn_apply_kernel<1, false, 4>));
  LOCO_CARVE((gn_stats_kernel<0, true>)); LOCO_CARVE((gn_stats_kernel<1, true>));
  LOCO_CARVE((gn_apply_kernel<0, true, kRC>)); LOCO_CARVE((gn_apply_kernel<1, true, 8>)); LOCO_CARVE((gn_apply_kernel<1, true, 4>));
  LOCO_CARVE2(edge_expand_kernel); LOCO_CARVE2(edge_reduce_kernel);
  LOCO_CARVE2(edge_expand128_kernel); LOCO_CARVE2(edge_reduce128_kernel);
  LOCO_CARVE2(upsample2x_kernel); LOCO_CARVE2(upsample2x_fast_kernel); LOCO_CARVE2(sumpool2x_kernel); LOCO_CARVE2(add_views_kernel);
  LOCO_CARVE(gn_apply_fwd16_kernel);
  LOCO_CARVE(temb_dense_kernel); LOCO_CARVE(temb_project_kernel); LOCO_CARVE(set_scalar_kernel);
  LOCO_CARVE(scale_shift_affine_kernel);
#undef LOCO_CARVE2
#undef LOCO_CARVE
  LOCO_TRY(edge_mma_init());
  // occupancy queries up front (never inside a stream capture)
  for (int b = 0; b < 2; ++b) {
    const int bd = b ? 192 : 256;
    gn_resident(gn_stats_kernel<0, false>, bd, res_slot(0, 0, b));
    gn_resident(gn_stats_kernel<1, false>, bd, res_slot(0, 1, b));
    gn_resident(gn_apply_kernel<0, false, kRC>, bd, res_slot(0, 2, b));
    gn_resident(gn_apply_kernel<1, false, 8>, bd, res_slot(0, 3, b));
    gn_resident(gn_apply_kernel<1, false, 4>, bd, res_slot(0, 4, b));
    gn_resident(gn_stats_kernel<0, true>, bd, res_slot(1, 0, b));
    gn_resident(gn_stats_kernel<1, true>, bd, res_slot(1, 1, b));
    gn_resident(gn_apply_kernel<0, true, kRC>, bd, res_slot(1, 2, b));
    gn_resident(gn_apply_kernel<1, true, 8>, bd, res_slot(1, 3, b));
    gn_resident(gn_apply_kernel<1, true, 4>, bd, res_slot(1, 4, b));
    gn_resident(gn_apply_fwd16_kernel, bd, &g_res16[current_device() < kMaxDevices ? current_device() : 0][b]);
  }
  return 0;
}

static int same_type(const View& a, const View& b, const char* what) {
  LOCO_REQUIRE(a.half == b.half, "%s: mixed fp16 / fp32 tensors", what);
  return 0;
}

__global__ void gn_stats_fold_pairs_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int group0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;      // (row, g < 16, which)
  if (e >= rows * (kGroups / 2) * 2) return;
  const int which = e & 1, g = (e >> 1) % (kGroups / 2), n = e / kGroups;
  const double* sp = src + ((long long)n * kGroups + 2 * g) * 2 + which;
  atomicAdd(&dst[((long long)n * kGroups + group0 + g) * 2 + which], sp[0] + sp[2]);
}
int gn_stats_fold_pairs(const double* src, double* dst, int rows, int group0, cudaStream_t s) {
  LOCO_REQUIRE(src && dst && rows > 0 && group0 >= 0 && group0 + kGroups / 2 <= kGroups, "gn_stats_fold_pairs: bad arguments");
  const int n = rows * kGroups;
  gn_stats_fold_pairs_kernel<<<(n + 127) / 128, 128, 0, s>>>(src, dst, rows, group0);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int gn_stats_fwd(View x, int n_primal, double* stats, cudaStream_t s) {
  LOCO_TRY(check_gn_view(x, "gn_stats_fwd"));
  const int h = x.half;
  const int vec = h ? 8 : 4;
  const int bd = gn_block_dim(x.C, vec);
  const int res = h ? gn_resident(gn_stats_kernel<0, true>, bd, res_slot(1, 0, bd == 192))
                    : gn_resident(gn_stats_kernel<0, false>, bd, res_slot(0, 0, bd == 192));
  const GnGeom g = gn_geom(x.C, vec, (long long)x.H * x.W, res);
  ProfScope prof(1, (h ? 2.0 : 4.0) * x.N * x.H * x.W * x.C, s);
  if (h) gn_stats_kernel<0, true><<<g.nblk, g.block, 0, s>>>(x, n_primal, x, nullptr, nullptr, nullptr, 0.f, 0, stats, gn_lin(x));
  else gn_stats_kernel<0, false><<<g.nblk, g.block, 0, s>>>(x, n_primal, x, nullptr, nullptr, nullptr, 0.f, 0, stats, gn_lin(x));
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int gn_apply_fwd(View x, int n_primal, const double* stats, const float* gamma, const float* beta,
                 float eps, int silu, int round_out, View y, cudaStream_t s) {
  LOCO_TRY(check_gn_view(x, "gn_apply_fwd"));
  LOCO_TRY(check_gn_view(y, "gn_apply_fwd(out)"));
  LOCO_TRY(same_type(x, y, "gn_apply_fwd"));
  LOCO_REQUIRE(x.N <= kGnMaxRows, "gn_apply_fwd: batch %d > %d rows", x.N, kGnMaxRows);
  const int h = x.half;
  const int vec = h ? 8 : 4;
  const int bd = gn_block_dim(x.C, vec);
  if (h && n_primal == x.N && x.C % 8 == 0 && (x.C / kGroups) % 4 == 0 && x.sW % 8 == 0 && x.sH % 8 == 0 &&
      x.sN % 8 == 0 && y.sW % 8 == 0 && y.sH % 8 == 0 && y.sN % 8 == 0) {
    // forward-only fp16: the 16-byte-vector kernel
    const int c8n = x.C / 8;
    const int block = (256 % c8n == 0) ? 256 : 192;
    LOCO_REQUIRE(block % c8n == 0, "gn_apply_fwd: C=%d unsupported by the fp16 kernel", x.C);
    int* rs = &g_res16[current_device() < kMaxDevices ? current_device() : 0][block == 192];
    gn_resident(gn_apply_fwd16_kernel, block, rs);
    // grid: (blocks per batch row, batch rows); about one resident wave in total
    const long long pix = (long long)x.H * x.W;
    const int pstep = block / c8n;
    long long nblk = (pix + (long long)pstep * kGn16Unroll - 1) / ((long long)pstep * kGn16Unroll);
    const long long per_row = std::max(1, *rs / x.N);
    if (nblk > per_row) nblk = per_row;
    ProfScope prof(1, 4.0 * x.N * x.H * x.W * x.C, s);
    gn_apply_fwd16_kernel<<<dim3((unsigned)nblk, (unsigned)x.N), block, 0, s>>>(x, stats, gamma, beta, eps, silu, y,
                                                                             gn_lin(x) && gn_lin(y));
    count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int res = h ? gn_resident(gn_apply_kernel<0, true, kRC>, bd, res_slot(1, 2, bd == 192))
                    : gn_resident(gn_apply_kernel<0, false, kRC>, bd, res_slot(0, 2, bd == 192));
  const GnGeom g = gn_geom(x.C, vec, (long long)x.H * x.W, res);
  ProfScope prof(1, (h ? 4.0 : 8.0) * x.N * x.H * x.W * x.C, s);
  if (h) gn_apply_kernel<0, true, kRC><<<g.nblk, g.block, 0, s>>>(x, n_primal, x, nullptr, stats, gamma, beta, eps,
                                                            silu, round_out, nullptr, 0, 0, 0, 0, y, gn_lin(x) && gn_lin(y));
  else gn_apply_kernel<0, false, kRC><<<g.nblk, g.block, 0, s>>>(x, n_primal, x, nullptr, stats, gamma, beta, eps,
                                                           silu, round_out, nullptr, 0, 0, 0, 0, y, gn_lin(x) && gn_lin(y));
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int gn_stats_vjp(View xp, const double* pstats, View gy, const float* gamma, const float* beta,
                 float eps, int silu, double* stats, cudaStream_t s) {
  LOCO_TRY(check_gn_view(xp, "gn_stats_vjp"));
  LOCO_TRY(check_gn_view(gy, "gn_stats_vjp(gy)"));
  LOCO_TRY(same_type(xp, gy, "gn_stats_vjp"));
  const int h = xp.half;
  const int vec = h ? 8 : 4;
  const int bd = gn_block_dim(xp.C, vec);
  const int res = h ? gn_resident(gn_stats_kernel<1, true>, bd, res_slot(1, 1, bd == 192))
                    : gn_resident(gn_stats_kernel<1, false>, bd, res_slot(0, 1, bd == 192));
  const GnGeom g = gn_geom(xp.C, vec, (long long)xp.H * xp.W, res);
  ProfScope prof(1, (h ? 2.0 : 4.0) * (gy.N + 1) * gy.H * gy.W * gy.C, s);
  if (h) gn_stats_kernel<1, true><<<g.nblk, g.block, 0, s>>>(xp, 0, gy, pstats, gamma, beta, eps, silu, stats, gn_lin(xp) && gn_lin(gy));
  else gn_stats_kernel<1, false><<<g.nblk, g.block, 0, s>>>(xp, 0, gy, pstats, gamma, beta, eps, silu, stats, gn_lin(xp) && gn_lin(gy));
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int gn_apply_vjp(View xp, const double* pstats, View gy, const double* stats, const float* gamma,
                 const float* beta, float eps, int silu, const View* addend, int accumulate,
                 int round_out, View gx, cudaStream_t s) {
  LOCO_TRY(check_gn_view(xp, "gn_apply_vjp"));
  LOCO_TRY(check_gn_view(gy, "gn_apply_vjp(gy)"));
  LOCO_TRY(check_gn_view(gx, "gn_apply_vjp(gx)"));
  LOCO_TRY(same_type(xp, gy, "gn_apply_vjp")); LOCO_TRY(same_type(gy, gx, "gn_apply_vjp(gx)"));
  if (addend) { LOCO_TRY(check_gn_view(*addend, "gn_apply_vjp(addend)")); LOCO_TRY(same_type(*addend, gx, "gn_apply_vjp(addend)")); }
  LOCO_REQUIRE(gy.N <= kGnMaxRows, "gn_apply_vjp: batch %d > %d rows", gy.N, kGnMaxRows);
  const int h = xp.half;
  const int vec = h ? 8 : 4;
  const int bd = gn_block_dim(xp.C, vec);
  const bool extra = addend != nullptr || accumulate != 0;
  int res;
  if (extra) res = h ? gn_resident(gn_apply_kernel<1, true, 4>, bd, res_slot(1, 4, bd == 192))
                     : gn_resident(gn_apply_kernel<1, false, 4>, bd, res_slot(0, 4, bd == 192));
  else res = h ? gn_resident(gn_apply_kernel<1, true, 8>, bd, res_slot(1, 3, bd == 192))
               : gn_resident(gn_apply_kernel<1, false, 8>, bd, res_slot(0, 3, bd == 192));
  const GnGeom g = gn_geom(xp.C, vec, (long long)xp.H * xp.W, res);
  const int lin = gn_lin(xp) && gn_lin(gy) && gn_lin(gx) && (!addend || gn_lin(*addend));
  ProfScope prof(1, (h ? 2.0 : 4.0) * gy.H * gy.W * gy.C * (1 + gy.N * (2 + (addend ? 1 : 0) + (accumulate ? 1 : 0))), s);
#define LOCO_GN_VJP(F16, RC)                                                                          \
  gn_apply_kernel<1, F16, RC><<<g.nblk, g.block, 0, s>>>(                                             \
      xp, 0, gy, pstats, stats, gamma, beta, eps, silu, round_out, addend ? addend->ptr : nullptr,    \
      addend ? addend->sN : 0, addend ? addend->sH : 0, addend ? addend->sW : 0, accumulate, gx, lin)
  if (h) { if (extra) LOCO_GN_VJP(true, 4); else LOCO_GN_VJP(true, 8); }
  else { if (extra) LOCO_GN_VJP(false, 4); else LOCO_GN_VJP(false, 8); }
#undef LOCO_GN_VJP
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- small sites: one launch per GroupNorm site (gn_small_kernel) ----
static long long gn_small_max() {
  static long long v = -1;
  if (v < 0) {
    const char* e = getenv("LOCO_GN_SMALL_MAX");
    v = e ? atoll(e) : kGnSmallMax;
  }
  return v;
}
// layout the kernel can address (any slice size)
static bool gn_small_supported(const View& x) {
  if (x.C % kGroups != 0) return false;
  const int cg = x.C / kGroups;
  if (cg % 4 != 0) return false;
  const int vw = (cg % 8 == 0) ? 8 : 4;
  return x.sW % vw == 0 && x.sH == (long long)x.W * x.sW && x.sN % vw == 0;
}
// ... and small enough for one block per (row, group) to beat the two-launch path (what the plans ask)
bool gn_small_eligible(const View& x) {
  return gn_small_supported(x) && (long long)x.H * x.W * (x.C / kGroups) <= gn_small_max();
}
template <int MODE>
static int gn_small_launch(View x, int n_primal, View gy, const double* pstats, double* stats, const float* gamma,
                           const float* beta, float eps, int silu, int round_out, const View* addend, int accumulate,
                           View out, cudaStream_t s) {
  const View& rows = MODE == 0 ? x : gy;
  const int cg = x.C / kGroups, vw = (cg % 8 == 0) ? 8 : 4;
  const long long nvec = (long long)x.H * x.W * (cg / vw);
  const int block = nvec >= 2048 ? 512 : (nvec >= 256 ? 256 : 128);
  const dim3 grid(kGroups, rows.N);
  const float* ap = addend ? addend->ptr : nullptr;
  const long long asN = addend ? addend->sN : 0, asW = addend ? addend->sW : 0;
  ProfScope prof(1, (x.half ? 2.0 : 4.0) * rows.N * x.H * x.W * x.C * (MODE == 0 ? 3 : 3 + (addend ? 1 : 0) + (accumulate ? 1 : 0)), s);
#define LOCO_GN_SMALL(F16, VW) \
  gn_small_kernel<MODE, F16, VW><<<grid, block, 0, s>>>(x, n_primal, gy, pstats, stats, gamma, beta, eps, silu, round_out, ap, asN, asW, accumulate, out)
  if (x.half) { if (vw == 8) LOCO_GN_SMALL(true, 8); else LOCO_GN_SMALL(true, 4); }
  else { if (vw == 8) LOCO_GN_SMALL(false, 8); else LOCO_GN_SMALL(false, 4); }
#undef LOCO_GN_SMALL
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int gn_small_fwd(View x, int n_primal, double* stats, const float* gamma, const float* beta, float eps, int silu,
                 int round_out, View y, cudaStream_t s) {
  LOCO_TRY(same_type(x, y, "gn_small_fwd"));
  LOCO_REQUIRE(gn_small_supported(x) && y.sW % 4 == 0 && y.sH == (long long)y.W * y.sW, "gn_small_fwd: layout not supported");
  LOCO_REQUIRE(n_primal == x.N || n_primal == 1, "gn_small_fwd: tangent rows need exactly one primal row");
  return gn_small_launch<0>(x, n_primal, x, nullptr, stats, gamma, beta, eps, silu, round_out, nullptr, 0, y, s);
}
int gn_small_vjp(View xp, const double* pstats, View gy, const float* gamma, const float* beta, float eps, int silu,
                 const View* addend, int accumulate, int round_out, View gx, cudaStream_t s) {
  LOCO_TRY(same_type(xp, gy, "gn_small_vjp")); LOCO_TRY(same_type(gy, gx, "gn_small_vjp(gx)"));
  LOCO_REQUIRE(gn_small_supported(xp) && gn_small_supported(gy) && gx.sH == (long long)gx.W * gx.sW &&
                   (!addend || addend->sH == (long long)addend->W * addend->sW),
               "gn_small_vjp: layout not supported");
  return gn_small_launch<1>(xp, 1, gy, pstats, nullptr, gamma, beta, eps, silu, round_out, addend, accumulate, gx, s);
}

int thin_pad(const float* in_nchw, int c, View out, const float* mix, int bias_rows, const float* scale_dev,
             int scale_from, int round_out, cudaStream_t s) {
  LOCO_REQUIRE(c >= 1 && c <= 4 && out.C == kThinPad, "thin_pad: %d channels into a %d-channel tensor", c, out.C);
  LOCO_REQUIRE((((uintptr_t)out.ptr) & 15) == 0 && out.sW % 8 == 0 && out.sH % 8 == 0 && out.sN % 8 == 0, "thin_pad: view not aligned");
  const long long total = (long long)out.N * out.H * out.W * (out.half ? kThinPad / 8 : kThinPad / 4);
  const int grid = grid_for(total, 256, num_sms() * 8);
  if (out.half) thin_pad_kernel<true><<<grid, 256, 0, s>>>(in_nchw, c, out, mix, bias_rows, scale_dev, scale_from, round_out);
  else thin_pad_kernel<false><<<grid, 256, 0, s>>>(in_nchw, c, out, mix, bias_rows, scale_dev, scale_from, round_out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int thin_extract(View in, int c, float* out_nchw, const float* mix, const float* scale_dev, int scale_from, cudaStream_t s) {
  LOCO_REQUIRE(c >= 1 && c <= 4 && in.C == kThinPad, "thin_extract: %d channels out of a %d-channel tensor", c, in.C);
  LOCO_REQUIRE((((uintptr_t)in.ptr) & 15) == 0 && in.sW % 8 == 0 && in.sH % 8 == 0 && in.sN % 8 == 0, "thin_extract: view not aligned");
  const long long total = (long long)in.N * in.H * in.W;
  const int grid = grid_for(total, 256, num_sms() * 8);
  if (in.half) thin_extract_kernel<true><<<grid, 256, 0, s>>>(in, c, out_nchw, mix, scale_dev, scale_from);
  else thin_extract_kernel<false><<<grid, 256, 0, s>>>(in, c, out_nchw, mix, scale_dev, scale_from);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int upsample2x(View in, View out, float scale, int accumulate, int round_out, cudaStream_t s) {
  LOCO_REQUIRE(out.H == 2 * in.H && out.W == 2 * in.W && out.C == in.C && out.N == in.N,
               "upsample2x: shape mismatch");
  const long long total = (long long)out.N * out.H * out.W * (out.C / 4);
  LOCO_TRY(same_type(in, out, "upsample2x"));
  const int vec = in.half ? 8 : 4;
  if (!accumulate && in.C % vec == 0 && in.sW % vec == 0 && in.sH % vec == 0 && in.sN % vec == 0 &&
      out.sW % vec == 0 && out.sH % vec == 0 && out.sN % vec == 0 && (((uintptr_t)in.ptr | (uintptr_t)out.ptr) & 15) == 0) {
    const long long tin = (long long)in.N * in.H * in.W * (in.C / vec);
    const int grid = grid_for(tin, 256, num_sms() * 8);
    if (in.half) upsample2x_fast_kernel<true><<<grid, 256, 0, s>>>(in, out, scale, round_out);
    else upsample2x_fast_kernel<false><<<grid, 256, 0, s>>>(in, out, scale, round_out);
    count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (in.half) upsample2x_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(in, out, scale, accumulate, round_out);
  else upsample2x_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(in, out, scale, accumulate, round_out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int sumpool2x(View in, View out, float scale, int accumulate, int round_out, cudaStream_t s) {
  LOCO_REQUIRE(in.H == 2 * out.H && in.W == 2 * out.W && out.C == in.C && out.N == in.N,
               "sumpool2x: shape mismatch");
  const long long total = (long long)out.N * out.H * out.W * (out.C / 4);
  LOCO_TRY(same_type(in, out, "sumpool2x"));
  if (in.half) sumpool2x_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(in, out, scale, accumulate, round_out);
  else sumpool2x_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(in, out, scale, accumulate, round_out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int add_views(View in, View out, int accumulate, cudaStream_t s) {
  LOCO_REQUIRE(in.H == out.H && in.W == out.W && out.C == in.C && out.N == in.N,
               "add_views: shape mismatch");
  const long long total = (long long)out.N * out.H * out.W * (out.C / 4);
  LOCO_TRY(same_type(in, out, "add_views"));
  if (in.half) add_views_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(in, out, accumulate);
  else add_views_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(in, out, accumulate);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int set_scalar(float* dst, float v, cudaStream_t s) {
  set_scalar_kernel<<<1, 1, 0, s>>>(dst, v);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int temb_forward(const float* t_dev, int ch, const float* w0, const float* b0, const float* w1,
                 const float* b1, float* scratch, int style, const float* cond, cudaStream_t s) {
  // scratch: [0,4ch) = silu(dense0(emb)), [4ch,8ch) = silu(dense1(.) + cond) = temb_act
  const int tch = 4 * ch;
  temb_dense_kernel<<<(tch + 7) / 8, 256, sizeof(float) * ch, s>>>(t_dev, style, 1, nullptr, ch, w0, b0, nullptr, tch, scratch);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  temb_dense_kernel<<<(tch + 7) / 8, 256, sizeof(float) * tch, s>>>(t_dev, style, 0, scratch, tch, w1, b1, cond, tch, scratch + tch);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int scale_shift_affine(const AffineSite* sites_dev, int n_sites, const float* weights,
                       const float* tproj, float* out, cudaStream_t s) {
  if (n_sites <= 0) return 0;
  scale_shift_affine_kernel<<<n_sites, 256, 0, s>>>(sites_dev, weights, tproj, out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int context_kv(const float* ctx, int n_tok, int dim, const float* w, const float* b, int cout, int rows, float* out,
               cudaStream_t s) {
  const long long outs = (long long)rows * cout;
  context_kv_kernel<<<(unsigned)((outs + 7) / 8), 256, 0, s>>>(ctx, n_tok, dim, w, b, cout, rows, out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int temb_project(const float* temb_act, int temb_ch, const float* w, const float* b, int cout,
                 float* out, cudaStream_t s) {
  temb_project_kernel<<<(cout + 7) / 8, 256, 0, s>>>(temb_act, temb_ch, w, b, cout, out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace loco
