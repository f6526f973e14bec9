// The two 3-channel "edge" convolutions of the U-Net (conv_in 3 -> 128 and conv_out 128 -> 3, and each
// other's data gradients in the VJP program) on the tensor cores, for fp16 activation tensors.
//
// The CUDA-core versions in layers.cu spend 3 456 multiply-adds per pixel on the FMA pipe: 1.0 ms for
// conv_out on a [40, 256, 256, 128] batch (0.66 TB/s, a tenth of what the tensor is worth in HBM
// time), 7 % of a B = 40 forward.  With 3 real channels on one side the GEMM is far too thin for a
// tcgen05 tile (N = 3 or K = 27), so these kernels use warp-level mma.sync.m16n8k16 (fp16 operands,
// fp32 accumulation) on a pixel tile staged once in shared memory:
//
//   reduce (128 -> 3): a CTA owns 8 x 32 pixels; the (8+2) x (32+2) halo of all 128 channels is
//     copied into shared memory once (cp.async, zero fill outside the image = the padding), pixel
//     pitch 272 B so that ldmatrix rows fall into distinct banks; a warp owns one pixel row = two
//     m16 tiles; per filter tap and 16-channel slab one ldmatrix.x4 per tile + one mma (N = 8 output
//     slots, 3 used); the weights sit in shared memory as ready-made B fragments.
//   expand (3 -> 128): the 27 inputs of a pixel (3 x 3 window x 3 planes, range-scaled) are the K
//     dimension (padded to 32) of an im2col tile built in shared memory; 16 n8 tiles give the 128
//     output channels; the fp16 result is staged through shared memory for 16-byte NHWC stores.
#include "layers.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace loco {

namespace {

constexpr int kTH = 8, kTW = 32;                 // pixel tile of a CTA (one warp per row)
constexpr int kPitch = 272;                      // bytes per pixel of a 128-channel fp16 smem row (+16 B pad)
constexpr int kHaloPix = (kTH + 2) * (kTW + 2);
constexpr int kRedSmem = kHaloPix * kPitch + 72 * 32 * 8;            // halo tile + B fragments
constexpr int kAPitch = 80;                      // bytes per pixel of the im2col row (32 fp16 + 16 B pad)
constexpr int kExpSmem = kTH * kTW * kAPitch + 2 * 16 * 32 * 8 + kTH * kTW * kPitch;   // A + B fragments + output staging

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&a)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(saddr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, bool valid) {
  const int bytes = valid ? 16 : 0;              // 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ---- 128 -> 3 ------------------------------------------------------------------------------------
// Wr: [tap][j][128] fp32 (pack_conv_edge); out3: [N, 3, H, W] fp32.  flip: data-gradient form.
__global__ void __launch_bounds__(256, 2)
edge_reduce_mma_kernel(View in, const float* __restrict__ Wr, const float* __restrict__ bias, int bias_rows,
                       float* __restrict__ out3, int flip, const float* __restrict__ scale_dev, int scale_from) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* tile = smem;
  uint2* bfrag = reinterpret_cast<uint2*>(smem + kHaloPix * kPitch);   // [72 k-steps][32 lanes]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = in.H, W = in.W;
  // B fragments: k-step ks = (window offset wo = 3 r + c, 16-channel slab c16); element (k, n):
  // k = channel c16 * 16 + kk, n = output slot (3 used).  Thread layout of mma.m16n8k16 B:
  // b0 = {(k = 2 (lane % 4), n = lane / 4), (k + 1, n)}, b1 = the same at k + 8.
  for (int e = threadIdx.x; e < 72 * 32; e += blockDim.x) {
    const int ks = e >> 5, l = e & 31;
    const int wo = ks >> 3, c16 = ks & 7;
    const int r = wo / 3, c = wo % 3;
    const int t = flip ? (2 - r) * 3 + (2 - c) : wo;     // filter tap read at window offset (r, c)
    const int n = l >> 2, k0 = c16 * 16 + (l & 3) * 2;
    uint2 v = make_uint2(0u, 0u);
    if (n < 3) {
      const float* w = Wr + (t * 3 + n) * 128;
      v.x = pack_half2(w[k0], w[k0 + 1]);
      v.y = pack_half2(w[k0 + 8], w[k0 + 9]);
    }
    bfrag[e] = v;
  }
  const int tiles_x = W / kTW, tiles_y = H / kTH;
  const long long ntiles = (long long)in.N * tiles_y * tiles_x;
  const __half* src = reinterpret_cast<const __half*>(in.ptr);
  const uint32_t tile_s = smem_u32(tile);
  for (long long ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const int tx = (int)(ti % tiles_x), ty = (int)((ti / tiles_x) % tiles_y), n = (int)(ti / ((long long)tiles_x * tiles_y));
    const int x0 = tx * kTW, y0 = ty * kTH;
    __syncthreads();                               // previous tile consumed (and the fragments written)
    for (int e = threadIdx.x; e < kHaloPix * 16; e += blockDim.x) {
      const int px = e >> 4, ch = e & 15;
      const int yy = y0 + px / (kTW + 2) - 1, xx = x0 + px % (kTW + 2) - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const __half* g = src + (long long)n * in.sN + (long long)(ok ? yy : 0) * in.sH + (long long)(ok ? xx : 0) * in.sW + ch * 8;
      cp_async16(tile_s + px * kPitch + ch * 16, g, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float acc[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][i] = 0.f;
    // ldmatrix row of this lane: pixel (lane & 15) of the m-tile, k half (lane >> 4)
    const uint32_t lrow = (uint32_t)((lane & 15) * kPitch + (lane >> 4) * 16);
#pragma unroll 1
    for (int wo = 0; wo < 9; ++wo) {
      const int r = wo / 3, c = wo % 3;
      const uint32_t base = tile_s + (uint32_t)(((warp + r) * (kTW + 2) + c) * kPitch) + lrow;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const uint2 b = bfrag[(wo * 8 + c16) * 32 + lane];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          uint32_t a[4];
          ldmatrix_x4(a, base + (uint32_t)(m * 16 * kPitch + c16 * 32));
          mma16816(acc[m], a, b.x, b.y);
        }
      }
    }
    // accumulator (row g = lane / 4 (+8), column 2 (lane % 4) (+1)): columns 0..2 are the outputs
    const float sc = (scale_dev != nullptr && n >= scale_from) ? scale_dev[1] : 1.f;
    const bool ub = bias != nullptr && n < bias_rows;
    const int col = (lane & 3) * 2;
    if (col < 3) {
      const int y = y0 + warp;
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int x = x0 + m * 16 + (lane >> 2) + hrow * 8;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int j = col + cc;
            if (j < 3) out3[((long long)n * 3 + j) * H * W + (long long)y * W + x] = acc[m][hrow * 2 + cc] * sc + (ub ? bias[j] : 0.f);
          }
        }
    }
  }
}

// ---- 3 -> 128 ------------------------------------------------------------------------------------
// in3: [N, 3, H, W] fp32; We: [tap][j][128] fp32; out: fp16 NHWC view with 128 channels.
__global__ void __launch_bounds__(256, 2)
edge_expand_mma_kernel(const float* __restrict__ in3, const float* __restrict__ We, const float* __restrict__ bias,
                       int bias_rows, View out, int flip, const float* __restrict__ scale_dev, int scale_from) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* atile = smem;                                                   // [256 px][kAPitch]
  uint2* bfrag = reinterpret_cast<uint2*>(smem + kTH * kTW * kAPitch);     // [2 k-steps][16 n-tiles][32 lanes]
  uint8_t* otile = smem + kTH * kTW * kAPitch + 2 * 16 * 32 * 8;           // [256 px][kPitch]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = out.H, W = out.W;
  // B fragments: k = (window offset wo) * 3 + plane j (27 used of 32), n = output channel
  for (int e = threadIdx.x; e < 2 * 16 * 32; e += blockDim.x) {
    const int l = e & 31, nt = (e >> 5) & 15, ks = e >> 9;
    const int n = nt * 8 + (l >> 2);
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = ks * 16 + (l & 3) * 2 + (i & 1) + (i >> 1) * 8;
      w[i] = 0.f;
      if (k < 27) {
        const int wo = k / 3, j = k % 3;
        const int r = wo / 3, c = wo % 3;
        const int t = flip ? (2 - r) * 3 + (2 - c) : wo;
        w[i] = We[(t * 3 + j) * 128 + n];
      }
    }
    bfrag[e] = make_uint2(pack_half2(w[0], w[1]), pack_half2(w[2], w[3]));
  }
  __syncthreads();                                 // the fragments are visible to every warp
  const int tiles_x = W / kTW, tiles_y = H / kTH;
  const long long ntiles = (long long)out.N * tiles_y * tiles_x;
  const uint32_t atile_s = smem_u32(atile);
  __half* dst = reinterpret_cast<__half*>(out.ptr);
  for (long long ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
    const int tx = (int)(ti % tiles_x), ty = (int)((ti / tiles_x) % tiles_y), n = (int)(ti / ((long long)tiles_x * tiles_y));
    const int x0 = tx * kTW, y0 = ty * kTH;
    const float sc = (scale_dev != nullptr && n >= scale_from) ? scale_dev[0] : 1.f;
    // a warp owns pixel row `warp` of the tile end to end: its 32 im2col rows, its 32 staging rows and their
    // stores are private to it, so the warps of a CTA run through the tiles decoupled (no block barrier in the loop)
    __syncwarp();                                  // previous tile's staging rows drained
    {
      // im2col row of pixel threadIdx.x: 27 range-scaled inputs, 5 zeros
      const int py = threadIdx.x / kTW, pxx = threadIdx.x % kTW;
      const float* plane = in3 + (long long)n * 3 * H * W;
      float v[32];
#pragma unroll
      for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
      for (int wo = 0; wo < 9; ++wo) {
        const int yy = y0 + py + wo / 3 - 1, xx = x0 + pxx + wo % 3 - 1;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
        for (int j = 0; j < 3; ++j) v[wo * 3 + j] = ok ? __ldg(plane + ((long long)j * H + yy) * W + xx) * sc : 0.f;
      }
      uint4* row = reinterpret_cast<uint4*>(atile + threadIdx.x * kAPitch);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        row[q] = make_uint4(pack_half2(v[8 * q], v[8 * q + 1]), pack_half2(v[8 * q + 2], v[8 * q + 3]),
                            pack_half2(v[8 * q + 4], v[8 * q + 5]), pack_half2(v[8 * q + 6], v[8 * q + 7]));
    }
    __syncwarp();
    const bool ub = bias != nullptr && n < bias_rows;
    const uint32_t lrow = (uint32_t)((lane & 15) * kAPitch + (lane >> 4) * 16);
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
      uint32_t a[2][4];
      const uint32_t base = atile_s + (uint32_t)((warp * kTW + m * 16) * kAPitch) + lrow;
      ldmatrix_x4(a[0], base);
      ldmatrix_x4(a[1], base + 32);
      // the 16 n-tiles in two halves of 8 (32 accumulators live at a time)
#pragma unroll 1
      for (int nh = 0; nh < 2; ++nh) {
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint2 b = bfrag[(ks * 16 + nh * 8 + nt) * 32 + lane];
            mma16816(acc[nt], a[ks], b.x, b.y);
          }
        }
        // (row g (+8), channels 2 (lane % 4) (+1) of n-tile) -> fp16 staging tile
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int ch = (nh * 8 + nt) * 8 + (lane & 3) * 2;
          float b0 = 0.f, b1 = 0.f;
          if (ub) { b0 = bias[ch]; b1 = bias[ch + 1]; }
#pragma unroll
          for (int hrow = 0; hrow < 2; ++hrow) {
            const int p = warp * kTW + m * 16 + (lane >> 2) + hrow * 8;
            *reinterpret_cast<uint32_t*>(otile + p * kPitch + ch * 2) =
                pack_half2(acc[nt][hrow * 2] + b0, acc[nt][hrow * 2 + 1] + b1);
          }
        }
      }
    }
    __syncwarp();
    // the warp's 32 pixels x 256 B -> NHWC, 16 bytes per thread and step (16 threads cover a pixel)
    for (int e = lane; e < kTW * 16; e += 32) {
      const int p = warp * kTW + (e >> 4), chunk = e & 15;
      const int y = y0 + p / kTW, x = x0 + p % kTW;
      *reinterpret_cast<uint4*>(dst + (long long)n * out.sN + (long long)y * out.sH + (long long)x * out.sW + chunk * 8) =
          *reinterpret_cast<const uint4*>(otile + p * kPitch + chunk * 16);
    }
  }
}

bool edge_mma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LOCO_EDGE_MMA");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

}  // namespace

int edge_mma_init() {
  static bool done[kMaxDevices] = {false};
  if (!first_time_on_device(done)) return 0;
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(edge_reduce_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRedSmem));
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(edge_expand_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kExpSmem));
  return 0;
}

bool edge_mma_eligible(const View& v) {
  return edge_mma_enabled() && v.half && v.C == 128 && v.H % kTH == 0 && v.W % kTW == 0 && v.sW % 8 == 0 && v.sH % 8 == 0 &&
         v.sN % 8 == 0 && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0;
}

int edge_conv_reduce_mma(View in, const float* Wr, const float* bias, int bias_rows, float* out3, int flip, cudaStream_t s,
                         const float* scale_dev, int scale_from) {
  const long long ntiles = (long long)in.N * (in.H / kTH) * (in.W / kTW);
  const int grid = (int)(ntiles < 2LL * num_sms() ? ntiles : 2LL * num_sms());
  ProfScope prof(2, 0, s);
  edge_reduce_mma_kernel<<<grid, 256, kRedSmem, s>>>(in, Wr, bias, bias_rows, out3, flip, scale_dev, scale_from);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int edge_conv_expand_mma(const float* in3, const float* We, const float* bias, int bias_rows, View out, int flip,
                         cudaStream_t s, const float* scale_dev, int scale_from) {
  const long long ntiles = (long long)out.N * (out.H / kTH) * (out.W / kTW);
  const int grid = (int)(ntiles < 2LL * num_sms() ? ntiles : 2LL * num_sms());
  ProfScope prof(2, 0, s);
  edge_expand_mma_kernel<<<grid, 256, kExpSmem, s>>>(in3, We, bias, bias_rows, out, flip, scale_dev, scale_from);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace loco
