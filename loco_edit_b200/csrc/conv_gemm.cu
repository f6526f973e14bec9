// Implicit-GEMM convolution for the LOCO-Edit U-Net hot path on Blackwell tensor cores.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap_offset, cin] * Wp[cout, tap*Cin + cin]
//
// * A tiles (128 pixels x 32 fp32 channels) are fetched by TMA straight from the channels-last
//   activation tensor as 4-D boxes; out-of-image coordinates are zero-filled by the TMA unit, which
//   implements the convolution padding (and the reference's asymmetric (0,1,0,1) stride-2 pad,
//   ddpm/diffusion.py:846-850) without any im2col buffer.
// * Weight tiles come from a pre-packed K-major matrix.  Both land in 128B-swizzled shared memory
//   and are consumed by tcgen05.mma.kind::tf32 with the fp32 accumulator in TMEM.
// * The batch dimension stacks the primal sample and the k probe tangents (JVP) or the k
//   cotangents (VJP): a convolution is linear, so every row of the batch shares the weight tile.
// * Warp-specialised persistent CTAs: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
//   allocator, warps 4-7 = epilogue (bias / timestep bias on primal rows, residual add, VJP
//   accumulation, tf32 rounding).  Two TMEM accumulator stages overlap epilogue and main loop.
#include "conv_gemm.cuh"

namespace loco {

namespace {

constexpr int kABytes = kConvBlockM * 128;       // 16 KB: 128 pixels x 32 fp32 channels
constexpr int kBBytes = kConvMaxBlockN * 128;    // 16 KB: 128 output channels x 32 fp32 channels
constexpr int kThreads = 256;
constexpr int kEpiPitch = 36;                    // floats per pixel row of a transpose buffer (16 B aligned, conflict-free)
constexpr int kEpiBytes = 4 * 32 * kEpiPitch * 4;   // per-epilogue-warp transpose buffers
// NT = M tiles (of 128 pixels) per work item.  NT = 2 shares every weight tile between two pixel
// tiles: 48 KB of operands per two MMA groups instead of 64 KB, which matters because the kernel
// is bound by the ~70 B/clk an SM can ingest from L2, not by the tensor pipe.
template <int NT> struct Cfg {
  static constexpr int kStages = NT == 1 ? 6 : 4;
  static constexpr int kStageBytes = NT * kABytes + kBBytes;
  static constexpr int kTmemCols = NT * 256;     // 2 accumulator stages x NT x 128 fp32 columns
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiBytes;
};

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle, rows of 128 bytes, 8-row
// groups 1024 bytes apart (cute::UMMA::SmemDescriptor layout; version = 1 for sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);   // start address
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                   // descriptor version
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}

// Instruction descriptor, fp32 accumulate, both operands K-major: kind::tf32 (operand format 2) or
// kind::f16 with fp16 operands (format 0).
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int in16) {
  uint32_t d = 0;
  d |= 1u << 4;                 // C format = F32
  if (!in16) {
    d |= 2u << 7;               // A format = TF32
    d |= 2u << 10;              // B format = TF32
  }                             // (F16 = 0 for kind::f16)
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}
// one K step of 32 bytes per operand row: 8 tf32 or 16 fp16 elements
// (the element types are template parameters of the kernels: the fp32 / tf32 instantiations are
// instruction-for-instruction the kernels of round 1, the epilogue runs one warp per SM
// sub-partition and its time is its instruction count)
template <bool IN16>
__device__ __forceinline__ void umma_any(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (IN16) umma_f16(d_tmem, a, b, idesc, acc); else umma_tf32(d_tmem, a, b, idesc, acc);
}
template <bool IN16>
__device__ __forceinline__ void umma_any_pair(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (IN16) umma_f16_pair(d_tmem, a, b, idesc, acc); else umma_tf32_pair(d_tmem, a, b, idesc, acc);
}

// L2 prefetch of the epilogue's read-side tile (residual addend and/or the accumulate target) of a
// 128-pixel tile: issued by the epilogue warps before they wait for the accumulator, so the DRAM
// latency of those reads is hidden behind the MMAs of the same tile.  512-byte pixel rows of
// block_n fp32 channels = block_n / 32 lines; thread `t` (0..127) takes pixel row t.
__device__ __forceinline__ void prefetch_l2(const float* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
template <bool OUT16>
__device__ __forceinline__ void prefetch_epilogue_tile(const ConvGemmParams& p, int tx, int ty, int tn,
                                                       int co0, int t) {
  if (p.addend == nullptr && !p.accumulate) return;
  const int R = t;
  const int x = tx * p.TW + (R & (p.TW - 1));
  const int y = ty * p.TH + ((R >> p.log_tw) & (p.TH - 1));
  const int n = tn * p.TN + (R >> (p.log_tw + p.log_th));
  if (n >= p.N || y >= p.Ho || x >= p.Wo) return;
  constexpr int esz = OUT16 ? 2 : 4;
  constexpr int line = 128 / esz;         // channels per 128-byte line
  if (p.addend) {
    const char* a = reinterpret_cast<const char*>(p.addend) +
                    ((long long)n * p.add_sN + (long long)y * p.add_sH + (long long)x * p.add_sW + co0) * esz;
    for (int c = 0; c < p.block_n; c += line) prefetch_l2(reinterpret_cast<const float*>(a + c * esz));
  }
  if (p.accumulate) {
    const char* o = reinterpret_cast<const char*>(p.out) +
                    ((long long)n * p.out_sN + (long long)y * p.out_sH + (long long)x * p.out_sW + co0) * esz;
    for (int c = 0; c < p.block_n; c += line) prefetch_l2(reinterpret_cast<const float*>(o + c * esz));
  }
}


// TMEM gives each thread one pixel row (32 consecutive channels); the padded shared-memory transpose
// turns that into "8 lanes x float4 = 128 contiguous bytes of one pixel" per load/store instruction.
// 8 STS.128 + 8 LDS.128 per thread, both bank-conflict-free at a pitch of 36 floats.
__device__ __forceinline__ void epilogue_transpose(const uint32_t (&r)[32], float* tbuf, int lane, int pr, int cq,
                                                   float4 (&v)[8]) {
  float4* trow = reinterpret_cast<float4*>(tbuf + lane * kEpiPitch);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    trow[c] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]),
                          __uint_as_float(r[4 * c + 3]));
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it)
    v[it] = *reinterpret_cast<const float4*>(tbuf + (it * 4 + pr) * kEpiPitch + cq * 4);
  __syncwarp();
}

// Everything after the accumulator values of one 32-channel chunk are in registers (v[it] = channels
// [ch + 4 cq, +4) of pixel row it*4 + pr of this warp's 32 rows): bias on the primal rows, residual,
// VJP accumulation, tf32 rounding, 128-byte-coalesced stores, fused GroupNorm statistics.  The common
// cases (whole tile valid, bias on all or none of the rows, no statistics) take branch-free paths:
// the epilogue runs one warp per SM sub-partition, so its time is its instruction count.
template <bool OUT16>
__device__ __forceinline__ void epilogue_chunk_tail(const ConvGemmParams& p, float4 (&v)[8],
                                                    const long long (&ooff)[8], const long long (&aoff)[8],
                                                    uint32_t vmask, uint32_t bmask, float4 bsum, int ch,
                                                    int co0, int cq, int pr, int nrow) {
  if (bmask == 0xFFu) {
#pragma unroll
    for (int it = 0; it < 8; ++it) { v[it].x += bsum.x; v[it].y += bsum.y; v[it].z += bsum.z; v[it].w += bsum.w; }
  } else if (bmask != 0u) {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      if ((bmask >> it) & 1u) { v[it].x += bsum.x; v[it].y += bsum.y; v[it].z += bsum.z; v[it].w += bsum.w; }
  }
  constexpr bool h = OUT16;
  if (p.addend) {
    float4 a[8];
#pragma unroll
    for (int it = 0; it < 8; ++it)
      a[it] = ((vmask >> it) & 1u) ? ld4t<OUT16>(p.addend, aoff[it] + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < 8; ++it) { v[it].x += a[it].x; v[it].y += a[it].y; v[it].z += a[it].z; v[it].w += a[it].w; }
  }
  if (p.accumulate) {
    float4 a[8];
#pragma unroll
    for (int it = 0; it < 8; ++it)
      a[it] = ((vmask >> it) & 1u) ? ld4t<OUT16>(p.out, ooff[it] + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < 8; ++it) { v[it].x += a[it].x; v[it].y += a[it].y; v[it].z += a[it].z; v[it].w += a[it].w; }
  }
  if (h) {
    // fp16 storage: what is stored (and what the fused statistics must see) is the fp16 rounding
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      v[it].x = round_f16(v[it].x); v[it].y = round_f16(v[it].y);
      v[it].z = round_f16(v[it].z); v[it].w = round_f16(v[it].w);
    }
  } else if (p.round_out) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      v[it].x = round_tf32(v[it].x); v[it].y = round_tf32(v[it].y);
      v[it].z = round_tf32(v[it].z); v[it].w = round_tf32(v[it].w);
    }
  }
  if (vmask == 0xFFu) {
#pragma unroll
    for (int it = 0; it < 8; ++it) st4t<OUT16>(p.out, ooff[it] + ch, v[it]);
  } else {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      if ((vmask >> it) & 1u) st4t<OUT16>(p.out, ooff[it] + ch, v[it]);
  }
  if (p.st_ptr[0] != nullptr || p.st_ptr[1] != nullptr) {
    // fused GroupNorm statistics of what was just stored: the 4 lanes that share a channel quad
    // (different pixels) combine, then one fp64 atomic pair per consumer
    float gs1 = 0.f, gs2 = 0.f;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if (!((vmask >> it) & 1u)) continue;
      const float4 o = v[it];
      gs1 += (o.x + o.y) + (o.z + o.w);
      gs2 += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
    }
    gs1 += __shfl_xor_sync(0xffffffffu, gs1, 8);  gs2 += __shfl_xor_sync(0xffffffffu, gs2, 8);
    gs1 += __shfl_xor_sync(0xffffffffu, gs1, 16); gs2 += __shfl_xor_sync(0xffffffffu, gs2, 16);
    if (pr == 0 && nrow < p.N) {
#pragma unroll
      for (int tg = 0; tg < 2; ++tg) {
        if (p.st_ptr[tg] == nullptr) continue;
        const int g = (p.st_choff[tg] + co0 + ch + cq * 4) / p.st_cg[tg];
        double* dst = p.st_ptr[tg] + ((long long)nrow * 32 + g) * 2;
        atomicAdd(dst, (double)gs1);
        atomicAdd(dst + 1, (double)gs2);
      }
    }
  }
}

template <int NT, bool IN16, bool OUT16>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tf32_kernel(const __grid_constant__ ConvGemmParams p) {
  constexpr int kStages = Cfg<NT>::kStages;
  constexpr int kStageBytes = Cfg<NT>::kStageBytes;
  constexpr int kTmemCols = Cfg<NT>::kTmemCols;
  constexpr int kBOff = NT * kABytes;            // weight tile offset inside a stage
  if (p.debug == 1) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* epi_smem = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.amap[0]);
    prefetch_tmap(&p.bmap);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                            // inputs of this launch are complete and visible from here on
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_n;
  const int units_m = tiles_m / NT;              // NT == 2 requires an even tile count
  const int total_units = units_m * p.tiles_co;
  const int kiters = p.ntaps * p.c_chunks;
  // Work item = (unit of NT pixel tiles x one output-channel tile, K split).  Layers with fewer
  // tiles than SMs split their K loop over several CTAs (NT == 1 only); the partial tiles meet in
  // L2 and every split CTA reduces + finishes its share of the tile.
  const int ksplit = NT == 1 ? p.ksplit : 1;
  const int total_items = p.debug == 2 ? 0 : total_units * ksplit;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = NT * kABytes + (uint32_t)p.block_n * 128u;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const int unit = w % total_units;
        const int split = w / total_units;
        const int k0 = (int)((long long)kiters * split / ksplit);
        const int k1 = (int)((long long)kiters * (split + 1) / ksplit);
        const int um = unit % units_m;
        const int co0 = (unit / units_m) * p.block_n;
        int x0[NT], y0[NT], n0[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int tm = um * NT + j;
          x0[j] = (tm % p.tiles_x) * p.TW;
          y0[j] = ((tm / p.tiles_x) % p.tiles_y) * p.TH;
          n0[j] = (tm / (p.tiles_x * p.tiles_y)) * p.TN;
        }
        int tap = k0 / p.c_chunks, cc = k0 % p.c_chunks;
        for (int it = k0; it < k1; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          if (p.debug == 3 || p.debug == 6) {
            mbar_arrive(&full_bar[stage]);
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          const CUtensorMap* am = &p.amap[p.tap_map[tap]];
#pragma unroll
          for (int j = 0; j < NT; ++j)
            tma_load_4d(sa + j * kABytes, am, &full_bar[stage], cc * (IN16 ? 2 * kConvBlockK : kConvBlockK), x0[j] + p.tap_dx[tap],
                        y0[j] + p.tap_dy[tap], n0[j]);
          tma_load_2d(sa + kBOff, &p.bmap, &full_bar[stage], p.tap_wk[tap] + cc * (IN16 ? 2 * kConvBlockK : kConvBlockK), co0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          if (++cc == p.c_chunks) { cc = 0; ++tap; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc(kConvBlockM, p.block_n, IN16);
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem));            // stage 0, pixel tile 0
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem) + kBOff);    // stage 0, weight tile
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const int split = w / total_units;
        const int nk = (int)((long long)kiters * (split + 1) / ksplit) - (int)((long long)kiters * split / ksplit);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NT) * kConvMaxBlockN;
        // The barrier wait of stage s+1 is issued before the last MMA of stage s, so its latency
        // (and the descriptor arithmetic) overlaps with MMAs that are still executing.
        if (nk > 0) mbar_wait(&full_bar[stage], phase);
        for (int it = 0; it < nk; ++it) {
          tc_fence_after();
          if (p.debug == 4) {
            mbar_arrive(&empty_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            if (it + 1 < nk) mbar_wait(&full_bar[stage], phase);
            continue;
          }
          const uint64_t so = (uint64_t)stage * (uint64_t)(kStageBytes >> 4);   // stage offset, 16 B units
          const uint64_t bdesc = bdesc0 + so;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == kStages) { nstage = 0; nphase ^= 1; }
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const uint64_t adesc = adesc0 + so + (uint64_t)(j * (kABytes >> 4));
#pragma unroll
            for (int k = 0; k < kConvBlockK / 8; ++k) {
              if (j == NT - 1 && k == kConvBlockK / 8 - 1 && it + 1 < nk) mbar_wait(&full_bar[nstage], nphase);
              // advance 8 tf32 = 32 bytes inside the 128-byte swizzle row (+2 in 16-byte units)
              umma_any<IN16>(d_tmem + j * kConvMaxBlockN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2),
                        idesc, (uint32_t)((it | k) != 0));
            }
          }
          umma_commit(&empty_bar[stage]);
          stage = nstage; phase = nphase;
        }
        if (p.debug == 4) mbar_arrive(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    // TMEM gives each thread one pixel row; a padded shared-memory transpose turns that into
    // "8 lanes x float4 = 128 contiguous bytes of one pixel" so every global load/store of the
    // residual / accumulate / output / split-K partial streams is a fully used 128-byte segment.
    const int q = warp & 3;              // TMEM lane quarter owned by this warp
    float* tbuf = epi_smem + q * (32 * kEpiPitch);
    const int pr = lane >> 3;            // pixel sub-row handled by this lane (0..3)
    const int cq = lane & 7;             // channel quad within the 32-column chunk
    const int lTW = p.log_tw, lTH = p.log_th;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      const int unit = w % total_units;
      const int split = w / total_units;
      const int um = unit % units_m;
      const int co0 = (unit / units_m) * p.block_n;
      if (ksplit == 1) {
#pragma unroll
        for (int jt = 0; jt < NT; ++jt) {
          const int tm = um * NT + jt;
          prefetch_epilogue_tile<OUT16>(p, tm % p.tiles_x, (tm / p.tiles_x) % p.tiles_y,
                                 tm / (p.tiles_x * p.tiles_y), co0, (int)threadIdx.x - 128);
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int jt = 0; jt < ((p.debug == 5 || p.debug == 6) ? 0 : NT); ++jt) {
        const int tm = um * NT + jt;
        const int tx = tm % p.tiles_x;
        const int ty = (tm / p.tiles_x) % p.tiles_y;
        const int tn = tm / (p.tiles_x * p.tiles_y);
        long long ooff[8], aoff[8];
        uint32_t vmask = 0, bmask = 0;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int R = q * 32 + it * 4 + pr;
          const int x = tx * p.TW + (R & (p.TW - 1));
          const int y = ty * p.TH + ((R >> lTW) & (p.TH - 1));
          const int n = tn * p.TN + (R >> (lTW + lTH));
          if (n < p.N && y < p.Ho && x < p.Wo) vmask |= 1u << it;
          if (n < p.bias_rows) bmask |= 1u << it;
          ooff[it] = (long long)n * p.out_sN + (long long)y * p.out_sH + (long long)x * p.out_sW + co0 + cq * 4;
          aoff[it] = (long long)n * p.add_sN + (long long)y * p.add_sH + (long long)x * p.add_sW + co0 + cq * 4;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                               (uint32_t)(acc * NT + jt) * kConvMaxBlockN;
        // partial tile of this (unit, split): [128 rows][block_n] fp32, row-major (split-K only)
        float* pbase = p.partial + ((long long)unit * ksplit) * (kConvBlockM * p.block_n);
        if (ksplit > 1) {
          // ---- phase 1: publish the raw accumulators of this K slice ----
          float* pmine = pbase + (long long)split * (kConvBlockM * p.block_n) +
                         (long long)(q * 32 + pr) * p.block_n + cq * 4;
          for (int ch = 0; ch < p.block_n; ch += 32) {
            uint32_t r[32];
            float4 t4[8];
            tmem_ld_32x32(taddr + ch, r);
            tmem_ld_wait();
            epilogue_transpose(r, tbuf, lane, pr, cq, t4);
#pragma unroll
            for (int it = 0; it < 8; ++it)
              __stcg(reinterpret_cast<float4*>(pmine + (long long)(it * 4) * p.block_n + ch), t4[it]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);     // accumulator stage is free again
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (threadIdx.x == 128) {
            // publish, then wait until every K slice of this tile is in L2.  All items of a split
            // layer are co-resident (items <= #SMs, one per CTA), so this cannot deadlock; the
            // bound turns a protocol bug into a launch error instead of a hang.
            atomicAdd(&p.counters[unit], 1);
            uint32_t spins = 0;
            while (*reinterpret_cast<volatile int*>(&p.counters[unit]) < ksplit) {
              __nanosleep(64);
              if (++spins > (1u << 22)) { printf("loco: split-K wait timeout unit %d\n", unit); __trap(); }
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          __threadfence();
        }
        for (int ch = 0; ch < p.block_n; ch += 32) {
          // split-K: the (column chunk, lane quarter) units of the tile are dealt round-robin to
          // the ksplit CTAs, so the final reduction + epilogue is spread over all of them
          if (ksplit > 1 && (((ch >> 5) * 4 + q) % ksplit) != split) continue;
          float4 v[8];
          if (ksplit > 1) {
            // ---- phase 2: fixed-order sum of the K slices (bit-reproducible) ----
#pragma unroll
            for (int it = 0; it < 8; ++it) v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* pr0 = pbase + (long long)(q * 32 + pr) * p.block_n + cq * 4 + ch;
            // the loads of four slices are in flight together (one L2 round trip per four slices instead
            // of one per slice: 16 slices took 7.5 of the kernel's 20 us), the sums keep their order
            const long long sstride = (long long)kConvBlockM * p.block_n;
            int sidx = 0;
            for (; sidx + 4 <= ksplit; sidx += 4) {
              float4 t[4][8];
#pragma unroll
              for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int it = 0; it < 8; ++it)
                  t[u][it] = __ldcg(reinterpret_cast<const float4*>(pr0 + (sidx + u) * sstride + (long long)(it * 4) * p.block_n));
#pragma unroll
              for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int it = 0; it < 8; ++it) { v[it].x += t[u][it].x; v[it].y += t[u][it].y; v[it].z += t[u][it].z; v[it].w += t[u][it].w; }
            }
            for (; sidx < ksplit; ++sidx) {
              float4 t[8];
#pragma unroll
              for (int it = 0; it < 8; ++it)
                t[it] = __ldcg(reinterpret_cast<const float4*>(pr0 + sidx * sstride + (long long)(it * 4) * p.block_n));
#pragma unroll
              for (int it = 0; it < 8; ++it) { v[it].x += t[it].x; v[it].y += t[it].y; v[it].z += t[it].z; v[it].w += t[it].w; }
            }
          } else {
            uint32_t r[32];
            tmem_ld_32x32(taddr + ch, r);
            tmem_ld_wait();
            epilogue_transpose(r, tbuf, lane, pr, cq, v);
          }
          float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bmask != 0u) {
            if (p.bias) bsum = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + ch + cq * 4));
            if (p.bias2) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias2 + co0 + ch + cq * 4));
              bsum.x += b.x; bsum.y += b.y; bsum.z += b.z; bsum.w += b.w;
            }
          }
          epilogue_chunk_tail<OUT16>(p, v, ooff, aoff, vmask, bmask, bsum, ch, co0, cq, pr,
                              tn * p.TN + ((q * 32) >> (lTW + lTH)));
        }
      }
      if (ksplit == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      } else {
        // last CTA out resets the counters for the next launch
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 128) {
          const int old = atomicAdd(&p.counters[p.counter_stride + unit], 1);
          if (old == ksplit - 1) {
            p.counters[unit] = 0;
            p.counters[p.counter_stride + unit] = 0;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// "Wide" variant for the large layers: operand roles swapped.  The weight tile (128 output channels
// x 32 input channels) is the M operand, 256 pixels (two pixel tiles) are the N operand of one
// tcgen05.mma 128x256x8.  Per 32-channel stage the tensor core then reads 4 x (4 KB + 8 KB) = 48 KB
// of shared memory instead of 8 x (4 KB + 4 KB) = 64 KB, which is what bounds this kernel (shared
// memory serves both the TMA fills and the MMA operand reads).  The accumulator is [cout][pixel]
// (TMEM lane = output channel); the epilogue's shared-memory transpose turns it back into
// channels-last 128-byte segments.
// ------------------------------------------------------------------------------------------------
//
// HALO = true ("halo" variant, 3x3 stride-1 fprop/dgrad): the work item is a 16 x 16 pixel block.
// Instead of one 128-pixel box per filter tap (9 x 32 KB of activations per 32-channel slab), the
// producer loads, per slab, three x-shifted copies of the (16 + 2)-row halo window (3 x 36 KB); the
// three taps of a filter column then address the same copy at row offsets 0 / 16 / 32 (2 KB steps,
// which keeps the 1024-byte swizzle atoms aligned).  L2 -> SM traffic per MMA clock drops from 96
// to 55 bytes (activations 2.67x less), below what the L2 can sustain per SM, so the kernel turns
// from L2-ingest-bound into tensor-pipe-bound.  Weights (16 KB per tap) ride a second, finer ring.
// ------------------------------------------------------------------------------------------------
constexpr int kHaloRows = 18 * 16;                  // (16 + 2) image rows x 16 pixels
constexpr int kHaloABytes = kHaloRows * 128;        // 36 KB per x-shifted copy and channel slab
constexpr int kHaloStagesA = 3;
constexpr int kHaloStagesB = 6;
constexpr int kHaloSmemBytes = kHaloStagesA * kHaloABytes + kHaloStagesB * kBBytes + 1024 + 256 + kEpiBytes;

template <bool HALO, bool IN16, bool OUT16>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tf32_wide_kernel(const __grid_constant__ ConvGemmParams p) {
  constexpr int NT = 2;
  constexpr int kStages = Cfg<2>::kStages;          // 4
  constexpr int kStageBytes = Cfg<2>::kStageBytes;  // 48 KB: [weights 16 KB | pixels 0 | pixels 1]
  constexpr int kTmemCols = 512;                    // 2 accumulator stages x 256 pixel columns
  constexpr int kPOff = kBBytes;                    // pixel tiles start after the weight tile
  constexpr int kRingBytes = HALO ? kHaloStagesA * kHaloABytes + kHaloStagesB * kBBytes
                                  : kStages * kStageBytes;
  constexpr int kNFull = HALO ? kHaloStagesA + kHaloStagesB : kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  // HALO: full_bar[0..3) / empty_bar[0..3) = activation ring, [3..9) = weight ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kRingBytes);
  uint64_t* empty_bar = full_bar + kNFull;
  uint64_t* tfull_bar = empty_bar + kNFull;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* epi_smem = reinterpret_cast<float*>(smem + kRingBytes + 256);
  uint8_t* smem_w = smem + kHaloStagesA * kHaloABytes;      // HALO: weight ring base

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(HALO ? &p.hmap : &p.amap[0]);
    prefetch_tmap(&p.bmap);
    if (HALO && p.c2_chunks > 0) { prefetch_tmap(&p.hmap2); prefetch_tmap(&p.bmap2); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kNFull; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                            // inputs of this launch are complete and visible from here on
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_n;
  const int units_m = tiles_m / NT;
  const int total_items = units_m * p.tiles_co;
  const int kiters = p.ntaps * p.c_chunks;
  const int half_y = p.tiles_y >> 1;     // HALO: 16-row blocks per image (TW = 16, TH = 8, TN = 1)

  if (warp == 0 && HALO) {
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const int um = w % units_m;
        const int co0 = (w / units_m) * 128;
        const int x0 = (um % p.tiles_x) * 16;
        const int y0 = ((um / p.tiles_x) % half_y) * 16;
        const int n0 = um / (p.tiles_x * half_y);
        for (int cc = 0; cc < p.c_chunks; ++cc) {
#pragma unroll 1
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&empty_bar[as], aph ^ 1);
            if (p.debug == 3 || p.debug == 7 || p.debug == 9) {   // profiling aid: no activation traffic
              mbar_arrive(&full_bar[as]);
            } else {
              mbar_arrive_expect_tx(&full_bar[as], kHaloABytes);
              tma_load_4d(smem + as * kHaloABytes, &p.hmap, &full_bar[as], cc * (IN16 ? 2 * kConvBlockK : kConvBlockK),
                          x0 + dxi - 1, y0 - 1, n0);
            }
            if (++as == kHaloStagesA) { as = 0; aph ^= 1; }
#pragma unroll 1
            for (int dyi = 0; dyi < 3; ++dyi) {
              uint64_t* fb = &full_bar[kHaloStagesA + bs];
              mbar_wait(&empty_bar[kHaloStagesA + bs], bph ^ 1);
              if (p.debug == 3 || p.debug == 8 || p.debug == 9) {   // profiling aid: no weight traffic
                mbar_arrive(fb);
              } else {
                mbar_arrive_expect_tx(fb, kBBytes);
                tma_load_2d(smem_w + bs * kBBytes, &p.bmap, fb, p.halo_wk[dxi * 3 + dyi] + cc * (IN16 ? 2 * kConvBlockK : kConvBlockK), co0);
              }
              if (++bs == kHaloStagesB) { bs = 0; bph ^= 1; }
            }
          }
        }
        // fused 1x1 shortcut: one 16 x 16 pixel box of the second input + one weight tile per slab
        for (int cc = 0; cc < p.c2_chunks; ++cc) {
          mbar_wait(&empty_bar[as], aph ^ 1);
          mbar_arrive_expect_tx(&full_bar[as], 2 * kABytes);
          tma_load_4d(smem + as * kHaloABytes, &p.hmap2, &full_bar[as], cc * (IN16 ? 2 * kConvBlockK : kConvBlockK), x0, y0, n0);
          if (++as == kHaloStagesA) { as = 0; aph ^= 1; }
          uint64_t* fb = &full_bar[kHaloStagesA + bs];
          mbar_wait(&empty_bar[kHaloStagesA + bs], bph ^ 1);
          mbar_arrive_expect_tx(fb, kBBytes);
          tma_load_2d(smem_w + bs * kBBytes, &p.bmap2, fb, cc * (IN16 ? 2 * kConvBlockK : kConvBlockK), co0);
          if (++bs == kHaloStagesB) { bs = 0; bph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && HALO) {
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc(128, 256, IN16);   // M = 128 (output channels), N = 256 (pixels)
      const uint64_t pdesc0 = make_smem_desc(smem_u32(smem));      // halo copies (N operand)
      const uint64_t wdesc0 = make_smem_desc(smem_u32(smem_w));    // weights     (M operand)
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        uint32_t first = 0;
        for (int cc = 0; cc < p.c_chunks; ++cc) {
#pragma unroll 1
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&full_bar[as], aph);
            const uint64_t pdesc = pdesc0 + (uint64_t)as * (uint64_t)(kHaloABytes >> 4);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              mbar_wait(&full_bar[kHaloStagesA + bs], bph);
              tc_fence_after();
              const uint64_t wdesc = wdesc0 + (uint64_t)bs * (uint64_t)(kBBytes >> 4);
              // the 256 pixels of tap row dyi are the 256 consecutive 128-byte rows that start
              // dyi image rows (16 pixels = 2 KB) into the halo copy
              const uint64_t pd = pdesc + (uint64_t)(dyi * ((16 * 128) >> 4));
              if (p.debug == 4) {                      // profiling aid: no tensor-core work
                mbar_arrive(&empty_bar[kHaloStagesA + bs]);
                if (dyi == 2) mbar_arrive(&empty_bar[as]);
                if (++bs == kHaloStagesB) { bs = 0; bph ^= 1; }
                continue;
              }
#pragma unroll
              for (int k = 0; k < kConvBlockK / 8; ++k) {
                umma_any<IN16>(d_tmem, wdesc + (uint64_t)(k * 2), pd + (uint64_t)(k * 2), idesc, first);
                first = 1;
              }
              umma_commit(&empty_bar[kHaloStagesA + bs]);
              if (++bs == kHaloStagesB) { bs = 0; bph ^= 1; }
            }
            if (p.debug != 4) umma_commit(&empty_bar[as]);
            if (++as == kHaloStagesA) { as = 0; aph ^= 1; }
          }
        }
        for (int cc = 0; cc < p.c2_chunks; ++cc) {        // fused 1x1 shortcut slabs
          mbar_wait(&full_bar[as], aph);
          mbar_wait(&full_bar[kHaloStagesA + bs], bph);
          tc_fence_after();
          const uint64_t pd = pdesc0 + (uint64_t)as * (uint64_t)(kHaloABytes >> 4);
          const uint64_t wdesc = wdesc0 + (uint64_t)bs * (uint64_t)(kBBytes >> 4);
#pragma unroll
          for (int k = 0; k < kConvBlockK / 8; ++k) {
            umma_any<IN16>(d_tmem, wdesc + (uint64_t)(k * 2), pd + (uint64_t)(k * 2), idesc, first);
            first = 1;
          }
          umma_commit(&empty_bar[kHaloStagesA + bs]);
          umma_commit(&empty_bar[as]);
          if (++bs == kHaloStagesB) { bs = 0; bph ^= 1; }
          if (++as == kHaloStagesA) { as = 0; aph ^= 1; }
        }
        if (p.debug == 4) mbar_arrive(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = NT * kABytes + kBBytes;     // block_n == 128 on this path
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const int um = w % units_m;
        const int co0 = (w / units_m) * 128;
        int x0[NT], y0[NT], n0[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int tm = um * NT + j;
          x0[j] = (tm % p.tiles_x) * p.TW;
          y0[j] = ((tm / p.tiles_x) % p.tiles_y) * p.TH;
          n0[j] = (tm / (p.tiles_x * p.tiles_y)) * p.TN;
        }
        int tap = 0, cc = 0;
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          const CUtensorMap* am = &p.amap[p.tap_map[tap]];
          tma_load_2d(sa, &p.bmap, &full_bar[stage], p.tap_wk[tap] + cc * (IN16 ? 2 * kConvBlockK : kConvBlockK), co0);
#pragma unroll
          for (int j = 0; j < NT; ++j)
            tma_load_4d(sa + kPOff + j * kABytes, am, &full_bar[stage], cc * (IN16 ? 2 * kConvBlockK : kConvBlockK),
                        x0[j] + p.tap_dx[tap], y0[j] + p.tap_dy[tap], n0[j]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          if (++cc == p.c_chunks) { cc = 0; ++tap; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc(128, 256, IN16);   // M = 128 (output channels), N = 256 (pixels)
      const uint64_t wdesc0 = make_smem_desc(smem_u32(smem));            // weights  (M operand)
      const uint64_t pdesc0 = make_smem_desc(smem_u32(smem) + kPOff);    // pixels   (N operand)
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        mbar_wait(&full_bar[stage], phase);
        for (int it = 0; it < kiters; ++it) {
          tc_fence_after();
          const uint64_t so = (uint64_t)stage * (uint64_t)(kStageBytes >> 4);
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == kStages) { nstage = 0; nphase ^= 1; }
#pragma unroll
          for (int k = 0; k < kConvBlockK / 8; ++k) {
            if (k == kConvBlockK / 8 - 1 && it + 1 < kiters) mbar_wait(&full_bar[nstage], nphase);
            umma_any<IN16>(d_tmem, wdesc0 + so + (uint64_t)(k * 2), pdesc0 + so + (uint64_t)(k * 2), idesc,
                      (uint32_t)((it | k) != 0));
          }
          umma_commit(&empty_bar[stage]);
          stage = nstage; phase = nphase;
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // epilogue: this warp owns output channels [32 q, 32 q + 32) of the tile (TMEM lanes) and walks
    // the 256 pixel columns in chunks of 32
    const int q = warp & 3;
    float* tbuf = epi_smem + q * (32 * 33);
    const int pr = lane >> 3;            // pixel sub-row (0..3)
    const int cq = lane & 7;             // channel quad inside this warp's 32 channels
    const int lTW = p.log_tw, lTH = p.log_th;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      const int um = w % units_m;
      const int co = (w / units_m) * 128 + q * 32 + cq * 4;    // first of this lane's 4 channels
      float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) bsum = __ldg(reinterpret_cast<const float4*>(p.bias + co));
      if (p.bias2) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias2 + co));
        bsum.x += b.x; bsum.y += b.y; bsum.z += b.z; bsum.w += b.w;
      }
#pragma unroll
      for (int jt = 0; jt < NT; ++jt) {
        int tx, ty, tn;
        if (HALO) {
          tx = um % p.tiles_x;
          ty = ((um / p.tiles_x) % half_y) * 2 + jt;
          tn = um / (p.tiles_x * half_y);
        } else {
          const int tm = um * NT + jt;
          tx = tm % p.tiles_x;
          ty = (tm / p.tiles_x) % p.tiles_y;
          tn = tm / (p.tiles_x * p.tiles_y);
        }
        prefetch_epilogue_tile<OUT16>(p, tx, ty, tn, (w / units_m) * 128, (int)threadIdx.x - 128);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
#pragma unroll 1
      for (int chn = ((p.debug == 5 || p.debug == 9) ? 8 : 0); chn < 8; ++chn) {   // 8 chunks of 32 pixels
        int tx, ty, tn;
        if (HALO) {       // the unit is a 16 x 16 block: two 16 x 8 tiles stacked vertically
          tx = um % p.tiles_x;
          ty = ((um / p.tiles_x) % half_y) * 2 + (chn >> 2);
          tn = um / (p.tiles_x * half_y);
        } else {
          const int tm = um * NT + (chn >> 2);
          tx = tm % p.tiles_x;
          ty = (tm / p.tiles_x) % p.tiles_y;
          tn = tm / (p.tiles_x * p.tiles_y);
        }
        uint32_t r[32];
        tmem_ld_32x32(taddr + chn * 32, r);
        tmem_ld_wait();
        // tbuf[channel = lane][pixel]
#pragma unroll
        for (int c = 0; c < 32; ++c) tbuf[lane * 33 + c] = __uint_as_float(r[c]);
        __syncwarp();
        float4 v[8];
        long long ooff[8], aoff[8];
        uint32_t vmask = 0, bmask = 0;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int pix = it * 4 + pr;                  // pixel inside the chunk
          const float* tp = tbuf + (cq * 4) * 33 + pix;
          v[it] = make_float4(tp[0], tp[33], tp[66], tp[99]);
          const int R = (chn & 3) * 32 + pix;           // pixel inside its 128-pixel tile
          const int x = tx * p.TW + (R & (p.TW - 1));
          const int y = ty * p.TH + ((R >> lTW) & (p.TH - 1));
          const int n = tn * p.TN + (R >> (lTW + lTH));
          if (n < p.N && y < p.Ho && x < p.Wo) vmask |= 1u << it;
          if (n < p.bias_rows) bmask |= 1u << it;
          ooff[it] = (long long)n * p.out_sN + (long long)y * p.out_sH + (long long)x * p.out_sW + co;
          aoff[it] = (long long)n * p.add_sN + (long long)y * p.add_sH + (long long)x * p.add_sW + co;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it)
          if ((bmask >> it) & 1u) { v[it].x += bsum.x; v[it].y += bsum.y; v[it].z += bsum.z; v[it].w += bsum.w; }
        constexpr bool h = OUT16;
        if (p.addend) {
          float4 a[8];
#pragma unroll
          for (int it = 0; it < 8; ++it)
            a[it] = ((vmask >> it) & 1u) ? ld4t<OUT16>(p.addend, aoff[it]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 8; ++it) { v[it].x += a[it].x; v[it].y += a[it].y; v[it].z += a[it].z; v[it].w += a[it].w; }
        }
        if (p.accumulate) {
          float4 a[8];
#pragma unroll
          for (int it = 0; it < 8; ++it)
            a[it] = ((vmask >> it) & 1u) ? ld4t<OUT16>(p.out, ooff[it]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 8; ++it) { v[it].x += a[it].x; v[it].y += a[it].y; v[it].z += a[it].z; v[it].w += a[it].w; }
        }
        float gs1 = 0.f, gs2 = 0.f;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if (!((vmask >> it) & 1u)) continue;
          float4 o = v[it];
          if (h) {
            o.x = round_f16(o.x); o.y = round_f16(o.y); o.z = round_f16(o.z); o.w = round_f16(o.w);
          } else if (p.round_out) {
            o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
          }
          st4t<OUT16>(p.out, ooff[it], o);
          gs1 += (o.x + o.y) + (o.z + o.w);
          gs2 += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
        }
        if (p.st_ptr[0] != nullptr || p.st_ptr[1] != nullptr) {
          gs1 += __shfl_xor_sync(0xffffffffu, gs1, 8);  gs2 += __shfl_xor_sync(0xffffffffu, gs2, 8);
          gs1 += __shfl_xor_sync(0xffffffffu, gs1, 16); gs2 += __shfl_xor_sync(0xffffffffu, gs2, 16);
          const int nrow = tn * p.TN + (((chn & 3) * 32) >> (lTW + lTH));   // uniform per chunk
          if (pr == 0 && nrow < p.N) {
#pragma unroll
            for (int tg = 0; tg < 2; ++tg) {
              if (p.st_ptr[tg] == nullptr) continue;
              const int g = (p.st_choff[tg] + co) / p.st_cg[tg];
              double* dst = p.st_ptr[tg] + ((long long)nrow * 32 + g) * 2;
              atomicAdd(dst, (double)gs1);
              atomicAdd(dst + 1, (double)gs2);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair halo variant (tcgen05 cta_group::2).  Two CTAs of a cluster (the two SMs of a TPC) work on
// neighbouring 16 x 16 pixel blocks of the same output-channel tile.  Pixels are the M operand again
// (128 per CTA and MMA, M = 256 over the pair), the 128 output channels the N operand, and each CTA
// holds only HALF of every weight tile (64 channels): the tensor cores of both SMs read both halves.
// Per SM the operand bytes per MMA clock drop from 55 to 40, below the ~43 B/clk an SM can ingest
// (profiles/README.md), which neither the halo layout alone nor a cluster multicast achieves.
//   * both CTAs run a TMA producer for their own activation windows and weight half; the bytes are
//     completed on the LEADER's full barriers (cta_group::2 TMA), whose single arrival is the
//     leader producer's expect_tx for both CTAs' bytes;
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 and releases stages / publishes
//     accumulators with multicast commits to the barriers of both CTAs;
//   * every CTA's epilogue drains its own TMEM (lane = pixel) and arrives on the leader's
//     accumulator-empty barrier (count = 8 warps).
// ------------------------------------------------------------------------------------------------
constexpr int kPairStagesA = 4;
constexpr int kPairStagesB = 7;
constexpr int kPairBBytes = 64 * 128;     // this CTA's half of a weight tile
constexpr int kPairSmemBytes = kPairStagesA * kHaloABytes + kPairStagesB * kPairBBytes + 1024 + 256 + kEpiBytes;
static_assert(kPairSmemBytes <= 232448 && kHaloSmemBytes <= 232448, "dynamic shared memory above the 227 KB limit");

// Epilogue of one 128-pixel x block_n tile whose accumulator has TMEM lane = pixel (shared by the
// pair kernel; same arithmetic and store order as the one-tile kernel's epilogue).
template <bool OUT16>
__device__ __forceinline__ void epilogue_pixel_tile(const ConvGemmParams& p, uint32_t taddr, float* tbuf,
                                                    int q, int lane, int tx, int ty, int tn, int co0) {
  const int pr = lane >> 3, cq = lane & 7;
  const int lTW = p.log_tw, lTH = p.log_th;
  long long ooff[8], aoff[8];
  uint32_t vmask = 0, bmask = 0;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int R = q * 32 + it * 4 + pr;
    const int x = tx * p.TW + (R & (p.TW - 1));
    const int y = ty * p.TH + ((R >> lTW) & (p.TH - 1));
    const int n = tn * p.TN + (R >> (lTW + lTH));
    if (n < p.N && y < p.Ho && x < p.Wo) vmask |= 1u << it;
    if (n < p.bias_rows) bmask |= 1u << it;
    ooff[it] = (long long)n * p.out_sN + (long long)y * p.out_sH + (long long)x * p.out_sW + co0 + cq * 4;
    aoff[it] = (long long)n * p.add_sN + (long long)y * p.add_sH + (long long)x * p.add_sW + co0 + cq * 4;
  }
  const int nrow = tn * p.TN + ((q * 32) >> (lTW + lTH));   // uniform per warp (TW*TH >= 32)
  for (int ch = 0; ch < p.block_n; ch += 32) {
    float4 v[8];
    uint32_t r[32];
    tmem_ld_32x32(taddr + ch, r);
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);      // overlaps the TMEM load
    if (bmask != 0u) {
      if (p.bias) bsum = __ldg(reinterpret_cast<const float4*>(p.bias + co0 + ch + cq * 4));
      if (p.bias2) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias2 + co0 + ch + cq * 4));
        bsum.x += b.x; bsum.y += b.y; bsum.z += b.z; bsum.w += b.w;
      }
    }
    tmem_ld_wait();
    epilogue_transpose(r, tbuf, lane, pr, cq, v);
    epilogue_chunk_tail<OUT16>(p, v, ooff, aoff, vmask, bmask, bsum, ch, co0, cq, pr, nrow);
  }
}

// ---- fp16 output: epilogue of one work item (two 128-pixel x 128-channel tiles) WITHOUT the
// shared-memory transpose.  TMEM gives thread t of the warp pixel row 32 q + t with 32 consecutive
// channels per load; in fp16 those are 64 contiguous bytes of the channels-last tensor, so the thread
// stores them itself as four 16-byte vectors (24 shared/global memory instructions per chunk
// become 4).  The residual tile is fetched into REGISTERS (2 tiles x 16 x 16 bytes per thread)
// before the wait for the accumulator, so its L2 / HBM latency is paid once per item, under the MMAs,
// instead of once per 32-channel chunk; the per-channel bias of the item is staged in shared memory.
// Measured on 40 x 256 x 256 x 128 -> 128 with residual + statistics (profiles/r2_conv_decomp_*.txt).
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
// Per-warpgroup GroupNorm-statistics accumulator of the direct epilogue: [target][group][sum, sum sq]
// in shared memory, flushed with one fp64 atomic per entry when the warpgroup moves on to another
// image row (and at kernel end).  All persistent CTAs work on the same image at the same time, so
// per-chunk global atomics serialise on that image's 64 addresses (+216 us on the 40 x 256^2 layer).
__device__ __forceinline__ void wg_bar(int bar) {       // named barrier 1 or 2, 128 threads
  if (bar == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
struct StatAcc {
  float* acc;      // shared: [2][64]
  int cur_n;       // image row the accumulator belongs to (-1: empty)
};
// called by the 128 threads of one epilogue warpgroup (named barrier `bar`, thread index t in it)
__device__ __forceinline__ void stat_flush(const ConvGemmParams& p, StatAcc& sa, int bar, int t) {
  wg_bar(bar);
  if (sa.cur_n >= 0) {
    const int tg = t >> 6, e = t & 63;
    const float v = sa.acc[t];
    if (p.st_ptr[tg] != nullptr && v != 0.f) atomicAdd(p.st_ptr[tg] + (long long)sa.cur_n * 64 + e, (double)v);
  }
  sa.acc[t] = 0.f;
  wg_bar(bar);
}

// ---- "direct" epilogue of ONE 128-pixel x 128-channel tile of the pair kernel, run by one warpgroup.
// TMEM gives thread t of warp q pixel row 32 q + t with 32 consecutive channels per load: 128 (fp32) or
// 64 (fp16) contiguous bytes of the channels-last tensor, which the thread stores itself with 256-bit
// stores -- no shared-memory transpose (24 shared/global memory instructions per chunk become 2-4).
//   * the residual tile is fetched into registers BEFORE the wait for the accumulator (fp16: the
//     whole 256-byte pixel row; fp32: the first chunk, then one chunk ahead), so its latency sits
//     under the MMAs; the next item's residual rows are L2-prefetched an item ahead;
//   * the per-channel bias of the item is staged in shared memory once per item;
//   * statistics go to the warpgroup's shared-memory accumulator.
template <bool OUT16>
__device__ __forceinline__ void tile_epilogue_direct(const ConvGemmParams& p, uint32_t tmem_tile_q, float* bias_smem,
                                                     StatAcc& sa, int bar, int q, int lane, int tx, int ty, int tn,
                                                     int co0, uint64_t* tfull, uint32_t tphase, bool skip,
                                                     long long next_aoff) {
  // geometry of the pair kernel: TW = 16, TH = 8, TN = 1, block_n = 128
  constexpr int ESZ = OUT16 ? 2 : 4;
  constexpr int CPV = 32 / ESZ;                  // channels per 256-bit vector: 16 (fp16) / 8 (fp32)
  constexpr int VPC = 32 / CPV;                  // vectors per 32-channel chunk: 2 / 4
  const int t = q * 32 + lane;                   // thread index inside the warpgroup = pixel row of the tile
  const int x = tx * 16 + (t & 15);
  const int y = ty * 8 + (t >> 4);
  const int n = tn;
  const bool want_stats = p.st_ptr[0] != nullptr || p.st_ptr[1] != nullptr;
  const bool valid = n < p.N && y < p.Ho && x < p.Wo;
  const long long ooff = (long long)n * p.out_sN + (long long)y * p.out_sH + (long long)x * p.out_sW + co0;
  const long long aoff = (long long)n * p.add_sN + (long long)y * p.add_sH + (long long)x * p.add_sW + co0;
  char* op = reinterpret_cast<char*>(p.out) + ooff * ESZ;
  const char* ap = reinterpret_cast<const char*>(p.addend) + aoff * ESZ;
  if (want_stats && n != sa.cur_n) {       // uniform over the warpgroup
    stat_flush(p, sa, bar, t);
    sa.cur_n = n;
  }
  const bool has_bias = (p.bias != nullptr || p.bias2 != nullptr) && n < p.bias_rows;
  if (has_bias) {      // the warpgroup's 128 threads stage the item's 128 bias values (uniform branch)
    float b = p.bias ? __ldg(p.bias + co0 + t) : 0.f;
    if (p.bias2) b += __ldg(p.bias2 + co0 + t);
    bias_smem[t] = b;
  }
  constexpr int NA = OUT16 ? 8 : 4;            // residual vectors kept in registers
  U32x8 A[NA];
  if (p.addend != nullptr) {
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      if (valid) A[j] = ldg256(ap + 32 * j);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) A[j].v[i] = 0u;
      }
    }
    // residual row of the pixel this thread handles in the CTA's NEXT item: into L2 now
    if (next_aoff >= 0) {
      const char* np = reinterpret_cast<const char*>(p.addend) + next_aoff * ESZ;
#pragma unroll
      for (int j = 0; j < 128 * ESZ / 128; ++j) prefetch_l2(reinterpret_cast<const float*>(np + 128 * j));
    }
  }
  wg_bar(bar);       // bias visible to the warpgroup
  mbar_wait(tfull, tphase);
  tc_fence_after();
  if (!skip) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_tile_q + (uint32_t)(c * 32), r);
      U32x8 An[4];                       // fp32: the next chunk's residual, one chunk ahead
      if (!OUT16 && p.addend != nullptr && c < 3) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (valid) An[j] = ldg256(ap + 128 * (c + 1) + 32 * j);
          else {
#pragma unroll
            for (int i = 0; i < 8; ++i) An[j].v[i] = 0u;
          }
        }
      }
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      if (has_bias) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = *reinterpret_cast<const float4*>(bias_smem + c * 32 + i * 4);
          v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
      }
      if (p.addend != nullptr) {
        if (OUT16) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 f = unpack_h2(A[c * 2 + j].v[i]);
              v[16 * j + 2 * i] += f.x; v[16 * j + 2 * i + 1] += f.y;
            }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * j + i] += __uint_as_float(A[j].v[i]);
        }
      }
      if (p.accumulate) {       // VJP fan-in: out += result
#pragma unroll
        for (int j = 0; j < VPC; ++j) {
          U32x8 a;
          if (valid) a = ld256(op + (c * 32 + CPV * j) * ESZ);
          else {
#pragma unroll
            for (int i = 0; i < 8; ++i) a.v[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (OUT16) {
              const float2 f = unpack_h2(a.v[i]);
              v[16 * j + 2 * i] += f.x; v[16 * j + 2 * i + 1] += f.y;
            } else {
              v[8 * j + i] += __uint_as_float(a.v[i]);
            }
          }
        }
      }
      if (OUT16) {
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = pack_h2(v[2 * i], v[2 * i + 1]);
        if (valid) {
          st256(op + c * 64, h);
          st256(op + c * 64 + 32, h + 8);
        }
      } else {
        if (p.round_out) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = round_tf32(v[i]);
        }
        if (valid) {
          uint32_t* w = reinterpret_cast<uint32_t*>(v);
#pragma unroll
          for (int j = 0; j < 4; ++j) st256(op + c * 128 + 32 * j, w + 8 * j);
        }
      }
      if (want_stats) {
        // fused GroupNorm statistics: per channel quad (sum, sum of squares) of this pixel, summed over the
        // warp's 32 pixels by a halving butterfly, then added to the warpgroup's shared accumulator.  fp16
        // storage: taken from the fp32 values before the final rounding (the difference averages out over
        // the >= 4096 elements of a group; the tf32 path rounds first, as before).
        float sv[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a0 = v[4 * i], a1 = v[4 * i + 1], a2 = v[4 * i + 2], a3 = v[4 * i + 3];
          sv[2 * i] = valid ? (a0 + a1) + (a2 + a3) : 0.f;
          sv[2 * i + 1] = valid ? (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3) : 0.f;
        }
        const float tot = butterfly16(sv, lane);
        if ((lane & 1) == 0) {
          const int idx = lane >> 1;          // value index = quad * 2 + (0: sum, 1: sum of squares)
#pragma unroll
          for (int tg = 0; tg < 2; ++tg) {
            if (p.st_ptr[tg] == nullptr) continue;
            const int g = (p.st_choff[tg] + co0 + c * 32 + (idx >> 1) * 4) / p.st_cg[tg];
            atomicAdd(&sa.acc[tg * 64 + g * 2 + (idx & 1)], tot);
          }
        }
      }
      if (!OUT16 && c < 3) {
#pragma unroll
        for (int j = 0; j < 4; ++j) A[j] = An[j];
      }
    }
  }
  wg_bar(bar);       // bias_smem may be overwritten by the next item
}

constexpr int kPairThreads = 384;      // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-7 and 8-11: epilogue
template <bool IN16, bool OUT16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_gemm_tf32_pair_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* smem_w = smem + kPairStagesA * kHaloABytes;
  constexpr int kRing = kPairStagesA * kHaloABytes + kPairStagesB * kPairBBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + kRing);
  uint64_t* b_full = a_full + kPairStagesA;
  uint64_t* a_empty = b_full + kPairStagesB;
  uint64_t* b_empty = a_empty + kPairStagesA;
  uint64_t* tfull_bar = b_empty + kPairStagesB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* epi_smem = reinterpret_cast<float*>(smem + kRing + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.hmap);
    prefetch_tmap(&p.bmap);
    if (p.c2_chunks > 0) { prefetch_tmap(&p.hmap2); prefetch_tmap(&p.bmap2); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kPairStagesA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kPairStagesB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);     // the eight epilogue warps of each CTA of the pair
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // the peer's barriers and TMEM exist from here on
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  // register re-allocation between the warpgroups (launch: 65536 / 384 = 168 per thread): the
  // producer / MMA warpgroup needs few, the two epilogue warpgroups keep a residual tile in registers

  const int half_y = p.tiles_y >> 1;
  const int units_m = p.tiles_x * half_y * p.tiles_n;        // 16 x 16 blocks per output-channel tile
  const int total_items = units_m * p.tiles_co;              // even; units_m even (host checks)

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const int um = w % units_m;
        const int co0 = (w / units_m) * 128;
        const int x0 = (um % p.tiles_x) * 16;
        const int y0 = ((um / p.tiles_x) % half_y) * 16;
        const int n0 = um / (p.tiles_x * half_y);
        for (int cc = 0; cc < p.c_chunks; ++cc) {
#pragma unroll 1
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&a_empty[as], aph ^ 1);
            if (p.debug == 3 || p.debug == 9) {            // profiling aid: no operand traffic
              if (leader) mbar_arrive(&a_full[as]);
            } else {
              if (leader) mbar_arrive_expect_tx(&a_full[as], 2 * kHaloABytes);
              tma_load_4d_pair(smem + as * kHaloABytes, &p.hmap, map_to_cta(&a_full[as], 0), cc * (IN16 ? 2 * kConvBlockK : kConvBlockK),
                               x0 + dxi - 1, y0 - 1, n0);
            }
            if (++as == kPairStagesA) { as = 0; aph ^= 1; }
#pragma unroll 1
            for (int dyi = 0; dyi < 3; ++dyi) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (p.debug == 3 || p.debug == 9) {
                if (leader) mbar_arrive(&b_full[bs]);
              } else {
                if (leader) mbar_arrive_expect_tx(&b_full[bs], 2 * kPairBBytes);
                tma_load_2d_pair(smem_w + bs * kPairBBytes, &p.bmap, map_to_cta(&b_full[bs], 0),
                                 p.halo_wk[dxi * 3 + dyi] + cc * (IN16 ? 2 * kConvBlockK : kConvBlockK), co0 + (int)rank * 64);
              }
              if (++bs == kPairStagesB) { bs = 0; bph ^= 1; }
            }
          }
        }
        for (int cc = 0; cc < p.c2_chunks; ++cc) {          // fused 1x1 shortcut slabs
          mbar_wait(&a_empty[as], aph ^ 1);
          if (leader) mbar_arrive_expect_tx(&a_full[as], 2 * 2 * kABytes);
          tma_load_4d_pair(smem + as * kHaloABytes, &p.hmap2, map_to_cta(&a_full[as], 0), cc * (IN16 ? 2 * kConvBlockK : kConvBlockK),
                           x0, y0, n0);
          if (++as == kPairStagesA) { as = 0; aph ^= 1; }
          mbar_wait(&b_empty[bs], bph ^ 1);
          if (leader) mbar_arrive_expect_tx(&b_full[bs], 2 * kPairBBytes);
          tma_load_2d_pair(smem_w + bs * kPairBBytes, &p.bmap2, map_to_cta(&b_full[bs], 0), cc * (IN16 ? 2 * kConvBlockK : kConvBlockK),
                           co0 + (int)rank * 64);
          if (++bs == kPairStagesB) { bs = 0; bph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && leader) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc(256, 128, IN16);   // M = 256 (pixels of both CTAs), N = 128 (output channels)
      const uint64_t adesc0 = make_smem_desc(smem_u32(smem));
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_w));
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        uint32_t first = 0;
        for (int cc = 0; cc < p.c_chunks; ++cc) {
#pragma unroll 1
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&a_full[as], aph);
            const uint64_t ad_s = adesc0 + (uint64_t)as * (uint64_t)(kHaloABytes >> 4);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint64_t bd = bdesc0 + (uint64_t)bs * (uint64_t)(kPairBBytes >> 4);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                // pixel tile j of this CTA's block = the 128 window rows that start (dyi + 8 j) image rows in
                const uint64_t ad = ad_s + (uint64_t)((dyi * 16 * 128 + j * kABytes) >> 4);
#pragma unroll
                for (int k = 0; k < kConvBlockK / 8; ++k)
                  umma_any_pair<IN16>(d_tmem + j * 128, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc,
                                 first | (uint32_t)(k != 0));
              }
              first = 1;
              umma_commit_pair(&b_empty[bs], 3);
              if (++bs == kPairStagesB) { bs = 0; bph ^= 1; }
            }
            umma_commit_pair(&a_empty[as], 3);
            if (++as == kPairStagesA) { as = 0; aph ^= 1; }
          }
        }
        for (int cc = 0; cc < p.c2_chunks; ++cc) {
          mbar_wait(&a_full[as], aph);
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          const uint64_t ad_s = adesc0 + (uint64_t)as * (uint64_t)(kHaloABytes >> 4);
          const uint64_t bd = bdesc0 + (uint64_t)bs * (uint64_t)(kPairBBytes >> 4);
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < kConvBlockK / 8; ++k)
              umma_any_pair<IN16>(d_tmem + j * 128, ad_s + (uint64_t)((j * kABytes) >> 4) + (uint64_t)(k * 2),
                             bd + (uint64_t)(k * 2), idesc, first | (uint32_t)(k != 0));
          first = 1;
          umma_commit_pair(&b_empty[bs], 3);
          umma_commit_pair(&a_empty[as], 3);
          if (++bs == kPairStagesB) { bs = 0; bph ^= 1; }
          if (++as == kPairStagesA) { as = 0; aph ^= 1; }
        }
        umma_commit_pair(&tfull_bar[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------ epilogue (both CTAs, own TMEM) ------------------------------
    // two warpgroups: warps 4-7 drain tile 0 (rows 0-7 of the 16 x 16 block), warps 8-11 tile 1
    const int q = warp & 3;
    const int jt = (warp - 4) >> 2;
    const int bar = 1 + jt;                              // named barrier of this warpgroup
    float* wg_smem = epi_smem + jt * 256;                // [0,128): bias, [128,256): statistics
    const uint32_t tempty_leader0 = map_to_cta(&tempty_bar[0], 0);
    const uint32_t tempty_leader1 = map_to_cta(&tempty_bar[1], 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    StatAcc sacc;
    sacc.acc = wg_smem + 128;
    sacc.cur_n = -1;
    sacc.acc[q * 32 + lane] = 0.f;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      const int um = w % units_m;
      const int co0 = (w / units_m) * 128;
      const int tx = um % p.tiles_x;
      const int ty0 = ((um / p.tiles_x) % half_y) * 2;
      const int tn = um / (p.tiles_x * half_y);
      long long next_aoff = -1;
      if (p.addend != nullptr && w + (int)gridDim.x < total_items) {
        const int w2 = w + (int)gridDim.x;
        const int um2 = w2 % units_m;
        const int t = q * 32 + lane;
        const int x2 = (um2 % p.tiles_x) * 16 + (t & 15);
        const int y2 = (((um2 / p.tiles_x) % half_y) * 2 + jt) * 8 + (t >> 4);
        const int n2 = um2 / (p.tiles_x * half_y);
        if (n2 < p.N && y2 < p.Ho && x2 < p.Wo)
          next_aoff = (long long)n2 * p.add_sN + (long long)y2 * p.add_sH + (long long)x2 * p.add_sW + (w2 / units_m) * 128;
      }
      tile_epilogue_direct<OUT16>(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 + jt) * 128u, wg_smem,
                                  sacc, bar, q, lane, tx, ty0 + jt, tn, co0, &tfull_bar[acc], acc_phase,
                                  p.debug == 5 || p.debug == 9, next_aoff);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.st_ptr[0] != nullptr || p.st_ptr[1] != nullptr) stat_flush(p, sacc, bar, q * 32 + lane);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // neither CTA frees TMEM / exits while the pair's MMAs or arrivals are in flight
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
      set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// `half`: fp16 elements (64 channels per 128-byte swizzle row), else fp32 (32 channels)
int encode_act_map(CUtensorMap* m, const float* base, int C, int W, int H, int N, long long sW,
                   long long sH, long long sN, int TW, int TH, int TN, int half) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 3;
  const cuuint64_t esz = half ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * esz, (cuuint64_t)sH * esz, (cuuint64_t)sN * esz};
  cuuint32_t box[4] = {(cuuint32_t)(128 / esz), (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  LOCO_REQUIRE(((uintptr_t)base & 15) == 0, "activation base not 16B aligned");
  LOCO_REQUIRE(strides[0] % 16 == 0 && strides[1] % 16 == 0 && strides[2] % 16 == 0,
               "activation strides must be multiples of 16 bytes");
  CUresult r = fn(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                  const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LOCO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(act) failed: %d (C=%d W=%d H=%d N=%d half=%d)",
               (int)r, C, W, H, N, half);
  return 0;
}

int encode_w_map(CUtensorMap* m, const float* base, int Ktot, int Cout, int block_n, int half) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 3;
  const cuuint64_t esz = half ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  LOCO_REQUIRE(((uintptr_t)base & 15) == 0, "weight base not 16B aligned");
  CUresult r = fn(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LOCO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: %d (K=%d Cout=%d half=%d)", (int)r,
               Ktot, Cout, half);
  return 0;
}

void pick_box(int W, int H, int* TW, int* TH, int* TN) {
  int tw = W < 16 ? W : 16;
  int th = kConvBlockM / tw;
  if (th > H) th = H;
  *TW = tw; *TH = th; *TN = kConvBlockM / (tw * th);
}

int fill_common(ConvGemmParams& p, const ConvProblem& prob, int N, int Ho, int Wo, float* out,
                long long osN, long long osH, long long osW, const float* add, long long asN,
                long long asH, long long asW) {
  p.in16 = prob.in.half; p.out16 = prob.out.half;
  p.kblock = p.in16 ? 2 * kConvBlockK : kConvBlockK;
  LOCO_REQUIRE(prob.Kc % p.kblock == 0, "conv: input channels %d not a multiple of %d", prob.Kc, p.kblock);
  LOCO_REQUIRE(prob.Ngemm % 32 == 0, "conv: output channels %d not a multiple of 32", prob.Ngemm);
  LOCO_REQUIRE(prob.addend == nullptr || prob.addend->half == prob.out.half,
               "conv: addend and output must have the same element type");
  LOCO_REQUIRE(prob.in2 == nullptr || prob.in2->half == prob.in.half,
               "conv: both inputs of a fused shortcut must have the same element type");
  p.c_chunks = prob.Kc / p.kblock;
  p.block_n = prob.Ngemm % 128 == 0 ? 128 : (prob.Ngemm % 64 == 0 ? 64 : 32);
  pick_box(Wo, Ho, &p.TW, &p.TH, &p.TN);
  LOCO_REQUIRE((p.TW & (p.TW - 1)) == 0 && (p.TH & (p.TH - 1)) == 0, "conv: tile dims must be powers of two");
  p.log_tw = 0; while ((1 << p.log_tw) < p.TW) ++p.log_tw;
  p.log_th = 0; while ((1 << p.log_th) < p.TH) ++p.log_th;
  LOCO_REQUIRE(Wo % p.TW == 0 && Ho % p.TH == 0, "conv: %dx%d not tileable", Ho, Wo);
  p.tiles_x = Wo / p.TW; p.tiles_y = Ho / p.TH; p.tiles_n = (N + p.TN - 1) / p.TN;
  p.tiles_co = prob.Ngemm / p.block_n;
  p.N = N; p.Ho = Ho; p.Wo = Wo; p.Cout = prob.Ngemm;
  p.out = out; p.out_sN = osN; p.out_sH = osH; p.out_sW = osW;
  p.addend = add; p.add_sN = asN; p.add_sH = asH; p.add_sW = asW;
  p.bias = prob.bias; p.bias2 = prob.bias2; p.bias_rows = prob.bias_rows;
  p.accumulate = prob.accumulate; p.round_out = prob.round_out;
  for (int i = 0; i < 2; ++i) {
    p.st_ptr[i] = prob.st_ptr[i]; p.st_cg[i] = prob.st_cg[i]; p.st_choff[i] = prob.st_choff[i];
    if (p.st_ptr[i]) LOCO_REQUIRE(p.TW * p.TH >= 32 && p.st_cg[i] % 4 == 0, "conv: fused statistics need >= 32-pixel image tiles");
  }
  LOCO_REQUIRE((osW % 4) == 0 && (osH % 4) == 0 && (osN % 4) == 0 && (((uintptr_t)out) & 15) == 0,
               "conv: output view not float4-aligned");
  if (add)
    LOCO_REQUIRE((asW % 4) == 0 && (asH % 4) == 0 && (asN % 4) == 0 && (((uintptr_t)add) & 15) == 0,
                 "conv: addend view not float4-aligned");
  return 0;
}

}  // namespace

static int conv_variant_cap() {
  // LOCO_CONV_NT caps the variant (profiling aid): 1 = one tile per item, 2 = two tiles sharing
  // the weight tile, 3 = "wide" (operand-swapped, N = 256 pixels), default 4 = halo where eligible
  const char* e = getenv("LOCO_CONV_NT");
  return e ? atoi(e) : 4;
}

// two-pixel-tile variants (NT = 2, wide, halo) are used from this many 128-pixel tiles per SM on
static int conv_two_tile_min() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("LOCO_CONV_MIN_TILES_PER_SM");
    v = e ? atoi(e) : 4;
    if (v < 1) v = 1;
  }
  return v;
}

bool conv_halo_eligible(int kind, int N, int H, int W, int Cout) {
  if (kind != CONV_3x3 && kind != CONV_3x3_DGRAD) return false;
  if (Cout % 128 != 0 || W % 16 != 0 || H % 16 != 0 || conv_variant_cap() < 4) return false;
  const long long tiles = (long long)N * (H / 8) * (W / 16) * (Cout / 128);
  const long long sms = num_sms();
  if (tiles >= (long long)conv_two_tile_min() * sms) return true;
  // Below that, weigh the wave quantisation of the two granularities: a halo item (256 pixels)
  // costs 1 unit, a one-tile item (128 pixels, twice the operand traffic per FLOP) 0.68 units
  // (measured on 64x64 and 256x256 layers, profiles/README.md); split-K territory stays one-tile.
  const long long items = tiles / 2;
  if (tiles * 2 <= sms || items * 2 < sms) return false;
  const double t_halo = (double)((items + sms - 1) / sms);
  const double t_one = 0.68 * (double)((tiles + sms - 1) / sms);
  return t_halo <= t_one;
}

int conv_prepare(const ConvProblem& prob, ConvLaunch* L) {
  memset(L, 0, sizeof(*L));
  const View& in = prob.in;
  const View& out = prob.out;
  const View* ad = prob.addend;
  LOCO_REQUIRE(in.C == prob.Kc && out.C == prob.Ngemm, "conv: channel mismatch (%d/%d, %d/%d)", in.C,
               prob.Kc, out.C, prob.Ngemm);
  LOCO_REQUIRE(in.N == out.N, "conv: batch mismatch");
  const int sms = num_sms();
  if (prob.kind == CONV_3x3 || prob.kind == CONV_1x1 || prob.kind == CONV_3x3_DGRAD) {
    LOCO_REQUIRE(in.H == out.H && in.W == out.W, "conv: stride-1 spatial mismatch");
    ConvGemmParams& p = L->p[0];
    LOCO_TRY(fill_common(p, prob, out.N, out.H, out.W, out.ptr, out.sN, out.sH, out.sW,
                         ad ? ad->ptr : nullptr, ad ? ad->sN : 0, ad ? ad->sH : 0, ad ? ad->sW : 0));
    LOCO_TRY(encode_act_map(&p.amap[0], in.ptr, in.C, in.W, in.H, in.N, in.sW, in.sH, in.sN, p.TW,
                            p.TH, p.TN, in.half));
    if (prob.kind == CONV_1x1) {
      p.ntaps = 1; p.tap_map[0] = 0; p.tap_dy[0] = 0; p.tap_dx[0] = 0; p.tap_wk[0] = 0;
    } else {
      p.ntaps = 9;
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          const int t = r * 3 + s;
          p.tap_map[t] = 0;
          // fprop reads x(y+r-1, x+s-1); dgrad reads dy(y-(r-1), x-(s-1))
          p.tap_dy[t] = prob.kind == CONV_3x3 ? r - 1 : 1 - r;
          p.tap_dx[t] = prob.kind == CONV_3x3 ? s - 1 : 1 - s;
          p.tap_wk[t] = t * prob.Kc;
        }
    }
    LOCO_TRY(encode_w_map(&p.bmap, prob.wpack, p.ntaps * prob.Kc, prob.Ngemm, p.block_n, prob.in.half));
    L->nlaunch = 1;
  } else if (prob.kind == CONV_3x3_S2) {
    LOCO_REQUIRE(in.H == 2 * out.H && in.W == 2 * out.W, "conv: stride-2 spatial mismatch");
    ConvGemmParams& p = L->p[0];
    LOCO_TRY(fill_common(p, prob, out.N, out.H, out.W, out.ptr, out.sN, out.sH, out.sW,
                         ad ? ad->ptr : nullptr, ad ? ad->sN : 0, ad ? ad->sH : 0, ad ? ad->sW : 0));
    // Four phase views of the input: x[:, pr::2, pc::2, :]
    for (int pr = 0; pr < 2; ++pr)
      for (int pc = 0; pc < 2; ++pc)
        LOCO_TRY(encode_act_map(&p.amap[pr * 2 + pc], elem_ptr(in, pr * in.sH + pc * in.sW), in.C,
                                in.W / 2, in.H / 2, in.N, 2 * in.sW, 2 * in.sH, in.sN, p.TW, p.TH,
                                p.TN, in.half));
    p.ntaps = 9;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        const int t = r * 3 + s;
        // out(y,x) = sum x(2y+r, 2x+s); row 2y+r = phase (r&1), phase-row y + (r>>1)
        p.tap_map[t] = (r & 1) * 2 + (s & 1);
        p.tap_dy[t] = r >> 1;
        p.tap_dx[t] = s >> 1;
        p.tap_wk[t] = t * prob.Kc;
      }
    LOCO_TRY(encode_w_map(&p.bmap, prob.wpack, 9 * prob.Kc, prob.Ngemm, p.block_n, prob.in.half));
    L->nlaunch = 1;
  } else if (prob.kind == CONV_3x3_S2_DGRAD) {
    // in = dy [N,h,w,Cout_fwd], out = dx [N,2h,2w,Cin_fwd]; one launch per output phase.
    LOCO_REQUIRE(out.H == 2 * in.H && out.W == 2 * in.W, "conv: stride-2 dgrad spatial mismatch");
    LOCO_REQUIRE(ad == nullptr, "conv: stride-2 dgrad does not take an addend");
    int li = 0;
    for (int pr = 0; pr < 2; ++pr)
      for (int pc = 0; pc < 2; ++pc) {
        ConvGemmParams& p = L->p[li];
        float* obase = elem_ptr(out, pr * out.sH + pc * out.sW);
        LOCO_TRY(fill_common(p, prob, out.N, in.H, in.W, obase, out.sN, 2 * out.sH, 2 * out.sW,
                             nullptr, 0, 0, 0));
        LOCO_TRY(encode_act_map(&p.amap[0], in.ptr, in.C, in.W, in.H, in.N, in.sW, in.sH, in.sN,
                                p.TW, p.TH, p.TN, in.half));
        int nt = 0;
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s) {
            if ((r & 1) != pr || (s & 1) != pc) continue;
            // dx(2y'+pr, 2x'+pc) += W[r,s]^T dy(y' - (r-pr)/2, x' - (s-pc)/2)
            p.tap_map[nt] = 0;
            p.tap_dy[nt] = -((r - pr) / 2);
            p.tap_dx[nt] = -((s - pc) / 2);
            p.tap_wk[nt] = (r * 3 + s) * prob.Kc;
            ++nt;
          }
        p.ntaps = nt;
        LOCO_TRY(encode_w_map(&p.bmap, prob.wpack, 9 * prob.Kc, prob.Ngemm, p.block_n, prob.in.half));
        ++li;
      }
    L->nlaunch = 4;
  } else {
    LOCO_REQUIRE(false, "conv: unknown kind %d", prob.kind);
  }
  for (int i = 0; i < L->nlaunch; ++i) {
    ConvGemmParams& p = L->p[i];
    const int tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.tiles_co;
    // split the K loop when the layer has fewer tiles than SMs (small-resolution layers)
    int ks = 1;
    if (prob.splitk_partial && prob.splitk_counters && tiles <= prob.splitk_max_tiles && tiles * 2 <= sms) {
      const int kiters = p.ntaps * p.c_chunks;
      ks = sms / tiles;
      if (ks > 16) ks = 16;
      if (ks > kiters / 4) ks = kiters / 4;
      // 1x1 convolutions have 4-16 K iterations: the partial-tile exchange costs more than the split
      // saves (fp16 sweep, profiles/README.md: 512->512 at 16^2, 6 rows: 14.4 us split in two, 10.3 us whole)
      if (p.ntaps == 1 && kiters < 32) ks = 1;
      if (ks < 1) ks = 1;
      while (ks > 1 && (long long)tiles * ks * kConvBlockM * p.block_n > prob.splitk_partial_floats) --ks;
    }
    // two pixel tiles per item (shared weight tile) when that still fills >= 2 waves of the GPU
    const bool halo_shape = (prob.kind == CONV_3x3 || prob.kind == CONV_3x3_DGRAD) && ks == 1 &&
                            conv_halo_eligible(prob.kind, p.N, p.Ho, p.Wo, p.Cout);
    p.nt = (ks == 1 && (p.tiles_x * p.tiles_y * p.tiles_n) % 2 == 0 &&
            (tiles >= conv_two_tile_min() * sms || halo_shape)) ? 2 : 1;
    {
      const int cap = conv_variant_cap();
      if (cap == 1) p.nt = 1;
      // 1x1 convolutions are HBM-bound (4-8 K iterations per tile): the one-tile kernel streams them
      // at 6.1 TB/s; two tiles per item only pay off when the weight tile is re-used a lot
      if (prob.kind == CONV_1x1 && p.nt == 2 && prob.Ngemm < 2 * prob.Kc) p.nt = 1;
      if (p.nt == 2 && p.block_n == 128 && cap >= 3) {
        // the operand-swapped "wide" kernel without the halo layout lost to the two-tile kernel once
        // the latter got the vectorised epilogue (1x1 N=10 256^2 128->256: 370 vs 233 us); it is
        // kept selectable (LOCO_CONV_NT=3) as the single-CTA base of the halo variant
        if (cap == 3) p.nt = 3;
        const bool s1 = prob.kind == CONV_3x3 || prob.kind == CONV_3x3_DGRAD;
        if (cap >= 4 && s1 && p.TW == 16 && p.TH == 8 && p.TN == 1 && p.tiles_y % 2 == 0) {
          // halo variant: (16+2)-row windows, one x-shifted copy per filter column
          const View& in = prob.in;
          EncodeTiledFn fn = get_encode_fn();
          if (!fn) return 3;
          cuuint64_t dims[4] = {(cuuint64_t)in.C, (cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.N};
          const cuuint64_t esz = in.half ? 2 : 4;
          const CUtensorMapDataType dt = in.half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
          cuuint64_t strides[3] = {(cuuint64_t)in.sW * esz, (cuuint64_t)in.sH * esz, (cuuint64_t)in.sN * esz};
          cuuint32_t box[4] = {(cuuint32_t)p.kblock, 16, 18, 1};
          cuuint32_t estr[4] = {1, 1, 1, 1};
          CUresult r = fn(&p.hmap, dt, 4, in.ptr, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          LOCO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(halo) failed: %d", (int)r);
          for (int dxi = 0; dxi < 3; ++dxi)
            for (int dyi = 0; dyi < 3; ++dyi)
              for (int t = 0; t < 9; ++t)
                if (p.tap_dx[t] == dxi - 1 && p.tap_dy[t] == dyi - 1) p.halo_wk[dxi * 3 + dyi] = p.tap_wk[t];
          p.nt = 4;
          LOCO_REQUIRE(conv_halo_eligible(prob.kind, p.N, p.Ho, p.Wo, p.Cout), "conv: halo eligibility mismatch");
          if (prob.in2 != nullptr) {
            const View& i2 = *prob.in2;
            LOCO_REQUIRE(i2.N == in.N && i2.H == in.H && i2.W == in.W && i2.C == prob.Kc2 &&
                             prob.Kc2 % p.kblock == 0 && prob.wpack2 != nullptr,
                         "conv: fused shortcut input mismatch");
            cuuint64_t d2[4] = {(cuuint64_t)i2.C, (cuuint64_t)i2.W, (cuuint64_t)i2.H, (cuuint64_t)i2.N};
            cuuint64_t s2[3] = {(cuuint64_t)i2.sW * esz, (cuuint64_t)i2.sH * esz, (cuuint64_t)i2.sN * esz};
            cuuint32_t b2[4] = {(cuuint32_t)p.kblock, 16, 16, 1};
            LOCO_REQUIRE(((uintptr_t)i2.ptr & 15) == 0 && s2[0] % 16 == 0 && s2[1] % 16 == 0 && s2[2] % 16 == 0,
                         "conv: fused shortcut input not 16B aligned");
            r = fn(&p.hmap2, dt, 4, i2.ptr, d2, s2, b2, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            LOCO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(shortcut) failed: %d", (int)r);
            LOCO_TRY(encode_w_map(&p.bmap2, prob.wpack2, prob.Kc2, prob.Ngemm, 128, prob.in.half));
            p.c2_chunks = prob.Kc2 / p.kblock;
          }
        }
      }
    }
    p.ksplit = ks;
    {
      const char* e = getenv("LOCO_CONV_DEBUG");
      p.debug = e ? atoi(e) : 0;
    }
    p.partial = prob.splitk_partial;
    p.counters = prob.splitk_counters;
    p.counter_stride = prob.splitk_max_tiles;
    const int items = tiles * ks / (p.nt >= 3 ? 2 : p.nt);
    L->grid[i] = items < sms ? items : sms;
    {
      // CTA-pair variant (nt == 5): consecutive items must share their weight tile (even number of
      // 16 x 16 blocks per output-channel tile) and the grid must be whole pairs
      const char* e = getenv("LOCO_CONV_PAIR");
      const int units = p.tiles_x * p.tiles_y * p.tiles_n / 2;
      if (p.nt == 4 && !(e && atoi(e) == 0) && units % 2 == 0 && items >= 2) {
        p.nt = 5;
        L->grid[i] &= ~1;
        LOCO_TRY(encode_w_map(&p.bmap, prob.wpack, p.ntaps * prob.Kc, prob.Ngemm, 64, prob.in.half));
        if (p.c2_chunks > 0) LOCO_TRY(encode_w_map(&p.bmap2, prob.wpack2, prob.Kc2, prob.Ngemm, 64, prob.in.half));
      }
    }
    L->flops += 2.0 * p.N * p.Ho * p.Wo * (double)p.Cout * (p.ntaps * p.c_chunks + p.c2_chunks) * p.kblock;
    LOCO_REQUIRE(prob.in2 == nullptr || p.c2_chunks > 0,
                 "conv: a fused 1x1 shortcut needs the halo variant (check conv_halo_eligible)");
  }
  return 0;
}

namespace {
typedef void (*ConvKernel)(const ConvGemmParams);
// kernel of variant nt (1, 2: pixel tiles per item; 3: wide; 4: halo; 5: CTA-pair halo) for the
// element types (in16, out16); the wide non-halo variant exists for fp32 only (profiling aid)
ConvKernel conv_kernel(int nt, int in16, int out16, int* smem) {
#define LOCO_PICK(K) (in16 ? (out16 ? (ConvKernel)K(true, true) : (ConvKernel)K(true, false)) \
                           : (out16 ? (ConvKernel)K(false, true) : (ConvKernel)K(false, false)))
#define K_PAIR(a, b) conv_gemm_tf32_pair_kernel<a, b>
#define K_HALO(a, b) conv_gemm_tf32_wide_kernel<true, a, b>
#define K_TWO(a, b) conv_gemm_tf32_kernel<2, a, b>
#define K_ONE(a, b) conv_gemm_tf32_kernel<1, a, b>
  switch (nt) {
    case 5: *smem = kPairSmemBytes; return LOCO_PICK(K_PAIR);
    case 4: *smem = kHaloSmemBytes; return LOCO_PICK(K_HALO);
    case 3: *smem = Cfg<2>::kSmemBytes; return (in16 || out16) ? nullptr : (ConvKernel)conv_gemm_tf32_wide_kernel<false, false, false>;
    case 2: *smem = Cfg<2>::kSmemBytes; return LOCO_PICK(K_TWO);
    default: *smem = Cfg<1>::kSmemBytes; return LOCO_PICK(K_ONE);
  }
#undef K_ONE
#undef K_TWO
#undef K_HALO
#undef K_PAIR
#undef LOCO_PICK
}
}  // namespace

int conv_init() {
  static bool attr_set[kMaxDevices] = {false};
  if (first_time_on_device(attr_set)) {
    for (int nt = 1; nt <= 5; ++nt)
      for (int i = 0; i < 2; ++i)
        for (int o = 0; o < 2; ++o) {
          int smem = 0;
          ConvKernel k = conv_kernel(nt, i, o, &smem);
          if (k) LOCO_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        }
  }
  return 0;
}

static bool conv_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LOCO_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int conv_run(const ConvLaunch& L, cudaStream_t stream) {
  LOCO_TRY(conv_init());
  {
    ProfScope prof(0, L.flops, stream);
    for (int i = 0; i < L.nlaunch; ++i) {
      int smem = 0;
      ConvKernel k = conv_kernel(L.p[i].nt, L.p[i].in16, L.p[i].out16, &smem);
      LOCO_REQUIRE(k != nullptr, "conv: variant %d has no fp16 instantiation", L.p[i].nt);
      // (the CTA-pair kernel carries __cluster_dims__(2,1,1): no launch attribute needed for it, even grid)
      const int threads = L.p[i].nt == 5 ? kPairThreads : kThreads;
      if (conv_pdl_enabled()) {
        // programmatic dependent launch: the prologue overlaps the tail of the producing kernel (pdl_wait)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(L.grid[i]); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        LOCO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k, L.p[i]));
      } else {
        k<<<L.grid[i], threads, smem, stream>>>(L.p[i]);
      }
    }
    count_launch(L.nlaunch);
  }
  LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace loco
