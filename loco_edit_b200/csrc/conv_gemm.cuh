// Implicit-GEMM convolution on tcgen05 (sm_100a): host-side problem description.
// One kernel serves every GEMM-shaped layer of the U-Net hot path (3x3 stride-1, 3x3 stride-2 with
// the (0,1,0,1) pad, 1x1, and the matching data-gradient convolutions of the VJP).
#pragma once
#include "common.cuh"

namespace loco {

constexpr int kConvMaxTaps = 9;
constexpr int kConvBlockM = 128;   // output pixels per tile (TMEM lanes)
constexpr int kConvBlockK = 32;    // fp32 channels per pipeline stage (128-byte swizzle row)
constexpr int kConvMaxBlockN = 128;

// Device-visible parameters (passed as one __grid_constant__ struct).
struct ConvGemmParams {
  CUtensorMap amap[4];   // activation maps: dims (C, W, H, N); [1..3] only for stride-2 phase views
  CUtensorMap bmap;      // packed weights: dims (Ktot, Cout), box (32, block_n)
  CUtensorMap hmap;      // halo variant: activation map with box (32, 16, 18, 1)
  int halo_wk[9];        // halo variant: weight column offset of tap (dx, dy) at [3 (dx+1) + (dy+1)]
  // halo variant, fused 1x1 shortcut (ResnetBlock nin_shortcut / skip_connection): the K loop is
  // extended by the channels of a second input through a second weight matrix, so that
  // out = conv3x3(in) + conv1x1(in2) is one accumulation and the shortcut tensor never exists
  CUtensorMap hmap2;     // second input, box (32, 16, 16, 1)
  CUtensorMap bmap2;     // its weights: dims (C2, Cout), box (32, 128)
  int c2_chunks;         // channels of the second input / 32 (0 = no fused shortcut)
  int ntaps;
  int tap_map[kConvMaxTaps];
  int tap_dy[kConvMaxTaps];
  int tap_dx[kConvMaxTaps];
  int tap_wk[kConvMaxTaps];   // column offset of the tap in the packed weight matrix
  int c_chunks;               // input channels / 32
  int TW, TH, TN;             // pixel box (TW*TH*TN == 128), TW and TH powers of two
  int log_tw, log_th;
  int tiles_x, tiles_y, tiles_n, tiles_co;
  int block_n;                // output channels per tile (multiple of 16, <= 128)
  int N, Ho, Wo, Cout;        // logical output grid
  float* out;
  long long out_sN, out_sH, out_sW;
  const float* addend;        // optional residual, same logical grid
  long long add_sN, add_sH, add_sW;
  const float* bias;          // [Cout] or null; applied to batch rows n < bias_rows
  const float* bias2;         // second per-channel bias (timestep projection), same rule
  int bias_rows;
  int debug;                  // profiling aid: 1 = exit at entry, 2 = setup/teardown only
  int nt;                     // kernel variant: 1 / 2 pixel tiles per item, 3 = wide, 4 = halo, 5 = CTA-pair halo (cta_group::2)
  int ksplit;                 // K-loop split factor (1 = none)
  float* partial;             // split-K partial tiles [tile][split][128][block_n]
  int* counters;              // split-K arrival [tile] and done [counter_stride + tile] counters
  int counter_stride;
  // optional fused GroupNorm statistics of the stored output (forward-only programs): per batch
  // row and group (sum, sum of squares) accumulated with fp64 atomics into up to two consumers
  double* st_ptr[2];
  int st_cg[2];               // channels per group of the consumer GroupNorm
  int st_choff[2];            // channel offset of this output inside the consumer's tensor
  int accumulate;             // out += result (VJP fan-in)
  int round_out;              // round stored values to tf32
  // element types: in16 = activations and packed weights are fp16 (tcgen05.mma kind::f16, 64
  // channels per 128-byte swizzle row), else fp32 consumed as tf32 (32 channels per row);
  // out16 = the output / addend / accumulate tensors are fp16, else fp32.  Accumulation is fp32.
  int in16, out16;
  int kblock;                 // channels per K stage: 64 (fp16) or 32 (fp32)
};

enum ConvKind {
  CONV_3x3 = 0,        // stride 1, pad 1
  CONV_1x1 = 1,
  CONV_3x3_S2 = 2,     // stride 2, pad (0,1,0,1)  (reference: ddpm/diffusion.py:834-853)
  CONV_3x3_DGRAD = 3,  // data gradient of CONV_3x3
  CONV_3x3_S2_DGRAD = 4  // data gradient of CONV_3x3_S2 (four output phases, four launches)
};

struct ConvProblem {
  int kind;
  View in;              // activations feeding the GEMM (x for fprop, dy for dgrad)
  View out;             // result (y for fprop, dx for dgrad)
  const float* wpack;   // [Ngemm][ntaps_total * Kc] K-major, tf32-rounded fp32 -- or fp16 data when in.half
  int Kc;               // channels of `in`
  int Ngemm;            // channels of `out`
  const float* bias = nullptr;
  const float* bias2 = nullptr;
  int bias_rows = 0;
  const View* addend = nullptr;
  // optional fused 1x1 shortcut (only honoured by the halo variant: check conv_halo_eligible first)
  const View* in2 = nullptr;
  const float* wpack2 = nullptr;   // [Ngemm][Kc2] K-major, tf32-rounded (fp16 data when in.half)
  int Kc2 = 0;
  int accumulate = 0;
  int round_out = 0;
  double* st_ptr[2] = {nullptr, nullptr};
  int st_cg[2] = {0, 0};
  int st_choff[2] = {0, 0};
  // optional split-K scratch shared by all launches of a stream (partials + zeroed counters)
  float* splitk_partial = nullptr;
  long long splitk_partial_floats = 0;
  int* splitk_counters = nullptr;
  int splitk_max_tiles = 0;
};

// A prepared launch: tensor maps encoded, grid sized. Valid while the buffers stay where they are.
struct ConvLaunch {
  ConvGemmParams p[4];
  int nlaunch = 0;
  int grid[4];
  double flops = 0;
};

int conv_init();   // one-time kernel attribute setup (must not happen inside a stream capture)
// true iff conv_prepare will pick the halo variant for a stride-1 3x3 (fprop or dgrad) of this
// shape: a pure function of the shape, so plan builders can decide fusions during the size query
bool conv_halo_eligible(int kind, int N, int H, int W, int Cout);
int conv_prepare(const ConvProblem& prob, ConvLaunch* out);
int conv_run(const ConvLaunch& l, cudaStream_t stream);

}  // namespace loco
