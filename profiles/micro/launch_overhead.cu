// Launch-overhead microbenchmark: empty kernels, varying dynamic smem / parameter size / PDL / graph.
#include <cuda_runtime.h>
#include <stdio.h>
struct Big { char b[1024]; };
__global__ void k_small(int x) { if (x == 12345) printf("x"); }
__global__ void k_bigparam(const __grid_constant__ Big p) { if (p.b[0] == 77) printf("x"); }
__global__ void __launch_bounds__(256, 1) k_smem(int x) { extern __shared__ char s[]; if (x == 12345) s[0] = 1; }
__global__ void __launch_bounds__(256, 1) k_pdl(int x) {
  extern __shared__ char s[];
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (x == 12345) s[0] = 1;
}
template <class F> float timeit(F f, int n, cudaStream_t st) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 20; ++i) f();
  cudaEventRecord(a, st);
  for (int i = 0; i < n; ++i) f();
  cudaEventRecord(b, st); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms * 1e3f / n;
}
int main() {
  cudaStream_t st; cudaStreamCreate(&st);
  Big bp; memset(&bp, 0, sizeof(bp));
  const int SM = 214 * 1024;
  cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  cudaFuncSetAttribute(k_pdl, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  printf("small, grid 1:        %.2f us\n", timeit([&] { k_small<<<1, 256, 0, st>>>(1); }, 2000, st));
  printf("small, grid 148:      %.2f us\n", timeit([&] { k_small<<<148, 256, 0, st>>>(1); }, 2000, st));
  printf("1KB params, grid 148: %.2f us\n", timeit([&] { k_bigparam<<<148, 256, 0, st>>>(bp); }, 2000, st));
  printf("214KB smem, grid 1:   %.2f us\n", timeit([&] { k_smem<<<1, 256, SM, st>>>(1); }, 2000, st));
  printf("214KB smem, grid 148: %.2f us\n", timeit([&] { k_smem<<<148, 256, SM, st>>>(1); }, 2000, st));
  printf("64KB smem, grid 148:  %.2f us\n", timeit([&] { k_smem<<<148, 256, 64 * 1024, st>>>(1); }, 2000, st));
  printf("alternating small/214KB: %.2f us per pair\n",
         timeit([&] { k_small<<<148, 256, 0, st>>>(1); k_smem<<<148, 256, SM, st>>>(1); }, 1000, st));
  // PDL launches
  auto pdl = [&] {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = SM; cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_pdl, 1);
  };
  printf("214KB smem + PDL, grid 148: %.2f us\n", timeit(pdl, 2000, st));
  // graph of 100 launches
  for (int variant = 0; variant < 3; ++variant) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < 100; ++i) {
      if (variant == 0) k_small<<<148, 256, 0, st>>>(1);
      else if (variant == 1) k_smem<<<148, 256, SM, st>>>(1);
      else pdl();
    }
    cudaStreamEndCapture(st, &g);
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("graph variant %d: instantiate failed %s\n", variant, cudaGetErrorString(e)); continue; }
    float us = timeit([&] { cudaGraphLaunch(ge, st); }, 50, st);
    printf("graph of 100 (%s): %.2f us per kernel\n", variant == 0 ? "small" : variant == 1 ? "214KB smem" : "214KB smem + PDL", us / 100);
  }
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
