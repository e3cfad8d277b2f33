// Micro-probe (not part of the library): fp64 pipe of the SM — DFMA throughput / dependent latency, the double butterfly
// reduction, fp64<->fp32 conversions.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_probe tools/probes/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_tp(double* out, int iters, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  if (s == 12345.678) out[0] = s;
}
template <int ILP>
__global__ void ffma_tp(float* out, int iters, float a, float b) {
  float v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fmaf(v[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;
}
__global__ void dfma_lat(long long* cyc, double* out, int iters, double a, double b) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) v = fma(v, a, b);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (v == 12345.678) out[0] = v;
}
__global__ void bfly_lat(long long* cyc, double* out, int iters) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (v == 12345.678) out[0] = v;
}
__global__ void cvt_lat(long long* cyc, double* out, int iters) {
  double v = threadIdx.x + 1.5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) { float f = (float)v; f = f * 1.0001f; v = (double)f; }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (v == 12345.678) out[0] = v;
}
__global__ void lds_tp(long long* cyc, double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 12; ++k) s += sm[(threadIdx.x & 31) + 32 * k + (threadIdx.x >> 5) * 392 % 3000];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (s == 12345.678) out[0] = s;
}

int main() {
  double* out; long long* cyc; float* outf;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 64); cudaMalloc(&outf, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148; float ms; long long h;
  const int iters = 20000;
  for (int warps : {1, 2, 4, 6, 8, 12, 16, 32}) {
    dfma_tp<8><<<sms, warps * 32>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0); dfma_tp<8><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double lanes = (double)warps * 32 * 8 * iters;   // DFMA lane-ops per SM
    printf("DFMA ILP8 %2d warps/SM: %.3f ms -> %.1f DFMA lanes/clk/SM at 1.9 GHz, %.1f TFLOP/s chip\n", warps, ms, lanes / (ms * 1e-3 * 1.9e9), lanes * sms * 2 / (ms * 1e-3) / 1e12);
  }
  for (int warps : {4, 16}) {
    cudaEventRecord(e0); ffma_tp<8><<<sms, warps * 32>>>(outf, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double lanes = (double)warps * 32 * 8 * iters;
    printf("FFMA ILP8 %2d warps/SM: %.3f ms -> %.1f FFMA lanes/clk/SM at 1.9 GHz\n", warps, ms, lanes / (ms * 1e-3 * 1.9e9));
  }
  dfma_lat<<<1, 32>>>(cyc, out, 10000, 1.0000001, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.1f cycles\n", h / 10000.0);
  bfly_lat<<<1, 32>>>(cyc, out, 2000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("double butterfly (5 x (2 SHFL + DADD)): %.1f cycles per reduction\n", h / 2000.0);
  cvt_lat<<<1, 32>>>(cyc, out, 5000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("F2F d->f, FMUL, F2F f->d chain: %.1f cycles per iteration\n", h / 5000.0);
  for (int warps : {6, 12}) {
    lds_tp<<<1, warps * 32, 4096 * 8>>>(cyc, out, 2000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.64 x12 per lane, %d warps: %.1f cycles per 12 loads (+12 DADD)\n", warps, h / 2000.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
