// pipe_microbench.cu -- measure per-SM issue throughput of the pipes the fused MC kernel leans on (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_microbench tools/pipe_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CHAINS = 8;

__global__ void k_dfma(double* out, double a, double b) {
  double v[CHAINS];
  for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = fma(v[c], a, b);
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, float a, float b) {
  float v[CHAINS];
  for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x * 1e-3f + c;
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = fmaf(v[c], a, b);
  float s = 0;
  for (int c = 0; c < CHAINS; ++c) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imadwide(uint64_t* out, uint32_t m) {
  uint32_t v[CHAINS];
  for (int c = 0; c < CHAINS; ++c) v[c] = threadIdx.x + c;
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      uint32_t hi, lo;
      asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(v[c]), "r"(m));
      v[c] = hi ^ lo;
    }
  uint64_t s = 0;
  for (int c = 0; c < CHAINS; ++c) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu(float* out) {
  float v[CHAINS];
  for (int c = 0; c < CHAINS; ++c) v[c] = 1.5f + threadIdx.x * 1e-3f + c;
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
  float s = 0;
  for (int c = 0; c < CHAINS; ++c) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f2f(double* out) {
  float v[CHAINS];
  double acc[CHAINS];
  for (int c = 0; c < CHAINS; ++c) { v[c] = 1.5f + threadIdx.x * 1e-3f + c; acc[c] = 0; }
  for (int i = 0; i < ITERS; ++i)
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(v[c])); acc[c] = d; v[c] = __int_as_float(__float_as_int(v[c]) ^ (i & 1)); }
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent-chain latency of DFMA: one chain per thread, one warp per SM sub-partition
__global__ void k_dfma_lat(double* out, double a, double b) {
  double v = threadIdx.x;
  for (int i = 0; i < ITERS * CHAINS; ++i) v = fma(v, a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  void* buf; cudaMalloc(&buf, 8ull * sms * 8 * 1024);
  const double ghz = khz / 1e6;
  printf("SMs %d, max clock %.3f GHz (rates below assume max clock)\n", sms, ghz);
  for (int wpsm : {4, 8, 16, 32}) {
    const int threads = 256, blocks = sms * wpsm * 32 / threads;
    const double n = (double)blocks * threads * ITERS * CHAINS;
    auto rep = [&](const char* name, float ms) {
      printf("%-10s warps/SM %2d: %8.3f ms  %7.2f lane-ops/clk/SM\n", name, wpsm, ms, n / (ms * 1e-3) / (ghz * 1e9) / sms);
    };
    rep("DFMA", timeit([&] { k_dfma<<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9); }));
    rep("FFMA", timeit([&] { k_ffma<<<blocks, threads>>>((float*)buf, 1.0000001f, 1e-9f); }));
    rep("IMAD.WIDE", timeit([&] { k_imadwide<<<blocks, threads>>>((uint64_t*)buf, 0xD2511F53u); }));
    rep("MUFU.LG2", timeit([&] { k_mufu<<<blocks, threads>>>((float*)buf); }));
    rep("F2F.64.32", timeit([&] { k_f2f<<<blocks, threads>>>((double*)buf); }));
  }
  float ms = timeit([&] { k_dfma_lat<<<sms, 128>>>((double*)buf, 1.0000001, 1e-9); });
  printf("DFMA dependent-chain latency: %.2f cycles\n", ms * 1e-3 * ghz * 1e9 / (ITERS * CHAINS));
  return 0;
}
