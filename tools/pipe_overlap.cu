// pipe_overlap.cu -- do the fp64, FMA-heavy (IMAD.WIDE) and XU (MUFU) pipes overlap on B200 when one warp interleaves them?
// Per loop iteration: ND DFMA (4 chains) + NI IMAD.WIDE (4 chains) + NM MUFU.LG2 (2 chains), the per-step mix of the MC slice kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
template <int ND, int NI, int NM, int HL = 0>
__global__ void k_mix(double* out, double a, double b, uint32_t m, uint32_t zero = 0) {
  double d[4]; uint32_t v[4]; float f[2];
  for (int c = 0; c < 4; ++c) { d[c] = threadIdx.x * 1e-3 + c; v[c] = threadIdx.x + c; }
  f[0] = 1.5f + threadIdx.x * 1e-3f; f[1] = 2.5f + threadIdx.x * 1e-3f;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int k = 0; k < 28; ++k) {
      if (k < ND) d[k & 3] = fma(d[k & 3], a, b);
      if (k < NI) {
        uint32_t hi, lo;
        if (HL == 0) asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(v[k & 3]), "r"(m));
        else if (HL == 1) { asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(v[k & 3]), "r"(m)); asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(v[k & 3]), "r"(m)); }
        else { const uint32_t v2 = v[k & 3] + zero; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(v[k & 3]), "r"(m)); asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(v2), "r"(m)); }
        v[k & 3] = hi ^ lo;
      }
      if (k < NM) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[k & 1]));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = d[0] + d[1] + d[2] + d[3] + v[0] + v[1] + v[2] + v[3] + f[0] + f[1];
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  void* buf; cudaMalloc(&buf, 8ull * sms * 2048);
  for (int wpsm : {8, 16, 32}) {
    const int threads = 256, blocks = sms * wpsm * 32 / threads;
    auto cyc = [&](float ms) { return ms * 1e-3 * khz * 1e3 / ITERS / (wpsm / 4.0); };   // cycles per iteration per warp per SMSP-slot
    float td = timeit([&] { k_mix<26, 0, 0><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float ti = timeit([&] { k_mix<0, 10, 0><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float tm = timeit([&] { k_mix<0, 0, 4><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float tx = timeit([&] { k_mix<26, 10, 4><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float tdi = timeit([&] { k_mix<26, 10, 0><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float tdm = timeit([&] { k_mix<26, 0, 4><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float tim = timeit([&] { k_mix<0, 10, 4><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float thl = timeit([&] { k_mix<0, 10, 0, 1><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float tdhl = timeit([&] { k_mix<26, 10, 0, 1><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    float tall = timeit([&] { k_mix<26, 10, 4, 1><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u); });
    printf("warps/SM %2d mul.hi+mul.lo instead of mul.wide: 10 pairs alone %.1f | with 26 DFMA %.1f | with DFMA and MUFU %.1f\n", wpsm, cyc(thl), cyc(tdhl), cyc(tall));
    float t2 = timeit([&] { k_mix<0, 10, 0, 2><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u, 0u); });
    float td2 = timeit([&] { k_mix<26, 10, 0, 2><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, 0xD2511F53u, 0u); });
    printf("warps/SM %2d UNFUSED IMAD.HI + IMAD (+IADD): 10 pairs alone %.1f | with 26 DFMA %.1f\n", wpsm, cyc(t2), cyc(td2));
    printf("warps/SM %2d pairs: DFMA+IMAD.WIDE %.1f (sum %.1f) | DFMA+MUFU %.1f (sum %.1f) | IMAD.WIDE+MUFU %.1f (sum %.1f)\n", wpsm, cyc(tdi),
           cyc(td) + cyc(ti), cyc(tdm), cyc(td) + cyc(tm), cyc(tim), cyc(ti) + cyc(tm));
    printf("warps/SM %2d: clk per iteration per warp-slot: 26 DFMA %.1f | 10 IMAD.WIDE %.1f | 4 MUFU %.1f | all three interleaved %.1f (sum %.1f, max %.1f)\n",
           wpsm, cyc(td), cyc(ti), cyc(tm), cyc(tx), cyc(td) + cyc(ti) + cyc(tm), fmaxf(cyc(td), fmaxf(cyc(ti), cyc(tm))));
  }
  return 0;
}
