// philox.cuh -- counter-based RNG for the fused MC kernels: Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
// "Parallel random numbers: as easy as 1, 2, 3", SC'11) + Box-Muller.
//
// The reference draws from Numba's process-global MT19937 (logsv_pricer.py:1025-1026) and pins no stream
// (SURVEY.md §8c), so the stream is this repo's to define.  It is defined ONCE here and restated in
// oracle/mc.py::device_normals so the fused kernel can be checked path by path:
//
//   key     = (seed_lo, seed_hi)
//   counter = (path_lo, path_hi, call, slice)       path = GLOBAL path id => results independent of the
//                                                   GPU count / grid / block size
//   gauss f64: call = step;      u1 = 2 - d(r0,r1|1) in (0,1), u2 = d(r2,r3) - 1 in [0,1),
//                                d(hi,lo) = double with exponent 0x3FF and mantissa (hi>>12):lo
//                                Z0 = R cos(2 pi u2), Z1 = R sin(2 pi u2), R = sqrt(-2 ln u1)      (gauss64.cuh: table + polynomial fp64)
//   gauss f32: call = step / 2;  even step uses (r0,r1), odd step (r2,r3);
//                                u1 = fma(float(ra), 2^-32, 2^-33) in (0,1], angle = int32(rb) * pi * 2^-31
//                                float Box-Muller through the SFU (lg2 / sin / cos approx), widened to Real
//   gauss f64 paired (check mode): the f32 stream's words and layout, u1 = (ra + 1/2) 2^-32 and the same angle evaluated in fp64
#pragma once
#include <cstdint>

#include "gauss64.cuh"

namespace b200sv {

enum GaussMode { kGaussF32 = 0, kGaussF64 = 1, kGaussF64Paired = 2 };

__device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  // one IMAD.WIDE.U32; written in PTX so the 64-bit product is split without zero-extension adds
  asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(a), "r"(b));
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(M0, c.x, hi0, lo0);
    mulhilo32(M1, c.z, hi1, lo1);
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// ---- float Box-Muller on the SFU -------------------------------------------------------------------
__device__ __forceinline__ void box_muller_f32(uint32_t ra, uint32_t rb, float& z0, float& z1) {
  const float u1 = fmaf(__uint2float_rn(ra), 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0, 1], >= 2^-33: never denormal
  float lg, rad;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));                                       // MUFU.LG2
  // R = sqrt(-2 ln u1) = sqrt(-2 ln2 * lg2(u1));  MUFU.SQRT(0) = 0 (u1 == 1)
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * lg));
  const float ang = __int2float_rn((int32_t)rb) * 1.4629180792671596e-09f;                        // pi * 2^-31 => [-pi, pi]
  float s, c;
  __sincosf(ang, &s, &c);                                                                         // MUFU.SIN / MUFU.COS
  z0 = rad * c;
  z1 = rad * s;
}

// Per-path generator that hands out one (Z0, Z1) pair per time step in the order defined above.
template <typename Real, int MODE>
struct StepNormals;

template <typename Real>
struct StepNormals<Real, kGaussF64> {
  uint2 key;
  uint32_t plo, phi, slice;
  __device__ __forceinline__ StepNormals(uint64_t seed, uint64_t path, uint32_t slice_)
      : key(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))), plo((uint32_t)path), phi((uint32_t)(path >> 32)), slice(slice_) {}
  // one step
  __device__ __forceinline__ void get(uint32_t step, Real& z0, Real& z1) {
    double a, b;
    box_muller_f64_fast(philox4x32_10(make_uint4(plo, phi, step, slice), key), a, b);
    z0 = (Real)a;
    z1 = (Real)b;
  }
};

template <typename Real, int MODE>
struct StepNormals {            // kGaussF32 and kGaussF64Paired: two steps per Philox call
  uint2 key;
  uint32_t plo, phi, slice;
  __device__ __forceinline__ StepNormals(uint64_t seed, uint64_t path, uint32_t slice_)
      : key(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))), plo((uint32_t)path), phi((uint32_t)(path >> 32)), slice(slice_) {}
  // two consecutive steps (2*call, 2*call+1) from one Philox call
  __device__ __forceinline__ void get2(uint32_t call, Real& a0, Real& a1, Real& b0, Real& b1) {
#if defined(B200SV_ABLATE) && (B200SV_ABLATE & 1)   // tuning only: replace Philox by a trivial mix (NOT a valid generator)
    uint4 r = make_uint4(plo * 0x9E3779B9u + call, phi ^ (call * 0x85EBCA6Bu), plo ^ (call * 0xC2B2AE35u), slice + call * 0x27D4EB2Fu);
#else
    const uint4 r = philox4x32_10(make_uint4(plo, phi, call, slice), key);
#endif
    if constexpr (MODE == kGaussF64Paired) {
      double x0, x1, y0, y1;
      box_muller_f64_u32(r.x, r.y, x0, x1);
      box_muller_f64_u32(r.z, r.w, y0, y1);
      a0 = (Real)x0;
      a1 = (Real)x1;
      b0 = (Real)y0;
      b1 = (Real)y1;
      return;
    }
    float x0, x1, y0, y1;
#if defined(B200SV_ABLATE) && (B200SV_ABLATE & 2)   // tuning only: no SFU, uniforms scaled to [-1.7, 1.7] (NOT normal)
    x0 = __int2float_rn((int)r.x) * 8e-10f; x1 = __int2float_rn((int)r.y) * 8e-10f;
    y0 = __int2float_rn((int)r.z) * 8e-10f; y1 = __int2float_rn((int)r.w) * 8e-10f;
#else
    box_muller_f32(r.x, r.y, x0, x1);
    box_muller_f32(r.z, r.w, y0, y1);
#endif
    a0 = (Real)x0;
    a1 = (Real)x1;
    b0 = (Real)y0;
    b1 = (Real)y1;
  }
};

}  // namespace b200sv
