// p2p.cuh -- the two per-maturity exchange steps of the sharded Monte Carlo chain over NVLink peer memory, fused into the kernels
// that produce / consume the values.
//
// Every rank owns a small MAILBOX in its own HBM: vals[2 slots][world][kmax] doubles + flags[2 slots][world] epochs, exported to the
// peers of the same node with CUDA IPC.  An exchange with epoch e uses slot e & 1:
//   producer (the fixed-order partial reduction kernel of rank r): writes its K local values into EVERY peer's mailbox at
//       vals[slot][r][:] with plain stores over NVLink, __threadfence_system(), then st.release.sys flags[slot][r] = e;
//   consumer (prologue of the payoff / finalize kernel of every rank): ld.acquire.sys-spins until flags[slot][s] >= e for all s, then
//       sums vals[slot][s][k] over s IN RANK ORDER (ld.cg, bypassing the non-coherent L1) => the same bits on every rank, independent
//       of any collective algorithm.
// No NCCL launch, no host synchronisation, no extra kernel: the 16 B (moments) and 24*J B (payoff sums) messages ride inside the
// compute kernels.  Double buffering is safe because a rank publishes epoch e+2 only after it consumed e+1, which needed every rank to
// have published e+1, which each does only after consuming e.  A spin that exceeds the context's spin limit poisons the result with NaN
// AND sets bit r of the mailbox's status word instead of hanging the GPU; the host reads the word after the chain's copy back
// (b200sv_p2p_status) and raises -- a timed-out exchange is an error, never a silent NaN price.
#pragma once
#include <cstdint>

namespace b200sv {

constexpr int kMaxPeers = 8;

struct P2pPublish {            // by value into the producer kernel; world == 0: disabled
  double* peer_vals[kMaxPeers];                  // &mailbox_s.vals[slot][my_rank][0] for every peer s (own mailbox included)
  unsigned long long* peer_flags[kMaxPeers];     // &mailbox_s.flags[slot][my_rank]
  int world;
  unsigned long long epoch;
};

struct P2pGather {             // by value into the consumer kernel; world == 0: disabled
  const double* vals;                            // &my_mailbox.vals[slot][0][0]
  const unsigned long long* flags;               // &my_mailbox.flags[slot][0]
  int world, kmax;
  unsigned long long epoch;
  unsigned int spin_limit;                       // x ~200 ns backoff; b200sv_p2p_set_spin_limit
  unsigned int* status;                          // device word in the own mailbox: bit r set <=> peer r never published (timeout)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr unsigned int kSpinLimit = 1u << 24;      // default: x ~200 ns backoff ~ 3 s

// value k of the exchange, summed over ranks in rank order (call from any thread; spins until every rank has published)
__device__ __forceinline__ double p2p_gather(const P2pGather& g, int k) {
  double s = 0.0;
  for (int r = 0; r < g.world; ++r) {
    unsigned int spins = 0;
    while (ld_acquire_sys(g.flags + r) < g.epoch) {
      __nanosleep(200);
      if (++spins > g.spin_limit) {     // peer never arrived: poison the value, RECORD it (the host raises after the chain's copy back)
        atomicOr(g.status, 1u << r);
        return __longlong_as_double(0x7ff8000000000000ll);
      }
    }
    s += __ldcg(g.vals + (size_t)r * g.kmax + k);
  }
  return s;
}

// called by ONE block after its threads wrote vals (each writer thread used p2p_store); publishes the epoch to every peer
__device__ __forceinline__ void p2p_store(const P2pPublish& p, int k, double v) {
  for (int s = 0; s < p.world; ++s) p.peer_vals[s][k] = v;
}
__device__ __forceinline__ void p2p_signal(const P2pPublish& p) {   // after __syncthreads(): one thread
  __threadfence_system();
  for (int s = 0; s < p.world; ++s) st_release_sys(p.peer_flags[s], p.epoch);
}

}  // namespace b200sv
