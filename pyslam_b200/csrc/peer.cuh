// peer.cuh -- the multi-GPU exchange of a landmark-sharded iteration over NVLink PEER MEMORY, inside the
// iteration's own kernels (north_star: "one all-reduce of the reduced camera matrix per LM
// iteration over NVLink").  No host code and no NCCL call sits inside an iteration: the whole sharded
// iteration is ONE CUDA graph per rank.
//
// Every rank owns an EXCHANGE REGION (one cudaMalloc, mapped into the peers through CUDA IPC, or plain
// device pointers when the ranks are handles of one process):
//     [ pack: structurally non-zero tiles of the rank's partial S | rhs | scalars ]   (bslam_packed_buffer)
//     [ mailbox: kMaxPeers x 4 records (value, epoch), slot r written by rank r ]
//     [ flags: 2 x kMaxPeers 64-bit epochs, slot r written by rank r ]
//
//   peer_pack_signal_kernel   after the rank's linearise + Schur kernels: gather the non-zero tiles into
//                             the region, then (last CTA) publish epoch e to flag `pre[rank]` of EVERY peer
//                             (st.release.sys over NVLink) and wait until all peers have published theirs.
//   chol_solve_kernel         (cholesky.cuh) is the fused all-reduce + factorisation: a tile task's first
//                             operand  S_ij = sum_r pack_r[slot(i,j)]  is read straight from the peers' regions
//                             (ld.global.cg on the mapped peer pointers, fixed rank order => bit-identical
//                             sums on every rank) while the dependency chain of the factorisation is in flight;
//                             the all-reduced matrix is never written anywhere.
//   peer_scalar_exchange_kernel  end of the iteration: the three partial scalars (cost at the linearisation
//                             point, cost at the new point, ||dx_p||^2) go to every peer's mailbox, epoch to
//                             flag `end[rank]`; wait for all peers, sum in rank order.  This second
//                             rendezvous also orders the next iteration's overwrite of a rank's pack
//                             after every peer's reads of it.
//
// Every wait is bounded (kPeerSpinLimit): a missing peer raises BSLAM_S_PEER_TIMEOUT instead of hanging the GPU.
#pragma once
#include "cholesky.cuh"
#include "common.cuh"

namespace bs {

constexpr int kMaxPeers = 8;
constexpr int kXchgMailbox = 8;                                          // doubles per mailbox slot: 4 records (value, epoch)
constexpr int kXchgTail = kMaxPeers * kXchgMailbox + 2 * kMaxPeers;      // doubles after the pack
constexpr long long kPeerSpinLimit = 40000000LL;                         // polls before a rendezvous is declared dead (~4 s)

struct PeerCtx {
  int world, rank;
  double* region[kMaxPeers];    // exchange region of every rank (own included)
  size_t pack_len;              // doubles before the mailbox
  long long* ctl;               // local: [0] epoch of the pre rendezvous, [1] of the end rendezvous, [2] CTAs done
};

// The rendezvous avoid every fence they can: a system-scope fence waits for all outstanding writes of the SM, and an
// acquire load is a load plus such a fence -- measured, a fence-per-poll rendezvous costs ~15 us, the protocol below ~3.
//   flags     : relaxed (volatile) system-scope stores / loads; ONE fence.sys before the flag of the bulk payload
//   mailbox   : NCCL-LL style records -- value and epoch travel in ONE 16-byte store, so no fence orders them
BS_D long long* peer_flags(double* region, size_t pack_len) {
  return reinterpret_cast<long long*>(region + pack_len + kMaxPeers * kXchgMailbox);
}
BS_D void st_flag(long long* p, long long v) {
  asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
BS_D long long ld_flag(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// wait until *flag >= epoch; false on timeout
BS_D bool peer_wait(const long long* flag, long long epoch) {
  for (long long spin = 0; ld_flag(flag) < epoch; ++spin)
    if (spin > kPeerSpinLimit) return false;
  return true;
}
BS_D void st_record(double* p, double value, long long epoch) {       // 16-byte aligned
  asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(value)), "l"(epoch) : "memory");
}
BS_D bool ld_record(const double* p, long long epoch, double& value) {
  long long v, e;
  for (long long spin = 0;; ++spin) {
    asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(e) : "l"(p) : "memory");
    if (e >= epoch) break;
    if (spin > kPeerSpinLimit) return false;
  }
  value = __longlong_as_double(v);
  return true;
}

// blocks [0, n_tiles): tile -> pack; block n_tiles: rhs | scalars -> pack tail; the last block to finish runs the rendezvous
__global__ void __launch_bounds__(256) peer_pack_signal_kernel(const double* __restrict__ S, int ld, int nt,
                                                               const int* __restrict__ tiles, int n_tiles,
                                                               const double* __restrict__ tail_src, int n_tail,
                                                               double* __restrict__ scalars, const PeerCtx pc) {
  double* pack = pc.region[pc.rank];
  if ((int)blockIdx.x < n_tiles) {
    const int id = tiles[blockIdx.x];
    const double* T = S + (size_t)(id / nt) * kNB * ld + (size_t)(id % nt) * kNB;
    double* P = pack + (size_t)blockIdx.x * kNB * kNB;
    for (int e = threadIdx.x; e < kNB * kNB / 2; e += 256) {
      const int r = e / (kNB / 2), c2 = (e % (kNB / 2)) << 1;
      *reinterpret_cast<double2*>(P + 2 * (size_t)e) = *reinterpret_cast<const double2*>(T + (size_t)r * ld + c2);
    }
  } else {
    double* P = pack + (size_t)n_tiles * kNB * kNB;
    for (int e = threadIdx.x; e < n_tail; e += 256) P[e] = tail_src[e];
  }
  __threadfence();                      // the CTA's tile is in L2 (the point of coherence the peers read through)
  __syncthreads();
  __shared__ int s_last;
  __shared__ long long s_epoch;
  if (threadIdx.x == 0) {
    const long long done = atomicAdd(reinterpret_cast<unsigned long long*>(pc.ctl + 2), 1ULL);
    s_last = done == (long long)gridDim.x - 1;
    if (s_last) {
      pc.ctl[2] = 0;
      s_epoch = ++pc.ctl[0];
      __threadfence_system();           // once: everything every CTA published is visible system-wide before the flags
    }
  }
  __syncthreads();
  if (!s_last) return;
  const int r = threadIdx.x;
  if (r < pc.world) st_flag(peer_flags(pc.region[r], pc.pack_len) + pc.rank, s_epoch);
  if (r < pc.world && !peer_wait(peer_flags(pc.region[pc.rank], pc.pack_len) + r, s_epoch)) scalars[5 /*PEER_TIMEOUT*/] = 1.0;
}

// one warp; with_data == 0: rendezvous only (bslam_peer_barrier).  Lane = (peer r = lane / 4, record k = lane % 4):
// records 0..2 carry the partial scalars, record 3 is the bare rendezvous.
__global__ void __launch_bounds__(32) peer_scalar_exchange_kernel(double* __restrict__ scalars, const PeerCtx pc, int with_data) {
  const int lane = threadIdx.x, r = lane >> 2, k = lane & 3;
  long long e = 0;
  if (lane == 0) e = ++pc.ctl[1];
  e = __shfl_sync(0xffffffffu, e, 0);
  const bool mine = r < pc.world && (with_data ? k < 3 : k == 3);
  if (mine) st_record(pc.region[r] + pc.pack_len + kXchgMailbox * pc.rank + 2 * k, with_data ? scalars[k] : 0.0, e);
  double v = 0.0;
  bool ok = true;
  if (mine) ok = ld_record(pc.region[pc.rank] + pc.pack_len + kXchgMailbox * r + 2 * k, e, v);
  if (!ok) scalars[5 /*PEER_TIMEOUT*/] = 1.0;
  if (with_data) {             // sum scalar k over the ranks in rank order (identical on every rank)
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) {
      const double t = __shfl_sync(0xffffffffu, v, 4 * q + k);
      if (q < pc.world) s += t;
    }
    if (lane < 3) scalars[lane] = s;
  }
}

}  // namespace bs
