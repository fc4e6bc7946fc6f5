// peer.cuh -- the multi-GPU exchange of a landmark-sharded iteration over NVLink PEER MEMORY, inside the
// iteration's own kernels (north_star: "one all-reduce of the reduced camera matrix per LM
// iteration over NVLink").  No host code and no NCCL call sits inside an iteration: the whole sharded
// iteration is ONE CUDA graph per rank.
//
// Every rank owns an EXCHANGE REGION (one cudaMalloc, mapped into the peers through CUDA IPC, or plain
// device pointers when the ranks are handles of one process):
//     [ pack: structurally non-zero tiles of the rank's partial S | rhs | scalars ]   (bslam_packed_buffer)
//     [ mailbox: kMaxPeers x 4 doubles, slot r written by rank r ]
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
// Every wait is bounded (kPeerTimeoutNs): a missing peer raises BSLAM_S_PEER_TIMEOUT instead of hanging the GPU.
#pragma once
#include "cholesky.cuh"
#include "common.cuh"

namespace bs {

constexpr int kMaxPeers = 8;
constexpr int kXchgMailbox = 4;                                          // doubles per mailbox slot
constexpr int kXchgTail = kMaxPeers * kXchgMailbox + 2 * kMaxPeers;      // doubles after the pack
constexpr long long kPeerTimeoutNs = 4000000000LL;

struct PeerCtx {
  int world, rank;
  double* region[kMaxPeers];    // exchange region of every rank (own included)
  size_t pack_len;              // doubles before the mailbox
  long long* ctl;               // local: [0] epoch of the pre rendezvous, [1] of the end rendezvous, [2] CTAs done
};

BS_D long long* peer_flags(double* region, size_t pack_len) {
  return reinterpret_cast<long long*>(region + pack_len + kMaxPeers * kXchgMailbox);
}
BS_D void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
BS_D long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// wait until *flag >= epoch; false on timeout
BS_D bool peer_wait(const long long* flag, long long epoch) {
  const long long t0 = gtime();
  while (ld_acquire_sys(flag) < epoch) {
    __nanosleep(64);
    if (gtime() - t0 > kPeerTimeoutNs) return false;
  }
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
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  __shared__ long long s_epoch;
  if (threadIdx.x == 0) {
    const long long done = atomicAdd(reinterpret_cast<unsigned long long*>(pc.ctl + 2), 1ULL);
    s_last = done == (long long)gridDim.x - 1;
    if (s_last) {
      pc.ctl[2] = 0;
      s_epoch = ++pc.ctl[0];
    }
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  const int r = threadIdx.x;
  if (r < pc.world) st_release_sys(peer_flags(pc.region[r], pc.pack_len) + pc.rank, s_epoch);
  if (r < pc.world && !peer_wait(peer_flags(pc.region[pc.rank], pc.pack_len) + r, s_epoch)) scalars[5 /*PEER_TIMEOUT*/] = 1.0;
}

// one warp; with_data == 0: rendezvous only (bslam_peer_barrier)
__global__ void __launch_bounds__(32) peer_scalar_exchange_kernel(double* __restrict__ scalars, const PeerCtx pc, int with_data) {
  const int r = threadIdx.x;
  long long e = 0;
  if (r == 0) e = ++pc.ctl[1];
  e = __shfl_sync(0xffffffffu, e, 0);
  if (r < pc.world) {
    if (with_data) {
      double* mb = pc.region[r] + pc.pack_len + kXchgMailbox * pc.rank;
      mb[0] = scalars[0]; mb[1] = scalars[1]; mb[2] = scalars[2];
    }
    __threadfence_system();
    st_release_sys(peer_flags(pc.region[r], pc.pack_len) + kMaxPeers + pc.rank, e);
  }
  bool ok = true;
  if (r < pc.world) ok = peer_wait(peer_flags(pc.region[pc.rank], pc.pack_len) + kMaxPeers + r, e);
  if (!ok) scalars[5 /*PEER_TIMEOUT*/] = 1.0;
  __syncwarp();
  if (with_data && r < 3) {    // lane r sums scalar r over the ranks, in rank order (identical on every rank)
    const double* mb = pc.region[pc.rank] + pc.pack_len + r;
    double s = 0.0;
    for (int q = 0; q < pc.world; ++q) s += __ldcg(mb + kXchgMailbox * q);
    scalars[r] = s;
  }
}

}  // namespace bs
