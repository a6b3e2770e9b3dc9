// peer.cuh — NVLink peer-memory communication fused into the CG kernels (one process per GPU, all
// GPUs of one NVSwitch node).  Replaces, inside the iteration, what comm.cu does with NCCL:
//
//   * halo exchange  : the kernel that PRODUCES the next SpMV input (p = z + beta p) also stores the
//                      boundary entries straight into the neighbours' ghost segments (st.global on
//                      CUDA-IPC mapped peer pointers) and then raises a per-neighbour epoch flag
//                      (st.release.sys).  The SpMV is ONE launch: interior slices first, boundary
//                      slices after an ld.acquire.sys wait on the flags.  No pack kernel, no send/recv
//                      kernels, no second SpMV launch.
//   * all-reduce     : the last block of a kernel with a fused dot product writes its rank's partial
//                      sums into every peer's inbox slot (payload + sequence stamp, release store);
//                      every block of the consumer kernel waits for the P stamps and adds the P
//                      payloads in rank order -> identical bits on every rank, no NCCL launch.
//
// Dependencies always point to a strictly earlier sequence number produced by an earlier kernel of
// the peer's stream, so there is no cyclic wait; every spin is bounded by a clock64() budget and
// raises I_STATUS = 3 instead of hanging.
#pragma once
#include "common.cuh"

namespace heat {

constexpr int kPeerMaxRanks = 8;       // one NVSwitch node
constexpr int kPeerMaxNbr = 16;        // neighbours per rank the peer path supports (else NCCL)
constexpr int kPeerSlots = 4;          // inbox ring depth (a rank can run at most one reduction ahead)
constexpr long long kPeerSpinBudget = 20000000000ll;   // ~10 s of SM clocks

// arena layout (unsigned long long words):  inbox[kPeerSlots][kPeerMaxRanks][4] | halo_flag[kPeerMaxNbr]
constexpr int kPeerInboxWords = kPeerSlots * kPeerMaxRanks * 4;
constexpr int kPeerArenaWords = kPeerInboxWords + kPeerMaxNbr + 16;

struct PeerRed {                       // passed by value to kernels; P == 0 disables the peer path
    int P = 0, rank = 0;
    unsigned long long *inbox[kPeerMaxRanks] = {};   // inbox base of every rank (own = local pointer)
};

struct PeerHalo {                      // consumer side: flags neighbours raise in MY arena
    int n_nbr = 0;
    const unsigned long long *flags = nullptr;       // [n_nbr]
    unsigned long long epoch = 0;
};

struct PeerPush {                      // producer side: where my boundary values go
    int n_nbr = 0;
    int n_blocks = 0;                  // blocks that take part in the push (ticket count)
    long long send_ptr[kPeerMaxNbr + 1] = {};
    const int32_t *send_idx = nullptr; // local rows, concatenated by neighbour
    double *dst[kPeerMaxNbr] = {};     // neighbour's ghost range for me (in the buffer being produced)
    unsigned long long *flag[kPeerMaxNbr] = {};      // neighbour's flag word for me
    unsigned long long epoch = 0;
    int *ticket = nullptr;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// one thread: publish this rank's partial sums under sequence number `seq` to every rank
__device__ __forceinline__ void peer_red_push(const PeerRed &pr, unsigned long long seq, double v0, double v1, double v2) {
    const int slot = (int)(seq % kPeerSlots);
    for (int q = 0; q < pr.P; ++q) {
        unsigned long long *e = pr.inbox[q] + ((size_t)slot * kPeerMaxRanks + pr.rank) * 4;
        double *d = reinterpret_cast<double *>(e);
        d[0] = v0; d[1] = v1; d[2] = v2;
        st_release_sys(e + 3, seq);
    }
}

// one thread: wait for all P contributions of `seq` in MY inbox and add them in rank order
__device__ __forceinline__ bool peer_red_wait(const PeerRed &pr, unsigned long long seq, double (&out)[3], int *I) {
    const int slot = (int)(seq % kPeerSlots);
    const unsigned long long *base = pr.inbox[pr.rank] + (size_t)slot * kPeerMaxRanks * 4;
    out[0] = 0.0; out[1] = 0.0; out[2] = 0.0;
    const long long t0 = clock64();
    for (int q = 0; q < pr.P; ++q) {
        const unsigned long long *e = base + (size_t)q * 4;
        while (ld_acquire_sys(e + 3) != seq) {
            if (clock64() - t0 > kPeerSpinBudget || ((volatile int *)I)[I_STATUS] == 3) { I[I_STATUS] = 3; return false; }
        }
        const double *d = reinterpret_cast<const double *>(e);
        out[0] += ld_relaxed_sys_f64(d); out[1] += ld_relaxed_sys_f64(d + 1); out[2] += ld_relaxed_sys_f64(d + 2);
    }
    return true;
}

// one thread: wait until every neighbour has delivered halo epoch >= h.epoch
__device__ __forceinline__ bool peer_halo_wait(const PeerHalo &h, int *I) {
    const long long t0 = clock64();
    for (int s = 0; s < h.n_nbr; ++s) {
        while (ld_acquire_sys(h.flags + s) < h.epoch) {
            if (clock64() - t0 > kPeerSpinBudget || ((volatile int *)I)[I_STATUS] == 3) { I[I_STATUS] = 3; return false; }
        }
    }
    return true;
}
#endif

}  // namespace heat
