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

// ---- optional timeline of the peer-path iteration (debug builds: make EXTRA=-DHEAT_PEER_TRACE) ----------
// per iteration and kernel (0 = SpMV, 1 = update_xr, 2 = update_p): {first block in, last block through its
// peer wait, last block out} in %globaltimer nanoseconds; dumped by solve_device to $HEAT_PEER_TRACE_FILE<rank>
#ifdef HEAT_PEER_TRACE
constexpr int kTraceIters = 2048;
struct TraceBuf { unsigned long long t[kTraceIters][3][3]; };
#endif

#ifdef __CUDACC__
#ifdef HEAT_PEER_TRACE
static __device__ TraceBuf *g_trace = nullptr;             // one copy per translation unit (no -rdc): see trace_set_*
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define HEAT_TRACE_MIN(it, k, slot) do { if (g_trace && (it) < kTraceIters) atomicMin(&g_trace->t[it][k][slot], trace_now()); } while (0)
#define HEAT_TRACE_MAX(it, k, slot) do { if (g_trace && (it) < kTraceIters) atomicMax(&g_trace->t[it][k][slot], trace_now()); } while (0)
#else
#define HEAT_TRACE_MIN(it, k, slot) do { } while (0)
#define HEAT_TRACE_MAX(it, k, slot) do { } while (0)
#endif
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

// lanes 0..P-1 of ONE warp: lane q publishes this rank's partial sums under sequence number `seq` to
// rank q (payload, then a release store of the stamp) -- P NVLink round trips in parallel
__device__ __forceinline__ void peer_red_push_warp(const PeerRed &pr, unsigned long long seq, double v0, double v1, double v2) {
    const int q = threadIdx.x & 31;
    if (q < pr.P) {
        const int slot = (int)(seq % kPeerSlots);
        unsigned long long *e = pr.inbox[q] + ((size_t)slot * kPeerMaxRanks + pr.rank) * 4;
        double *d = reinterpret_cast<double *>(e);
        d[0] = v0; d[1] = v1; d[2] = v2;
        st_release_sys(e + 3, seq);
    }
}

// ONE full warp: lane q waits for rank q's contribution of `seq` in MY inbox; the P payloads are then
// added in rank order by every lane (identical bits on every rank).  Returns false on timeout.
__device__ __forceinline__ bool peer_red_wait_warp(const PeerRed &pr, unsigned long long seq, double (&out)[3], int *I) {
    const int q = threadIdx.x & 31;
    const int slot = (int)(seq % kPeerSlots);
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    int ok = 1;
    if (q < pr.P) {
        const unsigned long long *e = pr.inbox[pr.rank] + ((size_t)slot * kPeerMaxRanks + q) * 4;
        const long long t0 = clock64();
        while (ld_acquire_sys(e + 3) != seq) {
            if (clock64() - t0 > kPeerSpinBudget || ((volatile int *)I)[I_STATUS] == 3) { I[I_STATUS] = 3; ok = 0; break; }
        }
        const double *d = reinterpret_cast<const double *>(e);
        v0 = ld_relaxed_sys_f64(d); v1 = ld_relaxed_sys_f64(d + 1); v2 = ld_relaxed_sys_f64(d + 2);
    }
    ok = __all_sync(0xffffffffu, ok);
    out[0] = 0.0; out[1] = 0.0; out[2] = 0.0;
    for (int r = 0; r < pr.P; ++r) {
        out[0] += __shfl_sync(0xffffffffu, v0, r);
        out[1] += __shfl_sync(0xffffffffu, v1, r);
        out[2] += __shfl_sync(0xffffffffu, v2, r);
    }
    return ok != 0;
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
