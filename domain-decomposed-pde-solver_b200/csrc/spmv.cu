// spmv.cu — fp64 SELL-C SpMV for sm_100a (HBM-bound; no tensor cores: nothing here is a dense
// contraction).  Replaces Tpetra::CrsMatrix::apply (called inside Belos on the reference path;
// explicit use at ExodusMatrixTest.cpp:101).
//
// Format: SELL-C with C = 64 rows per slice = one warp, R = 2 adjacent rows per lane.  Entry k of
// the two rows of a lane is one 16-byte (val) + one 8-byte (col) element, consecutive lanes are
// consecutive in memory, so a slice (or any run of k-entries of it) is ONE contiguous byte range of
// val and ONE of col.
//
// Two kernels over that format, bit-identical results (per-row fma() in CSR column order, the same
// order as the CPU oracle's oracle_spmv):
//   * sell_spmv_tma_kernel  — the matrix stream is staged through shared memory with TMA bulk copies
//     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), one private multi-stage ring per
//     warp, so ~100+ KB of matrix data per SM are in flight regardless of occupancy; the consumer side
//     reads val/col from shared memory and issues all x gathers of a chunk back to back (__ldg, L1).
//   * sell_spmv_kernel      — direct 128-bit ld.global.nc.L1::no_allocate loads (fallback when the
//     stage does not fit / tiny matrices, and the baseline the TMA kernel is measured against).
// Optional fused dot sum_i y_i*x_i (p.Ap of CG) is reduced deterministically (grid_sum).
//
// Algorithmic bytes per launch: 12*nnz + 16*n + 4*(n+1).
#include <cstdlib>

#include "device_utils.cuh"
#include "kernels.cuh"
#include "peer.cuh"

namespace heat {

__device__ __forceinline__ bool gate_done(const CgGate &g) {
    if (g.H == nullptr) return false;
    if (g.I[I_STATUS] != 0) return true;
    const double rr = g.H[g.it].rr, rr0 = g.H[0].rr;
    return !(rr > g.S[S_TOL2] * rr0);
}

// =================================================================================================
// direct-load kernel
// =================================================================================================
// DOT: 0 = none, 1 = sum_i y_i x_i, 2 = also sum_i y_i^2 (power method: lambda and ||A q||^2 in one pass)
template <int DOT>
__global__ void __launch_bounds__(kBlock)
sell_spmv_kernel(const int64_t *__restrict__ slice_ptr, const int32_t *__restrict__ col,
                 const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                 int64_t n_rows, const int32_t *__restrict__ slice_list, int64_t n_list, CgGate gate,
                 DotOut dot) {
    if (gate_done(gate)) return;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * kWarpsPerBlock;
    double dsum = 0.0, ysum = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); t < n_list; t += warps_total) {
        const int64_t s = slice_list ? (int64_t)slice_list[t] : t;
        const int64_t base = slice_ptr[s];
        const int w = (int)((slice_ptr[s + 1] - base) >> 6);       // entries per row in this slice
        const double *vp = val + base + 2 * lane;
        const int32_t *cp = col + base + 2 * lane;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
        for (int k = 0; k < w; ++k) {
            const double2 v = ld_stream_f64x2(vp + (int64_t)k * kSellChunk);
            const int2 c = ld_stream_s32x2(cp + (int64_t)k * kSellChunk);
            acc0 = fma(v.x, __ldg(x + c.x), acc0);
            acc1 = fma(v.y, __ldg(x + c.y), acc1);
        }
        const int64_t row = s * kSellChunk + 2 * lane;
        if (row + 1 < n_rows) {
            *reinterpret_cast<double2 *>(y + row) = make_double2(acc0, acc1);
            if (DOT) dsum += acc0 * __ldg(x + row) + acc1 * __ldg(x + row + 1);
            if (DOT == 2) ysum += acc0 * acc0 + acc1 * acc1;
        } else if (row < n_rows) {
            y[row] = acc0;
            if (DOT) dsum += acc0 * __ldg(x + row);
            if (DOT == 2) ysum += acc0 * acc0;
        }
    }
    if (DOT == 1) {
        double acc[1] = {dsum};
        double *const out[1] = {dot.out};
        grid_sum<1>(acc, dot.partials, dot.part_offset, dot.total_blocks, dot.counter, out);
    } else if (DOT == 2) {
        double acc[2] = {dsum, ysum};
        double *const out[2] = {dot.out, dot.out + 1};
        grid_sum<2>(acc, dot.partials, dot.part_offset, dot.total_blocks, dot.counter, out);
    }
}

// =================================================================================================
// TMA-staged kernel
// =================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, UBLKCP)
// The matrix stream is read exactly once: L2::evict_first keeps it from flushing the x vector (which
// every row gathers ~15 times, up to a whole k-plane apart) out of the 126 MB L2.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                            uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// CB = widest column entry of the matrix's stream in bytes: 1 (every slice table-indexed), 2 (table-indexed and int16
// slices mixed), 4 (int32 ids).  With CB <= 2 each warp also keeps NSTAGE table buffers (the issue cursor runs at
// most NSTAGE chunks, hence at most NSTAGE slices, ahead of the consume cursor).
template <int KC, int NSTAGE, int CB = 4>
struct TmaSmem {
    static constexpr int kValBytes = KC * kSellChunk * 8;
    static constexpr int kColBytes = KC * kSellChunk * CB;
    static constexpr int kStageBytes = kValBytes + kColBytes;
    static constexpr int kTabBytes = CB <= 2 ? kSellDictCap * 4 : 0;
    static constexpr int kRing = NSTAGE + 2;                       // slices the issue cursor can be ahead of the consumer, + 1
    static constexpr int kRingBytes = ((kRing * 8) + 15) & ~15;    // {slice id, width} per ring slot
    static constexpr int kWarpBytes = NSTAGE * (kStageBytes + kTabBytes) + kRingBytes;
    static constexpr size_t total(int nwarps) { return (size_t)nwarps * kWarpBytes + (size_t)nwarps * NSTAGE * 8; }
};

template <int DOT, int KC, int NSTAGE, int NWARPS, bool PEER = false, int CB = 4,
          int MINB = ((NWARPS <= 8 && KC <= 8) ? 2 : 1)>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
sell_spmv_tma_kernel(const SliceMeta *__restrict__ meta, const int32_t *__restrict__ col,
                     const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                     int64_t n_rows, int64_t n_list, CgGate gate, DotOut dot, SpmvPeer peer, SellDict dict) {
    // Everything up to pdl_wait() touches only the matrix (constant during a solve) and kernel parameters, so under
    // programmatic dependent launch it overlaps the tail of the kernel that produces x: barriers are set up and the
    // first NSTAGE chunks of every warp are already in flight when the input vector becomes valid.
    if (PEER && threadIdx.x == 0) HEAT_TRACE_MIN(gate.it, 0, 0);
    using L = TmaSmem<KC, NSTAGE, CB>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *my = smem_raw + (size_t)warp * L::kWarpBytes;
    unsigned char *my_tabs = my + NSTAGE * L::kStageBytes;           // C8: NSTAGE offset tables of this warp
    int2 *ring = reinterpret_cast<int2 *>(my + NSTAGE * (L::kStageBytes + L::kTabBytes));   // {slice id, width} issue -> consume
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NWARPS * L::kWarpBytes) + warp * NSTAGE;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const uint64_t policy = l2_evict_first_policy();
    const int64_t W = (int64_t)gridDim.x * NWARPS;
    const int64_t first = (int64_t)blockIdx.x * NWARPS + warp;
    // Slice metadata: ONE 16-byte load per slice (sell.cu packs {offset, width, id} in processing order), issued
    // TWO slices ahead of its use by the issue cursor only — under a saturated memory system a load takes about
    // as long as a warp needs for one slice, and a stalled issue cursor is a bubble in the TMA pipeline.  The
    // consume cursor gets {id, width} through a small shared-memory ring written when the slice is first issued
    // (ordered by the mbarrier phase the consumer waits for anyway).
    // gate.dir: on odd iterations the (interior) slices are swept from the end of the list, so the kernel starts on
    // the part of the input vector its predecessor wrote last (kernels.cuh: CgGate::dir); boundary slices stay last
    const bool back = spmv_backward(gate);
    const int64_t n_rev = PEER ? (peer.n_interior < n_list ? peer.n_interior : n_list) : n_list;
    struct Meta { int64_t base, cb; int s, w, mode; };     // value offset (entries), column offset (bytes), id, width, column mode
    auto load_meta = [&](int64_t t) -> Meta {
        Meta m{0, 0, 0, 0, 0};
        if (t < n_list) {
            if (back && t < n_rev) t = n_rev - 1 - t;
            const int4 q = __ldg(reinterpret_cast<const int4 *>(meta + t));
            m.base = (int64_t)(unsigned)q.x << 6;
            m.cb = (int64_t)(unsigned)q.y << 6;
            m.w = q.z & 0xffffff; m.s = q.w;
            m.mode = CB == 1 ? kColModeU8 : CB == 4 ? kColModeI32 : (q.z >> 24);
        }
        return m;
    };

    // issue cursor (runs NSTAGE chunks ahead of the consume cursor)
    int64_t ti = first;
    int ki = 0, si = 0;                                              // si: ordinal of the slice being issued
    Meta mi = load_meta(ti), mi_next = load_meta(ti + W), mi_next2 = load_meta(ti + 2 * W);
    auto issue = [&](int stage) {
        const int kc = (mi.w - ki) < KC ? (mi.w - ki) : KC;
        if (lane == 0) {
            if (ki == 0) ring[si % L::kRing] = make_int2(mi.s, mi.w | (mi.mode << 24));
            if (kc > 0) {
                unsigned char *dst = my + stage * L::kStageBytes;
                const uint32_t ebytes = mi.mode == kColModeU8 ? 1u : mi.mode == kColModeI16 ? 2u : 4u;     // column bytes per entry
                const uint32_t vb = (uint32_t)kc * kSellChunk * 8, cb = (uint32_t)kc * kSellChunk * ebytes;
                const uint32_t tb = (mi.mode == kColModeU8 && ki == 0) ? (uint32_t)dict.tpad * 4 : 0;   // first chunk brings the table
                mbar_arrive_expect_tx(bars + stage, vb + cb + tb);
                tma_load_1d(dst, val + mi.base + (int64_t)ki * kSellChunk, vb, bars + stage, policy);
                tma_load_1d(dst + L::kValBytes, dict.cstream + mi.cb + (int64_t)ki * kSellChunk * ebytes, cb, bars + stage, policy);
                if (tb) tma_load_1d(my_tabs + (si % NSTAGE) * L::kTabBytes, dict.tab + (int64_t)mi.s * dict.tpad, tb, bars + stage, policy);
            } else {
                mbar_arrive(bars + stage);                              // empty slice: complete the phase
            }
        }
        ki += KC;
        if (ki >= mi.w) {
            ti += W; ki = 0; ++si;
            mi = mi_next; mi_next = mi_next2;
            mi_next2 = load_meta(ti + 2 * W);
        }
    };
    int n_issued = 0;
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s)
        if (ti < n_list) { issue(s); ++n_issued; }

    pdl_wait();                                                      // the producer of x (and of the gate's records) is complete
    pdl_launch_dependents();                                         // all CTAs of this grid are resident: the next kernel's may move in as SMs free up
    if (gate_done(gate)) {                                           // frozen solve: drain the loads in flight, then leave
        for (int s = 0; s < n_issued; ++s) mbar_wait(bars + s, 0);
        return;
    }

    // consume cursor
    int64_t tc = first;
    int kc0 = 0, sc = 0;                                             // sc: ordinal of the slice being consumed
    int2 mc = make_int2(0, 0);                                       // {slice id, width}: read from the ring per slice
    int cmode = CB == 1 ? kColModeU8 : kColModeI32;                  // column mode of the slice being consumed
    int stage = 0;
    uint32_t parity = 0;
    double acc0 = 0.0, acc1 = 0.0, dsum = 0.0, ysum = 0.0;
    bool halo_ready = !PEER;
    while (tc < n_list) {
        mbar_wait(bars + stage, parity);
        if (kc0 == 0) {
            mc = ring[sc % L::kRing];
            if (CB == 2) cmode = mc.y >> 24;
            mc.y &= 0xffffff;
        }
        const int kc = (mc.y - kc0) < KC ? (mc.y - kc0) : KC;
        const double *vs = reinterpret_cast<const double *>(my + stage * L::kStageBytes) + 2 * lane;
        const int32_t *cs = reinterpret_cast<const int32_t *>(my + stage * L::kStageBytes + L::kValBytes) + 2 * lane;
        const unsigned char *is = my + stage * L::kStageBytes + L::kValBytes + 2 * lane;
        const int32_t *tb = reinterpret_cast<const int32_t *>(my_tabs + (sc % NSTAGE) * L::kTabBytes);
        const int row0 = mc.x * kSellChunk + 2 * lane;
        // the two column ids of this lane at entry k of the staged chunk
        auto cols_at = [&](int k) -> int2 {
            if (CB == 1 || (CB == 2 && cmode == kColModeU8)) {
                const uchar2 ix = *reinterpret_cast<const uchar2 *>(is + k * kSellChunk);
                return make_int2(row0 + tb[ix.x], row0 + 1 + tb[ix.y]);
            }
            if (CB == 2) {                                           // int16 deltas: col = row + delta
                const short2 d = *reinterpret_cast<const short2 *>(is + 2 * (k * kSellChunk + 2 * lane) - 2 * lane);
                return make_int2(row0 + d.x, row0 + 1 + d.y);
            }
            return *reinterpret_cast<const int2 *>(cs + k * kSellChunk);
        };
        if (PEER && tc >= peer.n_interior) {
            // boundary slice: its ghost columns are written by the neighbours' GPUs (peer stores).
            // Wait once per warp for their epoch flags, then gather through L2 (coherent), not L1.
            if (!halo_ready) {
                if (lane == 0) peer_halo_wait(peer.halo, peer.I);
                if (lane == 0) HEAT_TRACE_MAX(gate.it, 0, 1);
                __syncwarp();
                halo_ready = true;
            }
            // only the GHOST columns were written by other GPUs (through L2): owned entries keep the L1 path
#pragma unroll 4
            for (int k = 0; k < kc; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(vs + k * kSellChunk);
                const int2 c = cols_at(k);
                const double x0 = c.x < n_rows ? __ldg(x + c.x) : __ldcg(x + c.x);
                const double x1 = c.y < n_rows ? __ldg(x + c.y) : __ldcg(x + c.y);
                acc0 = fma(v.x, x0, acc0);
                acc1 = fma(v.y, x1, acc1);
            }
        } else if (kc == KC) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(vs + k * kSellChunk);
                const int2 c = cols_at(k);
                acc0 = fma(v.x, __ldg(x + c.x), acc0);
                acc1 = fma(v.y, __ldg(x + c.y), acc1);
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < kc; ++k) {
                const double2 v = *reinterpret_cast<const double2 *>(vs + k * kSellChunk);
                const int2 c = cols_at(k);
                acc0 = fma(v.x, __ldg(x + c.x), acc0);
                acc1 = fma(v.y, __ldg(x + c.y), acc1);
            }
        }
        __syncwarp();                                    // every lane is done with this stage
        if (ti < n_list) issue(stage);                   // refill it NSTAGE chunks ahead
        kc0 += KC;
        if (kc0 >= mc.y) {                               // slice finished: write its 64 rows
            const int64_t row = (int64_t)mc.x * kSellChunk + 2 * lane;
            if (row + 1 < n_rows) {
                *reinterpret_cast<double2 *>(y + row) = make_double2(acc0, acc1);
                if (DOT) dsum += acc0 * __ldg(x + row) + acc1 * __ldg(x + row + 1);
                if (DOT == 2) ysum += acc0 * acc0 + acc1 * acc1;
            } else if (row < n_rows) {
                y[row] = acc0;
                if (DOT) dsum += acc0 * __ldg(x + row);
                if (DOT == 2) ysum += acc0 * acc0;
            }
            acc0 = 0.0; acc1 = 0.0;
            tc += W; kc0 = 0; ++sc;
        }
        if (++stage == NSTAGE) { stage = 0; parity ^= 1u; }
    }
    if (DOT == 2) {
        double acc[2] = {dsum, ysum};
        double *const out[2] = {dot.out, dot.out + 1};
        grid_sum<2, NWARPS>(acc, dot.partials, dot.part_offset, dot.total_blocks, dot.counter, out);
    } else if (DOT) {
        double acc[1] = {dsum};
        double *const out[1] = {dot.out};
        if (PEER) {        // p.Ap of this rank -> all ranks' inboxes, one lane per destination
            if (grid_sum_block<1, NWARPS>(acc, dot.partials, dot.part_offset, dot.total_blocks, dot.counter, out) && threadIdx.x < 32)
                peer_red_push_warp(peer.red, peer.seq_out, *dot.out, 0.0, 0.0);
        } else {
            grid_sum<1, NWARPS>(acc, dot.partials, dot.part_offset, dot.total_blocks, dot.counter, out);
        }
    }
    if (PEER && threadIdx.x == 0) HEAT_TRACE_MAX(gate.it, 0, 2);
}

#ifdef HEAT_PEER_TRACE
int trace_set_spmv(TraceBuf *buf) {
    HEAT_CUDA(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
    return 0;
}
#endif

// -------------------------------------------------------------------------------------------------
// launch
// -------------------------------------------------------------------------------------------------
// variant: 0 = direct loads; 1 = TMA KC=16 x 2 stages x 8 warps; 2 = TMA KC=8 x 3 stages x 12 warps;
//          3 = TMA KC=8 x 2 stages x 16 warps; 4 = TMA KC=4 x 3 stages x 24 warps; 5 = TMA KC=8 x 3 x 8
//          (2 CTAs/SM).  HEAT_SPMV_VARIANT overrides the default.
static int spmv_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HEAT_SPMV_VARIANT");
        v = e ? atoi(e) : 5;
        if (v < 0 || v > 5) v = 5;
    }
    return v;
}

// shapes of the byte-indexed kernel (HEAT_SPMV_C8CFG, tools/bench_spmv.py), KC x stages x warps x CTAs/SM:
// 0 = default 8x2x10x2 (its smaller stages leave room for 20 warps per SM: 2.94 ms at 512^3 against 3.11 ms
// for the int32 kernel's 8x2x8x2 shape, 3.58 for 8x3x6x2, 3.46 for 4x4x8x2, 3.16 for 8x2x5x3, 3.08 for 4x3x10x2);
// 1 = 8x2x8x2; 2 = 8x2x11x2; 3 = 6x2x12x2; 4 = 6x3x10x2; 5 = 4x3x10x2
static int c8_cfg() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HEAT_SPMV_C8CFG");
        v = e ? atoi(e) : 0;
        if (v < 0 || v > 5) v = 0;
    }
    return v;
}
constexpr int kC8Warps = 10;

static int tma_warps(int variant) { return variant == 1 ? 8 : variant == 2 ? 12 : variant == 3 ? 16 : variant == 4 ? 24 : 8; }

int spmv_grid(int64_t n_list, int sm_count) {
    const int v = spmv_variant();
    if (v == 0) {
        int64_t blocks = (n_list + kWarpsPerBlock - 1) / kWarpsPerBlock;
        return grid_for(blocks, sm_count, 8);
    }
    const int nw = tma_warps(v);
    int64_t blocks = (n_list + nw - 1) / nw;
    return grid_for(blocks, sm_count, v == 5 ? 2 : 1);    // persistent: one (or two) CTA per SM
}

// column stream the kernel reads: the compact one (table indices / int16 deltas) or the int32 ids themselves
static SellDict dict_of(const heat_matrix *A) {
    if (A->sell_cmode == 4) return SellDict{reinterpret_cast<const uint8_t *>(A->sell_col.p), nullptr, 0};
    return SellDict{A->sell_idx8.p, A->sell_tab.p, A->sell_tpad};
}

template <int DOT, int KC, int NSTAGE, int NWARPS, int CB = 4, int MINB = ((NWARPS <= 8 && KC <= 8) ? 2 : 1)>
static int launch_tma(const heat_matrix *A, const double *x, double *y, int64_t first, int64_t n_list,
                      CgGate gate, DotOut dot, int grid, cudaStream_t st) {
    auto kern = sell_spmv_tma_kernel<DOT, KC, NSTAGE, NWARPS, false, CB, MINB>;
    const size_t smem = TmaSmem<KC, NSTAGE, CB>::total(NWARPS);
    static bool configured[kMaxDevices] = {};            // function attributes are per DEVICE, not per process
    const int dev = A->ctx->device;
    if (dev < 0 || dev >= kMaxDevices || !configured[dev]) {
        HEAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < kMaxDevices) configured[dev] = true;
    }
    HEAT_CUDA(launch_kernel(kern, grid, NWARPS * 32, smem, st, gate.pdl != 0, (const SliceMeta *)(A->slice_meta.p + first),
                            (const int32_t *)A->sell_col.p, (const double *)A->sell_val.p, x, y, A->n_owned, n_list, gate, dot,
                            SpmvPeer(), dict_of(A)));
    HEAT_LAUNCHED();
    return 0;
}

// peer-memory mode: ONE launch over [interior slices | boundary slices]; needs the default TMA config
bool spmv_peer_supported() { return spmv_variant() == 5; }
bool spmv_compact_supported() { return spmv_variant() == 5; }

template <int CB, int DOT>
static int launch_spmv_peer_t(const heat_matrix *A, const double *x, double *y, CgGate gate, DotOut dot, SpmvPeer peer,
                              int grid, cudaStream_t st) {
    constexpr int NW = CB <= 2 ? kC8Warps : 8;
    auto kern = sell_spmv_tma_kernel<DOT, 8, 2, NW, true, CB, 2>;
    const size_t smem = TmaSmem<8, 2, CB>::total(NW);
    static bool configured[kMaxDevices] = {};
    const int dev = A->ctx->device;
    if (dev < 0 || dev >= kMaxDevices || !configured[dev]) {
        HEAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < kMaxDevices) configured[dev] = true;
    }
    const int64_t n_list = A->n_slices;                   // slice_meta order: interior slices, then boundary slices
    if (A->n_ghost == 0) peer.n_interior = n_list;
    HEAT_CUDA(launch_kernel(kern, grid, NW * 32, smem, st, gate.pdl != 0, (const SliceMeta *)A->slice_meta.p,
                            (const int32_t *)A->sell_col.p, (const double *)A->sell_val.p, x, y, A->n_owned, n_list, gate, dot, peer,
                            dict_of(A)));
    HEAT_LAUNCHED();
    return 0;
}
template <int DOT>
static int launch_spmv_peer_d(const heat_matrix *A, const double *x, double *y, CgGate gate, DotOut dot, SpmvPeer peer,
                              int grid, cudaStream_t st) {
    switch (A->sell_cmode) {
        case 1: return launch_spmv_peer_t<1, DOT>(A, x, y, gate, dot, peer, grid, st);
        case 2: return launch_spmv_peer_t<2, DOT>(A, x, y, gate, dot, peer, grid, st);
        default: return launch_spmv_peer_t<4, DOT>(A, x, y, gate, dot, peer, grid, st);
    }
}
int launch_spmv_peer(const heat_matrix *A, const double *x, double *y, CgGate gate, DotOut dot, SpmvPeer peer,
                     int grid, cudaStream_t st) {
    return dot.out ? launch_spmv_peer_d<1>(A, x, y, gate, dot, peer, grid, st)       // p.Ap of CG goes to every rank's inbox
                   : launch_spmv_peer_d<0>(A, x, y, gate, dot, peer, grid, st);      // polynomial steps: no dot
}

int launch_spmv(const heat_matrix *A, const double *x, double *y, int64_t first,
                int64_t n_list, CgGate gate, DotOut dot, int grid, cudaStream_t st) {
    if (n_list <= 0 && dot.out == nullptr) return 0;
    const int32_t *slice_list = A->n_ghost > 0 ? A->slices_all.p + first : nullptr;      // direct-load kernel only
    const int v = spmv_variant();
    const bool d = dot.out != nullptr;
    if (A->sell_cmode == 2 && v == 5) {  // table-indexed and int16-delta slices mixed (sell.cu)
        if (d && dot.with_yy) return launch_tma<2, 8, 2, kC8Warps, 2, 2>(A, x, y, first, n_list, gate, dot, grid, st);
        return d ? launch_tma<1, 8, 2, kC8Warps, 2, 2>(A, x, y, first, n_list, gate, dot, grid, st)
                 : launch_tma<0, 8, 2, kC8Warps, 2, 2>(A, x, y, first, n_list, gate, dot, grid, st);
    }
    if (A->sell_cmode == 1 && v == 5) {  // byte-indexed column stream (sell.cu: every slice has a small offset table)
        if (d && dot.with_yy) return launch_tma<2, 8, 2, kC8Warps, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st);
        switch (c8_cfg()) {
            case 1: return d ? launch_tma<1, 8, 2, 8, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st)
                             : launch_tma<0, 8, 2, 8, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st);
            case 2: return d ? launch_tma<1, 8, 2, 11, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st)
                             : launch_tma<0, 8, 2, 11, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st);
            case 3: return d ? launch_tma<1, 6, 2, 12, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st)
                             : launch_tma<0, 6, 2, 12, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st);
            case 4: return d ? launch_tma<1, 6, 3, 10, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st)
                             : launch_tma<0, 6, 3, 10, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st);
            case 5: return d ? launch_tma<1, 4, 3, 10, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st)
                             : launch_tma<0, 4, 3, 10, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st);
            default: break;
        }
        return d ? launch_tma<1, 8, 2, kC8Warps, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st)
                 : launch_tma<0, 8, 2, kC8Warps, 1, 2>(A, x, y, first, n_list, gate, dot, grid, st);
    }
    if (d && dot.with_yy) {      // x.y and y.y in one pass (power method): default TMA config or direct loads
        if (v != 0) return launch_tma<2, 8, 2, 8>(A, x, y, first, n_list, gate, dot, grid, st);
        sell_spmv_kernel<2><<<grid, kBlock, 0, st>>>(A->slice_ptr.p, A->sell_col.p, A->sell_val.p, x, y,
                                                    A->n_owned, slice_list, n_list, gate, dot);
        HEAT_LAUNCHED();
        return 0;
    }
    switch (v) {
        case 1: return d ? launch_tma<1, 16, 2, 8>(A, x, y, first, n_list, gate, dot, grid, st)
                         : launch_tma<0, 16, 2, 8>(A, x, y, first, n_list, gate, dot, grid, st);
        case 2: return d ? launch_tma<1, 8, 3, 12>(A, x, y, first, n_list, gate, dot, grid, st)
                         : launch_tma<0, 8, 3, 12>(A, x, y, first, n_list, gate, dot, grid, st);
        case 3: return d ? launch_tma<1, 8, 2, 16>(A, x, y, first, n_list, gate, dot, grid, st)
                         : launch_tma<0, 8, 2, 16>(A, x, y, first, n_list, gate, dot, grid, st);
        case 4: return d ? launch_tma<1, 4, 3, 24>(A, x, y, first, n_list, gate, dot, grid, st)
                         : launch_tma<0, 4, 3, 24>(A, x, y, first, n_list, gate, dot, grid, st);
        case 5: return d ? launch_tma<1, 8, 2, 8>(A, x, y, first, n_list, gate, dot, grid, st)
                         : launch_tma<0, 8, 2, 8>(A, x, y, first, n_list, gate, dot, grid, st);
        default: break;
    }
    if (d)
        sell_spmv_kernel<1><<<grid, kBlock, 0, st>>>(A->slice_ptr.p, A->sell_col.p, A->sell_val.p, x, y,
                                                    A->n_owned, slice_list, n_list, gate, dot);
    else
        sell_spmv_kernel<0><<<grid, kBlock, 0, st>>>(A->slice_ptr.p, A->sell_col.p, A->sell_val.p, x, y,
                                                    A->n_owned, slice_list, n_list, gate, dot);
    HEAT_LAUNCHED();
    return 0;
}

}  // namespace heat
