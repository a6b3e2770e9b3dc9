// device_utils.cuh — load/store and reduction helpers shared by the sm_100a kernels.
#pragma once
#include <utility>

#include "common.cuh"

namespace heat {

constexpr int kBlock = 256;            // threads per block of every streaming kernel
constexpr int kWarpsPerBlock = kBlock / 32;

// ---- streaming (read-once) 128/64-bit loads: bypass L1 so the x-gather keeps it ----------------
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ld_stream_s32x2(const int32_t *p) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream_s32x4(const int32_t *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int32_t ld_stream_s32(const int32_t *p) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// streaming store (written once, read by a later kernel)
__device__ __forceinline__ void st_stream_f64x2(double *p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// The kernels of the CG loop are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel's CTAs
// may become resident while its predecessor in the stream drains (after every predecessor CTA has executed
// pdl_launch_dependents or exited) and run their prologue — launch latency, barrier set-up, and for the SpMV the
// first TMA loads of the (constant) matrix stream — before pdl_wait() blocks until the predecessor grid has
// completed and its writes are visible.  Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- reductions ----------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum NV per-thread values over the block; result valid in thread 0.
template <int NV, int NW = kWarpsPerBlock>
__device__ __forceinline__ void block_sum(double (&v)[NV], double (*sm)[NW]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double s = warp_sum(v[q]);
        if (lane == 0) sm[q][warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            double s = (lane < NW) ? sm[q][lane] : 0.0;
            v[q] = warp_sum(s);
        }
    }
    __syncthreads();
}

// Deterministic grid reduction: every block stores its partial sums, the last block to arrive
// (ticket counter) adds all `total_blocks` partials in a fixed order and writes out[q].
// Returns true in thread 0 of that last block (after out[] is written) so the caller can run
// scalar epilogue logic.  Works across several launches sharing one counter (interior + boundary
// SpMV): `part_offset` places this launch's partials, `total_blocks` counts all launches.
template <int NV, int NW = kWarpsPerBlock>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double *partials, int part_offset,
                                         int total_blocks, int *counter, double *const (&out)[NV]) {
    static_assert(NW <= 32, "at most 32 warps per block");
    __shared__ double sm[NV][NW];
    __shared__ int is_last;
    block_sum<NV, NW>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) partials[q * kMaxPartials + part_offset + blockIdx.x] = v[q];
        __threadfence();
        int t = atomicAdd(counter, 1);
        is_last = (t == total_blocks - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        acc[q] = 0.0;
        for (int i = threadIdx.x; i < total_blocks; i += NW * 32) acc[q] += __ldcg(partials + q * kMaxPartials + i);
    }
    block_sum<NV, NW>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) *out[q] = acc[q];
        *counter = 0;
        __threadfence();
        return true;
    }
    return false;
}

// Same reduction, but EVERY thread of the last block gets `true` (after out[] is written and visible
// to the block), so a whole warp can take part in the epilogue (peer-memory pushes).
template <int NV, int NW = kWarpsPerBlock>
__device__ __forceinline__ bool grid_sum_block(double (&v)[NV], double *partials, int part_offset,
                                               int total_blocks, int *counter, double *const (&out)[NV]) {
    __shared__ int last_flag;
    if (threadIdx.x == 0) last_flag = 0;
    __syncthreads();
    if (grid_sum<NV, NW>(v, partials, part_offset, total_blocks, counter, out)) last_flag = 1;
    __syncthreads();
    return last_flag != 0;
}

inline int grid_for(int64_t work_items_per_thread_block, int sm_count, int blocks_per_sm) {
    int64_t cap = (int64_t)sm_count * blocks_per_sm;
    int64_t g = work_items_per_thread_block < cap ? work_items_per_thread_block : cap;
    if (g < 1) g = 1;
    if (g > kMaxPartials) g = kMaxPartials;
    return (int)g;
}

}  // namespace heat
