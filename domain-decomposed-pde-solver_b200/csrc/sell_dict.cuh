// sell_dict.cuh — the per-slice offset dictionary of the byte-indexed SELL column stream, as a warp-cooperative
// device routine shared by the CSR -> SELL converter (sell.cu) and the direct-to-SELL cube assembly (assemble.cu).
#pragma once
#include "common.cuh"

namespace heat {

// One step of the dictionary build for entry k of a slice: every lane brings the (col - row) offsets of its
// two rows; offsets not yet in the slice's table `tab` (shared memory, T entries, warp-uniform) are appended in
// first-seen order (lane ascending, the lane's first row first).  id[h] = table index of off[h].  Returns
// false when the table would exceed kSellDictCap entries (the slice cannot be byte-indexed).
// `hint`: table index tried first.  In first-seen order the k-th entry of a fully interior slice of a structured mesh
// IS table entry k, so hint = k turns the linear search (the bulk of the integer work of a slice) into one compare.
__device__ __forceinline__ bool sell_dict_step(int32_t *tab, int &T, const int (&off)[2], int (&id)[2], int hint = 0) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int found = -1;
        if (hint < T && tab[hint] == off[h]) found = hint;
        else
            for (int t = 0; t < T; ++t)
                if (tab[t] == off[h]) { found = t; break; }
        unsigned miss = __ballot_sync(0xffffffffu, found < 0);
        while (miss) {
            const int v = __shfl_sync(0xffffffffu, off[h], __ffs(miss) - 1);
            if (T >= kSellDictCap) return false;
            if (lane == 0) tab[T] = v;
            __syncwarp();
            if (found < 0 && off[h] == v) found = T;
            ++T;
            miss = __ballot_sync(0xffffffffu, found < 0);
        }
        id[h] = found;
    }
    return true;
}

}  // namespace heat
