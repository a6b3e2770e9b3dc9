// sell_dict.cuh — the per-slice offset dictionary of the byte-indexed SELL column stream, as a warp-cooperative
// device routine shared by the CSR -> SELL converter (sell.cu) and the direct-to-SELL cube assembly (assemble.cu).
#pragma once
#include "common.cuh"

namespace heat {

// One step of the dictionary build for entry k of a slice: every lane brings the (col - row) offsets of its
// two rows; offsets not yet in the slice's table `tab` (shared memory, T <= kSellDictCap entries, T warp-uniform)
// are appended in first-seen order (lane ascending, the lane's first row first).  id[h] = table index of off[h].
// Returns false when the table would exceed kSellDictCap entries (the slice cannot be byte-indexed).
// The look-up is done BY THE WARP: the lowest unresolved lane broadcasts its offset, every lane compares two table
// entries with it (one ballot finds the index), all lanes holding that offset are resolved at once.  On a
// structured mesh the 32 lanes hold one offset, so a step is one round — the first version had every lane walk
// the table entry by entry (T dependent shared-memory loads), which made the dictionary 40 % of the cube assembly.
__device__ __forceinline__ bool sell_dict_step(int32_t *tab, int &T, const int (&off)[2], int (&id)[2]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int found = -1;
        unsigned todo = 0xffffffffu;                         // lanes whose offset has not been looked up yet (warp-uniform)
        while (todo) {
            const int v = __shfl_sync(0xffffffffu, off[h], __ffs(todo) - 1);
            const unsigned b0 = __ballot_sync(0xffffffffu, lane < T && tab[lane] == v);
            const unsigned b1 = __ballot_sync(0xffffffffu, lane + 32 < T && tab[lane + 32] == v);
            int idx;
            if (b0) idx = __ffs(b0) - 1;
            else if (b1) idx = 32 + __ffs(b1) - 1;
            else {
                if (T >= kSellDictCap) return false;
                if (lane == 0) tab[T] = v;
                __syncwarp();
                idx = T++;
            }
            const bool mine = off[h] == v;
            if (mine) found = idx;
            todo &= ~__ballot_sync(0xffffffffu, mine);
        }
        id[h] = found;
    }
    return true;
}

}  // namespace heat
