// scan.cuh -- single-pass ordered compaction support: decoupled look-back prefix over tiles.
//
// Both ordered compactions of the event (broad-phase survivors -> candidate list, non-empty clip results ->
// fragment arrays) must preserve the reference's consumption order (event, cell, piece; Surtr.cpp:2133-2146).
// Each tile publishes its aggregate, then sums its predecessors' aggregates until it meets an inclusive prefix.
// Tile ids are handed out by an atomic counter so every predecessor of a running tile has already started.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace surtr
{
template <int NQ>
struct ScanState
{
    unsigned int* flags;        // per tile: 0 = nothing, 1 = aggregate published, 2 = inclusive prefix published
    unsigned long long* agg;    // [tile][NQ]
    unsigned long long* inc;    // [tile][NQ]
};

// Called by the first warp of a block (all 32 lanes).  my_agg / excl are per-lane copies of the tile values.
template <int NQ>
__device__ void tile_lookback(const ScanState<NQ>& st, int tile, const unsigned long long (&my_agg)[NQ],
                              unsigned long long (&excl)[NQ], int lane)
{
    volatile unsigned int* flags = st.flags;
    volatile unsigned long long* agg = st.agg;
    volatile unsigned long long* inc = st.inc;
#pragma unroll
    for (int k = 0; k < NQ; k++) excl[k] = 0ull;
    if (tile == 0)
    {
        if (lane == 0)
        {
#pragma unroll
            for (int k = 0; k < NQ; k++) inc[k] = my_agg[k];
            __threadfence();
            flags[0] = 2u;
        }
        return;
    }
    if (lane == 0)
    {
#pragma unroll
        for (int k = 0; k < NQ; k++) agg[(size_t)tile * NQ + k] = my_agg[k];
        __threadfence();
        flags[tile] = 1u;
    }
    int pos = tile - 1;
    while (true)
    {
        const int idx = pos - lane;
        unsigned int f = 2u;
        if (idx >= 0)
        {
            f = flags[idx];
            while (f == 0u) f = flags[idx];
        }
        __threadfence();
        // lanes past the start of the sequence report "inclusive prefix 0"; m2 == 0 means the whole window of
        // 32 predecessors only has aggregates yet: add them all and look further back
        const unsigned m2 = __ballot_sync(0xffffffffu, f == 2u);
        const int first = m2 ? __ffs(m2) - 1 : 31;
        unsigned long long v[NQ];
#pragma unroll
        for (int k = 0; k < NQ; k++)
        {
            v[k] = 0ull;
            if (idx >= 0 && lane <= first) v[k] = (f == 2u) ? inc[(size_t)idx * NQ + k] : agg[(size_t)idx * NQ + k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            excl[k] += v[k];
        }
        if (m2) break;
        pos -= 32;
    }
    if (lane == 0)
    {
#pragma unroll
        for (int k = 0; k < NQ; k++) inc[(size_t)tile * NQ + k] = excl[k] + my_agg[k];
        __threadfence();
        flags[tile] = 2u;
    }
}
} // namespace surtr
