// clip_warp.cuh -- what the three tiers of K3 share: the full warp mask, the clip status codes, the warp prefix sum
// and the fragment record of the moments.
//
// How a warp restates Poly::ClipPolyhedron (Src/Poly.cpp:265-500), in every tier:
//   * classification (Poly.cpp:303-319) is one signed distance per owned vertex + warp ballots;
//   * insertion of new vertices on straddling edges (Poly.cpp:332-363) is lane-parallel with the reference's
//     append order (i ascending, ring slot ascending) reproduced by a warp prefix sum, so vertex numbering,
//     ring contents and ring order come out IDENTICAL to the reference, not merely isomorphic;
//   * the topology patch (Poly.cpp:365-431) runs lane-parallel in the generic case (no in-plane vertex, every
//     face walk ends on a distinct new vertex): ring(new w) = [pusher, walked, kept] -- derived in DESIGN.md;
//     any other case (comp == 0 present, walk anomaly, non-injective walk targets) is replayed by lane 0 with the
//     reference's exact sequential loop over the same arrays, never on the CPU;
//   * degree-2 splice (Poly.cpp:433-462) belongs to that sequential replay;
//   * compaction (Poly.cpp:464-499) is a ballot/popc renumbering.
// The bounding-box shortcut (Poly.cpp:297-299) is folded into the classification: it changes the outcome only
// when every vertex is in-plane, and exactly that case evaluates the box.
// Tiers: clip_sub.cuh (<= 64 vertex slots, ring words, shared memory), clip_global.cuh (rolled loops over a workspace
// that lives in shared memory for <= 256 slots and in global memory beyond).
#pragma once

#include "surtr_math.cuh"

namespace surtr
{
constexpr unsigned FULL = 0xffffffffu;

enum ClipStatus : int { CLIP_OK = 0, CLIP_OVERFLOW = 2, CLIP_NEED_SLOTS = 3, CLIP_NEED_DEG = 4 };   // 3 / 4: only the vertex slots / the ring slots of a vertex ran out (a larger workspace would do)

__device__ __forceinline__ int warp_exscan(int v, int lane, int& total)
{
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    total = __shfl_sync(FULL, inc, 31);
    return inc - v;
}

// Fragment moments in the reference's exact accumulation order (Poly::ExtractFaces + Poly::Moments,
// Src/Poly.cpp:55-126) plus the inertia tensor about the centroid (unit density; replaces Surtr.cpp:2520).
struct Moments
{
    int n_faces;
    double volume;
    float cx, cy, cz;
    float inertia[6];
};

} // namespace surtr
