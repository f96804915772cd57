// kernels.cuh -- the four kernels of a fracture event (sm_100a) and their device-side records.
//
//   K1  kdop_extents_kernel      slab extents of every piece and every cell (warp-shuffle min/max)
//   K2  broadphase_mask_kernel   piece x cell k-DOP overlap -> one ballot word per (cell, 32 pieces)
//       compact_pairs_kernel     ordered compaction of the ballot words into the candidate pair list
//   K3  clip_sub_kernel / clip_shared_kernel / clip_global_kernel   one warp per candidate pair: half-space clipping
//   K4  assemble_kernel          ordered compaction of the non-empty results into the fragment arrays
//       kdop_arg_kernel          Kdop::KdopContainer::Calc(Polyhedron) with first-extremal-vertex semantics
#pragma once

#include "clip_sub.cuh"
#include "clip_fast.cuh"
#include "clip_duo.cuh"
#include "clip_global.cuh"

#include <type_traits>
#include "scan.cuh"
#include "../../include/surtr_b200.h"

namespace surtr
{
// Programmatic dependent launch (sm_90+): every kernel of an event lets its successor's CTAs be scheduled as soon
// as its own CTAs have all started, and waits for its predecessor's completion (and memory flush) before it
// touches any data.  This hides the launch latency between the dependent kernels of an event.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int KMAX = 13;
__constant__ float c_dirs[KMAX][3] = {
    { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 },                      // k = 3  : AABB
    { 1, 1, 1 }, { 1, -1, 1 }, { 1, 1, -1 }, { 1, -1, -1 },     // k = 7  : 14-DOP
    { 1, 1, 0 }, { 1, -1, 0 }, { 1, 0, 1 }, { 1, 0, -1 }, { 0, 1, 1 }, { 0, 1, -1 }   // k = 13 : 26-DOP
};

struct Ctl   // device-side counters of one event (zeroed before every event)
{
    unsigned long long n_cand;
    unsigned long long n_frag, n_fverts, n_fring;
    unsigned int n_ovf;         // pairs queued for the large on-chip tier (256 vertices, degree 16)
    unsigned int n_fail;        // pairs that cannot be cut: malformed rings, or beyond the workspace of the last tier
    unsigned int tile_a, tile_b;
    unsigned int n_seq_cuts;
    unsigned int n_ovf3;        // pairs queued for the global-memory tier
    unsigned int n_grow3;       // global-tier pairs whose workspace ran out of vertex slots (the host enlarges it and re-runs)
    unsigned int n_ovf2;        // pairs the 128-slot tier handed on to the large on-chip tier
    unsigned int n_growdeg;     // global-tier pairs with a ring that outgrew the workspace's ring slots (the host enlarges them and re-runs)
    unsigned int max_deg;       // largest ring such a pair needed at staging time
    unsigned int k3_ticket;     // next candidate of the small tier's persistent launch
    unsigned int max_big_verts; // largest vertex count among the fragments of the larger tiers (small-tier fragments have <= 64): picks the ring entry width of the output blob
};

struct BpTile   // one broad-phase tile: <= 256 pieces x <= 32 cells of one event
{
    uint32_t piece_begin, n_piece;
    uint32_t cell_begin, n_cell;
    uint32_t mask_base;   // index of the ballot word of (cell_begin, warp 0 of piece_begin)
    uint32_t n_w;         // ballot words per cell row in this event
};

struct CandRec   // result of K3 for one candidate pair
{
    uint32_t nv;        // 0 = empty
    uint32_t ne;        // ring entries
    uint32_t nf;
    uint32_t tier;      // 1 or 2 (ring entry width of the blob), 0 = pending in the large tier
    double volume;
    float centroid[3];
    float inertia[6];
    uint32_t pad;
    unsigned long long blob;   // byte offset of the result blob in its tier's scratch
};

// ---------------------------------------------------------------------------------------------- K1
// L lanes per object (pieces then cells, one launch for both): the objects of the BASELINE configs have 8-60 vertices,
// and the cost of an object is the butterfly over its 2K extents (log2 L steps), not its vertex stream -- L = 8 does a
// quarter of the shuffles of a full warp per object and keeps four objects' loads in flight per warp; L = 32 is used
// when a piece is large (the host picks: max piece vertices > 128).
template <int K, int L>
__global__ void kdop_extents_kernel(const float4* __restrict__ p_verts, const uint32_t* __restrict__ p_vert_off,
                                    uint32_t n_pieces, float* __restrict__ ext_p, const float4* __restrict__ c_verts,
                                    const uint32_t* __restrict__ c_vert_off, uint32_t n_cells, float* __restrict__ ext_c,
                                    int cells_unbounded)
{
    pdl_launch_dependents();
    pdl_wait();
    constexpr uint32_t PER_WARP = 32 / L;
    const int lane = threadIdx.x & 31, sl = lane % L;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_obj = n_pieces + n_cells;
    for (uint32_t obj0 = warp * PER_WARP; obj0 < n_obj; obj0 += nwarps * PER_WARP)   // (warp-uniform trip count: the shuffles use the full mask)
    {
        const uint32_t obj = obj0 + lane / L;
        const bool have = obj < n_obj;
        const bool is_cell = obj >= n_pieces;
        const uint32_t o = is_cell ? obj - n_pieces : obj;
        const float4* verts = is_cell ? c_verts : p_verts;
        const uint32_t* vert_off = is_cell ? c_vert_off : p_vert_off;
        float* ext = is_cell ? ext_c : ext_p;
        float mn[K], mx[K];
#pragma unroll
        for (int d = 0; d < K; d++) { mn[d] = 3.402823466e+38f; mx[d] = -3.402823466e+38f; }
        const bool unbounded = is_cell && cells_unbounded;
        if (have && !unbounded)
        {
            const uint32_t v0 = vert_off[o], v1 = vert_off[o + 1];
            for (uint32_t v = v0 + sl; v < v1; v += L)
            {
                const float4 p = __ldg(verts + v);   // coalesced float4 stream
#pragma unroll
                for (int d = 0; d < K; d++)
                {
                    const float t = p.x * c_dirs[d][0] + p.y * c_dirs[d][1] + p.z * c_dirs[d][2];
                    mn[d] = fminf(mn[d], t);
                    mx[d] = fmaxf(mx[d], t);
                }
            }
        }
#pragma unroll
        for (int s = L / 2; s > 0; s >>= 1)
#pragma unroll
            for (int d = 0; d < K; d++)
            {
                mn[d] = fminf(mn[d], __shfl_xor_sync(FULL, mn[d], s));
                mx[d] = fmaxf(mx[d], __shfl_xor_sync(FULL, mx[d], s));
            }
        if (unbounded)
        {
#pragma unroll
            for (int d = 0; d < K; d++) { mn[d] = -__int_as_float(0x7f800000); mx[d] = __int_as_float(0x7f800000); }
        }
        if (have)
        {
            // the L lanes of the object write its 2K extents together
#pragma unroll
            for (int d = 0; d < K; d++)
            {
                if (sl == (2 * d) % L) ext[(size_t)o * 2 * K + 2 * d] = mn[d];
                if (sl == (2 * d + 1) % L) ext[(size_t)o * 2 * K + 2 * d + 1] = mx[d];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- pattern placement
// Polygon3D::Scale(Vector3) then Translate(Vector3) over a whole cell set (VMACH.cpp:506-534), for n_place placements
// of one resident pattern: every face vertex becomes (v * s) + t (two roundings, like the two host passes) and every
// face plane is rebuilt from the face's first three moved vertices (PolygonFace::ConstructFacePlane, :303-310).
// One thread per (placement, face).  Output = the cell arrays of the event batch: placement p owns cells
// [p * n_cells, (p + 1) * n_cells).
__global__ void __launch_bounds__(256) place_pattern_kernel(const float4* __restrict__ pat_verts, const uint32_t* __restrict__ face_vert_off,
                                                            uint32_t n_faces, uint32_t n_fverts, const float* __restrict__ xform6,
                                                            uint32_t n_place, float4* __restrict__ c_planes, float4* __restrict__ c_verts)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n_faces * n_place) return;
    const uint32_t p = (uint32_t)(i / n_faces), f = (uint32_t)(i - (uint64_t)p * n_faces);
    const float sx = xform6[6 * p], sy = xform6[6 * p + 1], sz = xform6[6 * p + 2];
    const float tx = xform6[6 * p + 3], ty = xform6[6 * p + 4], tz = xform6[6 * p + 5];
    const uint32_t v0 = face_vert_off[f], v1 = face_vert_off[f + 1];
    float4* out = c_verts + (size_t)p * n_fverts;
    float q[9];
    for (uint32_t v = v0; v < v1; v++)
    {
        const float4 a = __ldg(pat_verts + v);
        const float x = __fadd_rn(__fmul_rn(a.x, sx), tx), y = __fadd_rn(__fmul_rn(a.y, sy), ty), z = __fadd_rn(__fmul_rn(a.z, sz), tz);
        out[v] = make_float4(x, y, z, 0.f);
        if (v - v0 < 3) { q[3 * (v - v0)] = x; q[3 * (v - v0) + 1] = y; q[3 * (v - v0) + 2] = z; }
    }
    c_planes[(size_t)p * n_faces + f] = plane_from_points(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8]);
}

// PCIe wire format (surtr_upload_pieces3 / surtr_upload_cells3): float3 stream -> resident float4 stream, w = 0.
__global__ void __launch_bounds__(256) widen3_kernel(const float* __restrict__ in3, float4* __restrict__ out4, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out4[i] = make_float4(__ldg(in3 + 3 * i), __ldg(in3 + 3 * i + 1), __ldg(in3 + 3 * i + 2), 0.f);
}

// surtr_upload_blob: the compact input blob -> the resident arrays, one launch.  Both float3 streams (pieces, cell
// vertices) are widened to float4; the ring offsets are rebuilt from one LENGTH byte per vertex (eight lanes per piece: its
// first ring entry comes with the blob, the rest is a prefix sum); RB = 1: the ring entries travel as bytes (every
// piece has at most 256 vertices) and are widened to the resident 16-bit entries.
template <int RB>
__global__ void __launch_bounds__(256) expand_blob_kernel(const float* __restrict__ a3, float4* __restrict__ a4, uint64_t na,
                                                          const float* __restrict__ b3, float4* __restrict__ b4, uint64_t nb,
                                                          const uint32_t* __restrict__ vert_off, const uint32_t* __restrict__ ring_base,
                                                          const uint8_t* __restrict__ ring_len, uint32_t n_pieces, uint32_t* __restrict__ ring_off,
                                                          const uint8_t* __restrict__ ring8, uint16_t* __restrict__ ring16, uint64_t n_ring)
{
    const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = t0; i < na + nb; i += stride)
    {
        if (i < na) a4[i] = make_float4(__ldg(a3 + 3 * i), __ldg(a3 + 3 * i + 1), __ldg(a3 + 3 * i + 2), 0.f);
        else { const uint64_t j = i - na; b4[j] = make_float4(__ldg(b3 + 3 * j), __ldg(b3 + 3 * j + 1), __ldg(b3 + 3 * j + 2), 0.f); }
    }
    // ring offsets: 8 lanes per piece (four pieces per warp: the pieces of the BASELINE configs have 8-60 vertices), an
    // 8-wide prefix sum of the length bytes per round; the trip count is the warp's largest piece (full-mask shuffles)
    const int lane = threadIdx.x & 31, sl = lane & 7;
    for (uint64_t pw = (t0 >> 5) * 4; pw < n_pieces; pw += (stride >> 5) * 4)
    {
        const uint64_t p = pw + (uint64_t)(lane >> 3);
        const bool have = p < n_pieces;
        const uint32_t v0 = have ? vert_off[p] : 0u, v1 = have ? vert_off[p + 1] : 0u;
        uint32_t run = have ? ring_base[p] : 0u;
        const uint32_t nmax = __reduce_max_sync(FULL, v1 - v0);
        for (uint32_t i = 0; i < nmax; i += 8)
        {
            const uint32_t v = v0 + i + (uint32_t)sl;
            const uint32_t len = v < v1 ? ring_len[v] : 0u;
            uint32_t inc = len;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1)
            {
                const uint32_t t = __shfl_up_sync(FULL, inc, o, 8);
                if (sl >= o) inc += t;
            }
            if (v < v1) ring_off[v] = run + inc - len;
            run += __shfl_sync(FULL, inc, 7, 8);
        }
        if (have && p + 1 == n_pieces && sl == 0) ring_off[v1] = ring_base[n_pieces];
    }
    if (n_pieces == 0 && t0 == 0) ring_off[0] = 0u;
    if (RB == 1)
    {
        // 16 entries per thread where both sides are aligned (both arrays start on a 256-byte boundary)
        const uint64_t n16 = n_ring / 16;
        const uint4* in16 = reinterpret_cast<const uint4*>(ring8);
        uint4* out16 = reinterpret_cast<uint4*>(ring16);
        for (uint64_t i = t0; i < n16; i += stride)
        {
            const uint4 x = __ldg(in16 + i);
            out16[2 * i] = make_uint4(__byte_perm(x.x, 0u, 0x4140), __byte_perm(x.x, 0u, 0x4342), __byte_perm(x.y, 0u, 0x4140), __byte_perm(x.y, 0u, 0x4342));
            out16[2 * i + 1] = make_uint4(__byte_perm(x.z, 0u, 0x4140), __byte_perm(x.z, 0u, 0x4342), __byte_perm(x.w, 0u, 0x4140), __byte_perm(x.w, 0u, 0x4342));
        }
        for (uint64_t i = 16 * n16 + t0; i < n_ring; i += stride) ring16[i] = ring8[i];
    }
}

// surtr_download_blob_async: the four fragment arrays into ONE contiguous device blob in the packed wire format
// (records copied, positions narrowed to float3, ring offsets turned into one length byte per vertex; RB = 1: no fragment
// of the event has more than 256 vertices and the ring entries travel as bytes, RB = 2: copied as they are).
template <int RB>
__global__ void __launch_bounds__(256) pack_blob_kernel(const uint4* __restrict__ rec16, uint64_t n_rec16, const float4* __restrict__ verts4,
                                                        const uint32_t* __restrict__ ring_off, uint64_t n_verts,
                                                        const uint16_t* __restrict__ ring, uint64_t n_ring, uint4* __restrict__ o_rec16,
                                                        float* __restrict__ o_verts3, uint8_t* __restrict__ o_len, void* __restrict__ o_ring)
{
    const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = t0; i < n_rec16; i += stride) o_rec16[i] = rec16[i];
    for (uint64_t i = t0; i < n_verts; i += stride)
    {
        const float4 v = verts4[i];
        o_verts3[3 * i] = v.x; o_verts3[3 * i + 1] = v.y; o_verts3[3 * i + 2] = v.z;
        o_len[i] = (uint8_t)(ring_off[i + 1] - ring_off[i]);
    }
    // ring entries: 16 bytes per store where both sides are aligned (f_ring is a cudaMalloc'd array, the section 256-byte aligned)
    if (RB == 2)
    {
        const uint64_t n8 = n_ring / 8;
        const uint4* r16 = reinterpret_cast<const uint4*>(ring);
        uint4* o16 = reinterpret_cast<uint4*>(o_ring);
        uint16_t* o1 = reinterpret_cast<uint16_t*>(o_ring);
        for (uint64_t i = t0; i < n8; i += stride) o16[i] = r16[i];
        for (uint64_t i = 8 * n8 + t0; i < n_ring; i += stride) o1[i] = ring[i];
    }
    else
    {
        const uint64_t n16 = n_ring / 16;
        const uint4* r16 = reinterpret_cast<const uint4*>(ring);
        uint4* o16 = reinterpret_cast<uint4*>(o_ring);
        uint8_t* o1 = reinterpret_cast<uint8_t*>(o_ring);
        for (uint64_t i = t0; i < n16; i += stride)
        {
            const uint4 a = r16[2 * i], b = r16[2 * i + 1];   // the low byte of each 16-bit entry
            o16[i] = make_uint4(__byte_perm(a.x, a.y, 0x6420), __byte_perm(a.z, a.w, 0x6420), __byte_perm(b.x, b.y, 0x6420), __byte_perm(b.z, b.w, 0x6420));
        }
        for (uint64_t i = 16 * n16 + t0; i < n_ring; i += stride) o1[i] = (uint8_t)ring[i];
    }
}

// PCIe wire format of the fragments (surtr_download_fragments_packed): float3 positions, one byte of ring length per vertex.
__global__ void __launch_bounds__(256) pack_fragments_kernel(const float4* __restrict__ verts4, const uint32_t* __restrict__ ring_off,
                                                             float* __restrict__ verts3, uint8_t* __restrict__ ring_len, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        const float4 v = verts4[i];
        verts3[3 * i] = v.x; verts3[3 * i + 1] = v.y; verts3[3 * i + 2] = v.z;
        ring_len[i] = (uint8_t)(ring_off[i + 1] - ring_off[i]);
    }
}

// Poly::Transform (Poly.cpp:580-585) over the resident pieces: position = XMVector3TransformCoord(position, M^T) with the
// piece's world matrix M (row-major as the caller holds it; the reference transposes before use), i.e. per output
// component c: ((z*M[c][2] + M[c][3]) + y*M[c][1]) + x*M[c][0], then a division by the w component.  One warp per piece.
__global__ void __launch_bounds__(256) transform_pieces_kernel(float4* __restrict__ p_verts, const uint32_t* __restrict__ p_vert_off,
                                                               uint32_t n_pieces, const float* __restrict__ matrices16,
                                                               const uint32_t* __restrict__ piece_matrix)
{
    const int lane = threadIdx.x & 31;
    const uint32_t piece = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (piece >= n_pieces) return;
    const float* M = matrices16 + 16 * (size_t)(piece_matrix ? piece_matrix[piece] : 0u);
    float m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = __ldg(M + k);
    for (uint32_t v = p_vert_off[piece] + lane; v < p_vert_off[piece + 1]; v += 32)
    {
        const float4 p = p_verts[v];
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
            o[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.z, m[4 * c + 2]), m[4 * c + 3]), __fmul_rn(p.y, m[4 * c + 1])), __fmul_rn(p.x, m[4 * c]));
        p_verts[v] = make_float4(__fdiv_rn(o[0], o[3]), __fdiv_rn(o[1], o[3]), __fdiv_rn(o[2], o[3]), p.w);
    }
}

// FP32 FMA microbenchmark (measurement only; SURVEY section 8d asks for the FP32 pipe peak measured in the same run as
// the bench): 8 independent FFMA chains per thread, explicit __fmaf_rn (the TU is built with -fmad=false).
__global__ void __launch_bounds__(256) fma_peak_kernel(float* __restrict__ out, int iters, float a, float b)
{
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = (float)(threadIdx.x + k) * 1.0e-3f;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = __fmaf_rn(x[k], a, b);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the chains alive
}

// ---------------------------------------------------------------------------------------------- K2
// Conservative separation test: a pair is culled only when some slab is separated by more than a few ulps of
// the extents' magnitude, so no pair the clipper would keep is ever dropped (slivers).
__device__ __forceinline__ bool slab_separated(float amin, float amax, float bmin, float bmax)
{
    const float scale = fmaxf(fmaxf(fabsf(amin), fabsf(amax)), fmaxf(fabsf(bmin), fabsf(bmax)));
    const float tol = 4.0e-6f * scale;
    return (amin - bmax > tol) || (bmin - amax > tol);
}

template <int K>
__global__ void __launch_bounds__(256) broadphase_mask_kernel(const BpTile* __restrict__ tiles,
                                                              const float* __restrict__ ext_p,
                                                              const float* __restrict__ ext_c,
                                                              unsigned int* __restrict__ masks)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_cell[32 * 2 * K];
    const BpTile t = tiles[blockIdx.x];
    for (int i = threadIdx.x; i < (int)t.n_cell * 2 * K; i += 256)
        s_cell[i] = ext_c[(size_t)t.cell_begin * 2 * K + i];
    float e[2 * K];
    const bool valid = threadIdx.x < t.n_piece;
    if (valid)
    {
#pragma unroll
        for (int i = 0; i < 2 * K; i++) e[i] = ext_p[(size_t)(t.piece_begin + threadIdx.x) * 2 * K + i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp * 32 >= (int)t.n_piece) return;
    for (int c = 0; c < (int)t.n_cell; c++)
    {
        bool hit = valid;
        if (valid)
        {
#pragma unroll
            for (int d = 0; d < K; d++)
                hit = hit && !slab_separated(e[2 * d], e[2 * d + 1], s_cell[c * 2 * K + 2 * d], s_cell[c * 2 * K + 2 * d + 1]);
        }
        const unsigned m = __ballot_sync(FULL, hit);
        if (lane == 0) masks[(size_t)t.mask_base + (size_t)c * t.n_w + warp] = m;
    }
}

struct EventTables
{
    const uint32_t* ev_mask_base;   // n_events + 1
    const uint32_t* ev_piece_off;   // n_events + 1
    const uint32_t* ev_cell_off;    // n_events + 1
    uint32_t n_events;
};

__device__ __forceinline__ void mask_to_pair_base(const EventTables& et, uint32_t m, uint32_t& cell, uint32_t& piece0)
{
    uint32_t e = 0;
    if (et.n_events > 1)
    {
        uint32_t lo = 0, hi = et.n_events;   // last e with ev_mask_base[e] <= m
        while (hi - lo > 1)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (et.ev_mask_base[mid] <= m) lo = mid; else hi = mid;
        }
        e = lo;
    }
    const uint32_t local = m - et.ev_mask_base[e];
    const uint32_t p0 = et.ev_piece_off[e];
    const uint32_t n_w = (et.ev_piece_off[e + 1] - p0 + 31u) >> 5;
    const uint32_t c_local = local / n_w;
    cell = et.ev_cell_off[e] + c_local;
    piece0 = p0 + ((local - c_local * n_w) << 5);
}

constexpr int CP_THREADS = 256;
constexpr int CP_ITEMS = 4;   // ballot words per thread
__global__ void __launch_bounds__(CP_THREADS) compact_pairs_kernel(const unsigned int* __restrict__ masks,
                                                                    uint32_t n_masks, EventTables et,
                                                                    ScanState<1> st, Ctl* ctl,
                                                                    uint2* __restrict__ cand, uint64_t cap_cand)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_tile;
    __shared__ unsigned int s_warp[CP_THREADS / 32];
    __shared__ unsigned long long s_excl;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctl->tile_a, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t m0 = (uint32_t)tile * (CP_THREADS * CP_ITEMS) + threadIdx.x * CP_ITEMS;
    unsigned int w[CP_ITEMS];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < CP_ITEMS; i++)
    {
        w[i] = (m0 + i < n_masks) ? masks[m0 + i] : 0u;
        cnt += __popc(w[i]);
    }
    int wtot;
    const int wex = warp_exscan(cnt, lane, wtot);
    if (lane == 31) s_warp[warp] = (unsigned)wtot;
    __syncthreads();
    unsigned int block_ex = 0, block_tot = 0;
#pragma unroll
    for (int i = 0; i < CP_THREADS / 32; i++)
    {
        if (i < warp) block_ex += s_warp[i];
        block_tot += s_warp[i];
    }
    if (warp == 0)
    {
        unsigned long long agg[1] = { block_tot }, ex[1];
        tile_lookback<1>(st, tile, agg, ex, lane);
        if (lane == 0)
        {
            s_excl = ex[0];
            if ((uint32_t)(tile + 1) * (CP_THREADS * CP_ITEMS) >= n_masks) ctl->n_cand = ex[0] + block_tot;
        }
    }
    __syncthreads();
    unsigned long long out = s_excl + block_ex + wex;
#pragma unroll
    for (int i = 0; i < CP_ITEMS; i++)
    {
        unsigned int m = w[i];
        if (m)
        {
            uint32_t cell, piece0;
            mask_to_pair_base(et, m0 + i, cell, piece0);
            while (m)
            {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                if (out < cap_cand) cand[out] = make_uint2(piece0 + b, cell);
                out++;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- K3
constexpr size_t FAST_BLOB_BYTES = 64 * 16 + 64 * 2 + 64 * 8;   // small-tier result blob: float4 verts[64] | u16 ring_start[64] | u8 ring[packed]

struct ClipArgs
{
    const float4* p_verts;
    const uint32_t* p_vert_off;
    const uint32_t* p_ring_off;
    const uint16_t* p_ring;
    const float4* c_planes;
    const uint32_t* c_plane_off;
    const float* ext_p;           // K1's slab extents of the pieces ([piece][2k]; slabs 0-2 = x, y, z): the small tier's plane prefilter
    int kdirs;
    const uint2* cand;
    uint64_t cap_cand;
    CandRec* rec;
    unsigned char* scratch;       // this tier's blob area
    uint64_t slot_bytes;
    unsigned char* scratch1;      // small-tier blob area (one slot per candidate): tier 2 hands small results to K4's moments
    uint32_t* ovf_list;           // tier 1 (64 slots) appends, tier 1b (128 slots) consumes
    uint32_t* ovf2_list;          // tier 1b appends, tier 2 consumes
    int skip_tier1b;              // test hook / round-1 kernel: tier 1 hands its overflows straight to tier 2
    uint64_t cap_tier2;           // slots available to tier 2
    uint32_t* ovf3_list;          // tiers 1 / 1b / 2 append, tier 3 consumes
    unsigned char* ws3;           // tier 3: one workspace per warp
    uint64_t ws3_stride;
    int cap3;                     // tier 3: vertex slots per workspace
    int gd3;                      // tier 3: ring slots per vertex of the workspace (grown by the host on demand, never a hard limit)
    uint64_t cap_tier3;           // result slots available to tier 3
    uint32_t* fail_list;          // candidates that cannot be cut (malformed rings): reported per pair, the rest of the event is valid
    Ctl* ctl;
    uint32_t* dbg;                // optional: 8 words per candidate (cycles per phase, cut counts); NULL = off
};

// K3, large on-chip tier: pieces / intermediate results of up to 256 vertex slots and ring degree 16, one block of
// T2_WARPS warps per pair, persistent blocks over the pairs the small tier handed on.  Same rolled code as the unbounded tier (clip_global.cuh)
// with the pair's workspace in shared memory: the unrolled register-array version this replaces was 179 KB of
// SASS and spent 16 stalled cycles per issued instruction on instruction fetch (profiles/README.md).
constexpr int T2_CAP = 256;
constexpr int T2_WARPS = 4;            // warps per pair (= per block) in the large tier
constexpr int T2_BLOCKS_PER_SM = 4;
__host__ __device__ constexpr size_t blob2_bytes() { return (size_t)T2_CAP * (16 + 2 + GD * 2); }   // float4 verts | u16 ring_start | u16 ring
__host__ __device__ constexpr size_t t2_ws_bytes() { return (global_poly_bytes(T2_CAP) + 15) / 16 * 16; }

__global__ void __launch_bounds__(T2_WARPS * 32) clip_shared_kernel(ClipArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int N = T2_WARPS * 32;
    __shared__ int s_scan[T2_WARPS + 1];
    __shared__ float s_cov[T2_WARPS * 10];
    const int tid = threadIdx.x, lane = threadIdx.x & 31;
    const unsigned long long n_items = a.ctl->n_ovf2;
    GlobalPoly g = global_poly_carve(smem_raw, T2_CAP);   // one pair per block: the block's T2_WARPS warps share the workspace
    const Grp<T2_WARPS> grp{ tid, lane, s_scan };
    unsigned seq_cuts = 0;
    for (unsigned long long it = blockIdx.x; it < n_items; it += gridDim.x)
    {
        const uint32_t q = a.ovf2_list[it];
        const uint2 pr = a.cand[q];
        const uint32_t v0 = a.p_vert_off[pr.x];
        int nv = (int)(a.p_vert_off[pr.x + 1] - v0);
        bool too_big = nv > T2_CAP, malformed = false;
        if (!too_big)
        {
            for (int v = tid; v < nv; v += N)
            {
                const float4 p = __ldg(a.p_verts + v0 + v);
                g.x[v] = p.x; g.y[v] = p.y; g.z[v] = p.z;
                const uint32_t r0 = a.p_ring_off[v0 + v];
                const int d = (int)(a.p_ring_off[v0 + v + 1] - r0);
                if (d > GD) too_big = true;
                else
                {
                    g.deg[v] = (uint16_t)d;
                    if (d == 0) malformed = true;   // a vertex without neighbours is not a polyhedron vertex
                    for (int j = 0; j < d; j++)
                    {
                        const int idx = a.p_ring[r0 + j];
                        if (idx >= nv) malformed = true;   // reported as a failed pair, never used as an index
                        g.ring[(size_t)v * GD + j] = (uint16_t)idx;
                    }
                }
            }
        }
        malformed = grp.any(malformed);
        too_big = grp.any(too_big);
        grp.sync();
        int status = CLIP_OVERFLOW;
        if (!too_big && !malformed)
        {
            const uint32_t pl0 = a.c_plane_off[pr.y];
            const int npl = (int)(a.c_plane_off[pr.y + 1] - pl0);
            status = global_clip_by_planes<T2_WARPS>(g, nv, a.c_planes + pl0, npl, grp, seq_cuts);
        }
        CandRec* rec = a.rec + q;
        if (status != CLIP_OK)
        {
            if (tid == 0)
            {
                rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 0;
                if (malformed) a.fail_list[atomicAdd(&a.ctl->n_fail, 1u)] = q;
                else a.ovf3_list[atomicAdd(&a.ctl->n_ovf3, 1u)] = q;   // on to the global-memory tier (more slots, wider rings)
            }
            grp.sync();
            continue;
        }
        if (nv == 0)
        {
            if (tid == 0) { rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 2; }
            grp.sync();
            continue;
        }
        // ring entries of the result, and whether it fits the small tier's blob (<= 64 vertices, degree <= 8): then K4's
        // gather computes its face count and moments like for every small fragment
        int ne = 0;
        bool wide = false;
        for (int base = 0; base < nv; base += N)
        {
            const int d = base + tid < nv ? g.deg[base + tid] : 0;
            wide |= d > 8;
            int tot;
            grp.exscan(d, tot);
            ne += tot;
        }
        const bool small = nv <= 64 && !grp.any(wide);
        Moments mo;
        if (!small) global_fragment_moments<T2_WARPS>(g, nv, grp, mo, s_cov);
        const bool room = small || it < a.cap_tier2;
        const unsigned long long blob = small ? (unsigned long long)q * FAST_BLOB_BYTES : it * a.slot_bytes;
        unsigned char* b = (small ? a.scratch1 : a.scratch) + blob;
        float4* bv = reinterpret_cast<float4*>(b);
        uint16_t* bo = reinterpret_cast<uint16_t*>(b + (size_t)(small ? 64 : T2_CAP) * 16);
        uint8_t* br8 = b + 64 * 18;
        uint16_t* br16 = reinterpret_cast<uint16_t*>(b + (size_t)T2_CAP * 18);
        int run = 0;
        for (int base = 0; base < nv; base += N)
        {
            const int v = base + tid;
            const int d = v < nv ? g.deg[v] : 0;
            int tot;
            const int off = run + grp.exscan(d, tot);
            run += tot;
            if (v < nv && room)
            {
                bv[v] = make_float4(g.x[v], g.y[v], g.z[v], 0.f);
                bo[v] = (uint16_t)off;
                if (small)
                    for (int j = 0; j < d; j++) br8[off + j] = (uint8_t)g.ring[(size_t)v * GD + j];
                else
                    for (int j = 0; j < d; j++) br16[off + j] = g.ring[(size_t)v * GD + j];
            }
        }
        if (tid == 0)
        {
            rec->nv = room ? (uint32_t)nv : 0u;
            rec->ne = (uint32_t)ne;
            rec->tier = small ? 1 : 2;
            rec->blob = blob;
            rec->nf = 0;
            if (!small)
            {
                rec->nf = (uint32_t)mo.n_faces;
                rec->volume = mo.volume;
                rec->centroid[0] = mo.cx; rec->centroid[1] = mo.cy; rec->centroid[2] = mo.cz;
#pragma unroll
                for (int k = 0; k < 6; k++) rec->inertia[k] = mo.inertia[k];
            }
            // (!room: the host sees n_ovf2 > cap_tier2, grows the result slots and re-runs the event)
        }
        grp.sync();
    }
    if (seq_cuts && tid == 0) atomicAdd(&a.ctl->n_seq_cuts, seq_cuts);
}

// K3, unbounded tier: persistent warps over the pairs the on-chip tiers handed on (clip_global.cuh).
constexpr int T3_WARPS = 8;            // warps per pair (= per block) in the unbounded tier
constexpr int T3_BLOCKS_PER_SM = 2;    // persistent blocks, one workspace each
__host__ __device__ constexpr size_t blob3_bytes(size_t cap, size_t gd) { return cap * (16 + 4 + gd * 2); }   // float4 verts | u32 ring_start | u16 ring

__global__ void __launch_bounds__(T3_WARPS * 32) clip_global_kernel(ClipArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    constexpr int N = T3_WARPS * 32;
    __shared__ int s_scan[T3_WARPS + 1];
    __shared__ float s_cov[T3_WARPS * 10];
    const int tid = threadIdx.x;
    const Grp<T3_WARPS> grp{ tid, tid & 31, s_scan };
    const unsigned long long n_items = a.ctl->n_ovf3;
    GlobalPoly g = global_poly_carve(a.ws3 + (size_t)blockIdx.x * a.ws3_stride, a.cap3, a.gd3);
    const size_t GS = (size_t)a.gd3;
    unsigned seq_cuts = 0;
    for (unsigned long long it = blockIdx.x; it < n_items; it += gridDim.x)
    {
        const uint32_t q = a.ovf3_list[it];
        const uint2 pr = a.cand[q];
        const uint32_t v0 = a.p_vert_off[pr.x];
        int nv = (int)(a.p_vert_off[pr.x + 1] - v0);
        bool bad = nv > g.cap, wide = false;
        unsigned need_deg = 0;
        if (!bad)
        {
            for (int v = tid; v < nv; v += N)
            {
                const float4 p = __ldg(a.p_verts + v0 + v);
                g.x[v] = p.x; g.y[v] = p.y; g.z[v] = p.z;
                const uint32_t r0 = a.p_ring_off[v0 + v];
                const int d = (int)(a.p_ring_off[v0 + v + 1] - r0);
                if (d == 0) bad = true;
                else if (d > g.gd) { wide = true; need_deg = max(need_deg, (unsigned)d); }   // the workspace's rings are too narrow: the host widens them
                else
                {
                    g.deg[v] = (uint16_t)d;
                    for (int j = 0; j < d; j++)
                    {
                        const int idx = a.p_ring[r0 + j];
                        if (idx >= nv) bad = true;
                        g.ring[v * GS + j] = (uint16_t)idx;
                    }
                }
            }
        }
        bad = grp.any(bad);
        wide = grp.any(wide);
        if (need_deg) atomicMax(&a.ctl->max_deg, need_deg);
        grp.sync();
        int status = nv > g.cap ? CLIP_NEED_SLOTS : (wide && !bad ? CLIP_NEED_DEG : CLIP_OVERFLOW);
        if (!bad && !wide)
        {
            const uint32_t pl0 = a.c_plane_off[pr.y];
            const int npl = (int)(a.c_plane_off[pr.y + 1] - pl0);
            status = global_clip_by_planes<T3_WARPS>(g, nv, a.c_planes + pl0, npl, grp, seq_cuts);
        }
        CandRec* rec = a.rec + q;
        const bool room = it < a.cap_tier3;
        if (status != CLIP_OK || !room)
        {
            if (tid == 0)
            {
                rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 0;
                // !room alone: the host grows the result slots and re-runs; out of vertex / ring slots: it grows the workspace
                // (at the end of the 16-bit index range nothing is left to grow: the pair is reported as failed)
                if (status == CLIP_NEED_SLOTS && g.cap < 65520) atomicAdd(&a.ctl->n_grow3, 1u);
                else if (status == CLIP_NEED_DEG && g.gd < 65520) atomicAdd(&a.ctl->n_growdeg, 1u);
                else if (status != CLIP_OK) a.fail_list[atomicAdd(&a.ctl->n_fail, 1u)] = q;
            }
            grp.sync();
            continue;
        }
        if (nv == 0)
        {
            if (tid == 0) { rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 3; }
            grp.sync();
            continue;
        }
        Moments mo;
        global_fragment_moments<T3_WARPS>(g, nv, grp, mo, s_cov);
        if (grp.bcast0(mo.n_faces) > 65535)
        {
            // surtr_fragment::n_faces is 16 bits wide: reported as a failed pair, never truncated
            if (tid == 0) { rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 0; a.fail_list[atomicAdd(&a.ctl->n_fail, 1u)] = q; }
            grp.sync();
            continue;
        }
        const unsigned long long blob = it * a.slot_bytes;
        unsigned char* b = a.scratch + blob;
        float4* bv = reinterpret_cast<float4*>(b);
        uint32_t* bo = reinterpret_cast<uint32_t*>(b + (size_t)g.cap * 16);
        uint16_t* br = reinterpret_cast<uint16_t*>(b + (size_t)g.cap * 20);
        int ne = 0;
        for (int base = 0; base < nv; base += N)
        {
            const int v = base + tid;
            const int d = v < nv ? g.deg[v] : 0;
            int tot;
            const int off = ne + grp.exscan(d, tot);
            ne += tot;
            if (v < nv)
            {
                bv[v] = make_float4(g.x[v], g.y[v], g.z[v], 0.f);
                bo[v] = (uint32_t)off;
                for (int j = 0; j < d; j++) br[off + j] = g.ring[v * GS + j];
            }
        }
        if (tid == 0)
        {
            rec->nv = (uint32_t)nv;
            rec->ne = (uint32_t)ne;
            rec->nf = (uint32_t)mo.n_faces;
            rec->tier = 3;
            rec->volume = mo.volume;
            rec->centroid[0] = mo.cx; rec->centroid[1] = mo.cy; rec->centroid[2] = mo.cz;
#pragma unroll
            for (int k = 0; k < 6; k++) rec->inertia[k] = mo.inertia[k];
            rec->blob = blob;
        }
        grp.sync();
    }
    if (seq_cuts && tid == 0) atomicAdd(&a.ctl->n_seq_cuts, seq_cuts);
}


// K3, small tier (round 2): ONE warp per candidate pair, clip_fast.cuh.  G = 2 (64 vertex slots) is the main launch: warp
// w of the grid cuts candidate w, no persistent loop (the hardware scheduler balances the very uneven pair costs).
// G = 4 (128 slots, LIST = true) is a second, small launch over the pairs the main launch could not finish -- pieces of
// 50-60 vertices whose cut transiently needs more than 64 slots (clipped vertices keep their slots until the patch has
// walked through them) -- so that they do not detour through the block-per-pair large tier.
// DBG: the per-candidate cycle counters of csrc/surtr_debug.h (tests/measure/gpu_cycles.py).  A separate instantiation: in
// the kernel that ships, three 64-bit time stamps and the cut counter would be live across the whole clip.
template <int G, bool LIST, bool DBG = false, bool LAT = false>
__device__ __forceinline__ void fast_pair(const ClipArgs& a, FastPoly<G>& sp, uint32_t q, int lane)
{
    constexpr int RG = LAT ? G : SURTR_K3_REG_GROUPS;   // vertex groups with positions in registers (clip_fast.cuh)
    constexpr int S = 32 * G;
    const long long t0 = DBG && a.dbg ? clock64() : 0;
    const uint2 pr = a.cand[q];
    const uint32_t v0 = a.p_vert_off[pr.x];
    int nv = (int)(a.p_vert_off[pr.x + 1] - v0);
    const uint32_t pl0 = a.c_plane_off[pr.y];
    const int npl = (int)(a.c_plane_off[pr.y + 1] - pl0);
    const int nv_in = nv;
    float px[G], py[G], pz[G];
    bool bad = nv > S;
    if (!bad)
    {
        // The piece's ring entries are ONE contiguous u16 stream (p_ring_off[v0] .. p_ring_off[v0 + nv]): the lanes load it
        // coalesced and stage it as bytes in old_ring (unused until a sequential replay), then every lane assembles the
        // 64-bit ring words of its vertices from shared memory -- instead of a chain of per-lane global loads, one per
        // ring entry (the long-scoreboard stall at the head of every pair in profiles/r2_k3_cfg4.txt).
        uint8_t* stage = reinterpret_cast<uint8_t*>(sp.old_ring);
        const uint32_t rb = __ldg(a.p_ring_off + v0), re = __ldg(a.p_ring_off + v0 + nv);
        int ne = (int)(re - rb);
        if (re < rb || ne > 8 * S) { bad = true; ne = 0; }
#pragma unroll 4
        for (int e = lane; e < ne; e += 32)
        {
            const int idx = __ldg(a.p_ring + rb + e);
            bad = bad || idx >= nv;
            stage[e] = (uint8_t)idx;
        }
        float4 pp[G];
        uint32_t r0[G], r1[G];
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            const int v = lane + 32 * g;
            pp[g] = make_float4(0.f, 0.f, 0.f, 0.f);
            r0[g] = r1[g] = rb;
            if (v < nv)
            {
                pp[g] = __ldg(a.p_verts + v0 + v);
                r0[g] = __ldg(a.p_ring_off + v0 + v);
                r1[g] = __ldg(a.p_ring_off + v0 + v + 1);
            }
        }
        __syncwarp();
        const uint32_t* stage32 = reinterpret_cast<const uint32_t*>(sp.old_ring);
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            const int v = lane + 32 * g;
            px[g] = pp[g].x; py[g] = pp[g].y; pz[g] = pp[g].z;
            if (v < nv)
            {
                sp.x[v] = px[g]; sp.y[v] = py[g]; sp.z[v] = pz[g];
                const int d = (int)(r1[g] - r0[g]);
                const uint32_t o = r0[g] - rb;
                u64 rw = ~0ull;
                if (r0[g] < rb || r1[g] > re || d > 8 || d <= 0) bad = true;
                else
                {
                    // eight bytes from byte offset o (unaligned): three aligned words through the funnel shifter; the
                    // read may run up to 11 bytes past the stream's end, which is still inside this FastPoly
                    const uint32_t w0 = stage32[o >> 2], w1 = stage32[(o >> 2) + 1], w2 = stage32[(o >> 2) + 2];
                    const unsigned sh = (o & 3u) * 8u;
                    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi2 = __funnelshift_r(w1, w2, sh);
                    rw = (u64)lo | ((u64)hi2 << 32);
                    if (d < 8) rw |= ~0ull << (8 * d);
                }
                sp.ring[v] = rw;
            }
        }
        __syncwarp();   // old_ring is free again (fast_seq_cut snapshots into it)
    }
    bad = __ballot_sync(FULL, bad) != 0u;
    __syncwarp();
    const long long t1 = DBG && a.dbg ? clock64() : 0;
    unsigned seq_cuts = 0, n_cuts = 0;
    unsigned live[G];
    int hi = 0, status = CLIP_OVERFLOW;
    // the piece's axis-aligned box for the plane prefilter: the first three slabs K1 wrote (x, y, z of every direction set)
    float box[6];
#pragma unroll
    for (int k = 0; k < 6; k++) box[k] = a.ext_p ? __ldg(a.ext_p + (size_t)pr.x * 2 * a.kdirs + k) : 0.f;
    if (!bad) status = fast_clip_by_planes<G, RG, LAT || SURTR_K3_PREFETCH != 0>(sp, live, hi, nv, px, py, pz, a.c_planes + pl0, npl, lane, seq_cuts, n_cuts, box, a.ext_p != nullptr);
    const long long t2 = DBG && a.dbg ? clock64() : 0;
    if (DBG && a.dbg && lane == 0)
    {
        uint32_t* d = a.dbg + (size_t)q * 8;
        d[0] = (uint32_t)(t1 - t0); d[1] = (uint32_t)(t2 - t1); d[2] = 0; d[3] = 0;
        d[4] = seq_cuts; d[5] = n_cuts; d[6] = (uint32_t)nv_in; d[7] = (uint32_t)npl;
    }
    if (seq_cuts && lane == 0) atomicAdd(&a.ctl->n_seq_cuts, seq_cuts);
    CandRec* rec = a.rec + q;
    if (status == CLIP_OK && nv > 64) status = CLIP_OVERFLOW;   // (G = 4 only) the result does not fit the small tier's blob
    if (status != CLIP_OK)
    {
        if (lane == 0)
        {
            // too large for this tier (or a ring outgrew 8 slots): hand the pair on
            rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 0;
            if (nv_in > T2_CAP) a.ovf3_list[atomicAdd(&a.ctl->n_ovf3, 1u)] = q;
            else if (LIST || a.skip_tier1b) a.ovf2_list[atomicAdd(&a.ctl->n_ovf2, 1u)] = q;
            else a.ovf_list[atomicAdd(&a.ctl->n_ovf, 1u)] = q;
        }
        return;
    }
    if (nv == 0)
    {
        if (lane == 0) { rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 1; }
        return;
    }
    // result blob, renumbered to the reference's final order (rank in the live mask):
    // float4 verts[64] | u16 ring_start[64] | u8 ring[packed].  Face count and moments are K4's (assemble_gather_kernel).
    const unsigned long long blob = (unsigned long long)q * FAST_BLOB_BYTES;
    unsigned char* b = a.scratch1 + blob;
    float4* bv = reinterpret_cast<float4*>(b);
    uint16_t* bo = reinterpret_cast<uint16_t*>(b + 64 * 16);
    uint8_t* br = b + 64 * 18;
#pragma unroll
    for (int g = 0; g < G; g++)
    {
        const int v = lane + 32 * g;
        if ((g == 0 || hi > 32 * g) && ((live[g] >> lane) & 1u)) sp.id[v] = (uint8_t)mrank<G>(live, v);
    }
    __syncwarp();
    int ne = 0;
#pragma unroll
    for (int g = 0; g < G; g++)
    {
        if (g == 0 || hi > 32 * g)
        {
            const int v = lane + 32 * g;
            const bool lv = (live[g] >> lane) & 1u;
            const u64 rw = lv ? sp.ring[v] : ~0ull;
            const int d = rdeg(rw);
            int tot;
            const int off = ne + warp_exscan(d, lane, tot);
            ne += tot;
            if (lv)
            {
                const int t = sp.id[v];
                bv[t] = g < RG ? make_float4(px[g], py[g], pz[g], 0.f) : make_float4(sp.x[v], sp.y[v], sp.z[v], 0.f);
                bo[t] = (uint16_t)off;
#pragma unroll 1
                for (int j = 0; j < d; j++) br[off + j] = sp.id[rget(rw, j)];
            }
        }
    }
    if (lane == 0)
    {
        rec->nv = (uint32_t)nv;
        rec->ne = (uint32_t)ne;
        rec->nf = 0;
        rec->tier = 1;
        rec->blob = blob;
        if (DBG && a.dbg) a.dbg[(size_t)q * 8 + 3] = (uint32_t)(clock64() - t2);
    }
}

constexpr int FAST_WARPS = 2;   // pairs per block: a block's slots are held until its slowest pair ends; 2 packs better than 4 (profiles/README.md)
#ifndef SURTR_K3_RESIDENT_WARPS
#define SURTR_K3_RESIDENT_WARPS 40   // resident warps per SM of the persistent small-tier launch = its register budget (40 -> 48 registers);
                                     // measured 24 / 28 / 32 / 36 / 40 warps: profiles/r3_k3_experiments.md, section 3
#endif
constexpr int FAST_RESIDENT_WARPS = SURTR_K3_RESIDENT_WARPS;
constexpr int FAST_PERSIST_WARPS = 4;   // warps per block of the resident launch (each warp pulls its own candidates; fewer, larger blocks = fewer block launches on a one-wave event)
// W = warps (= pairs) per block, PERSIST = resident warps with a ticket counter (the main launch; measured against one
// block per one / two pairs in profiles/r2_k3_launch_shape.md: 2.50 vs 4.04 / 3.41 ms on a 256-event config-4 batch -- the
// register file holds 32 of these warps per SM, a two-pair block keeps its slots until its slower pair is through, and
// one-pair blocks are bound by the block launch rate).
// LAT: the latency build of the clipper for events that fit one wave of warps (clip_fast.cuh); the host picks it from the
// candidate count of the context's previous event.
template <int G, bool LIST, int W = FAST_WARPS, bool PERSIST = false, bool DBG = false, bool LAT = false>
__global__ void __launch_bounds__(W * 32, G == 2 ? (PERSIST && !LAT ? FAST_RESIDENT_WARPS : 32) / W : 8) clip_fast_kernel(ClipArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ FastPoly<G> s_poly[W];
    const int lane = threadIdx.x & 31;
    FastPoly<G>& sp = s_poly[threadIdx.x >> 5];
    const unsigned long long wid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (!LIST && PERSIST)
    {
        // Resident warps (FAST_RESIDENT_WARPS per SM in blocks of four): warp w cuts candidate w, then pulls further candidates from a
        // ticket counter.  The next ticket is requested before the current pair is cut, so its round trip to the L2 hides
        // behind the cut; an event of a single wave (config 2: 4096 pairs) never touches the counter.  Against one block
        // per pair (SURTR_K3_WARPS=2) this removes the block launches -- the grid is sized by capacity, more than half
        // of its blocks find no pair -- and the ragged tail: 3.41 -> 2.50 ms on a 256-event config-4 batch
        // (profiles/r2_k3_launch_shape.md).
        unsigned long long n_items = a.ctl->n_cand;
        if (n_items > a.cap_cand) n_items = a.cap_cand;
        const unsigned nw = (gridDim.x * blockDim.x) >> 5;
        const bool more = nw < n_items;
        unsigned q = (unsigned)wid;
        while (q < n_items)
        {
            unsigned next = 0xffffffffu;
            if (more && lane == 0) next = nw + atomicAdd(&a.ctl->k3_ticket, 1u);
            fast_pair<G, LIST, DBG, LAT>(a, sp, q, lane);
            __syncwarp();
            q = __shfl_sync(FULL, next, 0);
        }
    }
    else if (!LIST)
    {
        unsigned long long n_items = a.ctl->n_cand;
        if (n_items > a.cap_cand) n_items = a.cap_cand;
        if (wid < n_items) fast_pair<G, LIST, DBG, LAT>(a, sp, (uint32_t)wid, lane);
    }
    else
    {
        const unsigned long long n_items = a.ctl->n_ovf, nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
        for (unsigned long long it = wid; it < n_items; it += nw)
        {
            fast_pair<G, LIST, DBG, LAT>(a, sp, a.ovf_list[it], lane);
            __syncwarp();
        }
    }
}

// K3, small tier, TWO pairs per warp (clip_duo.cuh): lanes 0-15 stage, clip and write candidate q, lanes 16-31 candidate
// q + 1 -- neighbours in the candidate list are neighbouring pieces against the same cell, so both halves read the same
// planes and tend to cut equally often.  Staging, status handling and the result blob are fast_pair's; every
// collective is executed by the whole warp.
__device__ __forceinline__ void duo_pair(const ClipArgs& a, FastPoly<2>& sp, uint32_t q, bool act, int lane)
{
    constexpr int S = 64;
    const int sl = lane & (DUO_L - 1), shift = lane & DUO_L;
    uint2 pr = make_uint2(0u, 0u);
    uint32_t v0 = 0, pl0 = 0;
    int nv = 0, npl = 0;
    if (act)
    {
        pr = a.cand[q];
        v0 = a.p_vert_off[pr.x];
        nv = (int)(a.p_vert_off[pr.x + 1] - v0);
        pl0 = a.c_plane_off[pr.y];
        npl = (int)(a.c_plane_off[pr.y + 1] - pl0);
    }
    const int nv_in = nv;
    bool bad = nv > S;
    uint32_t rb = 0, re = 0;
    if (act && !bad)
    {
        // the piece's ring entries: one contiguous u16 stream, loaded coalesced and staged as bytes in old_ring (fast_pair)
        uint8_t* stage = reinterpret_cast<uint8_t*>(sp.old_ring);
        rb = __ldg(a.p_ring_off + v0); re = __ldg(a.p_ring_off + v0 + nv);
        int ne = (int)(re - rb);
        if (re < rb || ne > 8 * S) { bad = true; ne = 0; }
#pragma unroll 4
        for (int e = sl; e < ne; e += DUO_L)
        {
            const int idx = __ldg(a.p_ring + rb + e);
            bad = bad || idx >= nv;
            stage[e] = (uint8_t)idx;
        }
    }
    __syncwarp();
    if (act && nv <= S)
    {
        const uint32_t* stage32 = reinterpret_cast<const uint32_t*>(sp.old_ring);
#pragma unroll 2
        for (int g = 0; g < DUO_G; g++)
        {
            const int v = sl + DUO_L * g;
            if (v < nv)
            {
                const float4 pp = __ldg(a.p_verts + v0 + v);
                const uint32_t r0 = __ldg(a.p_ring_off + v0 + v), r1 = __ldg(a.p_ring_off + v0 + v + 1);
                sp.x[v] = pp.x; sp.y[v] = pp.y; sp.z[v] = pp.z;
                const int d = (int)(r1 - r0);
                const uint32_t o = r0 - rb;
                u64 rw = ~0ull;
                if (r0 < rb || r1 > re || d > 8 || d <= 0) bad = true;
                else
                {
                    // eight bytes from byte offset o (unaligned): three aligned words through the funnel shifter (the read may
                    // run up to 11 bytes past the stream's end, still inside this FastPoly)
                    const uint32_t w0 = stage32[o >> 2], w1 = stage32[(o >> 2) + 1], w2 = stage32[(o >> 2) + 2];
                    const unsigned sh = (o & 3u) * 8u;
                    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi2 = __funnelshift_r(w1, w2, sh);
                    rw = (u64)lo | ((u64)hi2 << 32);
                    if (d < 8) rw |= ~0ull << (8 * d);
                }
                sp.ring[v] = rw;
            }
        }
    }
    bad = duo_half(__ballot_sync(FULL, bad), shift) != 0u;
    __syncwarp();   // old_ring is free again (duo_seq_cut snapshots into it)
    float box[6];
#pragma unroll
    for (int k = 0; k < 6; k++) box[k] = a.ext_p && act ? __ldg(a.ext_p + (size_t)pr.x * 2 * a.kdirs + k) : 0.f;
    DuoResult R;
    duo_clip_by_planes(sp, act && !bad, nv, a.c_planes + pl0, npl, box, a.ext_p != nullptr, lane, R);
    const int status = bad ? CLIP_OVERFLOW : R.status;
    nv = R.nv;
    CandRec* rec = a.rec + q;
    if (act && a.dbg && sl == 0)
    {
        uint32_t* d = a.dbg + (size_t)q * 8;
        d[0] = 0; d[1] = 0; d[2] = 0; d[3] = 0;
        d[4] = R.seq_cuts; d[5] = R.n_cuts; d[6] = (uint32_t)nv_in; d[7] = (uint32_t)npl;
    }
    const bool wr = act && status == CLIP_OK && nv > 0;
    __syncwarp();
    // result blob, renumbered to the reference's final order (rank in the live mask), as fast_pair writes it:
    // float4 verts[64] | u16 ring_start[64] | u8 ring[packed]
    const unsigned long long blob = (unsigned long long)q * FAST_BLOB_BYTES;
    unsigned char* b = a.scratch1 + blob;
    float4* bv = reinterpret_cast<float4*>(b);
    uint16_t* bo = reinterpret_cast<uint16_t*>(b + 64 * 16);
    uint8_t* br = b + 64 * 18;
    const int hw = wr ? R.hi : 0;
    const int himax = max(hw, __shfl_xor_sync(FULL, hw, DUO_L));
#pragma unroll
    for (int g = 0; g < DUO_G; g++)
    {
        const int v = sl + DUO_L * g;
        if ((g == 0 || himax > DUO_L * g) && wr && duo_own(R.live, g, sl)) sp.id[v] = (uint8_t)mrank<2>(R.live, v);
    }
    __syncwarp();
    int ne = 0;
#pragma unroll
    for (int g = 0; g < DUO_G; g++)
    {
        if (g == 0 || himax > DUO_L * g)   // warp-uniform
        {
            const int v = sl + DUO_L * g;
            const bool lv = wr && duo_own(R.live, g, sl);
            const u64 rw = lv ? sp.ring[v] : ~0ull;
            const int d = rdeg(rw);
            int inc = d;
#pragma unroll
            for (int o = 1; o < DUO_L; o <<= 1)
            {
                const int t = __shfl_up_sync(FULL, inc, o, DUO_L);
                if (sl >= o) inc += t;
            }
            const int off = ne + inc - d;
            ne += __shfl_sync(FULL, inc, DUO_L - 1, DUO_L);
            if (lv)
            {
                const int t = sp.id[v];
                bv[t] = make_float4(sp.x[v], sp.y[v], sp.z[v], 0.f);
                bo[t] = (uint16_t)off;
#pragma unroll 1
                for (int j = 0; j < d; j++) br[off + j] = sp.id[rget(rw, j)];
            }
            __syncwarp();
        }
    }
    if (wr && sl == 0)
    {
        rec->nv = (uint32_t)nv;
        rec->ne = (uint32_t)ne;
        rec->nf = 0;
        rec->tier = 1;
        rec->blob = blob;
    }
    if (act && R.seq_cuts && sl == 0) atomicAdd(&a.ctl->n_seq_cuts, R.seq_cuts);
    if (act && sl == 0)
    {
        if (status != CLIP_OK)
        {
            // too large for this tier (or a ring outgrew 8 slots): hand the pair on
            rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 0;
            if (nv_in > T2_CAP) a.ovf3_list[atomicAdd(&a.ctl->n_ovf3, 1u)] = q;
            else if (a.skip_tier1b) a.ovf2_list[atomicAdd(&a.ctl->n_ovf2, 1u)] = q;
            else a.ovf_list[atomicAdd(&a.ctl->n_ovf, 1u)] = q;
        }
        else if (nv == 0) { rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 1; }
    }
}

// Resident warps, two candidates per warp and ticket (see clip_fast_kernel for the launch shape).
__global__ void __launch_bounds__(FAST_PERSIST_WARPS * 32, 32 / FAST_PERSIST_WARPS) clip_duo_kernel(ClipArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ FastPoly<2> s_poly[FAST_PERSIST_WARPS][2];
    const int lane = threadIdx.x & 31;
    FastPoly<2>& sp = s_poly[threadIdx.x >> 5][lane >> 4];
    const unsigned long long wid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long n_items = a.ctl->n_cand;
    if (n_items > a.cap_cand) n_items = a.cap_cand;
    const unsigned nw = (gridDim.x * blockDim.x) >> 5;
    const bool more = 2ull * nw < n_items;
    unsigned long long q0 = 2ull * wid;
    while (q0 < n_items)
    {
        unsigned next = 0x7fffffffu;
        if (more && lane == 0) next = nw + atomicAdd(&a.ctl->k3_ticket, 1u);
        const unsigned long long q = q0 + (unsigned)(lane >> 4);
        duo_pair(a, sp, (uint32_t)q, q < n_items, lane);
        __syncwarp();
        q0 = 2ull * __shfl_sync(FULL, next, 0);
    }
}

// K3, small tier, round-1 version (kept for A/B profiles: SURTR_K3=sub selects it): L lanes per candidate pair
// (clip_sub.cuh), one pair per sub-warp, no persistent loop.
constexpr int FAST_LANES = 32;   // lanes per pair; 16 (two pairs per warp in lock step) is correct but measured slower, see DESIGN.md section 7
constexpr size_t FAST_BLOB = FAST_BLOB_BYTES;

template <int L>
__global__ void __launch_bounds__(FAST_WARPS * 32, L == 32 ? 32 / FAST_WARPS : 4) clip_sub_kernel(ClipArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    constexpr int G = Sub<L>::G;
    constexpr int PAIRS = FAST_WARPS * 32 / L;
    __shared__ SubPoly s_poly[PAIRS];
    SubPoly& sp = s_poly[threadIdx.x / L];
    const Sub<L> sub(threadIdx.x & 31);
    const unsigned long long q64 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) / L;
    unsigned long long n_items = a.ctl->n_cand;
    if (n_items > a.cap_cand) n_items = a.cap_cand;
    // All lanes of a warp stay in the kernel to the end (the pairs of a warp run in lock step); a sub-warp without
    // a pair, or whose pair is finished, just takes part in the collectives.
    const bool have = q64 < n_items;
    const uint32_t q = have ? (uint32_t)q64 : 0u;
    const long long t0 = a.dbg ? clock64() : 0;

    int nv = 0, npl = 0;
    uint32_t pl0 = 0;
    bool bad = false;
    if (have)
    {
        const uint2 pr = a.cand[q];
        const uint32_t v0 = a.p_vert_off[pr.x];
        nv = (int)(a.p_vert_off[pr.x + 1] - v0);
        pl0 = a.c_plane_off[pr.y];
        npl = (int)(a.c_plane_off[pr.y + 1] - pl0);
        bad = nv > 64;
        if (!bad)
        {
            for (int v = sub.sl; v < nv; v += L)
            {
                const float4 p = __ldg(a.p_verts + v0 + v);
                const uint32_t r0 = a.p_ring_off[v0 + v];
                const int d = (int)(a.p_ring_off[v0 + v + 1] - r0);
                sp.x[v] = p.x; sp.y[v] = p.y; sp.z[v] = p.z;
                u64 rw = ~0ull;
                if (d > 8 || d == 0) bad = true;
                else
                    for (int j = 0; j < d; j++)
                    {
                        const int idx = a.p_ring[r0 + j];
                        bad = bad || idx >= nv;
                        rw = rset(rw, j, idx);
                    }
                sp.ring[v] = rw;
            }
        }
    }
    bad = sub.ballot(bad) != 0u;
    sub.sync();
    const long long t1 = a.dbg ? clock64() : 0;
    const int nv_in = nv;
    unsigned seq_cuts = 0, n_cuts = 0;
    CutState cs;
    int status = sub_clip_by_planes<L>(sp, cs, nv, a.c_planes + pl0, npl, sub, have && !bad, seq_cuts, n_cuts);
    if (bad) status = CLIP_OVERFLOW;
    const long long t2 = a.dbg ? clock64() : 0;
    if (a.dbg && have && sub.sl == 0)
    {
        uint32_t* d = a.dbg + (size_t)q * 8;
        d[0] = (uint32_t)(t1 - t0); d[1] = (uint32_t)(t2 - t1); d[2] = 0; d[3] = 0;
        d[4] = seq_cuts; d[5] = n_cuts; d[6] = (uint32_t)nv_in; d[7] = (uint32_t)npl;
    }
    if (seq_cuts && sub.sl == 0) atomicAdd(&a.ctl->n_seq_cuts, seq_cuts);
    CandRec* rec = a.rec + q;
    if (have && status != CLIP_OK && sub.sl == 0)
    {
        // too large for this tier (or a ring outgrew 8 slots): queue the pair for the large tier, or straight for the
        // global-memory tier when the piece cannot fit the large tier's 256 slots either
        rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 0;
        if (nv_in > T2_CAP) a.ovf3_list[atomicAdd(&a.ctl->n_ovf3, 1u)] = q;
        else a.ovf2_list[atomicAdd(&a.ctl->n_ovf2, 1u)] = q;
    }
    if (have && status == CLIP_OK && nv == 0 && sub.sl == 0) { rec->nv = 0; rec->ne = 0; rec->nf = 0; rec->tier = 1; }
    const bool has = have && status == CLIP_OK && nv > 0;
    if (!sub.any_warp(has)) return;
    const long long t3 = a.dbg ? clock64() : 0;
    // Face count and moments of a small-tier fragment are computed by K4's gather (assemble_gather_kernel) from the
    // blob written below: the clip loop and the moments code together do not fit the SM's instruction cache, and with
    // 28 warps per SM each in a different phase they kept evicting each other (profiles/README.md).

    // result blob, renumbered to the reference's final order (rank in the live mask):
    // float4 verts[64] | u16 ring_start[64] | u8 ring[packed]
    const unsigned long long blob = (unsigned long long)q * FAST_BLOB;
    unsigned char* b = a.scratch + blob;
    float4* bv = reinterpret_cast<float4*>(b);
    uint16_t* bo = reinterpret_cast<uint16_t*>(b + 64 * 16);
    uint8_t* br = b + 64 * 18;
    const int gmax = L == 32 ? (cs.hi + L - 1) / L : sub.max_warp(has ? (cs.hi + L - 1) / L : 0);
    // final number of every live slot once (rank in the live mask), looked up per ring entry below
#pragma unroll
    for (int g = 0; g < G; g++)
    {
        const int v = sub.sl + L * g;
        if (g < gmax && has && bit64(cs.live, v)) sp.id[v] = (uint8_t)rank64(cs.live, v);
    }
    sub.sync();
    int ne = 0;
#pragma unroll 1
    for (int g = 0; g < G; g++)
    {
        if (g < gmax)
        {
            const int v = sub.sl + L * g;
            const bool live = has && bit64(cs.live, v);
            const u64 rw = live ? sp.ring[v] : ~0ull;
            const int d = rdeg(rw);
            int tot;
            const int off = ne + sub.exscan(d, tot);
            ne += tot;
            if (live)
            {
                const int t = sp.id[v];
                bv[t] = make_float4(sp.x[v], sp.y[v], sp.z[v], 0.f);
                bo[t] = (uint16_t)off;
#pragma unroll 1
                for (int j = 0; j < d; j++) br[off + j] = sp.id[rget(rw, j)];
            }
        }
    }
    if (has && sub.sl == 0)
    {
        rec->nv = (uint32_t)nv;
        rec->ne = (uint32_t)ne;
        rec->nf = 0;
        rec->tier = 1;
        rec->blob = blob;
        if (a.dbg)
        {
            a.dbg[(size_t)q * 8 + 2] = (uint32_t)(t3 - t2);
            a.dbg[(size_t)q * 8 + 3] = (uint32_t)(clock64() - t3);
        }
    }
}

// ---------------------------------------------------------------------------------------------- K4
// Fragment assembly = ordered compaction of the non-empty clip results.  Two launches: a scan over the candidate
// records (decoupled look-back, one thread per candidate) that hands every non-empty candidate its fragment
// index, first vertex and first ring entry, then a gather with one warp per candidate.
struct AssembleArgs
{
    const uint2* cand;
    const CandRec* rec;
    uint64_t cap_cand;
    const unsigned char* scratch1;
    const unsigned char* scratch2;
    const unsigned char* scratch3;
    int cap1, cap2, cap3;       // vertex capacity (blob layout) of tiers 1 / 2 / 3
    ScanState<3> st;
    Ctl* ctl;
    uint4* out_off;             // per candidate: fragment index, first vertex, first ring entry
    uint32_t* frag_cand;        // per fragment: its candidate (the gather runs over fragments, not candidates)
    surtr_fragment* f_rec;
    float4* f_verts;
    uint32_t* f_ring_off;
    uint16_t* f_ring;
    uint64_t cap_frag, cap_fverts, cap_fring;
};

constexpr int AS_THREADS = 256;
__global__ void __launch_bounds__(AS_THREADS) assemble_scan_kernel(AssembleArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_tile;
    __shared__ unsigned int s_w[AS_THREADS / 32][3];
    __shared__ unsigned long long s_excl[3];
    unsigned long long n_cand = a.ctl->n_cand;
    if (n_cand > a.cap_cand) n_cand = a.cap_cand;
    const unsigned long long n_tiles = (n_cand + AS_THREADS - 1) / AS_THREADS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (true)
    {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&a.ctl->tile_b, 1u);
        __syncthreads();
        const int tile = s_tile;
        if ((unsigned long long)tile >= n_tiles) break;
        const unsigned long long q = (unsigned long long)tile * AS_THREADS + threadIdx.x;
        uint32_t nv = 0, ne = 0;
        if (q < n_cand) { nv = a.rec[q].nv; ne = a.rec[q].ne; }
        const int has = nv > 0;
        int t0, t1, t2;
        const int e0 = warp_exscan(has, lane, t0);
        const int e1 = warp_exscan((int)nv, lane, t1);
        const int e2 = warp_exscan((int)ne, lane, t2);
        if (lane == 31) { s_w[warp][0] = t0; s_w[warp][1] = t1; s_w[warp][2] = t2; }
        __syncthreads();
        unsigned int bex[3] = { 0, 0, 0 }, btot[3] = { 0, 0, 0 };
#pragma unroll
        for (int i = 0; i < AS_THREADS / 32; i++)
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                if (i < warp) bex[k] += s_w[i][k];
                btot[k] += s_w[i][k];
            }
        if (warp == 0)
        {
            unsigned long long agg[3] = { btot[0], btot[1], btot[2] }, ex[3];
            tile_lookback<3>(a.st, tile, agg, ex, lane);
            if (lane == 0)
            {
                s_excl[0] = ex[0]; s_excl[1] = ex[1]; s_excl[2] = ex[2];
                if ((unsigned long long)tile + 1 == n_tiles)
                {
                    a.ctl->n_frag = ex[0] + btot[0];
                    a.ctl->n_fverts = ex[1] + btot[1];
                    a.ctl->n_fring = ex[2] + btot[2];
                    if (ex[1] + btot[1] <= a.cap_fverts) a.f_ring_off[ex[1] + btot[1]] = (uint32_t)(ex[2] + btot[2]);
                }
            }
        }
        __syncthreads();
        if (q < n_cand)
        {
            const unsigned long long fi = s_excl[0] + bex[0] + e0;
            a.out_off[q] = make_uint4((uint32_t)fi, (uint32_t)(s_excl[1] + bex[1] + e1), (uint32_t)(s_excl[2] + bex[2] + e2), 0u);
            if (has && fi < a.cap_frag) a.frag_cand[fi] = (uint32_t)q;
        }
    }
}

// L lanes per candidate (L = 16: two candidates per warp, in lock step -- every candidate that reaches the moments is a
// live fragment of similar size, so the lock step costs little and the face walks of sub_fragment_moments use twice the
// lanes).  All lanes stay to the end: the collectives of the moments use the full warp mask.
constexpr int GATHER_LANES = 16;
constexpr int GATHER_THREADS = 128;
// Staging of the blob (north star: TMA versus plain loads, by measurement): a variant that handed the positions of a
// fragment to the TMA unit -- cp.async.bulk global -> shared with an mbarrier, then shared -> global as a bulk group, one
// elected lane per fragment -- was built (commit 4cd4327, SURTR_K4_BULK=1) and measured against the plain LDG.128 / STS /
// STG.128 path below: 1.896 vs 1.597 ms on a 256-event config-4 batch, 73.8 vs 65.5 us on config 3, equal on config 2
// (profiles/r2_staging_ab.md).  The copy is 0.2-0.4 KB per fragment: the mbarrier round trip costs more than the loads it
// replaces, so the plain path ships.
template <int L>
__global__ void __launch_bounds__(GATHER_THREADS, 8) assemble_gather_kernel(AssembleArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ MomPoly2 s_poly[GATHER_THREADS / L];   // a small-tier fragment is rebuilt here for its face count and moments
    const Sub<L> sub(threadIdx.x & 31);
    const int lane = sub.sl;
    MomPoly2& sp = s_poly[threadIdx.x / L];
    unsigned long long n_frag = a.ctl->n_frag;
    if (n_frag > a.cap_frag) n_frag = a.cap_frag;
    // One sub-warp per FRAGMENT (the scan listed the candidates that produced one): every sub-warp has work, and the
    // fragments that share a warp are neighbours of similar size.  The grid is capped at four waves of blocks (8 resident
    // per SM) and strides over the fragments: the fragment count is only known on the device, and a grid sized by capacity
    // launched two empty blocks for every useful one (1.34 -> 1.26 ms on a 256-event config-4 batch); a single resident
    // wave was slower on small events (config 3: 57 -> 64 us -- all warps in the same phase at the same time).
    // The trip count is warp-uniform (the collectives below use the full mask).
    const unsigned long long per_warp = 32 / L, stride = (unsigned long long)gridDim.x * (GATHER_THREADS / L);
    for (unsigned long long fw = ((unsigned long long)blockIdx.x * (GATHER_THREADS / 32) + (threadIdx.x >> 5)) * per_warp; fw < n_frag; fw += stride)
    {
    const unsigned long long f = fw + (threadIdx.x & 31) / L;
    bool have = f < n_frag;
    const unsigned long long q = have ? a.frag_cand[f] : 0ull;
    const CandRec* r = a.rec + q;
    const int cnv = have ? (int)r->nv : 0;
    const int cne = have ? (int)r->ne : 0;
    have = have && cnv > 0;
    unsigned long long cfi = 0, cvb = 0, crb = 0;
    if (have)
    {
        const uint4 off = a.out_off[q];
        cfi = off.x; cvb = off.y; crb = off.z;
        if (cfi >= a.cap_frag || cvb + cnv > a.cap_fverts || crb + cne > a.cap_fring) have = false;   // the host grows and re-runs
    }
    const int tier = have ? (int)r->tier : 0;
    if (have && tier == 3)
    {
        const unsigned char* b = a.scratch3 + r->blob;
        const float4* bv = reinterpret_cast<const float4*>(b);
        const uint32_t* bo = reinterpret_cast<const uint32_t*>(b + (size_t)a.cap3 * 16);
        const uint16_t* br = reinterpret_cast<const uint16_t*>(b + (size_t)a.cap3 * 20);
        for (int v = lane; v < cnv; v += L)
        {
            a.f_verts[cvb + v] = bv[v];
            a.f_ring_off[cvb + v] = (uint32_t)(crb + bo[v]);
        }
        for (int k = lane; k < cne; k += L) a.f_ring[crb + k] = br[k];
    }
    else if (have && tier != 1)
    {
        const unsigned char* b = a.scratch2 + r->blob;
        const float4* bv = reinterpret_cast<const float4*>(b);
        const uint16_t* bo = reinterpret_cast<const uint16_t*>(b + (size_t)a.cap2 * 16);
        const uint16_t* br = reinterpret_cast<const uint16_t*>(b + (size_t)a.cap2 * 18);
        for (int v = lane; v < cnv; v += L)
        {
            a.f_verts[cvb + v] = bv[v];
            a.f_ring_off[cvb + v] = (uint32_t)(crb + bo[v]);
        }
        for (int k = lane; k < cne; k += L) a.f_ring[crb + k] = br[k];
    }
    // Small tier: copy out, and at the same time put positions and ring words back into shared memory, numbered as in the
    // result (live slots = 0..nv-1), for Poly::ExtractFaces' count + Poly::Moments + inertia (Poly.cpp:55-126).  The ring
    // bytes are one contiguous stream: copied out coalesced and staged in shared memory (the edge table's space), from
    // where every lane assembles its ring words -- not one global byte load per ring entry.
    const bool t1 = have && tier == 1;
    int r0v[64 / L], r1v[64 / L];
    if (t1)
    {
        const unsigned char* b = a.scratch1 + r->blob;
        const float4* bv = reinterpret_cast<const float4*>(b);
        const uint16_t* bo = reinterpret_cast<const uint16_t*>(b + (size_t)a.cap1 * 16);
        const uint8_t* br = b + (size_t)a.cap1 * 18;
        uint8_t* stage = reinterpret_cast<uint8_t*>(sp.en);
        for (int k = lane; k < cne; k += L)
        {
            const uint8_t x = br[k];
            a.f_ring[crb + k] = x;
            if (k < 512) stage[k] = x;
        }
#pragma unroll
        for (int g = 0; g < 64 / L; g++)
        {
            const int v = lane + L * g;
            r0v[g] = r1v[g] = 0;
            if (v < cnv)
            {
                const float4 p = bv[v];
                r0v[g] = bo[v];
                r1v[g] = v + 1 < cnv ? (int)bo[v + 1] : cne;
                a.f_verts[cvb + v] = p;
                a.f_ring_off[cvb + v] = (uint32_t)(crb + r0v[g]);
                sp.x[v] = p.x; sp.y[v] = p.y; sp.z[v] = p.z;
                sp.estart[v] = (uint16_t)r0v[g];
            }
        }
    }
    sub.sync();   // (all lanes of the warp: the staged bytes are visible)
    if (t1)
    {
        const uint32_t* stage32 = reinterpret_cast<const uint32_t*>(sp.en);
#pragma unroll
        for (int g = 0; g < 64 / L; g++)
        {
            const int v = lane + L * g;
            if (v < cnv)
            {
                // eight bytes from byte offset o (unaligned): three aligned words through the funnel shifter
                const int d = min(max(r1v[g] - r0v[g], 0), 8);
                const unsigned o = (unsigned)min(r0v[g], 511);
                const uint32_t w0 = stage32[o >> 2], w1 = stage32[(o >> 2) + 1], w2 = stage32[(o >> 2) + 2];
                const unsigned sh = (o & 3u) * 8u;
                u64 rw = (u64)__funnelshift_r(w0, w1, sh) | ((u64)__funnelshift_r(w1, w2, sh) << 32);
                if (d < 8) rw |= ~0ull << (8 * d);
                sp.ring[v] = rw;
            }
        }
        // (the sub.sync() before the moments orders these reads before phase 1 rewrites the edge table)
    }
    const bool do_mo = have && tier == 1;
    Moments mo;
    if (sub.any_warp(do_mo))
    {
        sub.sync();
        sub_fragment_moments2<L>(sp, do_mo ? cnv : 0, sub, do_mo, mo);
    }
    if (have && lane == 0)
    {
        const uint2 pr = a.cand[q];
        surtr_fragment f;
        f.cell = pr.y; f.piece = pr.x;
        f.vert_off = (uint32_t)cvb;
        f.n_verts = (uint16_t)cnv;
        if (!do_mo && cnv > 64) atomicMax(&a.ctl->max_big_verts, (unsigned)cnv);
        if (do_mo)
        {
            f.n_faces = (uint16_t)mo.n_faces;
            f.volume = mo.volume;
            f.centroid[0] = mo.cx; f.centroid[1] = mo.cy; f.centroid[2] = mo.cz;
#pragma unroll
            for (int k = 0; k < 6; k++) f.inertia[k] = mo.inertia[k];
        }
        else
        {
            f.n_faces = (uint16_t)r->nf;
            f.volume = r->volume;
            f.centroid[0] = r->centroid[0]; f.centroid[1] = r->centroid[1]; f.centroid[2] = r->centroid[2];
#pragma unroll
            for (int k = 0; k < 6; k++) f.inertia[k] = r->inertia[k];
        }
        f.n_ring = (uint32_t)cne;
        a.f_rec[cfi] = f;
    }
    __syncwarp();   // the workspace is reused by the next fragment
    }
}

// Last kernel of an event: the counters go to the host through mapped pinned memory (no copy-engine queueing behind
// other contexts' downloads) and the control block -- counters, tile tickets, scan flags -- is zeroed for the next event.
__global__ void __launch_bounds__(256) finish_event_kernel(const Ctl* __restrict__ ctl, Ctl* host_ctl, uint4* zero, unsigned n16)
{
    __shared__ Ctl s;
    if (threadIdx.x == 0) s = *ctl;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        *host_ctl = s;
        __threadfence_system();
    }
    for (unsigned i = threadIdx.x; i < n16; i += blockDim.x) zero[i] = make_uint4(0u, 0u, 0u, 0u);
}

// closes the per-vertex ring offsets (ring_off[n_fverts] = n_fring) and the piece tables used by recursion
__global__ void finish_offsets_kernel(const Ctl* ctl, uint32_t* f_ring_off, uint64_t cap_fverts)
{
    if (threadIdx.x == 0 && blockIdx.x == 0 && ctl->n_fverts <= cap_fverts)
        f_ring_off[ctl->n_fverts] = (uint32_t)ctl->n_fring;
}

// fragment records -> piece vert_off table (surtr_fragments_to_pieces)
__global__ void fragments_vert_off_kernel(const surtr_fragment* f, uint32_t n, uint32_t n_verts, uint32_t* vert_off)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vert_off[i] = f[i].vert_off;
    if (i == n) vert_off[n] = n_verts;
}

// surtr_fragments_to_pieces_per_event: first fragment of every event.  The fragment records are in (event, cell,
// piece) order and the events own consecutive cell ranges, so record i belongs to event e iff ev_cell_off[e] <=
// f[i].cell < ev_cell_off[e + 1]: one lower bound per event over the records' cell ids (strided 64-byte reads, log2 n
// steps; 4096 events x 23 steps for the deepest level of BASELINE config 5).
__global__ void event_fragment_off_kernel(const surtr_fragment* __restrict__ f, uint32_t n, const uint32_t* __restrict__ ev_cell_off,
                                          uint32_t n_events, uint32_t* __restrict__ ev_frag_off)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n_events) return;
    if (e == n_events) { ev_frag_off[e] = n; return; }
    const uint32_t c0 = ev_cell_off[e];
    uint32_t lo = 0, hi = n;
    while (lo < hi)
    {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (f[mid].cell < c0) lo = mid + 1; else hi = mid;
    }
    ev_frag_off[e] = lo;
}

// ------------------------------------------------------------------------------- Kdop::Calc(Polyhedron)
// One block per normal.  t = Dot(v, n) in the reference's float op order; strict < / > so the FIRST extremal
// vertex wins (Kdop.cpp:98-112); ties across threads are broken towards the lower vertex index.
__global__ void __launch_bounds__(256) kdop_arg_kernel(const float4* __restrict__ verts, uint32_t nv,
                                                       const float* __restrict__ normals, float* __restrict__ dist,
                                                       int32_t* __restrict__ arg, float4* __restrict__ planes)
{
    __shared__ float s_min[256], s_max[256];
    __shared__ int s_imin[256], s_imax[256];
    const int e = blockIdx.x;
    const float nx = normals[3 * e], ny = normals[3 * e + 1], nz = normals[3 * e + 2];
    float mn = 0.f, mx = 0.f;
    int imin = -1, imax = -1;
    for (uint32_t v = threadIdx.x; v < nv; v += 256)
    {
        const float4 p = __ldg(verts + v);
        const float t = dot3(p.x, p.y, p.z, nx, ny, nz);
        if (imin < 0 || mn > t) { mn = t; imin = (int)v; }
        if (imax < 0 || mx < t) { mx = t; imax = (int)v; }
    }
    s_min[threadIdx.x] = mn; s_max[threadIdx.x] = mx; s_imin[threadIdx.x] = imin; s_imax[threadIdx.x] = imax;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1)
    {
        if ((int)threadIdx.x < s)
        {
            const int o = threadIdx.x + s;
            const int bi = s_imin[o], ai = s_imin[threadIdx.x];
            if (bi >= 0 && (ai < 0 || s_min[o] < s_min[threadIdx.x] || (s_min[o] == s_min[threadIdx.x] && bi < ai)))
            { s_min[threadIdx.x] = s_min[o]; s_imin[threadIdx.x] = bi; }
            const int bj = s_imax[o], aj = s_imax[threadIdx.x];
            if (bj >= 0 && (aj < 0 || s_max[o] > s_max[threadIdx.x] || (s_max[o] == s_max[threadIdx.x] && bj < aj)))
            { s_max[threadIdx.x] = s_max[o]; s_imax[threadIdx.x] = bj; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        dist[2 * e] = s_min[0]; dist[2 * e + 1] = s_max[0];
        arg[2 * e] = s_imin[0]; arg[2 * e + 1] = s_imax[0];
        if (s_imin[0] >= 0)
        {
            const float4 a = verts[s_imin[0]], b = verts[s_imax[0]];
            planes[2 * e] = plane_from_point_normal(a.x, a.y, a.z, -nx, -ny, -nz);   // Plane(vert, -Normal)
            planes[2 * e + 1] = plane_from_point_normal(b.x, b.y, b.z, nx, ny, nz);  // Plane(vert, Normal)
        }
    }
}

// Batched Kdop::Calc(Polyhedron): one warp per (object, normal).  normal_obj[e] = object of normal e.
__global__ void __launch_bounds__(256) kdop_arg_batch_kernel(const float4* __restrict__ verts, const uint32_t* __restrict__ vert_off,
                                                             const float* __restrict__ normals, const uint32_t* __restrict__ normal_obj,
                                                             uint32_t n_normals, float* __restrict__ dist, int32_t* __restrict__ arg,
                                                             float4* __restrict__ planes)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t e = warp; e < n_normals; e += nwarps)
    {
        const uint32_t o = normal_obj[e];
        const uint32_t v0 = vert_off[o], nv = vert_off[o + 1] - v0;
        const float nx = normals[3 * e], ny = normals[3 * e + 1], nz = normals[3 * e + 2];
        float mn = 0.f, mx = 0.f;
        int imin = -1, imax = -1;
        for (uint32_t v = lane; v < nv; v += 32)
        {
            const float4 p = __ldg(verts + v0 + v);
            const float t = dot3(p.x, p.y, p.z, nx, ny, nz);
            if (imin < 0 || mn > t) { mn = t; imin = (int)v; }
            if (imax < 0 || mx < t) { mx = t; imax = (int)v; }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1)
        {
            const float omn = __shfl_xor_sync(FULL, mn, s), omx = __shfl_xor_sync(FULL, mx, s);
            const int oimin = __shfl_xor_sync(FULL, imin, s), oimax = __shfl_xor_sync(FULL, imax, s);
            if (oimin >= 0 && (imin < 0 || omn < mn || (omn == mn && oimin < imin))) { mn = omn; imin = oimin; }
            if (oimax >= 0 && (imax < 0 || omx > mx || (omx == mx && oimax < imax))) { mx = omx; imax = oimax; }
        }
        if (lane == 0)
        {
            dist[2 * e] = mn; dist[2 * e + 1] = mx;
            arg[2 * e] = imin; arg[2 * e + 1] = imax;
            if (imin >= 0)
            {
                const float4 a = verts[v0 + imin], b = verts[v0 + imax];
                planes[2 * e] = plane_from_point_normal(a.x, a.y, a.z, -nx, -ny, -nz);   // Plane(vert, -Normal)
                planes[2 * e + 1] = plane_from_point_normal(b.x, b.y, b.z, nx, ny, nz);  // Plane(vert, Normal)
            }
            else
            {
                planes[2 * e] = planes[2 * e + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}
} // namespace surtr
