// clip_global.cuh -- the large and the unbounded tier of K3: one warp or one block per pair, polyhedron in a workspace.
//
// Algorithm and exactness argument: see clip_warp.cuh and DESIGN.md section 5.  Written with strided (rolled) loops
// over arrays in memory instead of per-lane register arrays, so the vertex count is bounded only by the workspace
// and the code stays small (an unrolled register-array version of the large tier was 179 KB of SASS and stalled on
// instruction fetch).  The workspace (117 bytes per vertex slot) is reached through plain pointers:
//   * clip_shared_kernel carves it from shared memory (256 slots): pieces / intermediate results beyond the small
//     tier's 64 slots or ring degree 8 -- ACH-sized convex pieces, mesh fragments;
//   * clip_global_kernel carves it from global memory (L2-resident): non-convex piece MESHES (2.5K-5K vertices, ring
//     degree up to 13; m_fractureTask's second clip, Surtr.cpp:1470) and anything beyond 256 slots.
#pragma once

#include "clip_warp.cuh"

namespace surtr
{
constexpr int GD = 16;                  // ring slots per vertex in the large on-chip tier (shared-memory workspace)
constexpr uint16_t G_NONE = 0xffffu;    // the reference's "-1" ring mark
constexpr int8_t G_GONE = -2;           // comp of a slot whose vertex left the polyhedron in an EARLIER cut (lazy compaction)

struct GlobalPoly   // views into one group's workspace
{
    float *x, *y, *z;
    uint16_t *ring, *old_ring;   // gd per vertex (old_ring: snapshot of the sequential replay, and the target of a compaction)
    uint16_t *deg, *old_deg;
    int8_t* comp;
    uint16_t* id;                // renumbering / walk-target probe / face-start masks
    uint32_t* list;              // straddling half-edges (v | slot << 16), then walk targets; triangle bases
    float4* tri;                 // 2 per vertex slot: ordered fan-triangle records
    int cap;                     // vertex slots
    int gd;                      // ring slots per vertex: GD in the on-chip tier, whatever the largest ring needs in the global-memory tier
};

__host__ __device__ constexpr size_t global_poly_bytes(size_t cap, size_t gd = GD)
{
    return cap * (3 * 4 + 2 * gd * 2 + 2 * 2 + 1 + 2 + 4 + 2 * 16) + 256;
}

__device__ inline GlobalPoly global_poly_carve(unsigned char* base, int cap, int gd = GD)
{
    GlobalPoly g;
    g.cap = cap;
    g.gd = gd;
    unsigned char* p = base;
    g.tri = reinterpret_cast<float4*>(p); p += (size_t)cap * 32;
    g.x = reinterpret_cast<float*>(p); p += (size_t)cap * 4;
    g.y = reinterpret_cast<float*>(p); p += (size_t)cap * 4;
    g.z = reinterpret_cast<float*>(p); p += (size_t)cap * 4;
    g.list = reinterpret_cast<uint32_t*>(p); p += (size_t)cap * 4;
    g.ring = reinterpret_cast<uint16_t*>(p); p += (size_t)cap * gd * 2;
    g.old_ring = reinterpret_cast<uint16_t*>(p); p += (size_t)cap * gd * 2;
    g.id = reinterpret_cast<uint16_t*>(p); p += (size_t)cap * 2;
    g.deg = reinterpret_cast<uint16_t*>(p); p += (size_t)cap * 2;
    g.old_deg = reinterpret_cast<uint16_t*>(p); p += (size_t)cap * 2;
    g.comp = reinterpret_cast<int8_t*>(p);
    return g;
}

// FaceLoop (Src/Poly.cpp:34-41)
__device__ __forceinline__ int g_face_loop(const GlobalPoly& g, int v, int vprev)
{
    const uint16_t* r = g.ring + (size_t)v * g.gd;
    const int d = g.deg[v];
    if (d == 0) return vprev;
    int k = 0;
    while (k < d && r[k] != (uint16_t)vprev) k++;
    return k == 0 ? r[d - 1] : r[k - 1];
}

// Sequential replay of Poly.cpp:365-462 by lane 0 (patch, erase marks, splice) over the slots [0, nverts): the new
// vertices [nverts0, nverts) first, then the older ones ascending (the reference's visiting order); slots of vertices that
// left in earlier cuts (comp == G_GONE) are never touched.  false = a ring ran out of slots.
__device__ __noinline__ bool g_seq_patch_splice(GlobalPoly& g, int nverts0, int nverts, int nlive)
{
    const size_t GS = (size_t)g.gd;
    for (int ii = 0; ii < nverts; ii++)
    {
        const int i = (ii + nverts0) % nverts;
        const int ci = g.comp[i];
        if (!(ci == 0 || ci == 2)) continue;
        const int nneigh = g.deg[i];
        for (int j = 0; j < nneigh; j++)
        {
            const uint16_t jn = g.ring[(size_t)i * GS + j];
            if (jn == G_NONE || g.comp[jn] != -1) continue;
            int iprev = i, inext = jn, itmp, k = 0;
            while (g.comp[inext] == -1 && k++ < nlive)
            {
                itmp = inext;
                inext = g_face_loop(g, inext, iprev);
                iprev = itmp;
            }
            if (g.ring[(size_t)i * GS + (j + 1) % g.deg[i]] == (uint16_t)inext || inext == i)
            {
                g.ring[(size_t)i * GS + j] = G_NONE;
            }
            else
            {
                g.ring[(size_t)i * GS + j] = (uint16_t)inext;
                const int dn = g.deg[inext], od = g.old_deg[inext];
                if (dn >= g.gd || od >= g.gd) return false;
                uint16_t* rn = g.ring + (size_t)inext * GS;
                uint16_t* on = g.old_ring + (size_t)inext * GS;
                int off = 0;
                uint16_t mark = (uint16_t)i;
                if (g.comp[inext] == 2) mark = G_NONE;   // Poly.cpp:409 inserts -1 into the snapshot
                else while (off < od && on[off] != (uint16_t)iprev) off++;
                for (int q = dn; q > off; q--) rn[q] = rn[q - 1];
                rn[off] = (uint16_t)i;
                g.deg[inext] = (uint16_t)(dn + 1);
                for (int q = od; q > off; q--) on[q] = on[q - 1];
                on[off] = mark;
                g.old_deg[inext] = (uint16_t)(od + 1);
            }
        }
    }
    for (int i = 0; i < nverts; i++)   // Poly.cpp:426-431
    {
        if (g.comp[i] == G_GONE) continue;
        uint16_t* r = g.ring + (size_t)i * GS;
        int w = 0;
        const int d = g.deg[i];
        for (int k = 0; k < d; k++)
            if (r[k] != G_NONE) r[w++] = r[k];
        g.deg[i] = (uint16_t)w;
    }
    bool updated = true;   // Poly.cpp:433-462
    while (updated)
    {
        updated = false;
        for (int i = 0; i < nverts; i++)
        {
            if (g.comp[i] >= 0 && g.deg[i] == 2)
            {
                updated = true;
                const int iprev = g.ring[(size_t)i * GS], inext = g.ring[(size_t)i * GS + 1];
                int k = 0;
                while (k < g.deg[iprev] && g.ring[(size_t)iprev * GS + k] != (uint16_t)i) ++k;
                if (k < g.deg[iprev]) g.ring[(size_t)iprev * GS + k] = (uint16_t)inext;
                k = 0;
                while (k < g.deg[inext] && g.ring[(size_t)inext * GS + k] != (uint16_t)i) ++k;
                if (k < g.deg[inext]) g.ring[(size_t)inext * GS + k] = (uint16_t)iprev;
                g.comp[i] = -1;
            }
        }
    }
    return true;
}

// The NW warps that work on one pair: one warp (large tier, workspace in shared memory, several pairs per block) or a
// whole block of NW warps (unbounded tier, workspace in global memory, one pair per block -- a 2500-vertex mesh is 79
// strided iterations per phase for one warp, 10 for eight).  All group-wide decisions are uniform, so every thread of
// the group takes the same branches and meets the same barriers.
template <int NW>
struct Grp
{
    static constexpr int N = NW * 32;
    int tid, lane;
    int* scratch;   // NW > 1: shared int[NW + 1]

    __device__ __forceinline__ void sync() const
    {
        if (NW == 1) __syncwarp(); else __syncthreads();
    }
    __device__ __forceinline__ bool any(bool p) const
    {
        if (NW == 1) return __ballot_sync(FULL, p) != 0u;
        return __syncthreads_or(p ? 1 : 0) != 0;
    }
    __device__ __forceinline__ int exscan(int v, int& total) const   // exclusive prefix over the group, thread order
    {
        int wt;
        const int ex = warp_exscan(v, lane, wt);
        if (NW == 1) { total = wt; return ex; }
        const int w = tid >> 5;
        if (lane == 0) scratch[w] = wt;
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < NW; i++)
        {
            const int t = scratch[i];
            if (i < w) before += t;
            tot += t;
        }
        __syncthreads();
        total = tot;
        return before + ex;
    }
    __device__ __forceinline__ int bcast0(int v) const   // thread 0's value
    {
        if (NW == 1) return __shfl_sync(FULL, v, 0);
        if (tid == 0) scratch[NW] = v;
        __syncthreads();
        const int r = scratch[NW];
        __syncthreads();
        return r;
    }
};

// Every vertex in-plane: the reference's box test decides (Poly.cpp:297-299, 725-744).  Rare; the first warp does it.
template <int NW>
__device__ bool g_all_inplane_box_says_skip(const GlobalPoly& g, int nv, const float4& pl, const Grp<NW>& grp)
{
    int skip = 0;
    if (grp.tid < 32)
    {
        const int lane = grp.lane;
        float lo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
        float hi[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
        for (int v = lane; v < nv; v += 32)
        {
            if (g.comp[v] == G_GONE) continue;
            lo[0] = fminf(lo[0], g.x[v]); hi[0] = fmaxf(hi[0], g.x[v]);
            lo[1] = fminf(lo[1], g.y[v]); hi[1] = fmaxf(hi[1], g.y[v]);
            lo[2] = fminf(lo[2], g.z[v]); hi[2] = fmaxf(hi[2], g.z[v]);
        }
        for (int o = 16; o > 0; o >>= 1)
            for (int k = 0; k < 3; k++)
            {
                lo[k] = fminf(lo[k], __shfl_xor_sync(FULL, lo[k], o));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(FULL, hi[k], o));
            }
        const int k = lane & 7;
        const int c = classify(signed_dist(pl, (k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2]));
        skip = __ballot_sync(FULL, c == -1) == 0u ? 1 : 0;
    }
    return grp.bcast0(skip) != 0;
}

// Stable renumbering of the live slots to 0..n-1 (the reference's compaction, Poly.cpp:464-495): rings are rewritten
// into the OTHER ring array (old_ring, free outside the sequential replay) and the two views swap, so no thread holds a
// ring in registers and the stride may be anything.  Returns false if a live ring points at an erased vertex.
template <int NW>
__device__ bool g_compact(GlobalPoly& g, int hi, int& n_out, const Grp<NW>& grp)
{
    constexpr int N = Grp<NW>::N;
    const int tid = grp.tid;
    const size_t GS = (size_t)g.gd;
    int kept_before = 0;
    {
        // new number of a live slot = live slots below it: a contiguous run per thread, one scan over the threads
        const int chunk = (hi + N - 1) / N;
        const int vb = min(hi, tid * chunk), ve = min(hi, vb + chunk);
        int mine = 0;
        for (int v = vb; v < ve; v++) mine += g.comp[v] >= 0 ? 1 : 0;
        int rank = grp.exscan(mine, kept_before);
        for (int v = vb; v < ve; v++)
        {
            const bool live = g.comp[v] >= 0;
            g.id[v] = live ? (uint16_t)rank : G_NONE;
            rank += live ? 1 : 0;
        }
    }
    grp.sync();
    bool dangling = false;
    for (int v = tid; v < hi; v += N)   // rings: renumbered copy into the other array (no hazards: disjoint source and target)
    {
        if (g.comp[v] < 0) continue;
        const int d = g.deg[v], t = g.id[v];
        const uint16_t* src = g.ring + (size_t)v * GS;
        uint16_t* dst = g.old_ring + (size_t)t * GS;
        for (int j = 0; j < d; j++)
        {
            const uint16_t r = src[j] < hi ? g.id[src[j]] : G_NONE;
            dangling |= r == G_NONE;
            dst[j] = r;
        }
    }
    for (int base = 0; base < hi; base += N)   // positions, degrees: in place, slot t <= v, chunk by chunk in ascending order
    {
        const int v = base + tid;
        const bool live = v < hi && g.comp[v] >= 0;
        float vx = 0.f, vy = 0.f, vz = 0.f;
        int d = 0, t = 0;
        if (live) { vx = g.x[v]; vy = g.y[v]; vz = g.z[v]; d = g.deg[v]; t = g.id[v]; }
        grp.sync();
        if (live) { g.x[t] = vx; g.y[t] = vy; g.z[t] = vz; g.deg[t] = (uint16_t)d; }
        grp.sync();
    }
    uint16_t* sw = g.ring; g.ring = g.old_ring; g.old_ring = sw;
    for (int v = tid; v < kept_before; v += N) g.comp[v] = 1;
    grp.sync();
    n_out = kept_before;
    return !grp.any(dangling);
}

// Clip the polyhedron in the workspace (nv vertices in slots 0..nv-1) by planes[0..npl).  All threads of the group call
// this together.  Compaction is LAZY, as in the small tier: a clipped vertex keeps its slot (comp = G_GONE) and new
// vertices are appended; the slots are renumbered -- stably, so to exactly the reference's numbering -- only when they
// run out and once at the end.  On return the polyhedron is dense again: nv vertices in slots 0..nv-1.
template <int NW>
__device__ int global_clip_by_planes(GlobalPoly& g, int& nv, const float4* __restrict__ planes, int npl, const Grp<NW>& grp, unsigned& seq_cuts)
{
    constexpr int N = Grp<NW>::N;
    const int tid = grp.tid;
    const size_t GS = (size_t)g.gd;
    int hi = nv;          // allocated slots
    bool dense = true;    // no G_GONE slot below hi
    for (int v = tid; v < nv; v += N) g.comp[v] = 1;
    grp.sync();
    for (int kp = 0; kp < npl && nv > 0; kp++)
    {
        const float4 pl = __ldg(planes + kp);

        // classify (Poly.cpp:303-319)
        bool t_clip = false, t_keep = false, t_zero = false;
        // (unrolled a little: the iterations are independent loads -- memory-level parallelism)
#pragma unroll 4
        for (int v = tid; v < hi; v += N)
        {
            if (g.comp[v] == G_GONE) continue;
            const int c = classify(signed_dist(pl, g.x[v], g.y[v], g.z[v]));
            g.comp[v] = (int8_t)c;
            t_clip |= c == -1;
            t_keep |= c == 1;
            t_zero |= c == 0;
        }
        const bool any_keep = grp.any(t_keep);
        const bool any_clip = grp.any(t_clip);
        const bool any_zero = grp.any(t_zero);
        if (!any_keep)
        {
            if (!any_clip && g_all_inplane_box_says_skip<NW>(g, hi, pl, grp)) continue;
            nv = 0;
            break;
        }
        if (!any_clip) continue;

        // straddling half-edges in the reference's append order (vertex ascending, slot ascending).  Three passes and ONE
        // group scan (a scan per strided chunk cost two block barriers per 256 vertices: 22 per plane on the bunny mesh,
        // most of the tier's time): (1) strided, the straddle count of every clipped vertex -> id[] (free until the patch);
        // (2) blocked, every thread sums a contiguous run of counts, one scan over the threads, the run's counts become
        // list positions; (3) strided, the entries are written at their positions.
        // (the kept-neighbour test is a chain of two global loads per ring entry: four entries at a time, so that the
        // comp loads of a group are in flight together; the result is kept as a slot MASK in old_deg -- free until a
        // sequential replay -- and pass 3 expands the mask instead of walking the ring again)
        for (int v = tid; v < hi; v += N)
        {
            unsigned mask = 0u;
            int cnt = 0;
            if (g.comp[v] == -1)
            {
                const int d = g.deg[v];
                const uint16_t* r = g.ring + (size_t)v * GS;
                for (int j0 = 0; j0 < d; j0 += 4)
                {
                    // (GS is a multiple of 8: a ring row starts on a 16-byte boundary, four entries are one 8-byte load)
                    const uint2 e = *reinterpret_cast<const uint2*>(r + j0);
                    const int n0 = e.x & 0xffffu, n1 = e.x >> 16, n2 = e.y & 0xffffu, n3 = e.y >> 16;
                    const int8_t c0 = g.comp[n0];
                    const int8_t c1 = j0 + 1 < d ? g.comp[n1] : (int8_t)0;
                    const int8_t c2 = j0 + 2 < d ? g.comp[n2] : (int8_t)0;
                    const int8_t c3 = j0 + 3 < d ? g.comp[n3] : (int8_t)0;
                    const unsigned m4 = (c0 > 0 ? 1u : 0u) | (c1 > 0 ? 2u : 0u) | (c2 > 0 ? 4u : 0u) | (c3 > 0 ? 8u : 0u);
                    cnt += __popc(m4);
                    if (j0 < 16) mask |= m4 << j0;
                }
                g.old_deg[v] = (uint16_t)mask;
            }
            g.id[v] = (uint16_t)cnt;
        }
        grp.sync();
        int nnew = 0;
        bool redo = false;
        {
            const int chunk = (hi + N - 1) / N;
            const int vb = min(hi, tid * chunk), ve = min(hi, vb + chunk);
            int mine = 0;
            for (int v = vb; v < ve; v++) mine += g.id[v];
            int off = grp.exscan(mine, nnew);
            redo = hi + nnew > g.cap;   // (uniform)
            if (!redo)
                for (int v = vb; v < ve; v++)
                {
                    const int c = g.id[v];
                    g.id[v] = (uint16_t)off;
                    off += c;
                }
        }
        if (!redo)
        {
            grp.sync();
            for (int v = tid; v < hi; v += N)
            {
                if (g.comp[v] != -1) continue;
                int w = g.id[v];
                const int d = g.deg[v];
                unsigned mask = g.old_deg[v];
                while (mask) { const int j = __ffs((int)mask) - 1; mask &= mask - 1u; g.list[w++] = (uint32_t)v | ((uint32_t)j << 16); }
                for (int j = 16; j < d; j++)   // (rings wider than the mask: the global tier only)
                    if (g.comp[g.ring[(size_t)v * GS + j]] > 0) g.list[w++] = (uint32_t)v | ((uint32_t)j << 16);
            }
        }
        if (redo)
        {
            // out of slots: renumber the live vertices and apply this plane again -- unless they really do not fit
            if (dense) return CLIP_NEED_SLOTS;
            for (int v = tid; v < hi; v += N)
                if (g.comp[v] == 0 || g.comp[v] == -1) g.comp[v] = 1;   // every non-gone slot is live for the renumbering
            grp.sync();
            int n;
            if (!g_compact<NW>(g, hi, n, grp)) return CLIP_OVERFLOW;
            hi = n;
            dense = true;
            kp--;
            continue;
        }
        const int nverts0 = hi;
        const int nverts = nverts0 + nnew;
        grp.sync();
        // insert (Poly.cpp:345-354): one new vertex per thread and iteration
        for (int t = tid; t < nnew; t += N)
        {
            const uint32_t e = g.list[t];
            const int v = (int)(e & 0xffffu), j = (int)(e >> 16), w = nverts0 + t;
            const int jn = g.ring[(size_t)v * GS + j];
            const float ax = g.x[v], ay = g.y[v], az = g.z[v], bx = g.x[jn], by = g.y[jn], bz = g.z[jn];
            const float sa = signed_dist(pl, ax, ay, az), sb = signed_dist(pl, bx, by, bz);
            float ox, oy, oz;
            plane_line_intersection(ax, ay, az, sa, bx, by, bz, sb, ox, oy, oz);
            g.x[w] = ox; g.y[w] = oy; g.z[w] = oz;
            g.comp[w] = 2;
            g.deg[w] = 2;
            g.ring[(size_t)w * GS] = (uint16_t)v;
            g.ring[(size_t)w * GS + 1] = (uint16_t)jn;
            // several threads may patch the ring of the same kept vertex jn at once: each replaces only the entry holding
            // ITS clipped vertex v, and an entry another thread is rewriting (v' -> w') equals v neither before nor after
            // -- entry-disjoint by construction (compute-sanitizer racecheck warns at word level, profiles/r1_sanitizer.txt)
            uint16_t* rj = g.ring + (size_t)jn * GS;
            const int dj = g.deg[jn];
            int k = 0;
            while (k < dj && rj[k] != (uint16_t)v) k++;
            if (k < dj) rj[k] = (uint16_t)w;
            g.ring[(size_t)v * GS + j] = (uint16_t)w;
        }
        grp.sync();

        // patch (Poly.cpp:365-431)
        bool need_seq = any_zero;
        if (!need_seq)
        {
            bool ok = true;
            for (int t = tid; t < nnew; t += N)
            {
                const int w = nverts0 + t;
                int iprev = w, inext = g.ring[(size_t)w * GS], itmp, k = 0;
                while (g.comp[inext] == -1 && k++ < nverts)
                {
                    itmp = inext;
                    inext = g_face_loop(g, inext, iprev);
                    iprev = itmp;
                }
                const bool okt = g.comp[inext] == 2 && inext != w;
                if (okt) g.id[inext] = (uint16_t)w;
                g.list[t] = (uint32_t)inext;
                ok = ok && okt;
            }
            grp.sync();
            for (int t = tid; t < nnew; t += N)
                if (ok) ok = g.id[g.list[t]] == (uint16_t)(nverts0 + t);
            need_seq = grp.any(!ok);
            if (!need_seq)
            {
                for (int t = tid; t < nnew; t += N)   // ring(w) = [pusher, walked, kept]
                {
                    const int w = nverts0 + t;
                    const uint16_t kept = g.ring[(size_t)w * GS + 1];
                    g.ring[(size_t)w * GS] = g.id[w];
                    g.ring[(size_t)w * GS + 1] = (uint16_t)g.list[t];
                    g.ring[(size_t)w * GS + 2] = kept;
                    g.deg[w] = 3;
                }
            }
        }
        if (need_seq)
        {
            if (tid == 0) seq_cuts++;
            for (int v = tid; v < nverts; v += N)
            {
                if (g.comp[v] == G_GONE) continue;
                const int d = g.deg[v];
                g.old_deg[v] = (uint16_t)d;
                for (int j = 0; j < d; j++) g.old_ring[(size_t)v * GS + j] = g.ring[(size_t)v * GS + j];
            }
            grp.sync();
            int okflag = 1;
            if (tid == 0) okflag = g_seq_patch_splice(g, nverts0, nverts, nv + nnew) ? 1 : 0;
            okflag = grp.bcast0(okflag);
            if (!okflag) return CLIP_NEED_DEG;
        }
        grp.sync();

        // lazy compaction (Poly.cpp:464-499): clipped and spliced vertices only give up their slot's liveness
        int live_now = 0;
        {
            int mine = 0;
            for (int v = tid; v < nverts; v += N)
            {
                const int c = g.comp[v];
                if (c == -1) g.comp[v] = G_GONE;
                mine += c >= 0 ? 1 : 0;
            }
            grp.exscan(mine, live_now);   // (one scan for the whole polyhedron: only the total is needed)
        }
        hi = nverts;
        dense = false;
        nv = live_now < 4 ? 0 : live_now;   // Poly.cpp:498-499
        grp.sync();
    }
    if (nv > 0 && !dense)
    {
        for (int v = tid; v < hi; v += N)
            if (g.comp[v] == 0) g.comp[v] = 1;
        grp.sync();
        int n;
        if (!g_compact<NW>(g, hi, n, grp)) return CLIP_OVERFLOW;   // a live ring pointing at an erased vertex: not a polyhedron
        nv = n;
    }
    grp.sync();
    return CLIP_OK;
}

// Face count + moments in the reference's order (Poly.cpp:55-126) on the workspace; see sub_fragment_moments.
template <int NW>
__device__ void global_fragment_moments(GlobalPoly& g, int nv, const Grp<NW>& grp, Moments& out, float* fscratch)
{
    constexpr int N = Grp<NW>::N;
    const int tid = grp.tid, lane = grp.lane;
    const float ox = g.x[0], oy = g.y[0], oz = g.z[0];
    int n_faces = 0, n_tri = 0;
    // pass 1: face starts (bit mask per vertex in g.id) and the first triangle slot of every vertex (g.list)
    for (int base = 0; base < nv; base += N)
    {
        const int v = base + tid;
        int cnt = 0, faces = 0;
        unsigned mask = 0u;
        if (v < nv)
        {
            const int d = g.deg[v];
            for (int j = 0; j < d; j++)
            {
                int at = g.ring[(size_t)v * g.gd + j];
                if (at < v || (int)g.ring[(size_t)v * g.gd + (j + 1 == d ? 0 : j + 1)] < v) continue;
                int prev = v, n = 1;
                bool is_start = true;
                while (at != v)
                {
                    if (at < v || n > nv) { is_start = false; break; }
                    const int nxt = g_face_loop(g, at, prev);
                    prev = at;
                    at = nxt;
                    n++;
                }
                if (is_start) { mask |= 1u << j; faces++; cnt += max(n - 2, 0); }
            }
            g.id[v] = (uint16_t)mask;
        }
        int tot, ftot;
        const int ex = grp.exscan(cnt, tot);
        grp.exscan(faces, ftot);
        if (v < nv) g.list[v] = (uint32_t)(n_tri + ex);
        n_tri += tot;
        n_faces += ftot;
    }
    n_tri = min(n_tri, 2 * g.cap);
    grp.sync();
    float cov[10] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    for (int v = tid; v < nv; v += N)
    {
        unsigned m = g.id[v];
        if (!m) continue;
        int w = (int)g.list[v];
        const float p0x = __fsub_rn(g.x[v], ox), p0y = __fsub_rn(g.y[v], oy), p0z = __fsub_rn(g.z[v], oz);
        while (m)
        {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            int prev = v, at = g.ring[(size_t)v * g.gd + j];
            float p1x = __fsub_rn(g.x[at], ox), p1y = __fsub_rn(g.y[at], oy), p1z = __fsub_rn(g.z[at], oz);
            int nxt = g_face_loop(g, at, prev);
            prev = at;
            at = nxt;
            while (at != v)
            {
                const float p2x = __fsub_rn(g.x[at], ox), p2y = __fsub_rn(g.y[at], oy), p2z = __fsub_rn(g.z[at], oz);
                float cx, cy, cz;
                cross3(p1x, p1y, p1z, p2x, p2y, p2z, cx, cy, cz);
                const float dV = dot3(p0x, p0y, p0z, cx, cy, cz);
                const float sx = __fadd_rn(__fadd_rn(p0x, p1x), p2x);
                const float sy = __fadd_rn(__fadd_rn(p0y, p1y), p2y);
                const float sz = __fadd_rn(__fadd_rn(p0z, p1z), p2z);
                if (w < 2 * g.cap) g.tri[w] = make_float4(dV, __fmul_rn(sx, dV), __fmul_rn(sy, dV), __fmul_rn(sz, dV));
                w++;
                cov[0] += dV * (sx * sx + p0x * p0x + p1x * p1x + p2x * p2x);
                cov[1] += dV * (sy * sy + p0y * p0y + p1y * p1y + p2y * p2y);
                cov[2] += dV * (sz * sz + p0z * p0z + p1z * p1z + p2z * p2z);
                cov[3] += dV * (sx * sy + p0x * p0y + p1x * p1y + p2x * p2y);
                cov[4] += dV * (sx * sz + p0x * p0z + p1x * p1z + p2x * p2z);
                cov[5] += dV * (sy * sz + p0y * p0z + p1y * p1z + p2y * p2z);
                cov[6] += dV;
                cov[7] += dV * sx; cov[8] += dV * sy; cov[9] += dV * sz;
                p1x = p2x; p1y = p2y; p1z = p2z;
                nxt = g_face_loop(g, at, prev);
                prev = at;
                at = nxt;
            }
        }
    }
    grp.sync();
    // ordered accumulation (Poly.cpp:77-85) by the first four lanes of the group, fixed-shape reduction of the rest
    double zeroth = 0.0;
    float fsum = 0.f;
    if (tid < 4)
    {
        const float* comp = reinterpret_cast<const float*>(g.tri) + tid;
        for (int t = 0; t < n_tri; t++)
        {
            const float r = comp[4 * (size_t)t];
            zeroth += (double)r;
            fsum = __fadd_rn(fsum, r);
        }
    }
#pragma unroll
    for (int k = 0; k < 10; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            cov[k] += __shfl_xor_sync(FULL, cov[k], o);
    if (NW > 1)
    {
        // warp partials -> shared memory -> summed in warp order by every thread (same value everywhere, deterministic)
        if (lane == 0)
            for (int k = 0; k < 10; k++) fscratch[(tid >> 5) * 10 + k] = cov[k];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 10; k++)
        {
            float sum = 0.f;
            for (int w = 0; w < NW; w++) sum += fscratch[w * 10 + k];
            cov[k] = sum;
        }
        __syncthreads();
    }
    if (tid < 32)
    {
        zeroth = __shfl_sync(FULL, zeroth, 0) / 6.0;
        float fx = __shfl_sync(FULL, fsum, 1), fy = __shfl_sync(FULL, fsum, 2), fz = __shfl_sync(FULL, fsum, 3);
        {
            const double q = 24.0 * zeroth;
            const double inv = (q >= 0.0 ? 1.0 : -1.0) / fmax(1.0e-30, fabs(q));
            const float sc = (float)inv;
            fx = __fmul_rn(fx, sc); fy = __fmul_rn(fy, sc); fz = __fmul_rn(fz, sc);
        }
        out.n_faces = n_faces;
        out.volume = zeroth;
        out.cx = __fadd_rn(fx, ox); out.cy = __fadd_rn(fy, oy); out.cz = __fadd_rn(fz, oz);
        const float V = cov[6] * (1.f / 6.f);
        const float iv = V != 0.f ? 1.f / (24.f * V) : 0.f;
        const float c0 = cov[7] * iv, c1 = cov[8] * iv, c2 = cov[9] * iv;
        const float k120 = 1.f / 120.f;
        const float Cxx = cov[0] * k120 - V * c0 * c0, Cyy = cov[1] * k120 - V * c1 * c1, Czz = cov[2] * k120 - V * c2 * c2;
        out.inertia[0] = Cyy + Czz;
        out.inertia[1] = Cxx + Czz;
        out.inertia[2] = Cxx + Cyy;
        out.inertia[3] = -(cov[3] * k120 - V * c0 * c1);
        out.inertia[4] = -(cov[4] * k120 - V * c0 * c2);
        out.inertia[5] = -(cov[5] * k120 - V * c1 * c2);
    }
}
} // namespace surtr
