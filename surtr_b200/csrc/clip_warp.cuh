// clip_warp.cuh -- one warp clips one convex piece by one cell's plane list, entirely in shared memory.
//
// Restates Poly::ClipPolyhedron (Src/Poly.cpp:265-500) for a 32-lane warp:
//   * classification (Poly.cpp:303-319) is one signed distance per owned vertex + warp ballots;
//   * insertion of new vertices on straddling edges (Poly.cpp:332-363) is lane-parallel with the reference's
//     append order (i ascending, ring slot ascending) reproduced by a warp prefix sum, so vertex numbering,
//     ring contents and ring order come out IDENTICAL to the reference, not merely isomorphic;
//   * the topology patch (Poly.cpp:365-431) runs lane-parallel in the generic case (no in-plane vertex, every
//     face walk ends on a distinct new vertex): ring(new w) = [pusher, walked, kept] -- derived in DESIGN.md;
//     any other case (comp == 0 present, walk anomaly, non-injective walk targets) is replayed by lane 0 with the
//     reference's exact sequential loop over the same shared-memory arrays (seq_patch), never on the CPU;
//   * degree-2 splice (Poly.cpp:433-462) is detected in parallel and replayed by lane 0 (never observed);
//   * compaction (Poly.cpp:464-499) is a ballot/popc renumbering.
// The bounding-box shortcut (Poly.cpp:297-299) is folded into the classification: it changes the outcome only
// when every vertex is in-plane, and exactly that case evaluates the box (see all_inplane_box_says_skip).
#pragma once

#include "surtr_math.cuh"

namespace surtr
{
constexpr unsigned FULL = 0xffffffffu;

enum ClipStatus : int { CLIP_OK = 0, CLIP_OVERFLOW = 2 };

template <int CAP_, int DMAX_, typename IdxT_>
struct WarpPoly
{
    static constexpr int CAP = CAP_;
    static constexpr int DMAX = DMAX_;
    static constexpr int VPL = CAP_ / 32;   // vertices per lane: vertex v is owned by lane v % 32, group v / 32
    using IdxT = IdxT_;
    static constexpr IdxT NONE = (IdxT)~(IdxT)0;   // the reference's "-1" ring mark

    float x[CAP], y[CAP], z[CAP];
    IdxT ring[CAP * DMAX];
    IdxT old_ring[CAP * DMAX];   // snapshot for the sequential replay (Poly.cpp:367-369)
    uint8_t deg[CAP];
    uint8_t old_deg[CAP];
    int8_t comp[CAP];
    IdxT id[CAP];                // renumbering / injectivity scratch
    float4 tri[2 * CAP];         // ordered fan-triangle records for the moments (dV, mx, my, mz)
};

// FaceLoop (Src/Poly.cpp:34-41): ring entry just before vprev (wrapping); absent vprev -> last entry.
template <class P>
__device__ __forceinline__ int face_loop(const P& sp, int v, int vprev)
{
    const typename P::IdxT* r = sp.ring + v * P::DMAX;
    const int d = sp.deg[v];
    int k = 0;
    while (k < d && r[k] != (typename P::IdxT)vprev) k++;
    return k == 0 ? r[d - 1] : r[k - 1];
}

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

// Sequential replay of Poly.cpp:365-431 by one lane (patch + erase of -1 marks).  Returns false on ring overflow.
template <class P>
__device__ bool seq_patch(P& sp, int nverts0, int nverts)
{
    using IdxT = typename P::IdxT;
    constexpr int DMAX = P::DMAX;
    for (int ii = 0; ii < nverts; ii++)
    {
        const int i = (ii + nverts0) % nverts;
        const int ci = sp.comp[i];
        if (!(ci == 0 || ci == 2)) continue;
        const int nneigh = sp.deg[i];
        for (int j = 0; j < nneigh; j++)
        {
            const IdxT jn = sp.ring[i * DMAX + j];
            if (jn == P::NONE || sp.comp[jn] != -1) continue;
            int iprev = i, inext = jn, itmp, k = 0;
            while (sp.comp[inext] == -1 && k++ < nverts)
            {
                itmp = inext;
                inext = face_loop(sp, inext, iprev);
                iprev = itmp;
            }
            if (sp.ring[i * DMAX + (j + 1) % sp.deg[i]] == (IdxT)inext || inext == i)
            {
                sp.ring[i * DMAX + j] = P::NONE;
            }
            else
            {
                sp.ring[i * DMAX + j] = (IdxT)inext;
                const int dn = sp.deg[inext];
                if (dn >= DMAX) return false;
                IdxT* rn = sp.ring + inext * DMAX;
                IdxT* on = sp.old_ring + inext * DMAX;
                int off = 0;
                IdxT mark = (IdxT)i;
                if (sp.comp[inext] == 2)
                {
                    mark = P::NONE;   // Poly.cpp:409 inserts -1 into the snapshot
                }
                else
                {
                    const int od = sp.old_deg[inext];
                    while (off < od && on[off] != (IdxT)iprev) off++;
                }
                for (int q = dn; q > off; q--) rn[q] = rn[q - 1];
                rn[off] = (IdxT)i;
                sp.deg[inext] = (uint8_t)(dn + 1);
                const int od = sp.old_deg[inext];
                if (od >= DMAX) return false;
                for (int q = od; q > off; q--) on[q] = on[q - 1];
                on[off] = mark;
                sp.old_deg[inext] = (uint8_t)(od + 1);
            }
        }
    }
    for (int i = 0; i < nverts; i++)   // Poly.cpp:426-431
    {
        IdxT* r = sp.ring + i * DMAX;
        int w = 0;
        const int d = sp.deg[i];
        for (int k = 0; k < d; k++)
            if (r[k] != P::NONE) r[w++] = r[k];
        sp.deg[i] = (uint8_t)w;
    }
    return true;
}

// Sequential replay of Poly.cpp:433-462 by one lane.
template <class P>
__device__ void seq_splice(P& sp, int nverts)
{
    using IdxT = typename P::IdxT;
    constexpr int DMAX = P::DMAX;
    bool updated = true;
    while (updated)
    {
        updated = false;
        for (int i = 0; i < nverts; i++)
        {
            if (sp.comp[i] >= 0 && sp.deg[i] == 2)
            {
                updated = true;
                const int iprev = sp.ring[i * DMAX], inext = sp.ring[i * DMAX + 1];
                int k = 0;
                while (k < sp.deg[iprev] && sp.ring[iprev * DMAX + k] != (IdxT)i) ++k;
                if (k < sp.deg[iprev]) sp.ring[iprev * DMAX + k] = (IdxT)inext;
                k = 0;
                while (k < sp.deg[inext] && sp.ring[inext * DMAX + k] != (IdxT)i) ++k;
                if (k < sp.deg[inext]) sp.ring[inext * DMAX + k] = (IdxT)iprev;
                sp.comp[i] = -1;
            }
        }
    }
}

// Every vertex is in-plane: the reference's box test decides (Poly.cpp:297-299, 725-744): skip the plane when
// no box corner is clipped, otherwise the polyhedron is removed.
template <class P>
__device__ bool all_inplane_box_says_skip(const P& sp, int nv, const float4& pl, int lane)
{
    float lo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
    float hi[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
    for (int v = lane; v < nv; v += 32)
    {
        lo[0] = fminf(lo[0], sp.x[v]); hi[0] = fmaxf(hi[0], sp.x[v]);
        lo[1] = fminf(lo[1], sp.y[v]); hi[1] = fmaxf(hi[1], sp.y[v]);
        lo[2] = fminf(lo[2], sp.z[v]); hi[2] = fmaxf(hi[2], sp.z[v]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            lo[k] = fminf(lo[k], __shfl_xor_sync(FULL, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(FULL, hi[k], o));
        }
    const int k = lane & 7;
    const int c = classify(signed_dist(pl, (k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2]));
    return __ballot_sync(FULL, c == -1) == 0u;
}

// Clip the polyhedron held in `sp` (nv vertices) by planes[0..npl).  All 32 lanes call this together.
// On return nv is the surviving vertex count (0 = no fragment).  seq_cuts counts sequential replays.
template <class P>
__device__ int clip_by_planes(P& sp, int& nv, const float4* __restrict__ planes, int npl, int lane, unsigned& seq_cuts,
                              unsigned& n_cuts)
{
    using IdxT = typename P::IdxT;
    constexpr int DMAX = P::DMAX;
    constexpr int VPL = P::VPL;
    constexpr int CAP = P::CAP;

    float px[VPL], py[VPL], pz[VPL];
#pragma unroll
    for (int h = 0; h < VPL; h++)
    {
        const int v = lane + 32 * h;
        if (v < nv) { px[h] = sp.x[v]; py[h] = sp.y[v]; pz[h] = sp.z[v]; }
    }

    for (int kb = 0; kb < npl && nv > 0; kb += 32)
    {
        // lane l holds plane kb + l; planes are broadcast by shuffle (no memory latency inside the plane loop)
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kb + lane < npl) mine = __ldg(planes + kb + lane);
        const int kend = min(32, npl - kb);
        for (int kk = 0; kk < kend && nv > 0; kk++)
        {
            float4 pl;
            pl.x = __shfl_sync(FULL, mine.x, kk);
            pl.y = __shfl_sync(FULL, mine.y, kk);
            pl.z = __shfl_sync(FULL, mine.z, kk);
            pl.w = __shfl_sync(FULL, mine.w, kk);

            // ---- classify (Poly.cpp:303-319) ----
            float s[VPL];
            int c[VPL];
            bool any_clip = false, any_keep = false, any_zero = false;
            const int ngroups = (nv + 31) >> 5;
#pragma unroll
            for (int h = 0; h < VPL; h++)
            {
                c[h] = 3;
                if (h < ngroups)
                {
                    const int v = lane + 32 * h;
                    if (v < nv)
                    {
                        s[h] = signed_dist(pl, px[h], py[h], pz[h]);
                        c[h] = classify(s[h]);
                    }
                    any_clip |= __ballot_sync(FULL, c[h] == -1) != 0u;
                    any_keep |= __ballot_sync(FULL, c[h] == 1) != 0u;
                    any_zero |= __ballot_sync(FULL, c[h] == 0) != 0u;
                }
            }
            if (!any_keep)
            {
                // "below" (Poly.cpp:322-327) -- unless every vertex is in-plane and the box test says "above"
                if (!any_clip && all_inplane_box_says_skip(sp, nv, pl, lane)) continue;
                nv = 0;
                break;
            }
            if (!any_clip) continue;   // "above" (Poly.cpp:328)

            // ---- the plane cuts: insert new vertices (Poly.cpp:332-363) ----
            n_cuts++;
            const int nverts0 = nv;
#pragma unroll
            for (int h = 0; h < VPL; h++)
            {
                const int v = lane + 32 * h;
                if (v < nverts0) sp.comp[v] = (int8_t)c[h];
            }
            __syncwarp();

            int base[VPL];
            int nverts = nverts0;
#pragma unroll
            for (int h = 0; h < VPL; h++)
            {
                base[h] = 0;
                if (h < ngroups)
                {
                    const int v = lane + 32 * h;
                    int cnt = 0;
                    if (v < nverts0 && c[h] == -1)
                    {
                        const int d = sp.deg[v];
                        for (int j = 0; j < d; j++)
                            cnt += sp.comp[sp.ring[v * DMAX + j]] > 0;
                    }
                    int tot;
                    base[h] = nverts + warp_exscan(cnt, lane, tot);
                    nverts += tot;
                }
            }
            if (nverts > CAP) return CLIP_OVERFLOW;

#pragma unroll
            for (int h = 0; h < VPL; h++)
            {
                const int v = lane + 32 * h;
                if (h < ngroups && v < nverts0 && c[h] == -1)
                {
                    const int d = sp.deg[v];
                    int w = base[h];
                    for (int j = 0; j < d; j++)
                    {
                        const int jn = sp.ring[v * DMAX + j];
                        if (sp.comp[jn] > 0)
                        {
                            const float bx = sp.x[jn], by = sp.y[jn], bz = sp.z[jn];
                            const float sb = signed_dist(pl, bx, by, bz);
                            float ox, oy, oz;
                            plane_line_intersection(px[h], py[h], pz[h], s[h], bx, by, bz, sb, ox, oy, oz);
                            sp.x[w] = ox; sp.y[w] = oy; sp.z[w] = oz;
                            sp.comp[w] = 2;
                            sp.deg[w] = 2;
                            sp.ring[w * DMAX] = (IdxT)v;
                            sp.ring[w * DMAX + 1] = (IdxT)jn;
                            // several lanes may patch the ring of the same kept vertex jn at once: each replaces only the
                            // entry holding ITS clipped vertex v, and an entry another lane is rewriting (v' -> w') equals
                            // v neither before nor after -- entry-disjoint by construction (racecheck warns, word-level)
                            IdxT* rj = sp.ring + jn * DMAX;
                            const int dj = sp.deg[jn];
                            int k = 0;
                            while (k < dj && rj[k] != (IdxT)v) k++;
                            if (k < dj) rj[k] = (IdxT)w;
                            sp.ring[v * DMAX + j] = (IdxT)w;
                            w++;
                        }
                    }
                }
            }
            __syncwarp();

            // ---- patch the links to clipped vertices (Poly.cpp:365-431) ----
            const int nnew = nverts - nverts0;
            bool need_seq = any_zero;
            int X[VPL];
            if (!need_seq)
            {
                // every new vertex w = [i, jn] walks the face loop through clipped vertices (read-only rings)
                bool ok = true;
#pragma unroll
                for (int t = 0; t < VPL; t++)
                {
                    X[t] = -1;
                    const int w = nverts0 + lane + 32 * t;
                    if (32 * t < nnew && w < nverts)
                    {
                        int iprev = w, inext = sp.ring[w * DMAX], itmp, k = 0;
                        while (sp.comp[inext] == -1 && k++ < nverts)
                        {
                            itmp = inext;
                            inext = face_loop(sp, inext, iprev);
                            iprev = itmp;
                        }
                        X[t] = inext;
                        ok = ok && sp.comp[inext] == 2 && inext != w;
                        if (ok) sp.id[inext] = (IdxT)w;   // injectivity probe
                    }
                }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < VPL; t++)
                {
                    const int w = nverts0 + lane + 32 * t;
                    if (32 * t < nnew && w < nverts && ok) ok = sp.id[X[t]] == (IdxT)w;
                }
                need_seq = __ballot_sync(FULL, !ok) != 0u;
            }
            if (!need_seq)
            {
                // ring(w) = [pusher, walked, kept]; each new vertex is walked to by exactly one other
                IdxT kept[VPL];
#pragma unroll
                for (int t = 0; t < VPL; t++)
                {
                    const int w = nverts0 + lane + 32 * t;
                    if (32 * t < nnew && w < nverts) kept[t] = sp.ring[w * DMAX + 1];
                }
                __syncwarp();
#pragma unroll
                for (int t = 0; t < VPL; t++)
                {
                    const int w = nverts0 + lane + 32 * t;
                    if (32 * t < nnew && w < nverts)
                    {
                        sp.ring[w * DMAX + 2] = kept[t];
                        sp.ring[w * DMAX + 1] = (IdxT)X[t];
                        sp.ring[X[t] * DMAX] = (IdxT)w;
                        sp.deg[w] = 3;
                    }
                }
                __syncwarp();
            }
            else
            {
                seq_cuts++;
                for (int v = lane; v < nverts; v += 32)
                {
                    sp.old_deg[v] = sp.deg[v];
                    const int d = sp.deg[v];
                    for (int j = 0; j < d; j++) sp.old_ring[v * DMAX + j] = sp.ring[v * DMAX + j];
                }
                __syncwarp();
                int okflag = 1;
                if (lane == 0) okflag = seq_patch(sp, nverts0, nverts) ? 1 : 0;
                okflag = __shfl_sync(FULL, okflag, 0);
                __syncwarp();
                if (!okflag) return CLIP_OVERFLOW;
            }

            // ---- degree-2 splice (Poly.cpp:433-462) ----
            {
                bool two = false;
                for (int v = lane; v < nverts; v += 32) two |= sp.comp[v] >= 0 && sp.deg[v] == 2;
                if (__ballot_sync(FULL, two) != 0u)
                {
                    if (lane == 0) seq_splice(sp, nverts);
                    __syncwarp();
                }
            }

            // ---- compaction (Poly.cpp:464-499) ----
            {
                const int ng = (nverts + 31) >> 5;
                int kept_before = 0;
                for (int h = 0; h < ng; h++)
                {
                    const int v = lane + 32 * h;
                    const bool live = v < nverts && sp.comp[v] >= 0;
                    const unsigned m = __ballot_sync(FULL, live);
                    if (live) sp.id[v] = (IdxT)(kept_before + __popc(m & ((1u << lane) - 1u)));
                    kept_before += __popc(m);
                }
                __syncwarp();
                for (int h = 0; h < ng; h++)
                {
                    const int v = lane + 32 * h;
                    const bool live = v < nverts && sp.comp[v] >= 0;
                    float vx = 0.f, vy = 0.f, vz = 0.f;
                    IdxT r[DMAX];
                    int d = 0, t = 0;
                    if (live)
                    {
                        vx = sp.x[v]; vy = sp.y[v]; vz = sp.z[v];
                        d = sp.deg[v];
                        t = sp.id[v];
#pragma unroll
                        for (int j = 0; j < DMAX; j++)
                            if (j < d) r[j] = sp.id[sp.ring[v * DMAX + j]];
                    }
                    __syncwarp();
                    if (live)
                    {
                        sp.x[t] = vx; sp.y[t] = vy; sp.z[t] = vz;
                        sp.deg[t] = (uint8_t)d;
#pragma unroll
                        for (int j = 0; j < DMAX; j++)
                            if (j < d) sp.ring[t * DMAX + j] = r[j];
                    }
                    __syncwarp();
                }
                nv = kept_before < 4 ? 0 : kept_before;   // Poly.cpp:498-499
            }
#pragma unroll
            for (int h = 0; h < VPL; h++)
            {
                const int v = lane + 32 * h;
                if (v < nv) { px[h] = sp.x[v]; py[h] = sp.y[v]; pz[h] = sp.z[v]; }
            }
        }
    }
    return CLIP_OK;
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

template <class P>
__device__ void fragment_moments(P& sp, int nv, int lane, Moments& out)
{
    using IdxT = typename P::IdxT;
    constexpr int DMAX = P::DMAX;
    constexpr int VPL = P::VPL;
    const float ox = sp.x[0], oy = sp.y[0], oz = sp.z[0];
    const int ngroups = (nv + 31) >> 5;

    // pass 1: a directed edge (v, slot j) starts a face iff v is the smallest vertex of its loop -- the order in
    // which ExtractFaces (Poly.cpp:94-122) meets unvisited edges.  Count faces and fan triangles per lane.
    unsigned start_mask[VPL];
    int tri_base[VPL];
    int n_faces = 0, n_tri = 0;
#pragma unroll
    for (int h = 0; h < VPL; h++)
    {
        start_mask[h] = 0u;
        tri_base[h] = 0;
        if (h < ngroups)
        {
            const int v = lane + 32 * h;
            int cnt = 0, faces = 0;
            if (v < nv)
            {
                const int d = sp.deg[v];
                for (int j = 0; j < d; j++)
                {
                    int prev = v, at = sp.ring[v * DMAX + j], n = 1;
                    bool is_start = true;
                    while (at != v)
                    {
                        if (at < v || n > nv) { is_start = false; break; }
                        const int nxt = face_loop(sp, at, prev);
                        prev = at;
                        at = nxt;
                        n++;
                    }
                    if (is_start)
                    {
                        start_mask[h] |= 1u << j;
                        faces++;
                        cnt += n - 2;
                    }
                }
            }
            int tot;
            tri_base[h] = n_tri + warp_exscan(cnt, lane, tot);
            n_tri += tot;
            int ftot;
            warp_exscan(faces, lane, ftot);
            n_faces += ftot;
        }
    }

    // pass 2: emit the fan triangles at their position in the reference's accumulation order
    // xx yy zz xy xz yz second moments about the origin vertex, then 6V and the first moments (all double)
    double cov[10] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
    for (int h = 0; h < VPL; h++)
    {
        if (h < ngroups)
        {
            const int v = lane + 32 * h;
            if (v < nv)
            {
                int w = tri_base[h];
                const float p0x = __fsub_rn(sp.x[v], ox), p0y = __fsub_rn(sp.y[v], oy), p0z = __fsub_rn(sp.z[v], oz);
                const int d = sp.deg[v];
                for (int j = 0; j < d; j++)
                {
                    if (!(start_mask[h] >> j & 1u)) continue;
                    int prev = v, at = sp.ring[v * DMAX + j];
                    float p1x = __fsub_rn(sp.x[at], ox), p1y = __fsub_rn(sp.y[at], oy), p1z = __fsub_rn(sp.z[at], oz);
                    int nxt = face_loop(sp, at, prev);
                    prev = at;
                    at = nxt;
                    while (at != v)
                    {
                        const float p2x = __fsub_rn(sp.x[at], ox), p2y = __fsub_rn(sp.y[at], oy),
                                    p2z = __fsub_rn(sp.z[at], oz);
                        float cx, cy, cz;
                        cross3(p1x, p1y, p1z, p2x, p2y, p2z, cx, cy, cz);
                        const float dV = dot3(p0x, p0y, p0z, cx, cy, cz);
                        const float sx = __fadd_rn(__fadd_rn(p0x, p1x), p2x);
                        const float sy = __fadd_rn(__fadd_rn(p0y, p1y), p2y);
                        const float sz = __fadd_rn(__fadd_rn(p0z, p1z), p2z);
                        sp.tri[w++] = make_float4(dV, __fmul_rn(sx, dV), __fmul_rn(sy, dV), __fmul_rn(sz, dV));
                        {
                            // inertia: independent all-double integral over the same fan (exact differences)
                            const double a0 = (double)sp.x[v] - ox, a1 = (double)sp.y[v] - oy, a2 = (double)sp.z[v] - oz;
                            const double b0 = (double)sp.x[prev] - ox, b1 = (double)sp.y[prev] - oy, b2 = (double)sp.z[prev] - oz;
                            const double c0 = (double)sp.x[at] - ox, c1 = (double)sp.y[at] - oy, c2 = (double)sp.z[at] - oz;
                            const double dd = a0 * (b1 * c2 - b2 * c1) + a1 * (b2 * c0 - b0 * c2) + a2 * (b0 * c1 - b1 * c0);
                            const double s0 = a0 + b0 + c0, s1 = a1 + b1 + c1, s2 = a2 + b2 + c2;
                            cov[0] += dd * (s0 * s0 + a0 * a0 + b0 * b0 + c0 * c0);
                            cov[1] += dd * (s1 * s1 + a1 * a1 + b1 * b1 + c1 * c1);
                            cov[2] += dd * (s2 * s2 + a2 * a2 + b2 * b2 + c2 * c2);
                            cov[3] += dd * (s0 * s1 + a0 * a1 + b0 * b1 + c0 * c1);
                            cov[4] += dd * (s0 * s2 + a0 * a2 + b0 * b2 + c0 * c2);
                            cov[5] += dd * (s1 * s2 + a1 * a2 + b1 * b2 + c1 * c2);
                            cov[6] += dd;
                            cov[7] += dd * s0; cov[8] += dd * s1; cov[9] += dd * s2;
                        }
                        p1x = p2x; p1y = p2y; p1z = p2z;
                        nxt = face_loop(sp, at, prev);
                        prev = at;
                        at = nxt;
                    }
                }
            }
        }
    }
    __syncwarp();

    // pass 3: ordered accumulation (Poly.cpp:77-85) by one lane; everything else is a fixed-shape tree
    double zeroth = 0.0;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (lane == 0)
    {
        for (int t = 0; t < n_tri; t++)
        {
            const float4 r = sp.tri[t];
            zeroth += (double)r.x;
            fx = __fadd_rn(fx, r.y); fy = __fadd_rn(fy, r.z); fz = __fadd_rn(fz, r.w);
        }
        zeroth /= 6.0;
        const double q = 24.0 * zeroth;
        const double inv = (q >= 0.0 ? 1.0 : -1.0) / fmax(1.0e-30, fabs(q));   // safeInv, Poly.cpp:33
        const float sc = (float)inv;
        fx = __fmul_rn(fx, sc); fy = __fmul_rn(fy, sc); fz = __fmul_rn(fz, sc);
    }
#pragma unroll
    for (int k = 0; k < 10; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            cov[k] += __shfl_xor_sync(FULL, cov[k], o);
    zeroth = __shfl_sync(FULL, zeroth, 0);
    fx = __shfl_sync(FULL, fx, 0); fy = __shfl_sync(FULL, fy, 0); fz = __shfl_sync(FULL, fz, 0);

    out.n_faces = n_faces;
    out.volume = zeroth;
    out.cx = __fadd_rn(fx, ox); out.cy = __fadd_rn(fy, oy); out.cz = __fadd_rn(fz, oz);
    {
        // shift the second moments from the origin vertex to the centroid, then I = tr(C) 1 - C
        const double V = cov[6] / 6.0;
        const double iv = V != 0.0 ? 1.0 / (24.0 * V) : 0.0;
        const double c0 = cov[7] * iv, c1 = cov[8] * iv, c2 = cov[9] * iv;
        const double Cxx = cov[0] / 120.0 - V * c0 * c0, Cyy = cov[1] / 120.0 - V * c1 * c1,
                     Czz = cov[2] / 120.0 - V * c2 * c2, Cxy = cov[3] / 120.0 - V * c0 * c1,
                     Cxz = cov[4] / 120.0 - V * c0 * c2, Cyz = cov[5] / 120.0 - V * c1 * c2;
        out.inertia[0] = (float)(Cyy + Czz);
        out.inertia[1] = (float)(Cxx + Czz);
        out.inertia[2] = (float)(Cxx + Cyy);
        out.inertia[3] = (float)(-Cxy);
        out.inertia[4] = (float)(-Cxz);
        out.inertia[5] = (float)(-Cyz);
    }
    (void)sizeof(IdxT);
}
} // namespace surtr
