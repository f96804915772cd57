// clip_fast.cuh -- small-tier clipper: one warp, <= 64 vertex slots, ring degree <= 8, register/ballot based.
//
// Same algorithm and the same exactness argument as clip_warp.cuh (which stays as the large tier), restated for
// the common case with a layout chosen for the warp:
//   * a vertex ring is ONE 64-bit shared-memory word: eight u8 neighbour indices, 0xFF padded.  FaceLoop
//     (Poly.cpp:34-41), find-and-replace (Poly.cpp:350-353) and the degree are byte-compare instructions on a
//     register (__vcmpeq4 / __ffs / PRMT) after a single LDS.64 -- no dependent chain of shared-memory reads;
//   * the per-vertex classification `comp` (Poly.cpp:303-319) never touches memory: the warp ballots of
//     "clipped" and "kept" are held by every lane and comp(j) is a bit test;
//   * new vertices are created one per lane from a list of straddling half-edges written in the reference's
//     append order (vertex ascending, ring slot ascending; Poly.cpp:333-357);
//   * compaction (Poly.cpp:464-499) is LAZY: clipped vertices just leave the live mask and new ones are appended.
//     The reference's compaction is stable, so the relative order of live vertices is the same with or without
//     it; indices are renumbered (popc on the live mask) only when the 64 slots run out and once at the end,
//     which yields exactly the reference's numbering.
// Sequential replays (in-plane vertices, walk anomalies, degree-2 splice) run on lane 0 over the same words.
#pragma once

#include "clip_warp.cuh"

namespace surtr
{
typedef unsigned long long u64;

struct FastPoly   // one per warp in shared memory (4032 bytes)
{
    float x[64], y[64], z[64];
    u64 ring[64];         // 8 x u8, 0xFF = empty slot
    u64 old_ring[64];     // snapshot for the sequential replay (Poly.cpp:367-369)
    uint16_t list[64];    // straddling half-edges of the current cut: vertex | slot << 8; then walk targets
    uint8_t id[64];       // walk-target probe: id[X(w)] = w
    float4 tri[128];      // ordered fan-triangle records (dV, mx, my, mz)
};

constexpr int R_NONE = 0xff;   // empty slot
constexpr int R_MARK = 0xfe;   // the reference's "-1, to be removed" (Poly.cpp:400)

__device__ __forceinline__ int rdeg(u64 w)
{
    const unsigned mlo = __vcmpeq4((unsigned)w, 0xffffffffu);
    if (mlo) return (__ffs(mlo) - 1) >> 3;
    const unsigned mhi = __vcmpeq4((unsigned)(w >> 32), 0xffffffffu);
    return mhi ? 4 + ((__ffs(mhi) - 1) >> 3) : 8;
}
__device__ __forceinline__ int rfind(u64 w, int val)   // first slot holding val, 8 if absent
{
    const unsigned pat = (unsigned)val * 0x01010101u;
    const unsigned mlo = __vcmpeq4((unsigned)w, pat);
    if (mlo) return (__ffs(mlo) - 1) >> 3;
    const unsigned mhi = __vcmpeq4((unsigned)(w >> 32), pat);
    return mhi ? 4 + ((__ffs(mhi) - 1) >> 3) : 8;
}
__device__ __forceinline__ int rget(u64 w, int k) { return (int)(__byte_perm((unsigned)w, (unsigned)(w >> 32), (unsigned)k) & 0xffu); }
__device__ __forceinline__ u64 rset(u64 w, int k, int val)
{
    unsigned lo = (unsigned)w, hi = (unsigned)(w >> 32);
    const unsigned sh = (unsigned)(k & 3) * 8u, msk = 0xffu << sh, ins = (unsigned)val << sh;
    if (k < 4) lo = (lo & ~msk) | ins; else hi = (hi & ~msk) | ins;
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 rinsert(u64 w, int k, int val)   // shift slots >= k up by one (caller checks deg < 8)
{
    const u64 low = k == 0 ? 0ull : (w & (~0ull >> (64 - 8 * k)));
    const u64 high = k == 0 ? w : (w >> (8 * k));
    return low | ((u64)(unsigned)val << (8 * k)) | (k == 7 ? 0ull : (high << (8 * k + 8)));
}
// FaceLoop (Src/Poly.cpp:34-41) on a ring word: entry just before vprev (wrapping); absent vprev -> last entry.
__device__ __forceinline__ int rface_loop(u64 w, int vprev)
{
    const int d = rdeg(w);
    if (d == 0) return vprev;   // malformed input (vertex without neighbours): callers' loop bounds end the walk
    int k = rfind(w, vprev);
    if (k > d) k = d;
    return rget(w, (k == 0 ? d : k) - 1);
}

__device__ __forceinline__ bool bit64(unsigned m0, unsigned m1, int j) { return (((j & 32) ? m1 : m0) >> (j & 31)) & 1u; }
__device__ __forceinline__ unsigned lowmask(int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); }
__device__ __forceinline__ int rank64(unsigned s0, unsigned s1, int u)   // set bits with index < u
{
    return u < 32 ? __popc(s0 & lowmask(u)) : __popc(s0) + __popc(s1 & lowmask(u - 32));
}

struct CutState   // warp-uniform
{
    unsigned l0, l1;   // live vertices
    unsigned c0, c1;   // clipped by the current plane (comp == -1)
    unsigned k0, k1;   // kept by the current plane (comp == +1)
    int hi;            // allocated vertex slots
};

// comp of the reference for the sequential replays: 2 = new, -1 clipped / gone, +1 kept, 0 in-plane
__device__ __forceinline__ int comp_of(const CutState& s, int hi0, unsigned d0, unsigned d1, int j)
{
    if (bit64(d0, d1, j)) return -1;   // spliced away (Poly.cpp:459)
    if (j >= hi0) return 2;
    if (bit64(s.c0, s.c1, j) || !bit64(s.l0, s.l1, j)) return -1;
    return bit64(s.k0, s.k1, j) ? 1 : 0;
}

// Sequential replay of Poly.cpp:365-462 (patch, erase marks, degree-2 splice) by lane 0 after the new vertices
// have been inserted.  Visiting order = the reference's: new vertices first, then the pre-existing ones, both
// ascending.  Returns 0 on ring overflow; d0/d1 receive the vertices spliced away.
__device__ __noinline__ int fast_seq_cut(FastPoly& sp, const CutState s, int hi0, int nnew, int lane, unsigned& d0, unsigned& d1)
{
    const int hi1 = hi0 + nnew;
    if (lane < hi1) sp.old_ring[lane] = sp.ring[lane];
    if (lane + 32 < hi1) sp.old_ring[lane + 32] = sp.ring[lane + 32];
    __syncwarp();
    int ok = 1;
    unsigned dd0 = 0u, dd1 = 0u;
    if (lane == 0)
    {
        const int nverts = __popc(s.l0) + __popc(s.l1) + nnew;   // the reference's vertex count (walk bound)
        for (int ii = 0; ii < hi1 && ok; ii++)
        {
            const int i = ii < nnew ? hi0 + ii : ii - nnew;
            const int ci = comp_of(s, hi0, 0u, 0u, i);
            if (!(ci == 0 || ci == 2)) continue;
            const int nneigh = rdeg(sp.ring[i]);
            for (int j = 0; j < nneigh; j++)
            {
                const int jn = rget(sp.ring[i], j);
                if (jn >= R_MARK || comp_of(s, hi0, 0u, 0u, jn) != -1) continue;
                int iprev = i, inext = jn, itmp, k = 0;
                while (comp_of(s, hi0, 0u, 0u, inext) == -1 && k++ < nverts)
                {
                    itmp = inext;
                    inext = rface_loop(sp.ring[inext], iprev);
                    iprev = itmp;
                }
                const u64 wi = sp.ring[i];
                if (rget(wi, (j + 1) % rdeg(wi)) == inext || inext == i)
                {
                    sp.ring[i] = rset(wi, j, R_MARK);
                }
                else
                {
                    sp.ring[i] = rset(wi, j, inext);
                    const u64 wn = sp.ring[inext], on = sp.old_ring[inext];
                    if (rdeg(wn) >= 8 || rdeg(on) >= 8) { ok = 0; break; }
                    int off = 0, mark = i;
                    if (comp_of(s, hi0, 0u, 0u, inext) == 2) mark = R_MARK;   // Poly.cpp:409 inserts -1 in the snapshot
                    else { off = rfind(on, iprev); if (off > rdeg(on)) off = rdeg(on); }
                    sp.ring[inext] = rinsert(wn, off, i);
                    sp.old_ring[inext] = rinsert(on, off, mark);
                }
            }
        }
        for (int i = 0; i < hi1; i++)   // Poly.cpp:426-431
        {
            const u64 w = sp.ring[i];
            u64 o = ~0ull;
            int n = 0;
            for (int k = 0; k < 8; k++)
            {
                const int b = rget(w, k);
                if (b == R_NONE) break;
                if (b != R_MARK) o = rset(o, n++, b);
            }
            sp.ring[i] = o;
        }
        bool updated = ok != 0;   // Poly.cpp:433-462
        while (updated)
        {
            updated = false;
            for (int i = 0; i < hi1; i++)
            {
                if (comp_of(s, hi0, dd0, dd1, i) >= 0 && rdeg(sp.ring[i]) == 2)
                {
                    updated = true;
                    const int iprev = rget(sp.ring[i], 0), inext = rget(sp.ring[i], 1);
                    int k = rfind(sp.ring[iprev], i);
                    if (k < rdeg(sp.ring[iprev])) sp.ring[iprev] = rset(sp.ring[iprev], k, inext);
                    k = rfind(sp.ring[inext], i);
                    if (k < rdeg(sp.ring[inext])) sp.ring[inext] = rset(sp.ring[inext], k, iprev);
                    if (i & 32) dd1 |= 1u << (i & 31); else dd0 |= 1u << i;
                }
            }
        }
    }
    ok = __shfl_sync(FULL, ok, 0);
    d0 = __shfl_sync(FULL, dd0, 0);
    d1 = __shfl_sync(FULL, dd1, 0);
    __syncwarp();
    return ok;
}

// Renumber the live vertices to 0..n-1 keeping their order (the reference's compaction, Poly.cpp:464-495).
// Positions are read from / written to shared memory; the caller reloads its register copies.
__device__ __noinline__ void fast_compact(FastPoly& sp, CutState& s, int lane)
{
    u64 r[2] = { ~0ull, ~0ull };
    float vx[2], vy[2], vz[2];
    bool live[2];
#pragma unroll
    for (int g = 0; g < 2; g++)
    {
        const int v = lane + 32 * g;
        live[g] = bit64(s.l0, s.l1, v);
        if (live[g])
        {
            vx[g] = sp.x[v]; vy[g] = sp.y[v]; vz[g] = sp.z[v];
            const u64 rw = sp.ring[v];
            for (int j = 0; j < 8; j++)
            {
                const int b = rget(rw, j);
                if (b == R_NONE) break;
                r[g] = rset(r[g], j, rank64(s.l0, s.l1, b));
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 2; g++)
    {
        if (live[g])
        {
            const int t = rank64(s.l0, s.l1, lane + 32 * g);
            sp.x[t] = vx[g]; sp.y[t] = vy[g]; sp.z[t] = vz[g]; sp.ring[t] = r[g];
        }
    }
    __syncwarp();
    const int n = __popc(s.l0) + __popc(s.l1);
    s.hi = n;
    s.l0 = lowmask(n);
    s.l1 = lowmask(n - 32);
}

// Every vertex in-plane: the reference's box test decides (Poly.cpp:297-299, 725-744).
__device__ __noinline__ bool fast_all_inplane_box_says_skip(const FastPoly& sp, const CutState s, const float4 pl, int lane)
{
    float lo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
    float hi[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
    for (int v = lane; v < s.hi; v += 32)
    {
        if (!bit64(s.l0, s.l1, v)) continue;
        lo[0] = fminf(lo[0], sp.x[v]); hi[0] = fmaxf(hi[0], sp.x[v]);
        lo[1] = fminf(lo[1], sp.y[v]); hi[1] = fmaxf(hi[1], sp.y[v]);
        lo[2] = fminf(lo[2], sp.z[v]); hi[2] = fmaxf(hi[2], sp.z[v]);
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++)
        {
            lo[k] = fminf(lo[k], __shfl_xor_sync(FULL, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(FULL, hi[k], o));
        }
    const int k = lane & 7;
    const int c = classify(signed_dist(pl, (k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2]));
    return __ballot_sync(FULL, c == -1) == 0u;
}

// Clip the polyhedron in `sp` (nv vertices in slots 0..nv-1; lanes own slots lane and lane + 32) by
// planes[0..npl).  On return s.l0/s.l1 are the live slots (not renumbered) and nv their count (0 = no fragment).
__device__ int fast_clip_by_planes(FastPoly& sp, CutState& s, int& nv, const float4* __restrict__ planes, int npl, int lane,
                                   unsigned& seq_cuts, unsigned& n_cuts)
{
    float px[2] = { 0.f, 0.f }, py[2] = { 0.f, 0.f }, pz[2] = { 0.f, 0.f };
    s.hi = nv;
    s.l0 = lowmask(nv);
    s.l1 = lowmask(nv - 32);
    if (lane < nv) { px[0] = sp.x[lane]; py[0] = sp.y[lane]; pz[0] = sp.z[lane]; }
    if (lane + 32 < nv) { px[1] = sp.x[lane + 32]; py[1] = sp.y[lane + 32]; pz[1] = sp.z[lane + 32]; }

    for (int kb = 0; kb < npl && nv > 0; kb += 32)
    {
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kb + lane < npl) mine = __ldg(planes + kb + lane);   // lane l holds plane kb + l
        const int kend = min(32, npl - kb);
        for (int kk = 0; kk < kend && nv > 0; kk++)
        {
            float4 pl;
            pl.x = __shfl_sync(FULL, mine.x, kk);
            pl.y = __shfl_sync(FULL, mine.y, kk);
            pl.z = __shfl_sync(FULL, mine.z, kk);
            pl.w = __shfl_sync(FULL, mine.w, kk);

            // ---- classify (Poly.cpp:303-319): one distance per owned live vertex, two ballots per group ----
            int c0 = 3, c1 = 3;
            if ((s.l0 >> lane) & 1u) c0 = classify(signed_dist(pl, px[0], py[0], pz[0]));
            s.c0 = __ballot_sync(FULL, c0 == -1);
            s.k0 = __ballot_sync(FULL, c0 == 1);
            s.c1 = s.k1 = 0u;
            if (s.l1)
            {
                if ((s.l1 >> lane) & 1u) c1 = classify(signed_dist(pl, px[1], py[1], pz[1]));
                s.c1 = __ballot_sync(FULL, c1 == -1);
                s.k1 = __ballot_sync(FULL, c1 == 1);
            }
            if (!(s.k0 | s.k1))
            {
                // "below" (Poly.cpp:322-327) -- unless every vertex is in-plane and the box test says "above"
                if (!(s.c0 | s.c1) && fast_all_inplane_box_says_skip(sp, s, pl, lane)) continue;
                nv = 0;
                break;
            }
            if (!(s.c0 | s.c1)) continue;   // "above" (Poly.cpp:328)

            // ---- the plane cuts ----
            n_cuts++;
            __syncwarp();   // ring words composed by the previous cut are visible from here on
            // straddling half-edges (clipped vertex -> kept neighbour) in the reference's append order
            unsigned smask = 0u;   // bits 0-7: slots of vertex `lane`, bits 8-15: slots of vertex `lane + 32`
            int cnt = 0;           // low half: group 0, high half: group 1
#pragma unroll 1
            for (int g = 0; g < 2; g++)
            {
                if (((g ? s.c1 : s.c0) >> lane) & 1u)
                {
                    const u64 rw = sp.ring[lane + 32 * g];
                    for (int j = 0; j < 8; j++)
                    {
                        const int b = rget(rw, j);
                        if (b == R_NONE) break;
                        if (bit64(s.k0, s.k1, b)) { smask |= 1u << (j + 8 * g); cnt += 1 << (16 * g); }
                    }
                }
            }
            int tot;
            const int ex = warp_exscan(cnt, lane, tot);
            const int tot0 = tot & 0xffff, nnew = tot0 + (tot >> 16);
            if (s.hi + nnew > 64)
            {
                // out of slots: renumber the live vertices (exactly the reference's compaction) and redo this plane
                if (__popc(s.l0) + __popc(s.l1) + nnew > 64) return CLIP_OVERFLOW;
                fast_compact(sp, s, lane);
#pragma unroll
                for (int g = 0; g < 2; g++)
                {
                    const int v = lane + 32 * g;
                    if (v < s.hi) { px[g] = sp.x[v]; py[g] = sp.y[v]; pz[g] = sp.z[v]; }
                }
                n_cuts--;
                kk--;
                continue;
            }
            const int hi0 = s.hi;
            {
                int w = ex & 0xffff;
                unsigned m = smask & 0xffu;
                while (m) { const int j = __ffs(m) - 1; m &= m - 1; sp.list[w++] = (uint16_t)(lane | (j << 8)); }
                w = tot0 + (ex >> 16);
                m = smask >> 8;
                while (m) { const int j = __ffs(m) - 1; m &= m - 1; sp.list[w++] = (uint16_t)((lane + 32) | (j << 8)); }
            }
            __syncwarp();
            // insert: one new vertex per lane (Poly.cpp:345-354)
            for (int t = lane; t < nnew; t += 32)
            {
                const int e = sp.list[t], v = e & 0xff, j = e >> 8, w = hi0 + t;
                const int jn = rget(sp.ring[v], j);
                const float ax = sp.x[v], ay = sp.y[v], az = sp.z[v], bx = sp.x[jn], by = sp.y[jn], bz = sp.z[jn];
                const float sa = signed_dist(pl, ax, ay, az), sb = signed_dist(pl, bx, by, bz);
                float ox, oy, oz;
                plane_line_intersection(ax, ay, az, sa, bx, by, bz, sb, ox, oy, oz);
                sp.x[w] = ox; sp.y[w] = oy; sp.z[w] = oz;
                sp.ring[w] = 0xffffffffffff0000ull | (u64)(unsigned)v | ((u64)(unsigned)jn << 8);
                reinterpret_cast<uint8_t*>(&sp.ring[v])[j] = (uint8_t)w;
                const int k = rfind(sp.ring[jn], v);
                if (k < 8) reinterpret_cast<uint8_t*>(&sp.ring[jn])[k] = (uint8_t)w;
            }
            __syncwarp();

            // patch (Poly.cpp:365-431): walk from each new vertex through clipped vertices to the next new one
            const bool any_zero = ((s.l0 & ~(s.c0 | s.k0)) | (s.l1 & ~(s.c1 | s.k1))) != 0u;
            bool need_seq = any_zero;
            if (!need_seq)
            {
                bool ok = true;
                for (int t = lane; t < nnew; t += 32)
                {
                    const int w = hi0 + t;
                    int iprev = w, inext = rget(sp.ring[w], 0), itmp, k = 0;
                    while (inext < hi0 && bit64(s.c0, s.c1, inext) && k++ < 64)
                    {
                        itmp = inext;
                        inext = rface_loop(sp.ring[inext], iprev);
                        iprev = itmp;
                    }
                    const bool okt = inext >= hi0 && inext < hi0 + nnew && inext != w;
                    if (okt) sp.id[inext] = (uint8_t)w;
                    sp.list[t] = (uint16_t)inext;
                    ok = ok && okt;
                }
                __syncwarp();
                for (int t = lane; t < nnew; t += 32)
                    if (ok) ok = sp.id[sp.list[t]] == (uint8_t)(hi0 + t);
                need_seq = __ballot_sync(FULL, !ok) != 0u;
                if (!need_seq)
                {
                    // the walk targets are a permutation of the new vertices: ring(w) = [pusher, walked, kept]
                    for (int t = lane; t < nnew; t += 32)
                    {
                        const int w = hi0 + t;
                        const int kept = rget(sp.ring[w], 1);
                        sp.ring[w] = 0xffffffffff000000ull | (u64)sp.id[w] | ((u64)sp.list[t] << 8) | ((u64)(unsigned)kept << 16);
                    }
                }
            }
            unsigned dead0 = 0u, dead1 = 0u;
            if (need_seq)
            {
                seq_cuts++;
                if (!fast_seq_cut(sp, s, hi0, nnew, lane, dead0, dead1)) return CLIP_OVERFLOW;
            }
            // lazy compaction: clipped vertices leave the live set, new ones join it
            s.hi = hi0 + nnew;
            s.l0 = ((s.l0 & ~s.c0) | (lowmask(s.hi) & ~lowmask(hi0))) & ~dead0;
            s.l1 = ((s.l1 & ~s.c1) | (lowmask(s.hi - 32) & ~lowmask(hi0 - 32))) & ~dead1;
            nv = __popc(s.l0) + __popc(s.l1);
            if (nv < 4) nv = 0;   // Poly.cpp:498-499
#pragma unroll
            for (int g = 0; g < 2; g++)
            {
                const int v = lane + 32 * g;
                if (v >= hi0 && v < s.hi) { px[g] = sp.x[v]; py[g] = sp.y[v]; pz[g] = sp.z[v]; }
            }
        }
    }
    __syncwarp();
    return CLIP_OK;
}

// Poly::ExtractFaces + Poly::Moments in the reference's accumulation order (Poly.cpp:55-126) + inertia, on the
// live (not renumbered) slots: vertex order = slot order, origin = first live vertex.  See fragment_moments in
// clip_warp.cuh for the derivation.
__device__ void fast_fragment_moments(FastPoly& sp, const CutState& s, int lane, Moments& out)
{
    const int first = s.l0 ? __ffs(s.l0) - 1 : 32 + __ffs(s.l1) - 1;
    const float ox = sp.x[first], oy = sp.y[first], oz = sp.z[first];
    const int nv = __popc(s.l0) + __popc(s.l1);
    unsigned start_mask = 0u;   // bits 0-7 group 0, 8-15 group 1
    int cnt = 0, faces = 0;
#pragma unroll 1
    for (int g = 0; g < 2; g++)
    {
        const int v = lane + 32 * g;
        if (bit64(s.l0, s.l1, v))
        {
            const u64 rw = sp.ring[v];
            const int d = rdeg(rw);
            for (int j = 0; j < d; j++)
            {
                int at = rget(rw, j);
                // v can only be the smallest vertex of the loop through (v -> at) if both loop neighbours of v are
                // larger: `at`, and the vertex the loop arrives from = the ring entry after `at` (FaceLoop inverted)
                if (at < v || rget(rw, j + 1 == d ? 0 : j + 1) < v) continue;
                int prev = v, n = 1;
                bool is_start = true;
                while (at != v)
                {
                    if (at < v || n > nv) { is_start = false; break; }
                    const int nxt = rface_loop(sp.ring[at], prev);
                    prev = at;
                    at = nxt;
                    n++;
                }
                if (is_start)
                {
                    start_mask |= 1u << (j + 8 * g);
                    faces++;
                    cnt += max(n - 2, 0) << (16 * g);
                }
            }
        }
    }
    int tot;
    const int ex = warp_exscan(cnt, lane, tot);   // both groups in one scan
    const int n_tri0 = tot & 0xffff;
    const int n_faces = __reduce_add_sync(FULL, faces);
    const int n_tri = min(n_tri0 + (tot >> 16), 128);

    float cov[10] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };   // xx yy zz xy xz yz, 6V, first moments
#pragma unroll 1
    for (int g = 0; g < 2; g++)
    {
        unsigned m = (start_mask >> (8 * g)) & 0xffu;
        if (!m) continue;
        const int v = lane + 32 * g;
        int w = g ? n_tri0 + (ex >> 16) : (ex & 0xffff);
        const float p0x = __fsub_rn(sp.x[v], ox), p0y = __fsub_rn(sp.y[v], oy), p0z = __fsub_rn(sp.z[v], oz);
        const u64 rw = sp.ring[v];
        while (m)
        {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            int prev = v, at = rget(rw, j);
            float p1x = __fsub_rn(sp.x[at], ox), p1y = __fsub_rn(sp.y[at], oy), p1z = __fsub_rn(sp.z[at], oz);
            int nxt = rface_loop(sp.ring[at], prev);
            prev = at;
            at = nxt;
            while (at != v)
            {
                const float p2x = __fsub_rn(sp.x[at], ox), p2y = __fsub_rn(sp.y[at], oy), p2z = __fsub_rn(sp.z[at], oz);
                float cx, cy, cz;
                cross3(p1x, p1y, p1z, p2x, p2y, p2z, cx, cy, cz);
                const float dV = dot3(p0x, p0y, p0z, cx, cy, cz);
                const float sx = __fadd_rn(__fadd_rn(p0x, p1x), p2x);
                const float sy = __fadd_rn(__fadd_rn(p0y, p1y), p2y);
                const float sz = __fadd_rn(__fadd_rn(p0z, p1z), p2z);
                if (w < 128) sp.tri[w] = make_float4(dV, __fmul_rn(sx, dV), __fmul_rn(sy, dV), __fmul_rn(sz, dV));
                w++;
                // second moments of the tetrahedron (origin, p0, p1, p2): dV/120 * (s s^T + sum p p^T)
                cov[0] += dV * (sx * sx + p0x * p0x + p1x * p1x + p2x * p2x);
                cov[1] += dV * (sy * sy + p0y * p0y + p1y * p1y + p2y * p2y);
                cov[2] += dV * (sz * sz + p0z * p0z + p1z * p1z + p2z * p2z);
                cov[3] += dV * (sx * sy + p0x * p0y + p1x * p1y + p2x * p2y);
                cov[4] += dV * (sx * sz + p0x * p0z + p1x * p1z + p2x * p2z);
                cov[5] += dV * (sy * sz + p0y * p0z + p1y * p1z + p2y * p2z);
                cov[6] += dV;
                cov[7] += dV * sx; cov[8] += dV * sy; cov[9] += dV * sz;
                p1x = p2x; p1y = p2y; p1z = p2z;
                nxt = rface_loop(sp.ring[at], prev);
                prev = at;
                at = nxt;
            }
        }
    }
    __syncwarp();

    // ordered accumulation (Poly.cpp:77-85): lane c < 4 owns component c of the triangle records (dV, mx, my, mz)
    // and adds them in the reference's order -- dV into a double, the first moments in float.  Four lanes share one
    // instruction stream, so the serial chain costs a quarter of a single-lane loop.
    double zeroth = 0.0;
    float fsum = 0.f;
    if (lane < 4)
    {
        const float* comp = reinterpret_cast<const float*>(sp.tri) + lane;
        int t = 0;
        for (; t + 4 <= n_tri; t += 4)   // the loads do not depend on the accumulation chain
        {
            const float r0 = comp[4 * t], r1 = comp[4 * t + 4], r2 = comp[4 * t + 8], r3 = comp[4 * t + 12];
            zeroth += (double)r0; zeroth += (double)r1; zeroth += (double)r2; zeroth += (double)r3;
            fsum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(fsum, r0), r1), r2), r3);
        }
        for (; t < n_tri; t++)
        {
            const float r = comp[4 * t];
            zeroth += (double)r;
            fsum = __fadd_rn(fsum, r);
        }
    }
    zeroth = __shfl_sync(FULL, zeroth, 0) / 6.0;
    float fx = __shfl_sync(FULL, fsum, 1), fy = __shfl_sync(FULL, fsum, 2), fz = __shfl_sync(FULL, fsum, 3);
    {
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

    out.n_faces = n_faces;
    out.volume = zeroth;
    out.cx = __fadd_rn(fx, ox); out.cy = __fadd_rn(fy, oy); out.cz = __fadd_rn(fz, oz);
    {
        // shift from the origin vertex to the centroid (all from the same sums), then I = tr(C) 1 - C
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
