// clip_fast.cuh -- small-tier clipper, round 2: ONE warp clips one (piece, cell) pair; 32 * G vertex slots (G = 2: the
// main tier, G = 4: the 128-slot tier for pieces whose cut transiently needs more than 64 slots), ring degree <= 8.
//
// Same algorithm and the same exactness argument as clip_sub.cuh -- ring words, ballots instead
// of a per-vertex comp array, new vertices in the reference's append order, lazy compaction, sequential replay of the
// reference loop for in-plane / anomalous cuts -- but written for exactly one pair per warp, which is what K3 launches:
//   * every mask (live, clipped, kept) is G 32-bit words, word g = the ballot of vertex group g: no 64-bit shifts, no
//     sub-warp bookkeeping, and every branch on them is warp-uniform by construction -- no votes to agree on a branch;
//   * the plane of the current iteration is ONE broadcast LDG.128 (prefetched one plane ahead) instead of four shuffles;
//   * a plane that does not cut (4 of 5 on the Voronoi-on-Voronoi configs) costs a signed distance, two compares and
//     two ballots per vertex group and touches no shared memory and no __syncwarp;
//   * the prefix sum that places new vertices in append order is built from ballots of the count bits (popc against
//     the lane mask) instead of a five-step shuffle scan.
// Profiles: profiles/r2_k3_*.txt (instruction count per pair and per cut before / after).
#pragma once

#include "clip_sub.cuh"

namespace surtr
{
template <int G>
struct FastPoly   // one per warp in shared memory: 31 bytes per slot (G = 2: 1984 bytes, G = 4: 3968 bytes)
{
    static constexpr int S = 32 * G;
    float x[S], y[S], z[S];
    u64 ring[S];          // 8 x u8, 0xFF = empty slot
    u64 old_ring[S];      // snapshot for the sequential replay (Poly.cpp:367-369)
    uint16_t list[S];     // straddling half-edges of the current cut: slot | ring slot << 8; then walk targets
    uint8_t id[S];        // walk-target probe: id[X(w)] = w; final numbering at write-out
};

// Two builds of the clipper, chosen per launch (kernels.cuh, clip_fast_kernel<..., LAT>):
//   LAT = false (throughput; events of more than one wave of warps): 40 resident warps per SM = 48 registers; the positions
//     are read from shared memory in the classification (three LDS per vertex group against six registers) and the next
//     plane of the queue is not loaded ahead (four registers) -- the other warps hide those latencies.  Config 4: K3 2.38 ->
//     2.24 ms per 256-event batch against the 32-warp / 64-register build.
//   LAT = true (latency; an event that fits one wave, BASELINE config 2): 32 warps, 64 registers, positions of the lane's
//     own slots in registers, plane prefetch: the single warp's critical path decides (config 2 cold event 0.136 vs 0.149 ms).
// Measured over 24-40 warps, 0 / 1 / all register groups, prefetch on / off: profiles/r3_k3_experiments.md, section 3.
#ifndef SURTR_K3_REG_GROUPS
#define SURTR_K3_REG_GROUPS 0   // (throughput build) vertex groups whose positions the owner lane keeps in registers
#endif
#ifndef SURTR_K3_PREFETCH
#define SURTR_K3_PREFETCH 0     // (throughput build) 1: load the next plane of the queue one iteration ahead
#endif
constexpr int FAST_MAX_PLANES = 64;   // planes the prefilter keeps a bit for (cells beyond it: every plane takes the exact path)

// The planes the exact path still has to look at, as a bit set that is consumed from the bottom: bits 0..63 = planes
// 0..63 (cleared by the prefilter for planes it proved irrelevant), `tail` = the next plane >= 64 (cells with more than
// FAST_MAX_PLANES planes visit those one after the other).  Warp-uniform.
struct PlaneQueue
{
    unsigned w0, w1;
    int tail;
    __device__ __forceinline__ int peek(int npl) const   // smallest plane left (npl if none)
    {
        const int p = w0 ? __ffs((int)w0) - 1 : (w1 ? 31 + __ffs((int)w1) : tail);
        return p < npl ? p : npl;
    }
    __device__ __forceinline__ void pop()
    {
        if (w0) w0 &= w0 - 1u;
        else if (w1) w1 &= w1 - 1u;
        else tail++;
    }
};

template <int G>
struct FastMasks   // warp-uniform
{
    unsigned live[G], c[G], k[G];
};

// word g of a register-resident mask array, g not a compile-time constant (a dynamic index would spill the array)
template <int G>
__device__ __forceinline__ unsigned mword(const unsigned (&m)[G], int g)
{
    unsigned w = m[0];
#pragma unroll
    for (int i = 1; i < G; i++) w = g == i ? m[i] : w;
    return w;
}
template <int G>
__device__ __forceinline__ bool mbit(const unsigned (&m)[G], int v) { return (mword<G>(m, v >> 5) >> (v & 31)) & 1u; }
template <int G>
__device__ __forceinline__ int mrank(const unsigned (&m)[G], int v)   // set bits below v
{
    int r = 0;
#pragma unroll
    for (int i = 0; i < G; i++)
    {
        const int g = v >> 5;
        r += i < g ? __popc(m[i]) : (i == g ? __popc(m[i] & ((1u << (v & 31)) - 1u)) : 0);
    }
    return r;
}
template <int G>
__device__ __forceinline__ int mcount(const unsigned (&m)[G])
{
    int r = 0;
#pragma unroll
    for (int i = 0; i < G; i++) r += __popc(m[i]);
    return r;
}
__device__ __forceinline__ unsigned lowmask32(int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); }

// comp of the reference for the sequential replay: 2 = new, -1 clipped / gone, +1 kept, 0 in-plane
template <int G>
__device__ __forceinline__ int fast_comp_of(const FastMasks<G>& m, const unsigned (&dead)[G], int hi0, int j)
{
    if (mbit<G>(dead, j)) return -1;   // spliced away (Poly.cpp:459)
    if (j >= hi0) return 2;
    if (mbit<G>(m.c, j) || !mbit<G>(m.live, j)) return -1;
    return mbit<G>(m.k, j) ? 1 : 0;
}

// Sequential replay of Poly.cpp:365-462 (patch, erase marks, degree-2 splice) after the new vertices have been
// inserted.  Only the patch itself is order dependent: lane 0 replays it over the vertices it can touch (comp 2 = the
// new ones first, then comp 0 = the in-plane ones, both ascending -- the reference's visiting order); erasing the marks
// is per vertex (all lanes), and the splice loop runs only if some vertex was left with two neighbours (never seen on
// the BASELINE configs).  Returns ok = 0 on ring overflow; dead = the vertices spliced away.  Called by all lanes.
template <int G>
struct SeqResult   // by value: nothing of the caller's mask registers has its address taken
{
    int ok;
    unsigned dead[G];
};
template <int G>
__device__ __noinline__ SeqResult<G> fast_seq_cut(FastPoly<G>& sp, const FastMasks<G> m, int hi0, int nnew, int lane)
{
    const int hi1 = hi0 + nnew;
    for (int v = lane; v < hi1; v += 32) sp.old_ring[v] = sp.ring[v];
    __syncwarp();
    int ok = 1;
    unsigned none[G];
#pragma unroll
    for (int g = 0; g < G; g++) none[g] = 0u;
    if (lane == 0)
    {
        const int nverts = mcount<G>(m.live) + nnew;   // the reference's vertex count (walk bound)
        // visiting order: new vertices hi0 .. hi1-1, then the in-plane ones (live, neither clipped nor kept) ascending
        int i = hi0, g_in = 0;
        unsigned in_w = m.live[0] & ~(m.c[0] | m.k[0]);
        while (ok)
        {
            if (i >= hi0)
            {
                if (i >= hi1) i = -1;      // new vertices done: switch to the in-plane ones
            }
            if (i < 0)
            {
                while (!in_w && ++g_in < G) in_w = mword<G>(m.live, g_in) & ~(mword<G>(m.c, g_in) | mword<G>(m.k, g_in));
                if (!in_w) break;
                i = 32 * g_in + __ffs((int)in_w) - 1;
                in_w &= in_w - 1;
            }
            const int nneigh = rdeg(sp.ring[i]);
            for (int j = 0; j < nneigh; j++)
            {
                const int jn = rget(sp.ring[i], j);
                if (jn >= R_MARK || fast_comp_of<G>(m, none, hi0, jn) != -1) continue;
                int iprev = i, inext = jn, itmp, k = 0;
                while (fast_comp_of<G>(m, none, hi0, inext) == -1 && k++ < nverts)
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
                    if (fast_comp_of<G>(m, none, hi0, inext) == 2) mark = R_MARK;   // Poly.cpp:409 inserts -1 in the snapshot
                    else { off = rfind(on, iprev); if (off > rdeg(on)) off = rdeg(on); }
                    sp.ring[inext] = rinsert(wn, off, i);
                    sp.old_ring[inext] = rinsert(on, off, mark);
                }
            }
            i = i >= hi0 ? i + 1 : -1;
        }
    }
    ok = __shfl_sync(FULL, ok, 0);
    __syncwarp();
    bool two = false;   // a surviving vertex left with exactly two neighbours (Poly.cpp:433-462 would splice it)
    for (int i = lane; i < hi1; i += 32)   // Poly.cpp:426-431, per vertex
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
        two |= n == 2 && fast_comp_of<G>(m, none, hi0, i) >= 0;
    }
    unsigned dd[G];
#pragma unroll
    for (int g = 0; g < G; g++) dd[g] = 0u;
    if (__ballot_sync(FULL, two) != 0u && ok)
    {
        __syncwarp();
        if (lane == 0)
        {
            bool updated = true;   // Poly.cpp:433-462
            while (updated)
            {
                updated = false;
                for (int i = 0; i < hi1; i++)
                {
                    if (fast_comp_of<G>(m, dd, hi0, i) >= 0 && rdeg(sp.ring[i]) == 2)
                    {
                        updated = true;
                        const int iprev = rget(sp.ring[i], 0), inext = rget(sp.ring[i], 1);
                        int k = rfind(sp.ring[iprev], i);
                        if (k < rdeg(sp.ring[iprev])) sp.ring[iprev] = rset(sp.ring[iprev], k, inext);
                        k = rfind(sp.ring[inext], i);
                        if (k < rdeg(sp.ring[inext])) sp.ring[inext] = rset(sp.ring[inext], k, iprev);
#pragma unroll
                        for (int g = 0; g < G; g++)
                            if ((i >> 5) == g) dd[g] |= 1u << (i & 31);
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; g++) dd[g] = __shfl_sync(FULL, dd[g], 0);
    }
    SeqResult<G> res;
    res.ok = ok;
#pragma unroll
    for (int g = 0; g < G; g++) res.dead[g] = dd[g];
    __syncwarp();
    return res;
}

// Renumber the live vertices to 0..n-1 keeping their order (the reference's compaction, Poly.cpp:464-495).
template <int G>
__device__ __noinline__ int fast_compact(FastPoly<G>& sp, const FastMasks<G> m, int lane)   // returns the live count n: slots 0..n-1 are live afterwards
{
    unsigned live[G];
#pragma unroll
    for (int g = 0; g < G; g++) live[g] = m.live[g];
    u64 r[G];
    float vx[G], vy[G], vz[G];
#pragma unroll
    for (int g = 0; g < G; g++)
    {
        const int v = lane + 32 * g;
        r[g] = ~0ull;
        vx[g] = vy[g] = vz[g] = 0.f;
        if ((live[g] >> lane) & 1u)
        {
            vx[g] = sp.x[v]; vy[g] = sp.y[v]; vz[g] = sp.z[v];
            const u64 rw = sp.ring[v];
            for (int j = 0; j < 8; j++)
            {
                const int b = rget(rw, j);
                if (b == R_NONE) break;
                r[g] = rset(r[g], j, mrank<G>(live, b));
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < G; g++)
    {
        const int v = lane + 32 * g;
        if ((live[g] >> lane) & 1u)
        {
            const int t = mrank<G>(live, v);
            sp.x[t] = vx[g]; sp.y[t] = vy[g]; sp.z[t] = vz[g]; sp.ring[t] = r[g];
        }
    }
    __syncwarp();
    return mcount<G>(live);
}

// Every vertex in-plane: the reference's box test decides (Poly.cpp:297-299, 725-744).  Called by all lanes.
template <int G>
__device__ __noinline__ bool fast_all_inplane_box_says_skip(const FastPoly<G>& sp, const FastMasks<G> m, int hi, const float4 pl, int lane)
{
    float lo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
    float hv[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
    for (int v = lane; v < hi; v += 32)
    {
        if (!mbit<G>(m.live, v)) continue;
        lo[0] = fminf(lo[0], sp.x[v]); hv[0] = fmaxf(hv[0], sp.x[v]);
        lo[1] = fminf(lo[1], sp.y[v]); hv[1] = fmaxf(hv[1], sp.y[v]);
        lo[2] = fminf(lo[2], sp.z[v]); hv[2] = fmaxf(hv[2], sp.z[v]);
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++)
        {
            lo[k] = fminf(lo[k], __shfl_xor_sync(FULL, lo[k], o));
            hv[k] = fmaxf(hv[k], __shfl_xor_sync(FULL, hv[k], o));
        }
    const int k = lane & 7;
    const int c = classify(signed_dist(pl, (k & 1) ? hv[0] : lo[0], (k & 2) ? hv[1] : lo[1], (k & 4) ? hv[2] : lo[2]));
    return __ballot_sync(FULL, c == -1) == 0u;
}

// Clip the polyhedron in `sp` (nv vertices in slots 0..nv-1, positions of the lane's own slots also in px/py/pz) by
// planes[0..npl).  Called by the 32 lanes of the pair's warp.  On return live = the live slots (not renumbered), hi the
// allocated slots, nv the live count (0 = no fragment); returns the pair's status.
template <int G, int RG = SURTR_K3_REG_GROUPS, bool PF = (SURTR_K3_PREFETCH != 0)>
__device__ int fast_clip_by_planes(FastPoly<G>& sp, unsigned (&live)[G], int& hi, int& nv, float (&px)[G], float (&py)[G], float (&pz)[G],
                                   const float4* __restrict__ planes, int npl, int lane, unsigned& seq_cuts, unsigned& n_cuts,
                                   const float (&box)[6], bool use_box)
{
    constexpr int S = 32 * G;
    const unsigned lm = 1u << lane, lt = lm - 1u;
    FastMasks<G> m;
    hi = nv;
#pragma unroll
    for (int g = 0; g < G; g++) { m.live[g] = lowmask32(nv - 32 * g); m.c[g] = m.k[g] = 0u; }
    int status = CLIP_OK;

    // ---- plane prefilter against the piece's bounding box (lane = plane) ----
    // Every vertex a cut creates lies on an edge of the current polytope, hence (up to rounding, ~1e-6 of the coordinate
    // scale per generation) inside the piece's axis-aligned box [lo, hi] (K1 wrote it: the first three k-DOP slabs).  A
    // plane whose signed distance is below -margin over the WHOLE box, margin = 5e-4 * (sum |n_i| (|c_i| + h_i) + |d|)
    // -- hundreds of times that rounding -- classifies every vertex the reference could ever hold as kept: its loop would
    // find the plane "above" whenever it gets there, and the plane is skipped without touching the vertices.  Likewise a
    // plane above +margin over the whole box clips every such vertex: the result is empty (Poly.cpp:322-327) whatever the
    // planes before it did.  Planes that come within the margin of the box take the exact path below.
    unsigned visit[(FAST_MAX_PLANES + 31) / 32];
    {
        const float cx = __fmul_rn(0.5f, __fadd_rn(box[0], box[1])), hx = __fmul_rn(0.5f, __fsub_rn(box[1], box[0]));
        const float cy = __fmul_rn(0.5f, __fadd_rn(box[2], box[3])), hy = __fmul_rn(0.5f, __fsub_rn(box[3], box[2]));
        const float cz = __fmul_rn(0.5f, __fadd_rn(box[4], box[5])), hz = __fmul_rn(0.5f, __fsub_rn(box[5], box[4]));
        bool kill = false;
#pragma unroll
        for (int w = 0; w < (FAST_MAX_PLANES + 31) / 32; w++)
        {
            visit[w] = 0xffffffffu;
            if (32 * w < npl && npl <= FAST_MAX_PLANES && use_box)   // warp-uniform
            {
                const int q = 32 * w + lane;
                bool near = false, dead = false;
                if (q < npl)
                {
                    const float4 pl = __ldg(planes + q);
                    const float ax = fabsf(pl.x), ay = fabsf(pl.y), az = fabsf(pl.z);
                    const float mid = signed_dist(pl, cx, cy, cz);
                    const float ext = __fadd_rn(__fadd_rn(__fmul_rn(ax, hx), __fmul_rn(ay, hy)), __fmul_rn(az, hz));
                    const float scale = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ax, fabsf(cx)), __fmul_rn(ay, fabsf(cy))), __fmul_rn(az, fabsf(cz))),
                                                  __fadd_rn(ext, fabsf(pl.w)));
                    const float margin = __fmul_rn(5.0e-4f, scale);
                    near = !(__fadd_rn(mid, ext) < -margin);     // (a NaN anywhere: near)
                    dead = __fsub_rn(mid, ext) > margin;
                }
                visit[w] = __ballot_sync(FULL, near);
                kill |= __ballot_sync(FULL, dead) != 0u;
            }
        }
        if (kill) { nv = 0; npl = 0; }
    }

    PlaneQueue pq;
    pq.w0 = visit[0] & lowmask32(npl);
    pq.w1 = visit[1] & lowmask32(npl - 32);
    pq.tail = FAST_MAX_PLANES;
    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
    int p = pq.peek(npl);   // first plane to visit
    pq.pop();
    if (p < npl) cur = __ldg(planes + p);
    while (p < npl && nv > 0)
    {
        const float4 pl = cur;
        const int pn = pq.peek(npl);                                           // next plane the exact path has to look at
        float4 nxt = cur;
        if (PF) nxt = __ldg(planes + (pn < npl ? pn : p));                     // broadcast load, one plane ahead

        // ---- classify (Poly.cpp:303-319): one distance per owned live vertex, two ballots per vertex group ----
        unsigned anyc = 0u, anyk = 0u;
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            m.c[g] = m.k[g] = 0u;
            if (g == 0 || hi > 32 * g)   // warp-uniform
            {
                const float d = g < RG ? signed_dist(pl, px[g], py[g], pz[g]) : signed_dist(pl, sp.x[lane + 32 * g], sp.y[lane + 32 * g], sp.z[lane + 32 * g]);
                const bool off = (m.live[g] & lm) && !(fabsf(d) < __uint_as_float(0x2EDBE6FFu));   // live and not in-plane (a NaN distance is in-plane)
                m.c[g] = __ballot_sync(FULL, off && d > 0.f);
                m.k[g] = __ballot_sync(FULL, off && d < 0.f);
                anyc |= m.c[g];
                anyk |= m.k[g];
            }
        }
        if (!anyc)
        {
            // nothing clipped: "above" (Poly.cpp:328) -- unless every vertex is in-plane and the box test says "below"
            if (!anyk && !fast_all_inplane_box_says_skip<G>(sp, m, hi, pl, lane)) { nv = 0; break; }
            cur = PF ? nxt : __ldg(planes + (pn < npl ? pn : p));
            p = pn;
            pq.pop();
            continue;
        }
        if (!anyk) { nv = 0; break; }   // "below" (Poly.cpp:322-327)

        // ---- the plane cuts: straddling half-edges (clipped vertex -> kept neighbour) in the reference's append order ----
        unsigned smk = 0u;      // 8 ring-slot bits per owned vertex group
        int cnt[G];
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            cnt[g] = 0;
            if (m.c[g] & lm)
            {
                const u64 rw = sp.ring[lane + 32 * g];
#pragma unroll 1
                for (int j = 0; j < 8; j++)
                {
                    const int b = rget(rw, j);
                    if (b == R_NONE) break;
                    if (mbit<G>(m.k, b)) { smk |= 1u << (j + 8 * g); cnt[g]++; }
                }
            }
        }
        // exclusive prefix in (group, lane) order from ballots of the count bits
        int pos[G], nnew = 0;
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            pos[g] = nnew;
            if (m.c[g])   // warp-uniform: some vertex of this group is clipped
            {
                const unsigned b0 = __ballot_sync(FULL, cnt[g] & 1), b1 = __ballot_sync(FULL, cnt[g] & 2), b23 = __ballot_sync(FULL, cnt[g] & 12);
                pos[g] += __popc(b0 & lt) + 2 * __popc(b1 & lt);
                nnew += __popc(b0) + 2 * __popc(b1);
                if (b23)   // a clipped vertex with four or more kept neighbours: rare
                {
                    const unsigned b2 = __ballot_sync(FULL, cnt[g] & 4), b3 = __ballot_sync(FULL, cnt[g] & 8);
                    pos[g] += 4 * __popc(b2 & lt) + 8 * __popc(b3 & lt);
                    nnew += 4 * __popc(b2) + 8 * __popc(b3);
                }
            }
        }
        if (hi + nnew > S)
        {
            // out of slots: renumber the live vertices (exactly the reference's compaction) and redo this plane
            if (mcount<G>(m.live) + nnew > S) { status = CLIP_OVERFLOW; break; }
            hi = fast_compact<G>(sp, m, lane);
#pragma unroll
            for (int g = 0; g < G; g++) m.live[g] = lowmask32(hi - 32 * g);
#pragma unroll
            for (int g = 0; g < G; g++)
            {
                const int v = lane + 32 * g;
                if (g < RG && v < hi) { px[g] = sp.x[v]; py[g] = sp.y[v]; pz[g] = sp.z[v]; }
            }
            continue;
        }
        n_cuts++;
        const int hi0 = hi;
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            unsigned mm = (smk >> (8 * g)) & 0xffu;
            int w = pos[g];   // (already offset by the totals of the earlier groups)
            while (mm) { const int j = __ffs(mm) - 1; mm &= mm - 1; sp.list[w++] = (uint16_t)((lane + 32 * g) | (j << 8)); }
        }
        __syncwarp();
        // insert: one new vertex per lane (Poly.cpp:345-354).  Lanes may touch the same ring WORD concurrently, but never
        // the same BYTE: lane t replaces exactly slot j of ring[v] and the slot of ring[jn] that holds v (see clip_sub.cuh).
#pragma unroll 1
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
        bool inplane = false;
#pragma unroll
        for (int g = 0; g < G; g++) inplane |= (m.live[g] & ~(m.c[g] | m.k[g])) != 0u;
        bool need_seq = inplane;   // warp-uniform
        if (!need_seq)
        {
            bool ok = true;
#pragma unroll 1
            for (int t = lane; t < nnew; t += 32)
            {
                const int w = hi0 + t;
                // first step without a search: w sits in slot j of its clipped end point v, so FaceLoop(v, w) is simply
                // the slot before j (list[t] still holds v | j << 8 from the insertion)
                const int e = sp.list[t], v = e & 0xff, j = e >> 8;
                const u64 rv = sp.ring[v];
                int iprev = v, inext = rget(rv, (j == 0 ? rdeg(rv) : j) - 1), itmp, k = 1;
#pragma unroll 1
                while (inext < hi0 && mbit<G>(m.c, inext) && k++ < S)
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
#pragma unroll 1
            for (int t = lane; t < nnew; t += 32)
                if (ok) ok = sp.id[sp.list[t]] == (uint8_t)(hi0 + t);
            need_seq = __ballot_sync(FULL, !ok) != 0u;
            if (!need_seq)
            {
                // the walk targets are a permutation of the new vertices: ring(w) = [pusher, walked, kept]
#pragma unroll 1
                for (int t = lane; t < nnew; t += 32)
                {
                    const int w = hi0 + t;
                    const int kept = rget(sp.ring[w], 1);
                    sp.ring[w] = 0xffffffffff000000ull | (u64)sp.id[w] | ((u64)sp.list[t] << 8) | ((u64)(unsigned)kept << 16);
                }
            }
        }
        unsigned dead[G];
#pragma unroll
        for (int g = 0; g < G; g++) dead[g] = 0u;
        if (need_seq)
        {
            seq_cuts++;
            const SeqResult<G> sr = fast_seq_cut<G>(sp, m, hi0, nnew, lane);
            if (!sr.ok) { status = CLIP_OVERFLOW; break; }
#pragma unroll
            for (int g = 0; g < G; g++) dead[g] = sr.dead[g];
        }
        // lazy compaction: clipped vertices leave the live set, new ones join it
        hi = hi0 + nnew;
        nv = 0;
        const u64 fresh = ((nnew >= 64 ? 0ull : (1ull << nnew)) - 1ull) << (hi0 & 63);   // slots hi0 .. hi-1 (G = 2: as one 64-bit mask)
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            const unsigned nm = G == 2 ? (unsigned)(fresh >> (32 * (g & 1))) : (lowmask32(hi - 32 * g) & ~lowmask32(hi0 - 32 * g));
            m.live[g] = ((m.live[g] & ~m.c[g]) | nm) & ~dead[g];
            nv += __popc(m.live[g]);
            const int v = lane + 32 * g;
            if (g < RG && v >= hi0 && v < hi) { px[g] = sp.x[v]; py[g] = sp.y[v]; pz[g] = sp.z[v]; }
        }
        if (nv < 4) nv = 0;   // Poly.cpp:498-499
        __syncwarp();         // ring words composed above are visible to the next cut
        cur = PF ? nxt : __ldg(planes + (pn < npl ? pn : p));
        p = pn;
        pq.pop();
    }
#pragma unroll
    for (int g = 0; g < G; g++) live[g] = m.live[g];
    return status;
}
} // namespace surtr
