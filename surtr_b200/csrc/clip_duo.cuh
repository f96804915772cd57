// clip_duo.cuh -- small-tier clipper, TWO candidate pairs per warp: lanes 0-15 clip one pair, lanes 16-31 another, with
// the expensive phases of a cut executed by both halves TOGETHER.
//
// Why: with one warp per pair (clip_fast.cuh) the cut phases -- straddle list, vertex insertion, patch walk, ring
// composition -- run with 5-10 of 32 lanes active (lane = clipped vertex or new vertex; a cut of a Voronoi piece makes
// 4-7 new vertices) and are 55 % of K3's warp instructions (profiles/r2_k3_instruction_buckets.txt); the kernel is
// bound by instruction issue (79 % issue-active), so the only lever left is lanes per instruction.  Round 1's lock-step
// version (clip_sub.cuh, L = 16) did not win because both halves stepped through their plane lists index by index: a
// plane cuts in one half while the other half's plane does not, so the cut phases were hardly ever shared.  Here every
// half first ADVANCES through its queue of planes (the prefilter of clip_fast.cuh leaves ~3 of ~15, nearly all of them
// cutting) until it holds a plane that cuts -- classification only, cheap -- and then both halves run ONE cut together.
// The warp works in rounds: search (each half, until its next cutting plane), cut (both), until neither half has a
// plane left.
//
// Algorithm, arithmetic, vertex numbering and ring contents are exactly clip_fast.cuh's: same
// FastPoly<2> workspace per pair (64 vertex slots, ring degree <= 8), same ring words, same append order of new vertices,
// lazy compaction, sequential replay of the reference loop for in-plane / anomalous cuts.  What changes is the mapping:
//   * lane sl (0..15) of a half owns vertex slots sl, sl + 16, sl + 32, sl + 48; the masks (live / clipped / kept) are
//     two 32-bit words per HALF, bit v = slot v, assembled from the half's 16 bits of every full-warp ballot;
//   * every collective is executed by all 32 lanes with the full mask, every branch around a collective is decided by a
//     warp-wide vote; pair-specific work inside is predicated (`searching`, `cut`, `mine`);
//   * positions are read from shared memory in the classification (no per-lane register copies: four slots per lane
//     would cost twelve registers of a 64-register budget).
#pragma once

#include "clip_fast.cuh"

namespace surtr
{
constexpr int DUO_L = 16;   // lanes per pair
constexpr int DUO_G = 4;    // vertex slots per lane

// this half's 16 bits of a full-warp ballot
__device__ __forceinline__ unsigned duo_half(unsigned ballot, int shift) { return (ballot >> shift) & 0xffffu; }
// bit of slot v = sl + 16 g in a two-word mask: word g >> 1, bit sl + 16 (g & 1)
__device__ __forceinline__ bool duo_own(const unsigned (&m)[2], int g, int sl) { return (m[g >> 1] >> (sl + 16 * (g & 1))) & 1u; }

struct DuoResult   // per half (uniform within it)
{
    int status, hi, nv;
    unsigned live[2];
    unsigned seq_cuts, n_cuts;
};

// Sequential replay of Poly.cpp:365-462 for the halves with `mine` set (fast_seq_cut of clip_fast.cuh, one lane of the
// half replays the patch).  Called by all 32 lanes.
__device__ __noinline__ SeqResult<2> duo_seq_cut(FastPoly<2>& sp, const FastMasks<2> m, int hi0, int nnew, int sl, int shift, bool mine)
{
    const int hi1 = mine ? hi0 + nnew : 0;
    for (int v = sl; v < hi1; v += DUO_L) sp.old_ring[v] = sp.ring[v];
    __syncwarp();
    int ok = 1;
    unsigned none[2] = { 0u, 0u };
    if (mine && sl == 0)
    {
        const int nverts = mcount<2>(m.live) + nnew;   // the reference's vertex count (walk bound)
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
                while (!in_w && ++g_in < 2) in_w = mword<2>(m.live, g_in) & ~(mword<2>(m.c, g_in) | mword<2>(m.k, g_in));
                if (!in_w) break;
                i = 32 * g_in + __ffs((int)in_w) - 1;
                in_w &= in_w - 1;
            }
            const int nneigh = rdeg(sp.ring[i]);
            for (int j = 0; j < nneigh; j++)
            {
                const int jn = rget(sp.ring[i], j);
                if (jn >= R_MARK || fast_comp_of<2>(m, none, hi0, jn) != -1) continue;
                int iprev = i, inext = jn, itmp, k = 0;
                while (fast_comp_of<2>(m, none, hi0, inext) == -1 && k++ < nverts)
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
                    if (fast_comp_of<2>(m, none, hi0, inext) == 2) mark = R_MARK;   // Poly.cpp:409 inserts -1 in the snapshot
                    else { off = rfind(on, iprev); if (off > rdeg(on)) off = rdeg(on); }
                    sp.ring[inext] = rinsert(wn, off, i);
                    sp.old_ring[inext] = rinsert(on, off, mark);
                }
            }
            i = i >= hi0 ? i + 1 : -1;
        }
    }
    ok = __shfl_sync(FULL, ok, 0, DUO_L);
    __syncwarp();
    bool two = false;   // a surviving vertex left with exactly two neighbours (Poly.cpp:433-462 would splice it)
    for (int i = sl; i < hi1; i += DUO_L)   // Poly.cpp:426-431, per vertex
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
        two |= n == 2 && fast_comp_of<2>(m, none, hi0, i) >= 0;
    }
    unsigned dd[2] = { 0u, 0u };
    const bool splice = duo_half(__ballot_sync(FULL, two), shift) != 0u && ok && mine;
    if (__any_sync(FULL, splice))
    {
        __syncwarp();
        if (splice && sl == 0)
        {
            bool updated = true;   // Poly.cpp:433-462
            while (updated)
            {
                updated = false;
                for (int i = 0; i < hi1; i++)
                {
                    if (fast_comp_of<2>(m, dd, hi0, i) >= 0 && rdeg(sp.ring[i]) == 2)
                    {
                        updated = true;
                        const int iprev = rget(sp.ring[i], 0), inext = rget(sp.ring[i], 1);
                        int k = rfind(sp.ring[iprev], i);
                        if (k < rdeg(sp.ring[iprev])) sp.ring[iprev] = rset(sp.ring[iprev], k, inext);
                        k = rfind(sp.ring[inext], i);
                        if (k < rdeg(sp.ring[inext])) sp.ring[inext] = rset(sp.ring[inext], k, iprev);
                        dd[i >> 5] |= 1u << (i & 31);
                    }
                }
            }
        }
        dd[0] = __shfl_sync(FULL, dd[0], 0, DUO_L);
        dd[1] = __shfl_sync(FULL, dd[1], 0, DUO_L);
    }
    SeqResult<2> res;
    res.ok = ok;
    res.dead[0] = dd[0];
    res.dead[1] = dd[1];
    __syncwarp();
    return res;
}

// Renumber the live vertices of the halves with `mine` set to 0..n-1 keeping their order (Poly.cpp:464-495); returns n.
__device__ __noinline__ int duo_compact(FastPoly<2>& sp, const FastMasks<2> m, int sl, bool mine)
{
    unsigned live[2] = { mine ? m.live[0] : 0u, mine ? m.live[1] : 0u };
    u64 r[DUO_G];
    float vx[DUO_G], vy[DUO_G], vz[DUO_G];
#pragma unroll
    for (int g = 0; g < DUO_G; g++)
    {
        const int v = sl + DUO_L * g;
        r[g] = ~0ull;
        vx[g] = vy[g] = vz[g] = 0.f;
        if (duo_own(live, g, sl))
        {
            vx[g] = sp.x[v]; vy[g] = sp.y[v]; vz[g] = sp.z[v];
            const u64 rw = sp.ring[v];
            for (int j = 0; j < 8; j++)
            {
                const int b = rget(rw, j);
                if (b == R_NONE) break;
                r[g] = rset(r[g], j, mrank<2>(live, b));
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < DUO_G; g++)
    {
        const int v = sl + DUO_L * g;
        if (duo_own(live, g, sl))
        {
            const int t = mrank<2>(live, v);
            sp.x[t] = vx[g]; sp.y[t] = vy[g]; sp.z[t] = vz[g]; sp.ring[t] = r[g];
        }
    }
    __syncwarp();
    return mcount<2>(live);
}

// Every vertex in-plane: the reference's box test decides (Poly.cpp:297-299, 725-744).  Called by all 32 lanes; the
// answer is meaningful for the halves with `mine` set.
__device__ __noinline__ bool duo_all_inplane_box_says_skip(const FastPoly<2>& sp, const FastMasks<2> m, int hi, const float4 pl, int sl, int shift, bool mine)
{
    float lo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
    float hv[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
    for (int v = sl; v < (mine ? hi : 0); v += DUO_L)
    {
        if (!mbit<2>(m.live, v)) continue;
        lo[0] = fminf(lo[0], sp.x[v]); hv[0] = fmaxf(hv[0], sp.x[v]);
        lo[1] = fminf(lo[1], sp.y[v]); hv[1] = fmaxf(hv[1], sp.y[v]);
        lo[2] = fminf(lo[2], sp.z[v]); hv[2] = fmaxf(hv[2], sp.z[v]);
    }
    for (int o = DUO_L / 2; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++)
        {
            lo[k] = fminf(lo[k], __shfl_xor_sync(FULL, lo[k], o, DUO_L));
            hv[k] = fmaxf(hv[k], __shfl_xor_sync(FULL, hv[k], o, DUO_L));
        }
    const int k = sl & 7;
    const int c = classify(signed_dist(pl, (k & 1) ? hv[0] : lo[0], (k & 2) ? hv[1] : lo[1], (k & 4) ? hv[2] : lo[2]));
    return duo_half(__ballot_sync(FULL, c == -1), shift) == 0u;
}

// Clip the polyhedra of both halves: each half's FastPoly<2> `sp` holds its piece (nv vertices in slots 0..nv-1), its
// cell is planes[0..npl), box = the piece's axis-aligned box (K1).  `act` = this half has a pair at all.  Called by the
// 32 lanes of the warp; on return R holds the half's status, live slots (not renumbered), allocated slots and live count.
__device__ void duo_clip_by_planes(FastPoly<2>& sp, bool act, int nv, const float4* __restrict__ planes, int npl, const float (&box)[6],
                                   bool use_box, int lane, DuoResult& R)
{
    constexpr int S = 64;
    const int sl = lane & (DUO_L - 1), shift = lane & DUO_L;
    const unsigned lt = (1u << sl) - 1u;
    FastMasks<2> m;
    if (!act) { nv = 0; npl = 0; }
    int hi = nv, status = CLIP_OK;
    m.live[0] = lowmask32(nv); m.live[1] = lowmask32(nv - 32);
    m.c[0] = m.c[1] = m.k[0] = m.k[1] = 0u;
    unsigned seq_cuts = 0, n_cuts = 0;

    // ---- plane prefilter against the piece's bounding box (lane = plane; see clip_fast.cuh for the argument) ----
    unsigned visit0 = 0xffffffffu, visit1 = 0xffffffffu;
    {
        const bool pf = act && use_box && npl <= FAST_MAX_PLANES;
        const int npl_pf = pf ? npl : 0;
        const int npl_max = max(npl_pf, __shfl_xor_sync(FULL, npl_pf, DUO_L));
        const float cx = __fmul_rn(0.5f, __fadd_rn(box[0], box[1])), hx = __fmul_rn(0.5f, __fsub_rn(box[1], box[0]));
        const float cy = __fmul_rn(0.5f, __fadd_rn(box[2], box[3])), hy = __fmul_rn(0.5f, __fsub_rn(box[3], box[2]));
        const float cz = __fmul_rn(0.5f, __fadd_rn(box[4], box[5])), hz = __fmul_rn(0.5f, __fsub_rn(box[5], box[4]));
        bool kill = false;
        if (pf) visit0 = visit1 = 0u;
#pragma unroll
        for (int w = 0; w < FAST_MAX_PLANES / DUO_L; w++)
        {
            if (DUO_L * w < npl_max)   // warp-uniform
            {
                const int q = DUO_L * w + sl;
                bool near = false, dead = false;
                if (q < npl_pf)
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
                const unsigned bn = duo_half(__ballot_sync(FULL, near), shift), bd = duo_half(__ballot_sync(FULL, dead), shift);
                if (pf)
                {
                    if (w < 2) visit0 |= bn << (16 * (w & 1)); else visit1 |= bn << (16 * (w & 1));
                    kill |= bd != 0u;
                }
            }
        }
        if (kill) { nv = 0; npl = 0; }
    }

    PlaneQueue pq;
    pq.w0 = visit0 & lowmask32(npl);
    pq.w1 = visit1 & lowmask32(npl - 32);
    pq.tail = FAST_MAX_PLANES;
    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f), nxt = cur;
    int p = pq.peek(npl), pn = p;   // first plane to visit
    pq.pop();
    if (p < npl) cur = __ldg(planes + p);
    const float in_plane = __uint_as_float(0x2EDBE6FFu);   // 1e-10 as the float threshold (surtr_math.cuh)

    while (true)
    {
        // ---- search: every half that has planes left classifies until it holds a plane that cuts ----
        bool searching = status == CLIP_OK && p < npl && nv > 0;
        bool cut = false;
        if (!__any_sync(FULL, searching)) break;
        do
        {
            if (searching)
            {
                pn = pq.peek(npl);                                   // next plane the exact path has to look at
                nxt = __ldg(planes + (pn < npl ? pn : p));           // one plane ahead (the same address for the whole half)
            }
            const int hs = searching ? hi : 0;
            const int himax = max(hs, __shfl_xor_sync(FULL, hs, DUO_L));
            unsigned c0 = 0u, c1 = 0u, k0 = 0u, k1 = 0u;
#pragma unroll
            for (int g = 0; g < DUO_G; g++)
            {
                if (g == 0 || himax > DUO_L * g)   // warp-uniform
                {
                    const int v = sl + DUO_L * g;
                    float d = 0.f;
                    const bool lv = searching && duo_own(m.live, g, sl);
                    if (lv) d = signed_dist(cur, sp.x[v], sp.y[v], sp.z[v]);
                    const bool off = lv && !(fabsf(d) < in_plane);   // live and not in-plane (a NaN distance is in-plane)
                    const unsigned bc = duo_half(__ballot_sync(FULL, off && d > 0.f), shift);
                    const unsigned bk = duo_half(__ballot_sync(FULL, off && d < 0.f), shift);
                    if (g < 2) { c0 |= bc << (16 * (g & 1)); k0 |= bk << (16 * (g & 1)); }
                    else { c1 |= bc << (16 * (g & 1)); k1 |= bk << (16 * (g & 1)); }
                }
            }
            bool all_inplane = false, advance = false;
            if (searching)
            {
                m.c[0] = c0; m.c[1] = c1; m.k[0] = k0; m.k[1] = k1;
                if (!(c0 | c1))
                {
                    if (k0 | k1) advance = true;        // "above" (Poly.cpp:328)
                    else all_inplane = true;            // every vertex in-plane: the box test decides
                }
                else if (!(k0 | k1)) { nv = 0; searching = false; }   // "below" (Poly.cpp:322-327)
                else { cut = true; searching = false; }
            }
            if (__any_sync(FULL, all_inplane))
            {
                const bool skip = duo_all_inplane_box_says_skip(sp, m, hi, cur, sl, shift, all_inplane);
                if (all_inplane)
                {
                    if (skip) advance = true;
                    else { nv = 0; searching = false; }
                }
            }
            if (advance)
            {
                cur = nxt;
                p = pn;
                pq.pop();
                searching = p < npl;
            }
        } while (__any_sync(FULL, searching));

        if (!__any_sync(FULL, cut)) continue;   // (every half ran out of planes or died: the loop head ends it)

        // ---- the cut, both halves together: straddling half-edges (clipped vertex -> kept neighbour) in append order ----
        unsigned smk = 0u;      // 8 ring-slot bits per owned vertex group
        int cnt[DUO_G];
#pragma unroll
        for (int g = 0; g < DUO_G; g++)
        {
            cnt[g] = 0;
            if (cut && duo_own(m.c, g, sl))
            {
                const u64 rw = sp.ring[sl + DUO_L * g];
#pragma unroll 1
                for (int j = 0; j < 8; j++)
                {
                    const int b = rget(rw, j);
                    if (b == R_NONE) break;
                    if (mbit<2>(m.k, b)) { smk |= 1u << (j + 8 * g); cnt[g]++; }
                }
            }
        }
        // which vertex groups hold a clipped vertex in either half (warp-uniform): the others need no prefix ballots
        unsigned cw0 = cut ? m.c[0] : 0u, cw1 = cut ? m.c[1] : 0u;
        cw0 |= __shfl_xor_sync(FULL, cw0, DUO_L);
        cw1 |= __shfl_xor_sync(FULL, cw1, DUO_L);
        int pos[DUO_G], nnew = 0;
#pragma unroll
        for (int g = 0; g < DUO_G; g++)
        {
            pos[g] = nnew;
            if (((g < 2 ? cw0 : cw1) >> (16 * (g & 1))) & 0xffffu)   // warp-uniform
            {
                const unsigned f0 = __ballot_sync(FULL, cnt[g] & 1), f1 = __ballot_sync(FULL, cnt[g] & 2), f23 = __ballot_sync(FULL, cnt[g] & 12);
                const unsigned b0 = duo_half(f0, shift), b1 = duo_half(f1, shift);
                pos[g] += __popc(b0 & lt) + 2 * __popc(b1 & lt);
                nnew += __popc(b0) + 2 * __popc(b1);
                if (f23)   // a clipped vertex with four or more kept neighbours: rare
                {
                    const unsigned b2 = duo_half(__ballot_sync(FULL, cnt[g] & 4), shift), b3 = duo_half(__ballot_sync(FULL, cnt[g] & 8), shift);
                    pos[g] += 4 * __popc(b2 & lt) + 8 * __popc(b3 & lt);
                    nnew += 4 * __popc(b2) + 8 * __popc(b3);
                }
            }
        }
        const bool full = cut && hi + nnew > S;
        if (__any_sync(FULL, full))
        {
            // out of slots: renumber the live vertices (exactly the reference's compaction) and redo this plane
            const bool fits = full && mcount<2>(m.live) + nnew <= S;
            const int n = duo_compact(sp, m, sl, fits);
            if (full)
            {
                cut = false;
                nnew = 0;
                if (!fits) status = CLIP_OVERFLOW;
                else
                {
                    hi = n;
                    m.live[0] = lowmask32(hi); m.live[1] = lowmask32(hi - 32);
                }
            }
        }
        if (!cut) nnew = 0;
        if (cut) n_cuts++;
        const int hi0 = hi;
#pragma unroll
        for (int g = 0; g < DUO_G; g++)
        {
            unsigned mm = cut ? (smk >> (8 * g)) & 0xffu : 0u;
            int w = pos[g];
            while (mm) { const int j = __ffs(mm) - 1; mm &= mm - 1; sp.list[w++] = (uint16_t)((sl + DUO_L * g) | (j << 8)); }
        }
        __syncwarp();
        // insert: one new vertex per lane (Poly.cpp:345-354); lanes touch distinct BYTES of the ring words (clip_sub.cuh)
#pragma unroll 1
        for (int t = sl; t < nnew; t += DUO_L)
        {
            const int e = sp.list[t], v = e & 0xff, j = e >> 8, w = hi0 + t;
            const int jn = rget(sp.ring[v], j);
            const float ax = sp.x[v], ay = sp.y[v], az = sp.z[v], bx = sp.x[jn], by = sp.y[jn], bz = sp.z[jn];
            const float sa = signed_dist(cur, ax, ay, az), sb = signed_dist(cur, bx, by, bz);
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
        const bool inplane = ((m.live[0] & ~(m.c[0] | m.k[0])) | (m.live[1] & ~(m.c[1] | m.k[1]))) != 0u;
        bool need_seq = cut && inplane;   // uniform within the half
        bool ok = true;
#pragma unroll 1
        for (int t = sl; t < (need_seq ? 0 : nnew); t += DUO_L)
        {
            const int w = hi0 + t;
            // first step without a search: w sits in slot j of its clipped end point v, so FaceLoop(v, w) is the slot before j
            const int e = sp.list[t], v = e & 0xff, j = e >> 8;
            const u64 rv = sp.ring[v];
            int iprev = v, inext = rget(rv, (j == 0 ? rdeg(rv) : j) - 1), itmp, k = 1;
#pragma unroll 1
            while (inext < hi0 && mbit<2>(m.c, inext) && k++ < S)
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
        for (int t = sl; t < (need_seq ? 0 : nnew); t += DUO_L)
            if (ok) ok = sp.id[sp.list[t]] == (uint8_t)(hi0 + t);
        const unsigned walk_failed = duo_half(__ballot_sync(FULL, !ok), shift);   // (no short circuit around the collective)
        need_seq = need_seq || (cut && walk_failed != 0u);
        if (!need_seq)
        {
            // the walk targets are a permutation of the new vertices: ring(w) = [pusher, walked, kept]
#pragma unroll 1
            for (int t = sl; t < nnew; t += DUO_L)
            {
                const int w = hi0 + t;
                const int kept = rget(sp.ring[w], 1);
                sp.ring[w] = 0xffffffffff000000ull | (u64)sp.id[w] | ((u64)sp.list[t] << 8) | ((u64)(unsigned)kept << 16);
            }
        }
        unsigned dead0 = 0u, dead1 = 0u;
        if (__any_sync(FULL, need_seq))
        {
            const SeqResult<2> sr = duo_seq_cut(sp, m, hi0, nnew, sl, shift, need_seq);
            if (need_seq)
            {
                seq_cuts++;
                if (!sr.ok) { status = CLIP_OVERFLOW; cut = false; }
                dead0 = sr.dead[0]; dead1 = sr.dead[1];
            }
        }
        if (cut)
        {
            // lazy compaction: clipped vertices leave the live set, new ones join it
            hi = hi0 + nnew;
            const u64 fresh = ((nnew >= 64 ? 0ull : (1ull << nnew)) - 1ull) << (hi0 & 63);   // slots hi0 .. hi-1
            m.live[0] = ((m.live[0] & ~m.c[0]) | (unsigned)fresh) & ~dead0;
            m.live[1] = ((m.live[1] & ~m.c[1]) | (unsigned)(fresh >> 32)) & ~dead1;
            nv = __popc(m.live[0]) + __popc(m.live[1]);
            if (nv < 4) nv = 0;   // Poly.cpp:498-499
            cur = nxt;
            p = pn;
            pq.pop();
        }
        __syncwarp();         // ring words composed above are visible to the next cut
    }
    R.status = status;
    R.hi = hi;
    R.nv = nv;
    R.live[0] = m.live[0];
    R.live[1] = m.live[1];
    R.seq_cuts = seq_cuts;
    R.n_cuts = n_cuts;
}
} // namespace surtr
