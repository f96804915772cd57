// clip_sub.cuh -- small-tier clipper: L lanes (a warp, half-warp or quarter-warp) clip one pair; <= 64 vertex
// slots, ring degree <= 8, everything register/ballot based.
//
// Same algorithm and the same exactness argument as the other tiers (clip_warp.cuh):
//   * a vertex ring is ONE 64-bit shared-memory word: eight u8 neighbour indices, 0xFF padded.  FaceLoop
//     (Poly.cpp:34-41), find-and-replace (Poly.cpp:350-353) and the degree are byte-compare instructions on a
//     register (__vcmpeq4 / __ffs / PRMT) after a single LDS.64 -- no dependent chain of shared-memory reads;
//   * the per-vertex classification `comp` (Poly.cpp:303-319) never touches memory: the ballots of "clipped" and
//     "kept" are held by every lane as 64-bit masks and comp(j) is a bit test;
//   * new vertices are created one per lane from a list of straddling half-edges written in the reference's
//     append order (vertex ascending, ring slot ascending; Poly.cpp:333-357);
//   * compaction (Poly.cpp:464-499) is LAZY: clipped vertices just leave the live mask and new ones are appended.
//     The reference's compaction is stable, so the relative order of live vertices is the same with or without
//     it; indices are renumbered (popc on the live mask) only when the 64 slots run out and once at the end,
//     which yields exactly the reference's numbering.
// Sequential replays (in-plane vertices, walk anomalies, degree-2 splice) run on one lane over the same words.
//
// Lane count: ncu on the one-warp-per-pair version showed 6-14 of 32 lanes active in every phase but the
// classification (profiles/).  With L < 32 a warp works on 32/L pairs in LOCK STEP: every collective (ballot, shuffle,
// __syncwarp) is executed by all 32 lanes with the full member mask (shuffles confined to the pair's lanes by their
// width argument), every branch that contains a collective is decided by a warp-wide vote, and the pair-specific work
// inside is predicated.  (Per-pair member masks were measured first: they split the warp into serially executed
// groups at every collective and were slower than one warp per pair.)
#pragma once

#include "clip_warp.cuh"

namespace surtr
{
typedef unsigned long long u64;

struct SubPoly   // one per pair in shared memory (4032 bytes)
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

// Bit 7 of every zero byte of x (exact for the LOWEST zero byte, which is all the callers use; bytes above it may be
// flagged falsely).  The SIMD-in-a-word compare __vcmpeq4 is emulated on sm_100 (a dozen instructions per word).
__device__ __forceinline__ unsigned zero_bytes(unsigned x) { return (x - 0x01010101u) & ~x & 0x80808080u; }
__device__ __forceinline__ int first_flag(unsigned z) { return (__ffs((int)z) - 1) >> 3; }   // byte index of the lowest flag

__device__ __forceinline__ int rdeg(u64 w)
{
    // empty slots (0xFF) are the trailing bytes: the degree is the index of the first 0xFF byte; the low word decides for
    // every ring of up to three entries (the generic vertex)
    const unsigned zl = zero_bytes(~(unsigned)w);
    if (zl) return first_flag(zl);
    const unsigned zh = zero_bytes(~(unsigned)(w >> 32));
    return zh ? 4 + first_flag(zh) : 8;
}
__device__ __forceinline__ int rfind(u64 w, int val)   // first slot holding val, 8 if absent
{
    const unsigned pat = __byte_perm((unsigned)val, 0u, 0u);   // val in all four bytes
    const unsigned zl = zero_bytes((unsigned)w ^ pat);
    if (zl) return first_flag(zl);
    const unsigned zh = zero_bytes((unsigned)(w >> 32) ^ pat);
    return zh ? 4 + first_flag(zh) : 8;
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
// The degree is only needed for the wrap (vprev in slot 0) and for an absent vprev: two steps in three skip it.
__device__ __forceinline__ int rface_loop(u64 w, int vprev)
{
    int k = rfind(w, vprev);
    if (k == 0 || k == 8)
    {
        k = rdeg(w);
        if (k == 0) return vprev;   // malformed input (vertex without neighbours): callers' loop bounds end the walk
    }
    return rget(w, k - 1);
}

__device__ __forceinline__ bool bit64(u64 m, int j) { return (m >> j) & 1ull; }
// bits [0, n) set, 0 <= n <= 64, on 32-bit halves (a 64-bit shift costs a handful of instructions and is inlined at every use)
__device__ __forceinline__ u64 lowmask64(int n)
{
    const unsigned lo = n >= 32 ? 0xffffffffu : ((1u << (n & 31)) - 1u);
    const unsigned hi = n >= 64 ? 0xffffffffu : (n > 32 ? ((1u << (n & 31)) - 1u) : 0u);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ int rank64(u64 m, int u)   // set bits below u, 0 <= u < 64
{
    const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
    const unsigned below = (1u << (u & 31)) - 1u;
    return u < 32 ? __popc(lo & below) : __popc(lo) + __popc(hi & below);
}

struct CutState   // uniform across the lanes of one pair
{
    u64 live;   // live vertex slots
    u64 c;      // clipped by the current plane (comp == -1)
    u64 k;      // kept by the current plane (comp == +1)
    int hi;     // allocated vertex slots
};

// The L consecutive lanes of a warp that work on one pair.  All collectives use the full member mask (lock step).
template <int L>
struct Sub
{
    static constexpr int G = 64 / L;   // vertex slots per lane: lane sl owns slots sl, sl + L, ...
    int sl;
    unsigned shift;
    __device__ __forceinline__ explicit Sub(int lane)
    {
        sl = lane % L;
        shift = (unsigned)(lane / L) * L;
    }
    __device__ __forceinline__ unsigned ballot(bool p) const
    {
        const unsigned b = __ballot_sync(FULL, p);
        return L == 32 ? b : ((b >> shift) & ((1u << (L & 31)) - 1u));
    }
    __device__ __forceinline__ bool any_warp(bool p) const { return __any_sync(FULL, p); }   // warp-uniform vote
    __device__ __forceinline__ int max_warp(int v) const { return __reduce_max_sync(FULL, v); }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(FULL, v, src, L); }
    template <class T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(FULL, v, d, L); }
    template <class T> __device__ __forceinline__ T shfl_xor(T v, int d) const { return __shfl_xor_sync(FULL, v, d, L); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    template <class T> __device__ __forceinline__ T exscan(T v, T& total) const
    {
        T inc = v;
#pragma unroll
        for (int o = 1; o < L; o <<= 1)
        {
            const T t = shfl_up(inc, o);
            if (sl >= o) inc += t;
        }
        total = shfl(inc, L - 1);
        return inc - v;
    }
    __device__ __forceinline__ int sum(int v) const
    {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) v += shfl_xor(v, o);
        return v;
    }
};

// comp of the reference for the sequential replays: 2 = new, -1 clipped / gone, +1 kept, 0 in-plane
__device__ __forceinline__ int comp_of(const CutState& s, int hi0, u64 dead, int j)
{
    if (bit64(dead, j)) return -1;   // spliced away (Poly.cpp:459)
    if (j >= hi0) return 2;
    if (bit64(s.c, j) || !bit64(s.live, j)) return -1;
    return bit64(s.k, j) ? 1 : 0;
}

// Sequential replay of Poly.cpp:365-462 (patch, erase marks, degree-2 splice) by one lane after the new vertices
// have been inserted.  Visiting order = the reference's: new vertices first, then the pre-existing ones, both
// ascending.  Called by all lanes of the warp; only pairs with `pred` do anything.  Returns 0 on ring overflow;
// `dead` receives the vertices spliced away.
template <int L>
__device__ __noinline__ int sub_seq_cut(SubPoly& sp, const CutState s, int hi0, int nnew, const Sub<L> sub, bool pred, u64& dead)
{
    const int hi1 = pred ? hi0 + nnew : 0;
    for (int v = sub.sl; v < hi1; v += L) sp.old_ring[v] = sp.ring[v];
    sub.sync();
    int ok = 1;
    unsigned dlo = 0u, dhi = 0u;
    if (pred && sub.sl == 0)
    {
        u64 dd = 0ull;
        const int nverts = __popcll(s.live) + nnew;   // the reference's vertex count (walk bound)
        for (int ii = 0; ii < hi1 && ok; ii++)
        {
            const int i = ii < nnew ? hi0 + ii : ii - nnew;
            const int ci = comp_of(s, hi0, 0ull, i);
            if (!(ci == 0 || ci == 2)) continue;
            const int nneigh = rdeg(sp.ring[i]);
            for (int j = 0; j < nneigh; j++)
            {
                const int jn = rget(sp.ring[i], j);
                if (jn >= R_MARK || comp_of(s, hi0, 0ull, jn) != -1) continue;
                int iprev = i, inext = jn, itmp, k = 0;
                while (comp_of(s, hi0, 0ull, inext) == -1 && k++ < nverts)
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
                    if (comp_of(s, hi0, 0ull, inext) == 2) mark = R_MARK;   // Poly.cpp:409 inserts -1 in the snapshot
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
                if (comp_of(s, hi0, dd, i) >= 0 && rdeg(sp.ring[i]) == 2)
                {
                    updated = true;
                    const int iprev = rget(sp.ring[i], 0), inext = rget(sp.ring[i], 1);
                    int k = rfind(sp.ring[iprev], i);
                    if (k < rdeg(sp.ring[iprev])) sp.ring[iprev] = rset(sp.ring[iprev], k, inext);
                    k = rfind(sp.ring[inext], i);
                    if (k < rdeg(sp.ring[inext])) sp.ring[inext] = rset(sp.ring[inext], k, iprev);
                    dd |= 1ull << i;
                }
            }
        }
        dlo = (unsigned)dd;
        dhi = (unsigned)(dd >> 32);
    }
    ok = sub.shfl(ok, 0);
    dlo = sub.shfl(dlo, 0);
    dhi = sub.shfl(dhi, 0);
    dead = ((u64)dhi << 32) | dlo;
    sub.sync();
    return ok;
}

// Renumber the live vertices to 0..n-1 keeping their order (the reference's compaction, Poly.cpp:464-495).
// Called by all lanes of the warp; only pairs with `pred` do anything.  Positions are read from / written to shared
// memory; the caller reloads its register copies.
template <int L>
__device__ __noinline__ void sub_compact(SubPoly& sp, CutState& s, const Sub<L> sub, bool pred)
{
    constexpr int G = Sub<L>::G;
    u64 r[G];
    float vx[G], vy[G], vz[G];
#pragma unroll
    for (int g = 0; g < G; g++)
    {
        const int v = sub.sl + L * g;
        r[g] = ~0ull;
        vx[g] = vy[g] = vz[g] = 0.f;
        if (pred && bit64(s.live, v))
        {
            vx[g] = sp.x[v]; vy[g] = sp.y[v]; vz[g] = sp.z[v];
            const u64 rw = sp.ring[v];
            for (int j = 0; j < 8; j++)
            {
                const int b = rget(rw, j);
                if (b == R_NONE) break;
                r[g] = rset(r[g], j, rank64(s.live, b));
            }
        }
    }
    sub.sync();
#pragma unroll
    for (int g = 0; g < G; g++)
    {
        const int v = sub.sl + L * g;
        if (pred && bit64(s.live, v))
        {
            const int t = rank64(s.live, v);
            sp.x[t] = vx[g]; sp.y[t] = vy[g]; sp.z[t] = vz[g]; sp.ring[t] = r[g];
        }
    }
    sub.sync();
    if (pred)
    {
        const int n = __popcll(s.live);
        s.hi = n;
        s.live = lowmask64(n);
    }
}

// Every vertex in-plane: the reference's box test decides (Poly.cpp:297-299, 725-744).  Called by all lanes.
template <int L>
__device__ __noinline__ bool sub_all_inplane_box_says_skip(const SubPoly& sp, const CutState s, const float4 pl, const Sub<L> sub)
{
    float lo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
    float hi[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
    for (int v = sub.sl; v < s.hi; v += L)
    {
        if (!bit64(s.live, v)) continue;
        lo[0] = fminf(lo[0], sp.x[v]); hi[0] = fmaxf(hi[0], sp.x[v]);
        lo[1] = fminf(lo[1], sp.y[v]); hi[1] = fmaxf(hi[1], sp.y[v]);
        lo[2] = fminf(lo[2], sp.z[v]); hi[2] = fmaxf(hi[2], sp.z[v]);
    }
    for (int o = L / 2; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++)
        {
            lo[k] = fminf(lo[k], sub.shfl_xor(lo[k], o));
            hi[k] = fmaxf(hi[k], sub.shfl_xor(hi[k], o));
        }
    const int k = sub.sl & 7;   // L >= 8: every corner is tested by at least one lane
    const int c = classify(signed_dist(pl, (k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2]));
    return sub.ballot(c == -1) == 0u;
}

// Clip the polyhedron in `sp` (nv vertices in slots 0..nv-1) by planes[0..npl).  Called by all 32 lanes; a pair
// with alive == false just takes part in the collectives.  On return s.live are the live slots (not renumbered),
// nv their count (0 = no fragment) and the return value the pair's status.
template <int L>
__device__ int sub_clip_by_planes(SubPoly& sp, CutState& s, int& nv, const float4* __restrict__ planes, int npl, const Sub<L> sub,
                                  bool alive, unsigned& seq_cuts, unsigned& n_cuts)
{
    constexpr int G = Sub<L>::G;
    constexpr int FW = G <= 4 ? 32 / G : 8;                               // bits per group in the packed scans
    using ScanT = typename std::conditional<(G <= 4), unsigned, u64>::type;
    constexpr ScanT FM = (ScanT)((FW == 32) ? ~0u : ((1ull << (FW & 63)) - 1ull));
    float px[G], py[G], pz[G];
    if (!alive) { nv = 0; npl = 0; }
    s.hi = nv;
    s.live = lowmask64(nv);
    s.c = s.k = 0ull;
#pragma unroll
    for (int g = 0; g < G; g++)
    {
        const int v = sub.sl + L * g;
        px[g] = py[g] = pz[g] = 0.f;
        if (v < nv) { px[g] = sp.x[v]; py[g] = sp.y[v]; pz[g] = sp.z[v]; }
    }

    int status = CLIP_OK;
    float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
    int p = 0, loaded = -1;
    while (true)
    {
        const bool run = p < npl && nv > 0 && status == CLIP_OK;   // this pair still has a plane to apply
        if (!sub.any_warp(run)) break;
        if (run && p / L != loaded)
        {
            loaded = p / L;
            mine = make_float4(0.f, 0.f, 0.f, 0.f);
            if (loaded * L + sub.sl < npl) mine = __ldg(planes + loaded * L + sub.sl);   // lane sl holds plane loaded*L + sl
        }
        float4 pl;
        pl.x = sub.shfl(mine.x, p % L);
        pl.y = sub.shfl(mine.y, p % L);
        pl.z = sub.shfl(mine.z, p % L);
        pl.w = sub.shfl(mine.w, p % L);

        // ---- classify (Poly.cpp:303-319): one distance per owned live vertex, two ballots per group ----
        const int gmax = L == 32 ? (s.hi + L - 1) / L : sub.max_warp(run ? (s.hi + L - 1) / L : 0);
        s.c = s.k = 0ull;
#pragma unroll
        for (int g = 0; g < G; g++)
        {
            if (g < gmax)
            {
                int c = 3;
                if (run && bit64(s.live, sub.sl + L * g)) c = classify(signed_dist(pl, px[g], py[g], pz[g]));
                s.c |= (u64)sub.ballot(c == -1) << (g * L);
                s.k |= (u64)sub.ballot(c == 1) << (g * L);
            }
        }
        // "below" (Poly.cpp:322-327) -- unless every vertex is in-plane and the box test says "above"
        const bool need_box = run && !s.k && !s.c;
        bool box_skip = false;
        if (sub.any_warp(need_box)) box_skip = sub_all_inplane_box_says_skip<L>(sp, s, pl, sub);
        if (run && !s.k && !(need_box && box_skip)) nv = 0;
        const bool cut = run && s.k != 0ull && s.c != 0ull;   // otherwise "above" (Poly.cpp:328) or removed
        bool redo = false;

        if (sub.any_warp(cut))
        {
            // ---- the plane cuts (some pair of this warp) ----
            sub.sync();   // ring words composed by the previous cut are visible from here on
            // straddling half-edges (clipped vertex -> kept neighbour) in the reference's append order
            u64 smk = 0ull;    // 8 slot bits per owned group
            ScanT cnt = 0;     // FW-bit counter per owned group
            if (cut)
            {
#pragma unroll
                for (int g = 0; g < G; g++)
                {
                    const int v = sub.sl + L * g;
                    if (g * L < s.hi && bit64(s.c, v))
                    {
                        const u64 rw = sp.ring[v];
                        for (int j = 0; j < 8; j++)
                        {
                            const int b = rget(rw, j);
                            if (b == R_NONE) break;
                            if (bit64(s.k, b)) { smk |= 1ull << (j + 8 * g); cnt += (ScanT)1 << (FW * g); }
                        }
                    }
                }
            }
            ScanT tot;
            const ScanT ex = sub.exscan(cnt, tot);
            int nnew = 0;
#pragma unroll
            for (int g = 0; g < G; g++) nnew += (int)((tot >> (FW * g)) & FM);
            bool docut = cut;
            const bool full = cut && s.hi + nnew > 64;
            if (sub.any_warp(full))
            {
                // out of slots: renumber the live vertices (exactly the reference's compaction) and redo this plane
                const bool ovf = full && __popcll(s.live) + nnew > 64;
                if (ovf) status = CLIP_OVERFLOW;
                sub_compact<L>(sp, s, sub, full && !ovf);
                if (full && !ovf)
                {
#pragma unroll
                    for (int g = 0; g < G; g++)
                    {
                        const int v = sub.sl + L * g;
                        if (v < s.hi) { px[g] = sp.x[v]; py[g] = sp.y[v]; pz[g] = sp.z[v]; }
                    }
                    redo = true;
                }
                docut = cut && !full;
            }
            if (!docut) nnew = 0;
            if (docut) n_cuts++;
            const int hi0 = s.hi;
            if (docut && smk)
            {
                int gbase = 0;
#pragma unroll
                for (int g = 0; g < G; g++)
                {
                    unsigned m = (unsigned)(smk >> (8 * g)) & 0xffu;
                    int w = gbase + (int)((ex >> (FW * g)) & FM);
                    while (m) { const int j = __ffs(m) - 1; m &= m - 1; sp.list[w++] = (uint16_t)((sub.sl + L * g) | (j << 8)); }
                    gbase += (int)((tot >> (FW * g)) & FM);
                }
            }
            sub.sync();
            // insert: one new vertex per lane (Poly.cpp:345-354).  Lanes may touch the same ring WORD concurrently, but
            // never the same BYTE: lane t replaces exactly slot j of ring[v] and the slot of ring[jn] that holds v, and no
            // other lane reads or writes those two slots (another lane's search in ring[jn] looks for its own v' != v and
            // can only meet values that are neither v' before nor after this lane's store).  compute-sanitizer racecheck
            // reports these word-level overlaps as warnings (profiles/r1_sanitizer.txt); memcheck and synccheck are clean.
            for (int t = sub.sl; t < nnew; t += L)
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
            sub.sync();

            // patch (Poly.cpp:365-431): walk from each new vertex through clipped vertices to the next new one
            bool need_seq = docut && (s.live & ~(s.c | s.k)) != 0ull;   // an in-plane vertex
            bool ok = true;
            if (!need_seq)
            {
                for (int t = sub.sl; t < nnew; t += L)
                {
                    const int w = hi0 + t;
                    // first step without a search: w sits in slot j of its clipped end point v, so FaceLoop(v, w)
                    // is simply the slot before j (list[t] still holds v | j << 8 from the insertion)
                    const int e = sp.list[t], v = e & 0xff, j = e >> 8;
                    const u64 rv = sp.ring[v];
                    int iprev = v, inext = rget(rv, (j == 0 ? rdeg(rv) : j) - 1), itmp, k = 1;
                    while (inext < hi0 && bit64(s.c, inext) && k++ < 64)
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
            }
            sub.sync();
            if (!need_seq)
            {
#pragma unroll 1
                for (int t = sub.sl; t < nnew; t += L)
                    if (ok) ok = sp.id[sp.list[t]] == (uint8_t)(hi0 + t);
            }
            const bool anomaly = sub.ballot(!ok) != 0u;
            need_seq = need_seq || (docut && anomaly);
            if (docut && !need_seq)
            {
                // the walk targets are a permutation of the new vertices: ring(w) = [pusher, walked, kept]
                for (int t = sub.sl; t < nnew; t += L)
                {
                    const int w = hi0 + t;
                    const int kept = rget(sp.ring[w], 1);
                    sp.ring[w] = 0xffffffffff000000ull | (u64)sp.id[w] | ((u64)sp.list[t] << 8) | ((u64)(unsigned)kept << 16);
                }
            }
            u64 dead = 0ull;
            if (sub.any_warp(need_seq))
            {
                const int r = sub_seq_cut<L>(sp, s, hi0, nnew, sub, need_seq, dead);
                if (need_seq)
                {
                    seq_cuts++;
                    if (!r) status = CLIP_OVERFLOW;
                }
                else dead = 0ull;
            }
            if (docut)
            {
                // lazy compaction: clipped vertices leave the live set, new ones join it
                s.hi = hi0 + nnew;
                s.live = ((s.live & ~s.c) | (lowmask64(s.hi) & ~lowmask64(hi0))) & ~dead;
                nv = __popcll(s.live);
                if (nv < 4) nv = 0;   // Poly.cpp:498-499
#pragma unroll
                for (int g = 0; g < G; g++)
                {
                    const int v = sub.sl + L * g;
                    if (v >= hi0 && v < s.hi) { px[g] = sp.x[v]; py[g] = sp.y[v]; pz[g] = sp.z[v]; }
                }
            }
        }
        if (run && !redo) p++;
    }
    sub.sync();
    return status;
}

// Shared-memory image of one finished small-tier fragment for its face count and moments (K4 gather).
struct MomPoly   // 4096 bytes
{
    float x[64], y[64], z[64];
    u64 ring[64];          // 8 x u8, 0xFF = empty slot
    float4 tri[128];       // ordered fan-triangle records (dV, mx, my, mz)
    uint16_t flist[128];   // one entry per face, in Poly::ExtractFaces order: start vertex | start slot << 6 | first triangle << 9
    uint8_t fcnt[512];     // [vertex * 8 + slot]: fan triangles of the face that starts there
};

// Poly::ExtractFaces + Poly::Moments in the reference's accumulation order (Poly.cpp:55-126) + inertia, on the
// live (not renumbered) slots: vertex order = slot order, origin = first live vertex.  See DESIGN.md section 5
// for the derivation.  Called by all lanes of the warp; pairs without a fragment (`has` false) idle.
// Three steps: (1) lane = vertex: which ring slots start a face (the vertex is the smallest of the loop) and how many
// fan triangles it has; (2) two prefix sums give every face its position in ExtractFaces order and its first triangle
// slot, and the faces are LISTED; (3) lane = face: the fan triangles of a face are written to their slots.  Listing the
// faces balances step 3 -- the lowest-numbered vertices start three faces each and used to walk them one after another.
template <int L>
__device__ void sub_fragment_moments(MomPoly& sp, const CutState& s, const Sub<L> sub, bool has, Moments& out)
{
    constexpr int G = Sub<L>::G;
    const int first = has ? __ffsll((long long)s.live) - 1 : 0;
    const int gmax = L == 32 ? (s.hi + L - 1) / L : sub.max_warp(has ? (s.hi + L - 1) / L : 0);
    const float ox = sp.x[first], oy = sp.y[first], oz = sp.z[first];
    const int nv = has ? __popcll(s.live) : 0;
    u64 start_mask = 0ull;   // 8 slot bits per owned group
    int cntg[G], facg[G];    // fan triangles / faces started per owned group (one lane = one scan entry per group)
#pragma unroll
    for (int g = 0; g < G; g++) cntg[g] = facg[g] = 0;
    // (rolled: the body is large and runs once per fragment -- copies of it only cost instruction-cache space)
#pragma unroll 1
    for (int g = 0; g < G; g++)
    {
        int tris = 0, faces = 0;
        const int v = sub.sl + L * g;
        if (has && g * L < s.hi && bit64(s.live, v))
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
                    start_mask |= 1ull << (j + 8 * g);
                    faces++;
                    tris += max(n - 2, 0);
                    sp.fcnt[v * 8 + j] = (uint8_t)max(n - 2, 0);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < G; h++)
            if (h == g) { cntg[h] = tris; facg[h] = faces; }
    }
    // face and triangle order = vertex order = group-major, lane-minor: one scan pair per group that has vertices
    int n_tri = 0, n_faces = 0;
#pragma unroll 1
    for (int g = 0; g < G; g++)
    {
        if (g < gmax)
        {
            int tris = 0, faces = 0;
#pragma unroll
            for (int h = 0; h < G; h++)
                if (h == g) { tris = cntg[h]; faces = facg[h]; }
            int ttot, ftot;
            int tpos = n_tri + sub.exscan(tris, ttot);
            int fpos = n_faces + sub.exscan(faces, ftot);
            n_tri += ttot;
            n_faces += ftot;
            unsigned m = (unsigned)(start_mask >> (8 * g)) & 0xffu;
            const int v = sub.sl + L * g;
            while (m)
            {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                if (fpos < 128) sp.flist[fpos] = (uint16_t)(v | (j << 6) | (min(tpos, 127) << 9));
                fpos++;
                tpos += sp.fcnt[v * 8 + j];
            }
        }
    }
    n_tri = min(n_tri, 128);
    sub.sync();

    float cov[10] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };   // xx yy zz xy xz yz, 6V, first moments
    const int n_listed = has ? min(n_faces, 128) : 0;
#pragma unroll 1
    for (int t = sub.sl; t < n_listed; t += L)
    {
        const unsigned e = sp.flist[t];
        const int v = (int)(e & 63u), j = (int)((e >> 6) & 7u);
        int w = (int)(e >> 9);
        const float p0x = __fsub_rn(sp.x[v], ox), p0y = __fsub_rn(sp.y[v], oy), p0z = __fsub_rn(sp.z[v], oz);
        int prev = v, at = rget(sp.ring[v], j);
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
    sub.sync();

    // ordered accumulation (Poly.cpp:77-85): lane c < 4 owns component c of the triangle records (dV, mx, my, mz)
    // and adds them in the reference's order -- dV into a double, the first moments in float.  Four lanes share one
    // instruction stream, so the serial chain costs a quarter of a single-lane loop.
    double zeroth = 0.0;
    float fsum = 0.f;
    if (sub.sl < 4)
    {
        const float* comp = reinterpret_cast<const float*>(sp.tri) + sub.sl;
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
    zeroth = sub.shfl(zeroth, 0) / 6.0;
    float fx = sub.shfl(fsum, 1), fy = sub.shfl(fsum, 2), fz = sub.shfl(fsum, 3);
    {
        const double q = 24.0 * zeroth;
        const double inv = (q >= 0.0 ? 1.0 : -1.0) / fmax(1.0e-30, fabs(q));   // safeInv, Poly.cpp:33
        const float sc = (float)inv;
        fx = __fmul_rn(fx, sc); fy = __fmul_rn(fy, sc); fz = __fmul_rn(fz, sc);
    }
#pragma unroll
    for (int k = 0; k < 10; k++)
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1)
            cov[k] += sub.shfl_xor(cov[k], o);

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

// ------------------------------------------------------------------------------------------------------------------
// Round 2: face count + moments with every FaceLoop search done ONCE per directed edge.
//
// sub_fragment_moments walks face loops three times over (every local-minimum vertex probes its loops for a smaller
// vertex, then the fan pass walks each face again) and every step is a byte search in a ring word (rfind + rget:
// ~25 instructions): ~170 searches for a 16-vertex fragment, 40 % of K4's instructions (profiles/r2_k4_lines.txt).
// Here the successor of every directed edge on its face loop is computed once (phase 1: 3 V searches) into a table in
// shared memory -- edge id = the vertex's ring start + slot, entry = next edge | source vertex << 10 -- and the probe
// and the fan pass follow table entries (one LDS.U16 per step).  Order, operands and accumulation are unchanged:
// faces in Poly::ExtractFaces order (a face starts at its smallest vertex; vertices ascending, ring slots ascending),
// fan triangles (p0, p_k, p_k+1) written to their slot in that order, ordered accumulation by four lanes.
struct __align__(16) MomPoly2   // 3456 bytes per fragment: eight 128-thread blocks of the gather per SM
{
    float x[64], y[64], z[64];
    float4 tri[64];         // a window of 64 ordered fan-triangle records (dV, mx, my, mz); its first 512 bytes double as per-edge triangle counts before
    u64 ring[64];           // 8 x u8, 0xFF = empty slot (phases 1-2); then, as 256 u16: [0, 128) flist, one entry per face in
                            // Poly::ExtractFaces order, start edge | first triangle << 9; [128, 256) corners (v_k | v_k+1 << 8) of
                            // every fan-triangle slot
    uint16_t en[512];       // directed edge e = (v -> ring[v][j]), e = estart[v] + j: next edge of the face loop | v << 10
                            // (before phase 1: the staging area of the fragment's ring bytes, assemble_gather_kernel)
    uint16_t estart[64];    // first directed-edge id of vertex v = its ring start (phases 1-3); then, as 128 bytes: corner v0 of every fan-triangle slot
};

template <int L>
__device__ void sub_fragment_moments2(MomPoly2& sp, int nv, const Sub<L> sub, bool has, Moments& out)
{
    constexpr int G = Sub<L>::G;
    if (!has) nv = 0;
    const int gmax = L == 32 ? (nv + L - 1) / L : sub.max_warp((nv + L - 1) / L);
    const float ox = sp.x[0], oy = sp.y[0], oz = sp.z[0];
    uint8_t* ecnt = reinterpret_cast<uint8_t*>(sp.tri);   // fan triangles of the face that starts at edge e (phases 2-3 only)
    uint16_t* flist = reinterpret_cast<uint16_t*>(sp.ring);   // (the ring words are dead once phase 2 is through)

    // ---- phase 1: successor of every directed edge (FaceLoop, Poly.cpp:34-41): (v -> a) is followed by (a -> entry before v in ring[a]) ----
#pragma unroll 1
    for (int g = 0; g < G; g++)
    {
        const int v = sub.sl + L * g;
        if (g < gmax && v < nv)
        {
            const u64 rw = sp.ring[v];
            const int d = rdeg(rw), e0 = sp.estart[v];
#pragma unroll 1
            for (int j = 0; j < d; j++)
            {
                const int a = rget(rw, j);
                const u64 wa = sp.ring[a];
                int k = rfind(wa, v);
                if (k == 0 || k == 8) k = rdeg(wa);          // wrap, or v absent from ring[a] (malformed): the last entry, as FaceLoop does
                sp.en[e0 + j] = (uint16_t)((sp.estart[a] + (k > 0 ? k - 1 : 0)) | (v << 10));
            }
        }
    }
    sub.sync();

    // ---- phase 2: which edges start a face (their source is the smallest vertex of the loop), and its fan-triangle count ----
    u64 start_mask = 0ull;   // 8 slot bits per owned group
    int cntg[G], facg[G];
#pragma unroll
    for (int g = 0; g < G; g++) cntg[g] = facg[g] = 0;
#pragma unroll 1
    for (int g = 0; g < G; g++)
    {
        int tris = 0, faces = 0;
        const int v = sub.sl + L * g;
        if (g < gmax && v < nv)
        {
            const u64 rw = sp.ring[v];
            const int d = rdeg(rw), e0 = sp.estart[v];
#pragma unroll 1
            for (int j = 0; j < d; j++)
            {
                // v can only be the smallest vertex of the loop through (v -> a) if both loop neighbours of v are larger:
                // a, and the vertex the loop arrives from = the ring entry after a (FaceLoop inverted)
                if (rget(rw, j) < v || rget(rw, j + 1 == d ? 0 : j + 1) < v) continue;
                unsigned e = sp.en[e0 + j] & 1023u;
                int n = 1;
                bool is_start = true;
                while (true)
                {
                    const unsigned en = sp.en[e];
                    const int at = (int)(en >> 10);
                    if (at == v) break;
                    if (at < v || n > nv) { is_start = false; break; }
                    e = en & 1023u;
                    n++;
                }
                if (is_start)
                {
                    start_mask |= 1ull << (j + 8 * g);
                    faces++;
                    tris += max(n - 2, 0);
                    ecnt[e0 + j] = (uint8_t)max(n - 2, 0);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < G; h++)
            if (h == g) { cntg[h] = tris; facg[h] = faces; }
    }

    sub.sync();   // every lane is through with the ring words: the face list takes their place

    // ---- phase 3: list the faces in order (vertex order = group-major, lane-minor; slots ascending) with their first triangle slot ----
    int n_tri = 0, n_faces = 0;
#pragma unroll 1
    for (int g = 0; g < G; g++)
    {
        if (g < gmax)
        {
            int tris = 0, faces = 0;
#pragma unroll
            for (int h = 0; h < G; h++)
                if (h == g) { tris = cntg[h]; faces = facg[h]; }
            int ttot, ftot;
            int tpos = n_tri + sub.exscan(tris, ttot);
            int fpos = n_faces + sub.exscan(faces, ftot);
            n_tri += ttot;
            n_faces += ftot;
            unsigned m = (unsigned)(start_mask >> (8 * g)) & 0xffu;
            const int v = sub.sl + L * g;
            const int e0 = m ? sp.estart[v] : 0;
            while (m)
            {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                if (fpos < 128) flist[fpos] = (uint16_t)((e0 + j) | (min(tpos, 127) << 9));
                fpos++;
                tpos += ecnt[e0 + j];
            }
        }
    }
    n_tri = min(n_tri, 128);
    sub.sync();

    // ---- phase 4a: lane = face: one walk around the loop records the corners of every fan triangle at its slot ----
    // (v0 in the bytes of estart, (v_k, v_k+1) in the upper half of the ring words: both are dead by now.)  The fan
    // arithmetic itself then runs with lane = TRIANGLE (phase 4b): a face of eight vertices no longer holds the other
    // lanes up for six rounds of cross products.
    const int n_listed = has ? min(n_faces, 128) : 0;
    uint8_t* tv0 = reinterpret_cast<uint8_t*>(sp.estart);
    uint16_t* tpair = reinterpret_cast<uint16_t*>(sp.ring) + 128;
#pragma unroll 1
    for (int t = sub.sl; t < n_listed; t += L)
    {
        const unsigned f = flist[t];
        int w = (int)(f >> 9);
        unsigned en = sp.en[f & 511u];
        const int v = (int)(en >> 10);
        en = sp.en[en & 1023u];
        int prev = (int)(en >> 10);
        en = sp.en[en & 1023u];
        int at = (int)(en >> 10);
        int guard = 0;
        while (at != v && guard++ < 64)
        {
            if (w < 128) { tv0[w] = (uint8_t)v; tpair[w] = (uint16_t)(prev | (at << 8)); }
            w++;
            prev = at;
            en = sp.en[en & 1023u];
            at = (int)(en >> 10);
        }
    }
    sub.sync();

    // ---- phases 4b + 5, over windows of 64 triangle slots (one window for fragments of up to 34 vertices) ----
    // phase 4b: lane = fan triangle (p0, p_k, p_k+1): its record (dV, first moments) goes to its slot, second moments on the fly
    // phase 5: ordered accumulation (Poly.cpp:77-85), as sub_fragment_moments
    float cov[10] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };   // xx yy zz xy xz yz, 6V, first moments
    double zeroth = 0.0;
    float fsum = 0.f;
    const int n_win = L == 32 ? n_tri : sub.max_warp(n_tri);
#pragma unroll 1
    for (int base = 0; base == 0 || base < n_win; base += 64)
    {
        const int w_end = has ? min(n_tri, base + 64) : 0;
#pragma unroll 1
        for (int w = base + sub.sl; w < w_end; w += L)
        {
            const int v0 = tv0[w] & 63, pr = tpair[w], v1 = pr & 63, v2 = (pr >> 8) & 63;
            const float p0x = __fsub_rn(sp.x[v0], ox), p0y = __fsub_rn(sp.y[v0], oy), p0z = __fsub_rn(sp.z[v0], oz);
            const float p1x = __fsub_rn(sp.x[v1], ox), p1y = __fsub_rn(sp.y[v1], oy), p1z = __fsub_rn(sp.z[v1], oz);
            const float p2x = __fsub_rn(sp.x[v2], ox), p2y = __fsub_rn(sp.y[v2], oy), p2z = __fsub_rn(sp.z[v2], oz);
            float cx, cy, cz;
            cross3(p1x, p1y, p1z, p2x, p2y, p2z, cx, cy, cz);
            const float dV = dot3(p0x, p0y, p0z, cx, cy, cz);
            const float sx = __fadd_rn(__fadd_rn(p0x, p1x), p2x);
            const float sy = __fadd_rn(__fadd_rn(p0y, p1y), p2y);
            const float sz = __fadd_rn(__fadd_rn(p0z, p1z), p2z);
            sp.tri[w - base] = make_float4(dV, __fmul_rn(sx, dV), __fmul_rn(sy, dV), __fmul_rn(sz, dV));
            // Second moments of the tetrahedron (origin, p0, p1, p2): dV/120 * (s s^T + sum p p^T).  No reference arithmetic
            // to match here (PhysX is absent): fused multiply-adds are fine, unlike in dV and the
            // first moments above.
            cov[0] = fmaf(dV, fmaf(sx, sx, fmaf(p2x, p2x, fmaf(p1x, p1x, p0x * p0x))), cov[0]);
            cov[1] = fmaf(dV, fmaf(sy, sy, fmaf(p2y, p2y, fmaf(p1y, p1y, p0y * p0y))), cov[1]);
            cov[2] = fmaf(dV, fmaf(sz, sz, fmaf(p2z, p2z, fmaf(p1z, p1z, p0z * p0z))), cov[2]);
            cov[3] = fmaf(dV, fmaf(sx, sy, fmaf(p2x, p2y, fmaf(p1x, p1y, p0x * p0y))), cov[3]);
            cov[4] = fmaf(dV, fmaf(sx, sz, fmaf(p2x, p2z, fmaf(p1x, p1z, p0x * p0z))), cov[4]);
            cov[5] = fmaf(dV, fmaf(sy, sz, fmaf(p2y, p2z, fmaf(p1y, p1z, p0y * p0z))), cov[5]);
            cov[6] += dV;
            cov[7] = fmaf(dV, sx, cov[7]); cov[8] = fmaf(dV, sy, cov[8]); cov[9] = fmaf(dV, sz, cov[9]);
        }
        sub.sync();
        if (sub.sl < 4)
        {
            const float* comp = reinterpret_cast<const float*>(sp.tri) + sub.sl;
            const int n_here = min(64, n_tri - base);
            int t = 0;
            for (; t + 4 <= n_here; t += 4)   // the loads do not depend on the accumulation chain
            {
                const float r0 = comp[4 * t], r1 = comp[4 * t + 4], r2 = comp[4 * t + 8], r3 = comp[4 * t + 12];
                zeroth += (double)r0; zeroth += (double)r1; zeroth += (double)r2; zeroth += (double)r3;
                fsum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(fsum, r0), r1), r2), r3);
            }
            for (; t < n_here; t++)
            {
                const float r = comp[4 * t];
                zeroth += (double)r;
                fsum = __fadd_rn(fsum, r);
            }
        }
        if (base + 64 < n_win) sub.sync();   // the next window overwrites the records
    }
    zeroth = sub.shfl(zeroth, 0) / 6.0;
    float fx = sub.shfl(fsum, 1), fy = sub.shfl(fsum, 2), fz = sub.shfl(fsum, 3);
    {
        const double q = 24.0 * zeroth;
        const double inv = (q >= 0.0 ? 1.0 : -1.0) / fmax(1.0e-30, fabs(q));   // safeInv, Poly.cpp:33
        const float sc = (float)inv;
        fx = __fmul_rn(fx, sc); fy = __fmul_rn(fy, sc); fz = __fmul_rn(fz, sc);
    }
#pragma unroll
    for (int k = 0; k < 10; k++)
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1)
            cov[k] += sub.shfl_xor(cov[k], o);

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
