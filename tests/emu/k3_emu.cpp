// (second entry point below: the large / unbounded tier, global_clip_by_planes<NW> + global_fragment_moments<NW> of
// clip_global.cuh, as a block of NW warps -- NW = 4 is clip_shared_kernel, NW = 8 clip_global_kernel, NW = 1 one warp)
// k3_emu.cpp -- TEST INFRASTRUCTURE: the small tier of K3 (sub_clip_by_planes<32>, surtr_b200/csrc/clip_sub.cuh) and the
// K4 moments (sub_fragment_moments<16>, two fragments per warp in lock step, as assemble_gather_kernel runs them) compiled
// for the HOST over the SIMT shim, behind a flat C interface for tests/test_k3_emulation.py.  The staging of a pair
// into the shared-memory image and the write-out by rank in the live mask restate the few lines of clip_sub_kernel /
// assemble_gather_kernel around those calls (surtr_b200/csrc/kernels.cuh).
#include "simt_shim.h"

#include <memory>

#define __noinline__ __attribute__((noinline))
#include "../../surtr_b200/csrc/clip_sub.cuh"
#include "../../surtr_b200/csrc/clip_fast.cuh"
#include "../../surtr_b200/csrc/clip_duo.cuh"
#include "../../surtr_b200/csrc/clip_global.cuh"
#undef __noinline__

using namespace surtr;

extern "C"
{
// Thread order between collectives: 0 ascending, 1 descending, other = seeded random permutation per pass.
void k3emu_set_schedule(unsigned mode) { simt::schedule() = mode; }

}   // extern "C"

// Which small-tier clipper k3emu_pair runs: 2 / 4 = fast_clip_by_planes<G> of clip_fast.cuh (64 / 128 vertex slots, the
// throughput build: positions from shared memory, no plane prefetch), 22 = its latency build (64 slots, positions of the
// lane's own slots in registers, next plane loaded one iteration ahead: clip_fast_kernel<..., LAT = true>); (the
// kernels that ship, clip_fast_kernel<2,false> and <4,true>), 16 = duo_clip_by_planes of clip_duo.cuh (two pairs per warp,
// clip_duo_kernel), 0 = sub_clip_by_planes<32> of clip_sub.cuh (round 1).
static int g_variant = 2;
extern "C" void k3emu_set_variant(int v) { g_variant = v; }

// The staging, clip and write-out of fast_pair<G> (kernels.cuh) for one pair.  Returns < 0 on an emulation inconsistency.
template <int G, int RG = SURTR_K3_REG_GROUPS, bool PF = (SURTR_K3_PREFETCH != 0)>
static int fast_pair_emu(const float* verts4, const uint32_t* ring_off, const uint16_t* ring, int nv_in, const std::vector<float4>& planes,
                         int npl, float* out_verts4, uint32_t* out_ring_off, uint16_t* out_ring, int* out_info)
{
    constexpr int S = 32 * G;
    auto sp = std::make_unique<FastPoly<G>>();
    bool bad = nv_in > S;
    float px[32][G], py[32][G], pz[32][G];
    if (!bad)
        for (int v = 0; v < nv_in; v++)
        {
            const int d = (int)(ring_off[v + 1] - ring_off[v]);
            sp->x[v] = verts4[4 * v]; sp->y[v] = verts4[4 * v + 1]; sp->z[v] = verts4[4 * v + 2];
            u64 rw = ~0ull;
            if (d > 8 || d == 0) bad = true;
            else
                for (int j = 0; j < d; j++)
                {
                    const int idx = ring[ring_off[v] + j];
                    bad = bad || idx >= nv_in;
                    rw = rset(rw, j, idx);
                }
            sp->ring[v] = rw;
        }
    if (bad) { out_info[0] = CLIP_OVERFLOW; return 0; }
    for (int l = 0; l < 32; l++)
        for (int g = 0; g < G; g++)
        {
            const int v = l + 32 * g;
            px[l][g] = py[l][g] = pz[l][g] = 0.f;
            if (v < nv_in) { px[l][g] = sp->x[v]; py[l][g] = sp->y[v]; pz[l][g] = sp->z[v]; }
        }
    unsigned live[32][G];
    int hi[32], nv[32], status[32];
    unsigned seq[32], cuts[32];
    // K1's box of the piece (kdop_extents_kernel: plain min / max of the coordinates)
    float box[6] = { 3.402823466e+38f, -3.402823466e+38f, 3.402823466e+38f, -3.402823466e+38f, 3.402823466e+38f, -3.402823466e+38f };
    for (int v = 0; v < nv_in; v++)
        for (int k = 0; k < 3; k++)
        {
            box[2 * k] = std::min(box[2 * k], verts4[4 * v + k]);
            box[2 * k + 1] = std::max(box[2 * k + 1], verts4[4 * v + k]);
        }
    const unsigned long n_coll = simt::run_warp([&](int lane) {
        nv[lane] = nv_in;
        seq[lane] = cuts[lane] = 0;
        hi[lane] = 0;
        status[lane] = fast_clip_by_planes<G, RG, PF>(*sp, live[lane], hi[lane], nv[lane], px[lane], py[lane], pz[lane], planes.data(), npl, lane,
                                              seq[lane], cuts[lane], box, true);
    });
    for (int l = 1; l < 32; l++)   // warp-uniform by construction
    {
        if (hi[l] != hi[0] || nv[l] != nv[0] || status[l] != status[0] || seq[l] != seq[0]) return -1;
        if (status[0] == CLIP_OK && nv[0] > 0)
            for (int g = 0; g < G; g++)
                if (live[l][g] != live[0][g]) return -1;
    }
    out_info[0] = status[0];
    out_info[3] = (int)seq[0];
    out_info[4] = (int)cuts[0];
    out_info[5] = (int)n_coll;
    if (status[0] == CLIP_OK && nv[0] > 64) out_info[0] = CLIP_OVERFLOW;   // fast_pair: the result must fit the small blob
    if (out_info[0] != CLIP_OK || nv[0] == 0) return 0;
    // write-out of fast_pair: final number of a live slot = its rank in the live mask; positions from the owner lane's registers
    int ne = 0, n = 0;
    for (int v = 0; v < hi[0]; v++)
    {
        if (!mbit<G>(live[0], v)) continue;
        const int t = mrank<G>(live[0], v);
        if (t != n) return -2;
        const int l = v & 31, g = v >> 5;
        if (g < RG && px[l][g] != sp->x[v] && !(px[l][g] != px[l][g])) return -6;   // the register copy is the shared-memory copy
        if (g < RG) { out_verts4[4 * t] = px[l][g]; out_verts4[4 * t + 1] = py[l][g]; out_verts4[4 * t + 2] = pz[l][g]; }
        else { out_verts4[4 * t] = sp->x[v]; out_verts4[4 * t + 1] = sp->y[v]; out_verts4[4 * t + 2] = sp->z[v]; }
        out_verts4[4 * t + 3] = 0.f;
        out_ring_off[t] = (uint32_t)ne;
        const u64 rw = sp->ring[v];
        const int d = rdeg(rw);
        for (int j = 0; j < d; j++)
        {
            const int nb = rget(rw, j);
            if (nb >= S || !mbit<G>(live[0], nb)) return -3;   // a ring entry that points at a dead slot
            out_ring[ne++] = (uint16_t)mrank<G>(live[0], nb);
        }
        n++;
    }
    out_ring_off[n] = (uint32_t)ne;
    if (n != nv[0]) return -4;
    out_info[1] = n;
    out_info[2] = ne;
    return 0;
}

// The two-pairs-per-warp clipper (duo_clip_by_planes of clip_duo.cuh, clip_duo_kernel).  The pair handed in shares the
// emulated warp with the pair of the PREVIOUS call (alternating between the lower and the upper half), so every pair is
// cut twice -- next to two different neighbours, once in each half -- and the second result must equal the first bit for
// bit; the caller compares the first with the oracle.  Staging and write-out restate duo_pair (kernels.cuh).
namespace
{
struct DuoCase
{
    bool valid = false;
    std::vector<float> verts4;
    std::vector<uint32_t> ring_off;
    std::vector<uint16_t> ring;
    std::vector<float4> planes;
    int nv_in = 0, npl = 0;
    // result
    int status = 0, nv = 0, ne = 0;
    unsigned seq = 0, cuts = 0;
    std::vector<float> out_verts4;
    std::vector<uint32_t> out_ring_off;
    std::vector<uint16_t> out_ring;
};
DuoCase g_duo_prev;
unsigned g_duo_calls = 0;

bool duo_stage(const DuoCase& c, FastPoly<2>& sp, float (&box)[6])
{
    bool bad = c.nv_in > 64;
    if (!bad)
        for (int v = 0; v < c.nv_in; v++)
        {
            const int d = (int)(c.ring_off[v + 1] - c.ring_off[v]);
            sp.x[v] = c.verts4[4 * v]; sp.y[v] = c.verts4[4 * v + 1]; sp.z[v] = c.verts4[4 * v + 2];
            u64 rw = ~0ull;
            if (d > 8 || d == 0) bad = true;
            else
                for (int j = 0; j < d; j++)
                {
                    const int idx = c.ring[c.ring_off[v] + j];
                    bad = bad || idx >= c.nv_in;
                    rw = rset(rw, j, idx);
                }
            sp.ring[v] = rw;
        }
    const float big = 3.402823466e+38f;
    box[0] = box[2] = box[4] = big; box[1] = box[3] = box[5] = -big;
    for (int v = 0; v < c.nv_in; v++)
        for (int k = 0; k < 3; k++)
        {
            box[2 * k] = std::min(box[2 * k], c.verts4[4 * v + k]);
            box[2 * k + 1] = std::max(box[2 * k + 1], c.verts4[4 * v + k]);
        }
    return !bad;
}

// write-out of duo_pair: final number of a live slot = its rank in the live mask
int duo_write(DuoCase& c, const FastPoly<2>& sp, const DuoResult& R)
{
    c.status = R.status; c.nv = 0; c.ne = 0; c.seq = R.seq_cuts; c.cuts = R.n_cuts;
    c.out_verts4.assign(64 * 4, 0.f); c.out_ring_off.assign(65, 0u); c.out_ring.assign(512, 0);
    if (R.status != CLIP_OK || R.nv == 0) return 0;
    int ne = 0, n = 0;
    for (int v = 0; v < R.hi; v++)
    {
        if (!mbit<2>(R.live, v)) continue;
        const int t = mrank<2>(R.live, v);
        if (t != n) return -2;
        c.out_verts4[4 * t] = sp.x[v]; c.out_verts4[4 * t + 1] = sp.y[v]; c.out_verts4[4 * t + 2] = sp.z[v];
        c.out_ring_off[t] = (uint32_t)ne;
        const u64 rw = sp.ring[v];
        const int d = rdeg(rw);
        for (int j = 0; j < d; j++)
        {
            const int nb = rget(rw, j);
            if (nb >= 64 || !mbit<2>(R.live, nb)) return -3;   // a ring entry that points at a dead slot
            c.out_ring[ne++] = (uint16_t)mrank<2>(R.live, nb);
        }
        n++;
    }
    c.out_ring_off[n] = (uint32_t)ne;
    if (n != R.nv) return -4;
    c.nv = n; c.ne = ne;
    return 0;
}

int duo_pair_emu(const float* verts4, const uint32_t* ring_off, const uint16_t* ring, int nv_in, const std::vector<float4>& planes, int npl,
                 float* out_verts4, uint32_t* out_ring_off, uint16_t* out_ring, int* out_info)
{
    DuoCase cur;
    cur.valid = true;
    cur.nv_in = nv_in; cur.npl = npl;
    cur.verts4.assign(verts4, verts4 + 4 * (size_t)nv_in);
    cur.ring_off.assign(ring_off, ring_off + nv_in + 1);
    cur.ring.assign(ring, ring + ring_off[nv_in]);
    cur.planes = planes;
    const int hc = (int)(g_duo_calls++ & 1u);        // the half the new pair runs in; the previous pair takes the other
    DuoCase* cs[2];
    DuoCase again = g_duo_prev;
    cs[hc] = &cur; cs[1 - hc] = &again;
    auto sp = std::make_unique<FastPoly<2>[]>(2);
    float box[2][6] = {};
    bool act[2];
    for (int h = 0; h < 2; h++) act[h] = cs[h]->valid && duo_stage(*cs[h], sp[h], box[h]);
    DuoResult R[32];
    const unsigned long n_coll = simt::run_warp([&](int lane) {
        const int h = lane >> 4;
        duo_clip_by_planes(sp[h], act[h], cs[h]->nv_in, cs[h]->planes.data(), cs[h]->npl, box[h], true, lane, R[lane]);
    });
    for (int h = 0; h < 2; h++)
        for (int l = 1; l < 16; l++)   // uniform within a half by construction
        {
            const DuoResult &a = R[16 * h], &b = R[16 * h + l];
            if (a.status != b.status || a.hi != b.hi || a.nv != b.nv || a.seq_cuts != b.seq_cuts || a.n_cuts != b.n_cuts) return -1;
            if (a.status == CLIP_OK && a.nv > 0 && (a.live[0] != b.live[0] || a.live[1] != b.live[1])) return -1;
        }
    for (int h = 0; h < 2; h++)
    {
        if (!cs[h]->valid) continue;
        if (!act[h]) { cs[h]->status = CLIP_OVERFLOW; cs[h]->nv = cs[h]->ne = 0; continue; }
        const int rc = duo_write(*cs[h], sp[h], R[16 * h]);
        if (rc) return rc;
    }
    if (again.valid)   // the previous pair, cut again next to a different neighbour and in the other half: the same fragment
    {
        const DuoCase& p = g_duo_prev;
        if (again.status != p.status || again.nv != p.nv || again.ne != p.ne || again.seq != p.seq) return -7;
        if (again.status == CLIP_OK && again.nv > 0 &&
            (std::memcmp(again.out_verts4.data(), p.out_verts4.data(), 16 * (size_t)p.nv) || again.out_ring_off != p.out_ring_off || again.out_ring != p.out_ring))
            return -7;
    }
    out_info[0] = cur.status;
    out_info[3] = (int)cur.seq;
    out_info[4] = (int)cur.cuts;
    out_info[5] = (int)n_coll;
    if (cur.status == CLIP_OK && cur.nv > 0)
    {
        std::memcpy(out_verts4, cur.out_verts4.data(), 16 * (size_t)cur.nv);
        std::memcpy(out_ring_off, cur.out_ring_off.data(), 4 * ((size_t)cur.nv + 1));
        std::memcpy(out_ring, cur.out_ring.data(), 2 * (size_t)cur.ne);
        out_info[1] = cur.nv;
        out_info[2] = cur.ne;
    }
    g_duo_prev = cur;
    return 0;
}
} // namespace

extern "C"
{
// One (piece, plane list) pair through the device code of the small tier.
//   verts4[nv_in][4], ring_off[nv_in + 1] (relative to ring[0]), ring[...]: the piece;  planes4[npl][4]: the cell.
//   out_verts4[64][4], out_ring_off[65], out_ring[512]: the fragment, numbered as the kernel writes it.
//   out_info[8]: status (0 ok, 2 overflow = the pair would go to the large tier), nv, ne, sequential cuts, cuts,
//                collectives executed, n_faces, 0.
//   out_volume[1], out_centroid[3], out_inertia[6]: the K4 record fields (only when nv > 0).
int k3emu_pair(const float* verts4, const uint32_t* ring_off, const uint16_t* ring, int nv_in, const float* planes4, int npl,
               float* out_verts4, uint32_t* out_ring_off, uint16_t* out_ring, int* out_info, double* out_volume, float* out_centroid,
               float* out_inertia)
{
    for (int k = 0; k < 8; k++) out_info[k] = 0;
    std::vector<float4> planes(std::max(npl, 1));
    for (int p = 0; p < npl; p++) planes[p] = make_float4(planes4[4 * p], planes4[4 * p + 1], planes4[4 * p + 2], planes4[4 * p + 3]);
    int n = 0;
    if (g_variant != 0)
    {
        const int rc = g_variant == 22 ? fast_pair_emu<2, 2, true>(verts4, ring_off, ring, nv_in, planes, npl, out_verts4, out_ring_off, out_ring, out_info)
                       : g_variant == 16 ? duo_pair_emu(verts4, ring_off, ring, nv_in, planes, npl, out_verts4, out_ring_off, out_ring, out_info)
                       : g_variant == 4 ? fast_pair_emu<4>(verts4, ring_off, ring, nv_in, planes, npl, out_verts4, out_ring_off, out_ring, out_info)
                                      : fast_pair_emu<2>(verts4, ring_off, ring, nv_in, planes, npl, out_verts4, out_ring_off, out_ring, out_info);
        if (rc) return rc;
        if (out_info[0] != CLIP_OK || out_info[1] == 0) return 0;
        n = out_info[1];
    }
    else
    {
    auto sp = std::make_unique<SubPoly>();
    bool bad = nv_in > 64;
    if (!bad)
        for (int v = 0; v < nv_in; v++)   // clip_sub_kernel, staging (kernels.cuh): positions + one ring word per vertex
        {
            const int d = (int)(ring_off[v + 1] - ring_off[v]);
            sp->x[v] = verts4[4 * v]; sp->y[v] = verts4[4 * v + 1]; sp->z[v] = verts4[4 * v + 2];
            u64 rw = ~0ull;
            if (d > 8 || d == 0) bad = true;
            else
                for (int j = 0; j < d; j++)
                {
                    const int idx = ring[ring_off[v] + j];
                    bad = bad || idx >= nv_in;
                    rw = rset(rw, j, idx);
                }
            sp->ring[v] = rw;
        }
    if (bad) { out_info[0] = CLIP_OVERFLOW; return 0; }

    CutState cs[32];
    int nv[32], status[32];
    unsigned seq[32], cuts[32];
    const unsigned long n_coll = simt::run_warp([&](int lane) {
        const Sub<32> sub(lane);
        nv[lane] = nv_in;
        seq[lane] = cuts[lane] = 0;
        status[lane] = sub_clip_by_planes<32>(*sp, cs[lane], nv[lane], planes.data(), npl, sub, true, seq[lane], cuts[lane]);
    });
    for (int l = 1; l < 32; l++)   // warp-uniform by construction
        if (cs[l].live != cs[0].live || cs[l].hi != cs[0].hi || nv[l] != nv[0] || status[l] != status[0] || seq[l] != seq[0])
            return -1;
    out_info[0] = status[0];
    out_info[3] = (int)seq[0];
    out_info[4] = (int)cuts[0];
    out_info[5] = (int)n_coll;
    if (status[0] != CLIP_OK || nv[0] == 0) return 0;

    // write-out of clip_sub_kernel: final number of a live slot = its rank in the live mask
    const u64 live = cs[0].live;
    int ne = 0;
    for (int v = 0; v < cs[0].hi; v++)
    {
        if (!bit64(live, v)) continue;
        const int t = rank64(live, v);
        if (t != n) return -2;
        out_verts4[4 * t] = sp->x[v]; out_verts4[4 * t + 1] = sp->y[v]; out_verts4[4 * t + 2] = sp->z[v]; out_verts4[4 * t + 3] = 0.f;
        out_ring_off[t] = (uint32_t)ne;
        const u64 rw = sp->ring[v];
        const int d = rdeg(rw);
        for (int j = 0; j < d; j++)
        {
            const int nb = rget(rw, j);
            if (nb >= 64 || !bit64(live, nb)) return -3;   // a ring entry that points at a dead slot
            out_ring[ne++] = (uint16_t)rank64(live, nb);
        }
        n++;
    }
    out_ring_off[n] = (uint32_t)ne;
    if (n != nv[0]) return -4;
    out_info[1] = n;
    out_info[2] = ne;
    }

    // K4 (assemble_gather_kernel, tier 1): the fragment is rebuilt in shared memory numbered 0..nv-1 and its face count,
    // volume, centroid and inertia come from sub_fragment_moments<16>, two fragments per warp in lock step.  Both halves
    // of the emulated warp get this fragment; they must agree.
    auto mp = std::make_unique<MomPoly2[]>(2);
    for (int h = 0; h < 2; h++)
        for (int v = 0; v < n; v++)
        {
            mp[h].x[v] = out_verts4[4 * v]; mp[h].y[v] = out_verts4[4 * v + 1]; mp[h].z[v] = out_verts4[4 * v + 2];
            u64 rw = ~0ull;
            const int r0 = (int)out_ring_off[v], r1 = (int)out_ring_off[v + 1];
            for (int j = 0; j < r1 - r0 && j < 8; j++) rw = rset(rw, j, out_ring[r0 + j]);
            mp[h].ring[v] = rw;
            mp[h].estart[v] = (uint16_t)r0;
        }
    Moments mo[32];
    simt::run_warp([&](int lane) {
        const Sub<16> sub(lane);
        if (sub.any_warp(true))
        {
            sub.sync();
            sub_fragment_moments2<16>(mp[lane / 16], n, sub, true, mo[lane]);
        }
    });
    for (int l = 1; l < 32; l++)
        if (std::memcmp(&mo[l].volume, &mo[0].volume, 8) || mo[l].n_faces != mo[0].n_faces || std::memcmp(&mo[l].cx, &mo[0].cx, 12) ||
            std::memcmp(mo[l].inertia, mo[0].inertia, 24))
            return -5;
    out_info[6] = mo[0].n_faces;
    *out_volume = mo[0].volume;
    out_centroid[0] = mo[0].cx; out_centroid[1] = mo[0].cy; out_centroid[2] = mo[0].cz;
    for (int k = 0; k < 6; k++) out_inertia[k] = mo[0].inertia[k];
    return 0;
}
}

// K4's moments for TWO different fragments that share a warp (lanes 0-15 / 16-31 in lock step), as assemble_gather_kernel
// pairs neighbouring fragments.  n[2] vertices each (<= 64, ring degree <= 8; a fragment with n = 0 idles, `has` false);
// out: n_faces[2], volume[2], centroid[2][3], inertia[2][6].
extern "C" int k3emu_moments_two(const float* const* verts4, const uint32_t* const* ring_off, const uint16_t* const* ring, const int* n,
                                 int* out_faces, double* out_volume, float* out_centroid, float* out_inertia)
{
    auto mp = std::make_unique<MomPoly2[]>(2);
    for (int h = 0; h < 2; h++)
        for (int v = 0; v < n[h]; v++)
        {
            mp[h].x[v] = verts4[h][4 * v]; mp[h].y[v] = verts4[h][4 * v + 1]; mp[h].z[v] = verts4[h][4 * v + 2];
            u64 rw = ~0ull;
            const int r0 = (int)ring_off[h][v], r1 = (int)ring_off[h][v + 1];
            if (r1 - r0 > 8) return -1;
            for (int j = 0; j < r1 - r0; j++) rw = rset(rw, j, ring[h][r0 + j]);
            mp[h].ring[v] = rw;
            mp[h].estart[v] = (uint16_t)r0;
        }
    Moments mo[32];
    simt::run_warp([&](int lane) {
        const Sub<16> sub(lane);
        const int h = lane / 16;
        const bool do_mo = n[h] > 0;
        if (sub.any_warp(do_mo))     // assemble_gather_kernel: the warp enters when either fragment needs it
        {
            sub.sync();
            sub_fragment_moments2<16>(mp[h], do_mo ? n[h] : 0, sub, do_mo, mo[lane]);
        }
    });
    for (int h = 0; h < 2; h++)
    {
        if (n[h] <= 0) continue;
        for (int l = 1; l < 16; l++)
            if (std::memcmp(&mo[16 * h + l].volume, &mo[16 * h].volume, 8) || mo[16 * h + l].n_faces != mo[16 * h].n_faces) return -5;
        out_faces[h] = mo[16 * h].n_faces;
        out_volume[h] = mo[16 * h].volume;
        out_centroid[3 * h] = mo[16 * h].cx; out_centroid[3 * h + 1] = mo[16 * h].cy; out_centroid[3 * h + 2] = mo[16 * h].cz;
        for (int k = 0; k < 6; k++) out_inertia[6 * h + k] = mo[16 * h].inertia[k];
    }
    return 0;
}

static int g_large_gd = GD;   // ring slots per vertex of the emulated workspace (GD = the on-chip tier; the global tier's is a run-time value)
extern "C" void k3emu_set_gd(int gd) { g_large_gd = gd; }

namespace
{
template <int NW>
int run_large(const float* verts4, const uint32_t* ring_off, const uint16_t* ring, int nv_in, const float* planes4, int npl, int cap,
              float* out_verts4, uint32_t* out_ring_off, uint16_t* out_ring, int* out_info, double* out_volume, float* out_centroid,
              float* out_inertia)
{
    constexpr int N = NW * 32;
    for (int k = 0; k < 8; k++) out_info[k] = 0;
    const int gd = g_large_gd;
    std::vector<float4> ws((global_poly_bytes((size_t)cap, (size_t)gd) + 15) / 16 + 1);          // 16-byte aligned workspace
    GlobalPoly g0 = global_poly_carve(reinterpret_cast<unsigned char*>(ws.data()), cap, gd);
    bool bad = nv_in > cap;
    if (!bad)
        for (int v = 0; v < nv_in; v++)   // staging of clip_shared_kernel / clip_global_kernel (kernels.cuh)
        {
            g0.x[v] = verts4[4 * v]; g0.y[v] = verts4[4 * v + 1]; g0.z[v] = verts4[4 * v + 2];
            const int d = (int)(ring_off[v + 1] - ring_off[v]);
            if (d > gd || d == 0) { bad = true; continue; }
            g0.deg[v] = (uint16_t)d;
            for (int j = 0; j < d; j++)
            {
                const int idx = ring[ring_off[v] + j];
                if (idx >= nv_in) bad = true;
                g0.ring[(size_t)v * gd + j] = (uint16_t)idx;
            }
        }
    if (bad) { out_info[0] = nv_in > cap ? CLIP_NEED_SLOTS : CLIP_OVERFLOW; return 0; }
    std::vector<float4> planes(std::max(npl, 1));
    for (int p = 0; p < npl; p++) planes[p] = make_float4(planes4[4 * p], planes4[4 * p + 1], planes4[4 * p + 2], planes4[4 * p + 3]);

    int s_scan[NW + 1] = { 0 };
    float s_cov[NW * 10] = { 0.f };
    std::vector<int> nv(N), status(N);
    std::vector<unsigned> seq(N, 0u);
    std::vector<Moments> mo(N);
    GlobalPoly gfin = g0;
    const unsigned long n_coll = simt::run_block(N, [&](int tid) {
        GlobalPoly g = g0;                                  // every thread holds its own views, as in the kernels
        const Grp<NW> grp{ tid, tid & 31, s_scan };
        nv[tid] = nv_in;
        status[tid] = global_clip_by_planes<NW>(g, nv[tid], planes.data(), npl, grp, seq[tid]);
        if (status[tid] == CLIP_OK && nv[tid] > 0) global_fragment_moments<NW>(g, nv[tid], grp, mo[tid], s_cov);
        if (tid == 0) gfin = g;                             // (a compaction swaps the two ring views)
    });
    g0 = gfin;
    for (int t = 1; t < N; t++)
        if (nv[t] != nv[0] || status[t] != status[0]) return -1;
    out_info[0] = status[0];
    out_info[3] = (int)seq[0];
    out_info[5] = (int)n_coll;
    if (status[0] != CLIP_OK || nv[0] == 0) return 0;
    int ne = 0;
    for (int v = 0; v < nv[0]; v++)   // the tiers compact after every cut: the result is vertices 0 .. nv-1 of the workspace
    {
        out_verts4[4 * v] = g0.x[v]; out_verts4[4 * v + 1] = g0.y[v]; out_verts4[4 * v + 2] = g0.z[v]; out_verts4[4 * v + 3] = 0.f;
        out_ring_off[v] = (uint32_t)ne;
        for (int j = 0; j < g0.deg[v]; j++) out_ring[ne++] = g0.ring[(size_t)v * gd + j];
    }
    out_ring_off[nv[0]] = (uint32_t)ne;
    out_info[1] = nv[0];
    out_info[2] = ne;
    out_info[6] = mo[0].n_faces;       // thread 0 writes the record in the kernels
    *out_volume = mo[0].volume;
    out_centroid[0] = mo[0].cx; out_centroid[1] = mo[0].cy; out_centroid[2] = mo[0].cz;
    for (int k = 0; k < 6; k++) out_inertia[k] = mo[0].inertia[k];
    return 0;
}
} // namespace

extern "C" int k3emu_pair_large(int nw, int cap, const float* verts4, const uint32_t* ring_off, const uint16_t* ring, int nv_in,
                                const float* planes4, int npl, float* out_verts4, uint32_t* out_ring_off, uint16_t* out_ring,
                                int* out_info, double* out_volume, float* out_centroid, float* out_inertia)
{
    switch (nw)
    {
    case 1: return run_large<1>(verts4, ring_off, ring, nv_in, planes4, npl, cap, out_verts4, out_ring_off, out_ring, out_info, out_volume, out_centroid, out_inertia);
    case 4: return run_large<4>(verts4, ring_off, ring, nv_in, planes4, npl, cap, out_verts4, out_ring_off, out_ring, out_info, out_volume, out_centroid, out_inertia);
    case 8: return run_large<8>(verts4, ring_off, ring, nv_in, planes4, npl, cap, out_verts4, out_ring_off, out_ring, out_info, out_volume, out_centroid, out_inertia);
    default: return -9;
    }
}
