// surtr_engine.cu -- context, buffer management and the C ABI of libsurtr_b200.so (include/surtr_b200.h).
//
// Host side of the drop-in boundary: it owns device memory, orders K1 -> K2 -> K3 -> K4 on one stream with no
// host round trip inside an event, and re-runs an event once with larger buffers when a capacity estimate was
// too small.  There is no CPU implementation of any step behind this ABI.
#include "kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace surtr;

namespace
{

std::string g_create_error;

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may be scheduled while its
// predecessor in the stream is still draining; the kernels call griddepcontrol.wait before reading its output.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct DevBuf
{
    void* p = nullptr;
    size_t cap = 0;
    bool view = false;   // points into another buffer (a section of the input blob): never freed, never grown in place
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap && !view) return cudaSuccess;
        if (p && !view) cudaFree(p);
        p = nullptr;
        cap = 0;
        view = false;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p && !view) cudaFree(p);
        p = nullptr;
        cap = 0;
        view = false;
    }
    void set_view(void* ptr, size_t bytes)
    {
        release();
        p = ptr;
        cap = bytes;
        view = true;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
} // namespace

struct surtr_ctx
{
    int device = 0;
    int num_sm = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // asynchronous downloads run on their own stream so that the next event's uploads and K1-K3 need not queue behind
    // them; only K4 (which rewrites the fragment arrays) waits for the copies (copy_done)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done = nullptr;
    bool copy_pending = false;
    std::string err;
    int kdirs = 13;   // 26-DOP: 10 % faster events than the AABB on configs 3 and 4 (profiles/r1_kdop_sweep.json)

    // inputs
    uint32_t n_pieces = 0, n_cells = 0, n_events_p = 0, n_events_c = 0;
    uint64_t n_pverts = 0, n_pring = 0, n_planes = 0, n_cverts = 0;
    bool cells_bounded = false, have_pieces = false, have_cells = false, tables_dirty = true;
    DevBuf p_verts, p_vert_off, p_ring_off, p_ring;
    DevBuf c_planes, c_plane_off, c_verts, c_vert_off;
    std::vector<uint32_t> h_ev_piece_off, h_ev_cell_off;

    // resident pattern (surtr_upload_pattern / surtr_place_pattern)
    DevBuf pat_verts, pat_face_off, pat_xform, xf_mat, xf_idx;
    std::vector<uint32_t> h_pat_cell_face_off, h_pat_cell_vert_off, h_off_scratch;
    std::vector<float> h_xform;
    uint32_t pat_faces = 0, pat_cells = 0, pat_fverts = 0;

    // derived tables
    DevBuf d_tiles, d_ev_mask_base, d_ev_piece_off, d_ev_cell_off, d_ev_frag_off;
    uint32_t n_tiles = 0, n_masks = 0;
    uint64_t n_pairs = 0;

    // work buffers
    DevBuf ext_p, ext_c, masks, cand, cand_rec, scratch1, scratch2, ovf_list, ovf2_list, ovf3_list, fail_list, ws3, scratch3, ctl, dbg, out_off, frag_cand;
    bool debug = false;
    uint64_t cap_cand = 0, cap_tier2 = 0, cap_tier3 = 0;
    int cap3 = 0;                 // vertex slots of a tier-3 workspace (from the largest uploaded piece)
    int gd3 = 16;                 // ring slots per vertex of a tier-3 workspace (doubled whenever a ring outgrows them)
    uint32_t n_ws3 = 0;           // tier-3 workspaces = persistent blocks (fewer when the workspaces get large)
    std::vector<uint32_t> h_failed;   // candidates of the last event that could not be cut
    uint32_t max_piece_verts = 0;
    uint32_t n_tiles_a = 0, n_tiles_b = 0;

    // outputs
    DevBuf f_rec, f_verts, f_ring_off, f_ring;
    DevBuf wire_p3, wire_c3, wire_f3, wire_flen;   // float3 / u8 staging of the PCIe wire format
    DevBuf in_blob, out_blob;                      // one-copy transfers (surtr_upload_blob / surtr_download_blob_async)
    uint64_t cap_frag = 0, cap_fverts = 0, cap_fring = 0;

    Ctl* h_ctl = nullptr;       // pinned + mapped: the event's last kernel writes the counters straight into it
    Ctl* h_ctl_dev = nullptr;   // device-side alias of h_ctl
    // [0] start, [NPH] end; with profiling on also the phase boundaries in between (see surtr_last_event_phases)
    static constexpr int NPH = 8;
    cudaEvent_t ev[NPH + 1] = {};
    cudaEvent_t ev_done = nullptr;   // after the counters of the event have reached h_ctl
    int launches = 0;
    uint32_t ctl_layout_a = 0xffffffffu, ctl_layout_b = 0xffffffffu;
    void* ctl_ptr = nullptr;
    bool profile = false;         // per-kernel CUDA events inside an event (they serialise the PDL chain)
    bool profiled_last = false;
    bool tier1b_enabled = false;  // the 128-slot warp-per-pair tier is launched once an event needed it
    bool k3_round1 = false;       // SURTR_K3=sub: the round-1 small-tier kernel (A/B profiles only)
    uint32_t last_ring_bytes = 2; // ring entry width of the last event's output blob (surtr_download_blob_async)
    int k3_warps = 0;             // small tier's main launch: 0 = persistent warps + ticket (default); SURTR_K3_WARPS=2: one block of two pairs per two candidates (A/B)
    int k3_build = 0;             // SURTR_K3_BUILD=throughput / latency pins the small tier's build (A/B); default: by the last event's candidate count
    uint64_t last_n_cand = 0;     // candidates of the last resolved event
    bool k3_duo = false;          // SURTR_K3_DUO=1: two candidate pairs per warp (clip_duo.cuh)
    bool no_tier1b = false;       // SURTR_DEBUG_NO_TIER1B=1 (test hook): 64-slot overflows go straight to the large tier
    bool tier2_enabled = false;   // the large on-chip tier is launched once an event needed it
    bool tier3_enabled = false;   // likewise the global-memory tier
    bool event_launched = false, event_resolved = false;
    surtr_counts last{};
};

namespace
{
int fail(surtr_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CK(call)                                                                                       \
    do                                                                                                 \
    {                                                                                                  \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? SURTR_ERR_NOMEM : SURTR_ERR_CUDA,       \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                          \
    } while (0)

size_t ctl_bytes(const surtr_ctx* c)
{
    return sizeof(Ctl) + sizeof(unsigned int) * ((size_t)c->n_tiles_a + c->n_tiles_b + 2);
}

int rebuild_tables(surtr_ctx* ctx)
{
    if (!ctx->tables_dirty) return SURTR_OK;
    if (ctx->n_events_p != ctx->n_events_c)
        return fail(ctx, SURTR_ERR_INVALID, "pieces and cells were uploaded with different event counts");
    const uint32_t ne = ctx->n_events_p;
    std::vector<BpTile> tiles;
    std::vector<uint32_t> mask_base(ne + 1, 0);
    uint64_t masks = 0, pairs = 0;
    for (uint32_t e = 0; e < ne; e++)
    {
        mask_base[e] = (uint32_t)masks;
        const uint32_t p0 = ctx->h_ev_piece_off[e], p1 = ctx->h_ev_piece_off[e + 1];
        const uint32_t c0 = ctx->h_ev_cell_off[e], c1 = ctx->h_ev_cell_off[e + 1];
        const uint32_t np = p1 - p0, nc = c1 - c0;
        const uint32_t n_w = (np + 31) / 32;
        pairs += (uint64_t)np * nc;
        for (uint32_t cb = 0; cb < nc; cb += 32)
            for (uint32_t pb = 0; pb < np; pb += 256)
            {
                BpTile t;
                t.piece_begin = p0 + pb;
                t.n_piece = std::min(256u, np - pb);
                t.cell_begin = c0 + cb;
                t.n_cell = std::min(32u, nc - cb);
                t.mask_base = (uint32_t)(masks + (uint64_t)cb * n_w + pb / 32);
                t.n_w = n_w;
                tiles.push_back(t);
            }
        masks += (uint64_t)nc * n_w;
        if (masks > 0xfffffff0ull) return fail(ctx, SURTR_ERR_INVALID, "batch too large: split the events");
    }
    mask_base[ne] = (uint32_t)masks;
    ctx->n_tiles = (uint32_t)tiles.size();
    ctx->n_masks = (uint32_t)masks;
    ctx->n_pairs = pairs;
    CK(ctx->d_tiles.reserve(sizeof(BpTile) * std::max<size_t>(1, tiles.size())));
    CK(ctx->d_ev_mask_base.reserve(4 * (ne + 1)));
    CK(ctx->d_ev_piece_off.reserve(4 * (ne + 1)));
    CK(ctx->d_ev_cell_off.reserve(4 * (ne + 1)));
    // the tables are tiny; synchronous copies keep the std::vectors' lifetime trivial
    CK(cudaMemcpyAsync(ctx->d_tiles.p, tiles.data(), sizeof(BpTile) * tiles.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_ev_mask_base.p, mask_base.data(), 4 * (ne + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_ev_piece_off.p, ctx->h_ev_piece_off.data(), 4 * (ne + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_ev_cell_off.p, ctx->h_ev_cell_off.data(), 4 * (ne + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->tables_dirty = false;
    // first capacity estimates (grown on demand by surtr_event_counts)
    const uint64_t est = std::min<uint64_t>(pairs, std::max<uint64_t>(4096, 8ull * (ctx->n_pieces + ctx->n_cells)));
    ctx->cap_cand = std::max(ctx->cap_cand, est);
    return SURTR_OK;
}

int ensure_capacity(surtr_ctx* ctx)
{
    const int K = ctx->kdirs;
    ctx->cap_frag = std::max(ctx->cap_frag, ctx->cap_cand);
    ctx->cap_fverts = std::max(ctx->cap_fverts, ctx->cap_frag * 24);
    ctx->cap_fring = std::max(ctx->cap_fring, ctx->cap_fverts * 3 + 1024);
    ctx->cap_tier2 = std::max<uint64_t>(ctx->cap_tier2, 16);
    ctx->n_tiles_a = (ctx->n_masks + CP_THREADS * CP_ITEMS - 1) / (CP_THREADS * CP_ITEMS);
    ctx->n_tiles_b = (uint32_t)((ctx->cap_cand + AS_THREADS - 1) / AS_THREADS);
    CK(ctx->ext_p.reserve(sizeof(float) * 2 * K * std::max(1u, ctx->n_pieces)));
    CK(ctx->ext_c.reserve(sizeof(float) * 2 * K * std::max(1u, ctx->n_cells)));
    CK(ctx->masks.reserve(4 * std::max<size_t>(1, ctx->n_masks)));
    CK(ctx->cand.reserve(sizeof(uint2) * ctx->cap_cand));
    CK(ctx->cand_rec.reserve(sizeof(CandRec) * ctx->cap_cand));
    CK(ctx->ovf_list.reserve(4 * ctx->cap_cand));
    CK(ctx->ovf2_list.reserve(4 * ctx->cap_cand));
    CK(ctx->out_off.reserve(16 * ctx->cap_cand));
    if (ctx->debug) CK(ctx->dbg.reserve(32 * ctx->cap_cand));
    CK(ctx->scratch1.reserve(FAST_BLOB * ctx->cap_cand));
    CK(ctx->scratch2.reserve(blob2_bytes() * ctx->cap_tier2));
    CK(ctx->ovf3_list.reserve(4 * ctx->cap_cand));   // tiers 1 and 2 hand pairs on to the global tier through this list
    CK(ctx->fail_list.reserve(4 * ctx->cap_cand));
    if (ctx->tier3_enabled)
    {
        // workspace: twice the largest piece (a cut adds at most one vertex per straddling edge), at least 4096 slots
        // (a multiple of 16 keeps every array of the workspace and of the result blobs 16-byte aligned)
        uint64_t want = std::min<uint64_t>(65520, (std::max<uint64_t>(4096, 2ull * ctx->max_piece_verts + 1024) + 15) / 16 * 16);
        if (const char* e = std::getenv("SURTR_DEBUG_CAP3"))   // test hook: start with a small workspace to exercise its growth
            want = std::min<uint64_t>(65520, (std::max<uint64_t>(64, std::strtoull(e, nullptr, 10)) + 15) / 16 * 16);
        ctx->cap3 = std::max<int>(ctx->cap3, (int)want);
        ctx->cap_tier3 = std::max<uint64_t>(ctx->cap_tier3, 8);
        const size_t stride = (global_poly_bytes((size_t)ctx->cap3, (size_t)ctx->gd3) + 255) / 256 * 256;
        // one workspace per persistent block; at most 8 GB of them (a 65 520-slot workspace with 1024-wide rings is 270 MB)
        ctx->n_ws3 = (uint32_t)std::max<size_t>(1, std::min<size_t>((size_t)ctx->num_sm * T3_BLOCKS_PER_SM, ((size_t)8 << 30) / stride));
        CK(ctx->ws3.reserve(stride * ctx->n_ws3));
        CK(ctx->scratch3.reserve(blob3_bytes((size_t)ctx->cap3, (size_t)ctx->gd3) * ctx->cap_tier3));
    }
    // Ctl | flagsA | flagsB, then the (never zeroed) aggregate / inclusive arrays
    const size_t zero_bytes = (ctl_bytes(ctx) + 15) / 16 * 16;
    CK(ctx->ctl.reserve(zero_bytes + 8 * 2 * ((size_t)ctx->n_tiles_a + 1) + 8 * 6 * ((size_t)ctx->n_tiles_b + 1)));
    CK(ctx->f_rec.reserve(sizeof(surtr_fragment) * ctx->cap_frag));
    CK(ctx->frag_cand.reserve(4 * ctx->cap_frag));
    CK(ctx->f_verts.reserve(16 * ctx->cap_fverts));
    CK(ctx->f_ring_off.reserve(4 * (ctx->cap_fverts + 1)));
    CK(ctx->f_ring.reserve(2 * ctx->cap_fring));
    return SURTR_OK;
}

template <int K>
void launch_extents(surtr_ctx* ctx)
{
    const uint64_t n_obj = (uint64_t)ctx->n_pieces + ctx->n_cells;
    if (!n_obj) return;
    const int threads = 256;
    const bool wide = ctx->max_piece_verts > 128;   // a full warp per object only when a piece is large (kernels.cuh)
    const uint64_t lanes = wide ? 32 : 8;
    const int blocks = (int)std::min<uint64_t>((n_obj * lanes + threads - 1) / threads, (uint64_t)ctx->num_sm * 8);
    if (wide)
        launch_pdl(kdop_extents_kernel<K, 32>, dim3(blocks), dim3(threads), 0, ctx->stream, ctx->p_verts.as<float4>(),
                   ctx->p_vert_off.as<uint32_t>(), ctx->n_pieces, ctx->ext_p.as<float>(), ctx->c_verts.as<float4>(),
                   ctx->c_vert_off.as<uint32_t>(), ctx->n_cells, ctx->ext_c.as<float>(), ctx->cells_bounded ? 0 : 1);
    else
        launch_pdl(kdop_extents_kernel<K, 8>, dim3(blocks), dim3(threads), 0, ctx->stream, ctx->p_verts.as<float4>(),
                   ctx->p_vert_off.as<uint32_t>(), ctx->n_pieces, ctx->ext_p.as<float>(), ctx->c_verts.as<float4>(),
                   ctx->c_vert_off.as<uint32_t>(), ctx->n_cells, ctx->ext_c.as<float>(), ctx->cells_bounded ? 0 : 1);
    ctx->launches++;
}

template <int K>
void launch_masks(surtr_ctx* ctx)
{
    if (!ctx->n_tiles) return;
    launch_pdl(broadphase_mask_kernel<K>, dim3(ctx->n_tiles), dim3(256), 0, ctx->stream, ctx->d_tiles.as<BpTile>(),
               ctx->ext_p.as<float>(), ctx->ext_c.as<float>(), ctx->masks.as<unsigned int>());
    ctx->launches++;
}

int launch_event(surtr_ctx* ctx)
{
    int rc = rebuild_tables(ctx);
    if (rc) return rc;
    rc = ensure_capacity(ctx);
    if (rc) return rc;
    ctx->launches = 0;

    unsigned char* ctl_base = ctx->ctl.as<unsigned char>();
    const size_t zero_bytes = (ctl_bytes(ctx) + 15) / 16 * 16;
    Ctl* d_ctl = reinterpret_cast<Ctl*>(ctl_base);
    unsigned int* flags_a = reinterpret_cast<unsigned int*>(ctl_base + sizeof(Ctl));
    unsigned int* flags_b = flags_a + ctx->n_tiles_a + 1;
    unsigned long long* agg_a = reinterpret_cast<unsigned long long*>(ctl_base + zero_bytes);
    unsigned long long* inc_a = agg_a + ctx->n_tiles_a + 1;
    unsigned long long* agg_b = inc_a + ctx->n_tiles_a + 1;
    unsigned long long* inc_b = agg_b + 3 * ((size_t)ctx->n_tiles_b + 1);

    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    // The counters and scan flags are zeroed at the END of every event (off the next event's critical path);
    // only a fresh / re-laid-out control block is zeroed here.
    if (ctx->ctl_layout_a != ctx->n_tiles_a || ctx->ctl_layout_b != ctx->n_tiles_b || ctx->ctl_ptr != ctx->ctl.p)
    {
        CK(cudaMemsetAsync(ctl_base, 0, zero_bytes, ctx->stream));
        ctx->ctl_layout_a = ctx->n_tiles_a;
        ctx->ctl_layout_b = ctx->n_tiles_b;
        ctx->ctl_ptr = ctx->ctl.p;
    }

    // K1
    switch (ctx->kdirs)
    {
    case 3: launch_extents<3>(ctx); if (ctx->profile) CK(cudaEventRecord(ctx->ev[1], ctx->stream)); launch_masks<3>(ctx); break;
    case 7: launch_extents<7>(ctx); if (ctx->profile) CK(cudaEventRecord(ctx->ev[1], ctx->stream)); launch_masks<7>(ctx); break;
    default: launch_extents<13>(ctx); if (ctx->profile) CK(cudaEventRecord(ctx->ev[1], ctx->stream)); launch_masks<13>(ctx); break;
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    // K2 compaction
    if (ctx->n_tiles_a)
    {
        EventTables et{ ctx->d_ev_mask_base.as<uint32_t>(), ctx->d_ev_piece_off.as<uint32_t>(),
                        ctx->d_ev_cell_off.as<uint32_t>(), ctx->n_events_p };
        ScanState<1> st{ flags_a, agg_a, inc_a };
        launch_pdl(compact_pairs_kernel, dim3(ctx->n_tiles_a), dim3(CP_THREADS), 0, ctx->stream, ctx->masks.as<unsigned int>(),
                   ctx->n_masks, et, st, d_ctl, ctx->cand.as<uint2>(), ctx->cap_cand);
        ctx->launches++;
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->ev[3], ctx->stream));

    // K3
    ClipArgs ca;
    ca.p_verts = ctx->p_verts.as<float4>();
    ca.p_vert_off = ctx->p_vert_off.as<uint32_t>();
    ca.p_ring_off = ctx->p_ring_off.as<uint32_t>();
    ca.p_ring = ctx->p_ring.as<uint16_t>();
    ca.c_planes = ctx->c_planes.as<float4>();
    ca.c_plane_off = ctx->c_plane_off.as<uint32_t>();
    ca.ext_p = ctx->ext_p.as<float>();
    ca.kdirs = ctx->kdirs;
    ca.cand = ctx->cand.as<uint2>();
    ca.cap_cand = ctx->cap_cand;
    ca.rec = ctx->cand_rec.as<CandRec>();
    ca.ovf_list = ctx->ovf_list.as<uint32_t>();
    ca.ovf2_list = ctx->ovf2_list.as<uint32_t>();
    ca.skip_tier1b = ctx->no_tier1b || ctx->k3_round1 ? 1 : 0;
    ca.cap_tier2 = ctx->cap_tier2;
    ca.ovf3_list = ctx->ovf3_list.as<uint32_t>();
    ca.ws3 = ctx->ws3.as<unsigned char>();
    ca.ws3_stride = (global_poly_bytes((size_t)std::max(1, ctx->cap3), (size_t)ctx->gd3) + 255) / 256 * 256;
    ca.cap3 = ctx->cap3;
    ca.gd3 = ctx->gd3;
    ca.fail_list = ctx->fail_list.as<uint32_t>();
    ca.cap_tier3 = ctx->cap_tier3;
    ca.ctl = d_ctl;
    ca.dbg = ctx->debug ? ctx->dbg.as<uint32_t>() : nullptr;
    {
        ca.scratch = ctx->scratch1.as<unsigned char>();
        ca.scratch1 = ctx->scratch1.as<unsigned char>();
        ca.slot_bytes = FAST_BLOB;
        constexpr uint64_t pairs_per_block = FAST_WARPS * 32 / FAST_LANES;
        const uint64_t blocks = std::max<uint64_t>(1, (ctx->cap_cand + pairs_per_block - 1) / pairs_per_block);
        if (ctx->k3_round1) launch_pdl(clip_sub_kernel<FAST_LANES>, dim3((unsigned)blocks), dim3(FAST_WARPS * 32), 0, ctx->stream, ca);
        else if (ctx->k3_duo)
            launch_pdl(clip_duo_kernel,
                       dim3((unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (ctx->cap_cand + 2 * FAST_PERSIST_WARPS - 1) / (2 * FAST_PERSIST_WARPS)), (uint64_t)ctx->num_sm * (32 / FAST_PERSIST_WARPS))),
                       dim3(FAST_PERSIST_WARPS * 32), 0, ctx->stream, ca);
        else if (ctx->k3_warps == 0)
        {
            // an event that fits one wave of warps is bound by its slowest pair, not by issue slots: the latency build
            // (32 warps per SM, 64 registers); decided from the candidate count of this context's previous event
            const bool lat = !ctx->debug && ctx->k3_build != 1 &&
                             (ctx->k3_build == 2 || (ctx->last_n_cand > 0 && ctx->last_n_cand <= (uint64_t)ctx->num_sm * 32));
            const unsigned per_sm = (lat ? 32 : FAST_RESIDENT_WARPS) / FAST_PERSIST_WARPS;
            launch_pdl(ctx->debug ? clip_fast_kernel<2, false, FAST_PERSIST_WARPS, true, true>
                       : lat      ? clip_fast_kernel<2, false, FAST_PERSIST_WARPS, true, false, true>
                                  : clip_fast_kernel<2, false, FAST_PERSIST_WARPS, true>,
                       dim3((unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (ctx->cap_cand + FAST_PERSIST_WARPS - 1) / FAST_PERSIST_WARPS), (uint64_t)ctx->num_sm * per_sm)),
                       dim3(FAST_PERSIST_WARPS * 32), 0, ctx->stream, ca);
        }
        else launch_pdl(clip_fast_kernel<2, false, 2>, dim3((unsigned)blocks), dim3(FAST_WARPS * 32), 0, ctx->stream, ca);
        ctx->launches++;
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (ctx->tier1b_enabled)
    {
        // the pairs the 64-slot launch could not finish, in 128 slots (still one warp per pair)
        launch_pdl(clip_fast_kernel<4, true>, dim3(ctx->num_sm * 2), dim3(FAST_WARPS * 32), 0, ctx->stream, ca);
        ctx->launches++;
    }
    if (ctx->tier2_enabled)
    {
        ca.scratch = ctx->scratch2.as<unsigned char>();
        ca.slot_bytes = blob2_bytes();
        launch_pdl(clip_shared_kernel, dim3(ctx->num_sm * T2_BLOCKS_PER_SM), dim3(T2_WARPS * 32), t2_ws_bytes(), ctx->stream, ca);
        ctx->launches++;
    }
    if (ctx->tier3_enabled)
    {
        ca.scratch = ctx->scratch3.as<unsigned char>();
        ca.slot_bytes = blob3_bytes((size_t)ctx->cap3, (size_t)ctx->gd3);
        launch_pdl(clip_global_kernel, dim3(ctx->n_ws3), dim3(T3_WARPS * 32), 0, ctx->stream, ca);
        ctx->launches++;
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->ev[5], ctx->stream));

    // K4
    {
        AssembleArgs aa;
        aa.cand = ctx->cand.as<uint2>();
        aa.rec = ctx->cand_rec.as<CandRec>();
        aa.cap_cand = ctx->cap_cand;
        aa.scratch1 = ctx->scratch1.as<unsigned char>();
        aa.scratch2 = ctx->scratch2.as<unsigned char>();
        aa.scratch3 = ctx->scratch3.as<unsigned char>();
        aa.cap3 = ctx->cap3;
        aa.cap1 = 64;
        aa.cap2 = T2_CAP;
        aa.st = ScanState<3>{ flags_b, agg_b, inc_b };
        aa.ctl = d_ctl;
        aa.f_rec = ctx->f_rec.as<surtr_fragment>();
        aa.f_verts = ctx->f_verts.as<float4>();
        aa.f_ring_off = ctx->f_ring_off.as<uint32_t>();
        aa.f_ring = ctx->f_ring.as<uint16_t>();
        aa.cap_frag = ctx->cap_frag;
        aa.cap_fverts = ctx->cap_fverts;
        aa.cap_fring = ctx->cap_fring;
        aa.out_off = ctx->out_off.as<uint4>();
        aa.frag_cand = ctx->frag_cand.as<uint32_t>();
        const int blocks = (int)std::max<uint32_t>(1, std::min<uint32_t>(ctx->n_tiles_b, (uint32_t)ctx->num_sm * 4));
        if (ctx->copy_pending)
        {
            // K4 rewrites the fragment arrays: the asynchronous download of the previous event must have left them
            CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
            ctx->copy_pending = false;
        }
        launch_pdl(assemble_scan_kernel, dim3(blocks), dim3(AS_THREADS), 0, ctx->stream, aa);
        ctx->launches++;
        if (ctx->profile) CK(cudaEventRecord(ctx->ev[6], ctx->stream));
        constexpr uint64_t cand_per_block = GATHER_THREADS / GATHER_LANES;
        const uint64_t gblocks = std::min<uint64_t>(std::max<uint64_t>(1, (std::min(ctx->cap_cand, ctx->cap_frag) + cand_per_block - 1) / cand_per_block),
                                                    (uint64_t)ctx->num_sm * 32);  // at most four waves of blocks, striding over the fragments
        launch_pdl(assemble_gather_kernel<GATHER_LANES>, dim3((unsigned)gblocks), dim3(GATHER_THREADS), 0, ctx->stream, aa);
        ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev[7], ctx->stream));
    // counters -> mapped host memory and reset for the next event, by a kernel: a D2H copy of 64 bytes would queue on
    // the copy engine behind megabytes of other contexts' downloads, and the host waits on exactly this read-back
    finish_event_kernel<<<1, 256, 0, ctx->stream>>>(d_ctl, ctx->h_ctl_dev, reinterpret_cast<uint4*>(ctl_base), (unsigned)(zero_bytes / 16));
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev[surtr_ctx::NPH], ctx->stream));
    CK(cudaEventRecord(ctx->ev_done, ctx->stream));
    ctx->profiled_last = ctx->profile;
    ctx->event_launched = true;
    ctx->event_resolved = false;
    return SURTR_OK;
}

int resolve_event(surtr_ctx* ctx)
{
    if (!ctx->event_launched) return fail(ctx, SURTR_ERR_INVALID, "no fracture event has been launched");
    CK(cudaSetDevice(ctx->device));   // a re-run allocates and launches: every entry point that lands here may be on another device
    if (ctx->event_resolved)
    {
        if (ctx->copy_pending) { CK(cudaStreamSynchronize(ctx->copy_stream)); ctx->copy_pending = false; }
        return SURTR_OK;
    }
    // Growth reasons surface one after another (candidates, tier 2, tier 3, its workspace up to four doublings, the
    // fragment arrays): re-run until an event asks for nothing more.  Every re-run strictly enlarges something that is
    // bounded (by the pair count, 65520 slots, or device memory -> SURTR_ERR_NOMEM), so the bound below is never the
    // reason a legitimate event fails.
    for (int attempt = 0; attempt < 64; attempt++)
    {
        CK(cudaEventSynchronize(ctx->ev_done));
        const Ctl c = *ctx->h_ctl;
        bool grow = false;
        if (c.n_cand > ctx->cap_cand) { ctx->cap_cand = c.n_cand + c.n_cand / 8 + 64; grow = true; }
        if (c.n_ovf && !ctx->tier1b_enabled && !ctx->no_tier1b && !ctx->k3_round1) { ctx->tier1b_enabled = true; grow = true; }   // re-run with the 128-slot tier
        if (c.n_ovf2 > ctx->cap_tier2) { ctx->cap_tier2 = (uint64_t)c.n_ovf2 + c.n_ovf2 / 4 + 16; grow = true; }
        if (c.n_ovf2 && !ctx->tier2_enabled) { ctx->tier2_enabled = true; grow = true; }   // re-run with the large tier
        if (c.n_ovf3 > ctx->cap_tier3) { ctx->cap_tier3 = (uint64_t)c.n_ovf3 + c.n_ovf3 / 4 + 8; grow = true; }
        if (c.n_ovf3 && !ctx->tier3_enabled) { ctx->tier3_enabled = true; grow = true; }  // re-run with the global tier
        if (c.n_grow3 && ctx->cap3 < 65520)   // a global-tier workspace ran out of vertex slots: double it (up to the u16 index range)
        {
            ctx->cap3 = (int)std::min<uint64_t>(65520, 2ull * (uint64_t)ctx->cap3);
            grow = true;
        }
        if (c.n_growdeg && ctx->gd3 < 65520)  // a ring outgrew the workspace's ring slots: widen them (a vertex may have any number of neighbours)
        {
            ctx->gd3 = (int)std::min<uint64_t>(65520, std::max<uint64_t>(2ull * (uint64_t)ctx->gd3, ((uint64_t)c.max_deg + 8 + 7) / 8 * 8));
            grow = true;
        }
        if (!grow)
        {
            if (c.n_frag > ctx->cap_frag) { ctx->cap_frag = c.n_frag + c.n_frag / 8 + 64; grow = true; }
            if (c.n_fverts > ctx->cap_fverts) { ctx->cap_fverts = c.n_fverts + c.n_fverts / 8 + 64; grow = true; }
            if (c.n_fring > ctx->cap_fring) { ctx->cap_fring = c.n_fring + c.n_fring / 8 + 64; grow = true; }
        }
        if (!grow)
        {
            // what is left are pairs that no workspace can help: malformed rings, or beyond the 16-bit index range.  They
            // are reported PER PAIR (surtr_failed_pairs); every other fragment of the event is valid and can be read.
            const uint32_t n_bad = c.n_fail;
            ctx->h_failed.clear();
            if (c.n_fail)
            {
                std::vector<uint32_t> q(c.n_fail);
                CK(cudaMemcpy(q.data(), ctx->fail_list.p, 4 * (size_t)c.n_fail, cudaMemcpyDeviceToHost));
                for (uint32_t i = 0; i < c.n_fail; i++)
                {
                    uint2 one;
                    CK(cudaMemcpy(&one, ctx->cand.as<uint2>() + q[i], sizeof(uint2), cudaMemcpyDeviceToHost));
                    ctx->h_failed.push_back(one.x);
                    ctx->h_failed.push_back(one.y);
                }
            }
            ctx->last.n_failed = n_bad;
            if (n_bad)
                ctx->err = std::to_string(n_bad) + " pair(s) cannot be cut (malformed rings, or more than 65520 vertex / ring slots needed); "
                           "see surtr_failed_pairs -- all other fragments of the event are valid";
            ctx->last.n_pairs = ctx->n_pairs;
            ctx->last.n_candidates = c.n_cand;
            ctx->last_n_cand = c.n_cand;
            ctx->last.n_fragments = c.n_frag;
            ctx->last.n_verts = c.n_fverts;
            ctx->last.n_ring = c.n_fring;
            ctx->last.n_seq_cuts = c.n_seq_cuts;
            ctx->last.n_tier2 = c.n_ovf2;
            ctx->last.n_tier3 = c.n_ovf3;
            ctx->last.n_tier1b = c.n_ovf;
            ctx->last_ring_bytes = c.max_big_verts <= 256u ? 1u : 2u;   // (small-tier fragments have at most 64 vertices)
            ctx->event_resolved = true;
            return SURTR_OK;
        }
        const int rc = launch_event(ctx);
        if (rc) return rc;
    }
    return fail(ctx, SURTR_ERR_NOMEM, "buffers still too small after 64 rounds of growth");
}

// Event layout of an upload: n_events + 1 offsets, first 0, non-decreasing, last == n.  Checked BEFORE the context is
// touched, so a rejected call leaves it exactly as it was.
bool make_layout(const uint32_t* ev_off, uint32_t n_events, uint32_t n, std::vector<uint32_t>& layout)
{
    if (ev_off && n_events) layout.assign(ev_off, ev_off + n_events + 1);
    else layout = { 0u, n };
    if (layout.front() != 0u || layout.back() != n) return false;
    for (size_t i = 1; i < layout.size(); i++)
        if (layout[i] < layout[i - 1]) return false;
    return true;
}

int upload(surtr_ctx* ctx, DevBuf& b, const void* src, size_t bytes)
{
    CK(b.reserve(std::max<size_t>(bytes, 16)));
    if (bytes && src) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SURTR_OK;
}
// float3 stream from the host -> staging buffer -> widened to the resident float4 buffer on the context stream
int upload3(surtr_ctx* ctx, DevBuf& stage, DevBuf& dst4, const float* src3, uint64_t n)
{
    CK(dst4.reserve(std::max<size_t>(16 * n, 16)));
    if (!n || !src3) return SURTR_OK;
    const int rc = upload(ctx, stage, src3, 12 * n);
    if (rc) return rc;
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->num_sm * 8);
    widen3_kernel<<<blocks, 256, 0, ctx->stream>>>(stage.as<float>(), dst4.as<float4>(), n);
    CK(cudaGetLastError());
    return SURTR_OK;
}
} // namespace

static int upload_pieces_impl(surtr_ctx* ctx, const float* verts, bool packed3, const uint32_t* vert_off, const uint32_t* ring_off,
                              const uint16_t* ring, uint32_t n_pieces, const uint32_t* ev_piece_off, uint32_t n_events);
static int upload_cells_impl(surtr_ctx* ctx, const float* planes4, const uint32_t* plane_off, const float* cell_verts, bool packed3,
                             const uint32_t* cvert_off, uint32_t n_cells, const uint32_t* ev_cell_off, uint32_t n_events);
static int download_impl(surtr_ctx* ctx, surtr_fragment* fragments, void* verts, void* ring_off_or_len, uint16_t* ring, bool packed);

extern "C"
{
const char* surtr_version(void) { return "surtr_b200 0.1 (sm_100a)"; }

const char* surtr_last_error(const surtr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int surtr_ctx_create(int device, void* stream, surtr_ctx** out)
{
    surtr_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, SURTR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, SURTR_ERR_NO_DEVICE,
                    std::string("no CUDA device (") + cudaGetErrorString(e) + "): this library has no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, SURTR_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, SURTR_ERR_NO_DEVICE, "device is not compute capability 10.x: kernels are built for sm_100a only");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SURTR_ERR_CUDA, "cudaSetDevice failed");
    ctx = new surtr_ctx();
    ctx->device = device;
    ctx->num_sm = prop.multiProcessorCount;
    if (stream) ctx->stream = (cudaStream_t)stream;
    else
    {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
        {
            delete ctx;
            return fail(nullptr, SURTR_ERR_CUDA, "cudaStreamCreate failed");
        }
        ctx->own_stream = true;
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming);
    if (cudaHostAlloc(&ctx->h_ctl, sizeof(Ctl), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&ctx->h_ctl_dev, ctx->h_ctl, 0) != cudaSuccess)
    {
        delete ctx;
        return fail(nullptr, SURTR_ERR_NOMEM, "cudaHostAlloc (mapped) failed");
    }
    cudaFuncSetAttribute(clip_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t2_ws_bytes());
    if (const char* e = std::getenv("SURTR_K3")) ctx->k3_round1 = std::string(e) == "sub";
    if (const char* e = std::getenv("SURTR_DEBUG_NO_TIER1B")) ctx->no_tier1b = e[0] == '1';
    if (const char* e = std::getenv("SURTR_K3_WARPS")) ctx->k3_warps = e[0] == '2' ? 2 : 0;
    if (const char* e = std::getenv("SURTR_K3_DUO")) ctx->k3_duo = e[0] == '1';
    if (const char* e = std::getenv("SURTR_K3_BUILD")) ctx->k3_build = e[0] == 't' ? 1 : (e[0] == 'l' ? 2 : 0);
    *out = ctx;
    return SURTR_OK;
}

void surtr_ctx_destroy(surtr_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    DevBuf* all[] = { &ctx->p_verts, &ctx->p_vert_off, &ctx->p_ring_off, &ctx->p_ring, &ctx->c_planes, &ctx->c_plane_off,
                      &ctx->c_verts, &ctx->c_vert_off, &ctx->d_tiles, &ctx->d_ev_mask_base, &ctx->d_ev_piece_off,
                      &ctx->d_ev_cell_off, &ctx->ext_p, &ctx->ext_c, &ctx->masks, &ctx->cand, &ctx->cand_rec,
                      &ctx->scratch1, &ctx->scratch2, &ctx->scratch3, &ctx->ws3, &ctx->ovf_list, &ctx->ovf2_list, &ctx->ovf3_list, &ctx->fail_list, &ctx->ctl, &ctx->dbg, &ctx->out_off, &ctx->f_rec, &ctx->f_verts,
                      &ctx->f_ring_off, &ctx->f_ring, &ctx->wire_p3, &ctx->wire_c3, &ctx->wire_f3, &ctx->wire_flen, &ctx->in_blob, &ctx->out_blob, &ctx->frag_cand, &ctx->pat_verts, &ctx->pat_face_off, &ctx->pat_xform, &ctx->xf_mat,
                      &ctx->xf_idx, &ctx->d_ev_frag_off };
    for (DevBuf* b : all) b->release();
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int surtr_set_kdop_directions(surtr_ctx* ctx, int k)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (k != 3 && k != 7 && k != 13) return fail(ctx, SURTR_ERR_INVALID, "k must be 3, 7 or 13");
    ctx->kdirs = k;
    return SURTR_OK;
}

int surtr_set_clip_build(surtr_ctx* ctx, int mode)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (mode < 0 || mode > 2) return fail(ctx, SURTR_ERR_INVALID, "mode must be 0 (by event size), 1 (throughput) or 2 (latency)");
    ctx->k3_build = mode;
    return SURTR_OK;
}

int surtr_upload_pieces(surtr_ctx* ctx, const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off,
                        const uint16_t* ring, uint32_t n_pieces, const uint32_t* ev_piece_off, uint32_t n_events)
{
    return upload_pieces_impl(ctx, verts4, false, vert_off, ring_off, ring, n_pieces, ev_piece_off, n_events);
}

int surtr_upload_pieces3(surtr_ctx* ctx, const float* verts3, const uint32_t* vert_off, const uint32_t* ring_off,
                         const uint16_t* ring, uint32_t n_pieces, const uint32_t* ev_piece_off, uint32_t n_events)
{
    return upload_pieces_impl(ctx, verts3, true, vert_off, ring_off, ring, n_pieces, ev_piece_off, n_events);
}

int surtr_upload_cells(surtr_ctx* ctx, const float* planes4, const uint32_t* plane_off, const float* cell_verts4,
                       const uint32_t* cvert_off, uint32_t n_cells, const uint32_t* ev_cell_off, uint32_t n_events)
{
    return upload_cells_impl(ctx, planes4, plane_off, cell_verts4, false, cvert_off, n_cells, ev_cell_off, n_events);
}

int surtr_upload_cells3(surtr_ctx* ctx, const float* planes4, const uint32_t* plane_off, const float* cell_verts3,
                        const uint32_t* cvert_off, uint32_t n_cells, const uint32_t* ev_cell_off, uint32_t n_events)
{
    return upload_cells_impl(ctx, planes4, plane_off, cell_verts3, true, cvert_off, n_cells, ev_cell_off, n_events);
}
} // extern "C"

static int upload_pieces_impl(surtr_ctx* ctx, const float* verts4, bool packed3, const uint32_t* vert_off, const uint32_t* ring_off,
                              const uint16_t* ring, uint32_t n_pieces, const uint32_t* ev_piece_off, uint32_t n_events)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!vert_off || (vert_off[n_pieces] && (!verts4 || !ring_off || !ring))) return fail(ctx, SURTR_ERR_INVALID, "NULL piece array");
    std::vector<uint32_t> layout;
    if (!make_layout(ev_piece_off, n_events, n_pieces, layout))
        return fail(ctx, SURTR_ERR_INVALID, "ev_piece_off must start at 0, be non-decreasing and end at n_pieces");
    for (uint32_t i = 0; i < n_pieces; i++)
        if (vert_off[i + 1] < vert_off[i] || vert_off[i + 1] - vert_off[i] > 65520u)
            return fail(ctx, SURTR_ERR_INVALID, "vert_off must be non-decreasing and a piece may have at most 65520 vertices (ring entries are 16-bit local indices)");
    CK(cudaSetDevice(ctx->device));
    const uint64_t nv = vert_off[n_pieces];
    const uint64_t ne = nv ? ring_off[nv] : 0;
    int rc;
    if ((rc = packed3 ? upload3(ctx, ctx->wire_p3, ctx->p_verts, verts4, nv) : upload(ctx, ctx->p_verts, verts4, 16 * nv))) return rc;
    if ((rc = upload(ctx, ctx->p_vert_off, vert_off, 4 * ((size_t)n_pieces + 1)))) return rc;
    if ((rc = upload(ctx, ctx->p_ring_off, ring_off, 4 * (nv + 1)))) return rc;
    if ((rc = upload(ctx, ctx->p_ring, ring, 2 * ne))) return rc;
    ctx->n_pieces = n_pieces;
    ctx->n_pverts = nv;
    ctx->n_pring = ne;
    ctx->max_piece_verts = 0;
    for (uint32_t i = 0; i < n_pieces; i++) ctx->max_piece_verts = std::max(ctx->max_piece_verts, vert_off[i + 1] - vert_off[i]);
    // the tile tables depend only on the event layout: an upload of the same shape (the steady state of a caller
    // that streams events) keeps them, and with them the whole upload -> event sequence free of host synchronisation
    if (layout != ctx->h_ev_piece_off)
    {
        ctx->h_ev_piece_off.swap(layout);
        ctx->tables_dirty = true;
    }
    ctx->n_events_p = (uint32_t)ctx->h_ev_piece_off.size() - 1;
    ctx->have_pieces = true;
    ctx->event_launched = false;
    return SURTR_OK;
}

static int upload_cells_impl(surtr_ctx* ctx, const float* planes4, const uint32_t* plane_off, const float* cell_verts4, bool packed3,
                             const uint32_t* cvert_off, uint32_t n_cells, const uint32_t* ev_cell_off, uint32_t n_events)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!plane_off || (plane_off[n_cells] && !planes4)) return fail(ctx, SURTR_ERR_INVALID, "NULL cell array");
    std::vector<uint32_t> layout;
    if (!make_layout(ev_cell_off, n_events, n_cells, layout))
        return fail(ctx, SURTR_ERR_INVALID, "ev_cell_off must start at 0, be non-decreasing and end at n_cells");
    CK(cudaSetDevice(ctx->device));
    const uint64_t np = plane_off[n_cells];
    int rc;
    if ((rc = upload(ctx, ctx->c_planes, planes4, 16 * np))) return rc;
    if ((rc = upload(ctx, ctx->c_plane_off, plane_off, 4 * ((size_t)n_cells + 1)))) return rc;
    ctx->cells_bounded = cell_verts4 && cvert_off;
    if (ctx->cells_bounded)
    {
        const uint64_t ncv = cvert_off[n_cells];
        if ((rc = packed3 ? upload3(ctx, ctx->wire_c3, ctx->c_verts, cell_verts4, ncv) : upload(ctx, ctx->c_verts, cell_verts4, 16 * ncv))) return rc;
        if ((rc = upload(ctx, ctx->c_vert_off, cvert_off, 4 * ((size_t)n_cells + 1)))) return rc;
        ctx->n_cverts = ncv;
    }
    ctx->n_cells = n_cells;
    ctx->n_planes = np;
    if (layout != ctx->h_ev_cell_off)
    {
        ctx->h_ev_cell_off.swap(layout);
        ctx->tables_dirty = true;
    }
    ctx->n_events_c = (uint32_t)ctx->h_ev_cell_off.size() - 1;
    ctx->have_cells = true;
    ctx->event_launched = false;
    return SURTR_OK;
}

extern "C"
{
int surtr_fracture_event(surtr_ctx* ctx)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!ctx->have_pieces || !ctx->have_cells) return fail(ctx, SURTR_ERR_INVALID, "upload pieces and cells first");
    CK(cudaSetDevice(ctx->device));
    return launch_event(ctx);
}

int surtr_event_counts(surtr_ctx* ctx, surtr_counts* out)
{
    if (!ctx) return SURTR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    if (out) *out = ctx->last;
    return ctx->last.n_failed ? SURTR_ERR_OVERFLOW : SURTR_OK;   // (the event IS complete: counts are filled in, downloads work)
}

int surtr_failed_pairs(surtr_ctx* ctx, uint32_t* piece_cell, uint64_t capacity_pairs, uint64_t* n_pairs)
{
    if (!ctx || !n_pairs) return SURTR_ERR_INVALID;
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    *n_pairs = ctx->h_failed.size() / 2;
    if (piece_cell)
        for (size_t i = 0; i < ctx->h_failed.size() && i < 2 * capacity_pairs; i++) piece_cell[i] = ctx->h_failed[i];
    return SURTR_OK;
}

int surtr_download_fragments_async(surtr_ctx* ctx, surtr_fragment* fragments, float* verts4, uint32_t* ring_off, uint16_t* ring)
{
    return download_impl(ctx, fragments, verts4, ring_off, ring, false);
}

int surtr_download_fragments_packed_async(surtr_ctx* ctx, surtr_fragment* fragments, float* verts3, uint8_t* ring_len, uint16_t* ring)
{
    return download_impl(ctx, fragments, verts3, ring_len, ring, true);
}

int surtr_download_fragments_packed(surtr_ctx* ctx, surtr_fragment* fragments, float* verts3, uint8_t* ring_len, uint16_t* ring)
{
    const int rc = download_impl(ctx, fragments, verts3, ring_len, ring, true);
    return rc ? rc : surtr_sync(ctx);
}
} // extern "C"

static int download_impl(surtr_ctx* ctx, surtr_fragment* fragments, void* verts4, void* ring_off, uint16_t* ring, bool packed)
{
    if (!ctx) return SURTR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->event_resolved && ctx->copy_pending) ctx->copy_pending = false;   // (re-issued below; the stream keeps them ordered)
    const int rc = resolve_event(ctx);   // the event is complete on the main stream from here on
    if (rc) return rc;
    const surtr_counts& c = ctx->last;
    cudaStream_t cs = ctx->copy_stream ? ctx->copy_stream : ctx->stream;
    if (fragments && c.n_fragments)
        CK(cudaMemcpyAsync(fragments, ctx->f_rec.p, sizeof(surtr_fragment) * c.n_fragments, cudaMemcpyDeviceToHost, cs));
    if (packed && c.n_verts && (verts4 || ring_off))
    {
        // wire format: float3 + one byte of ring length per vertex, packed on the copy stream (the event is complete)
        CK(ctx->wire_f3.reserve(12 * c.n_verts));
        CK(ctx->wire_flen.reserve(c.n_verts));
        const unsigned blocks = (unsigned)std::min<uint64_t>((c.n_verts + 255) / 256, (uint64_t)ctx->num_sm * 8);
        pack_fragments_kernel<<<blocks, 256, 0, cs>>>(ctx->f_verts.as<float4>(), ctx->f_ring_off.as<uint32_t>(), ctx->wire_f3.as<float>(),
                                                      ctx->wire_flen.as<uint8_t>(), c.n_verts);
        CK(cudaGetLastError());
        if (verts4) CK(cudaMemcpyAsync(verts4, ctx->wire_f3.p, 12 * c.n_verts, cudaMemcpyDeviceToHost, cs));
        if (ring_off) CK(cudaMemcpyAsync(ring_off, ctx->wire_flen.p, c.n_verts, cudaMemcpyDeviceToHost, cs));
    }
    if (!packed && verts4 && c.n_verts)
        CK(cudaMemcpyAsync(verts4, ctx->f_verts.p, 16 * c.n_verts, cudaMemcpyDeviceToHost, cs));
    if (!packed && ring_off)
        CK(cudaMemcpyAsync(ring_off, ctx->f_ring_off.p, 4 * (c.n_verts + 1), cudaMemcpyDeviceToHost, cs));
    if (ring && c.n_ring)
        CK(cudaMemcpyAsync(ring, ctx->f_ring.p, 2 * c.n_ring, cudaMemcpyDeviceToHost, cs));
    if (ctx->copy_stream)
    {
        CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
        ctx->copy_pending = true;
    }
    return SURTR_OK;
}

extern "C"
{
static inline uint64_t blob_align(uint64_t x) { return (x + 255ull) & ~255ull; }

int surtr_input_blob_layout(uint32_t n_pieces, uint64_t n_piece_verts, uint64_t n_piece_ring, uint32_t n_cells, uint64_t n_planes,
                            uint64_t n_cell_verts, uint32_t n_events, uint32_t ring_entry_bytes, surtr_in_layout* out)
{
    if (!out || (ring_entry_bytes != 1 && ring_entry_bytes != 2)) return SURTR_ERR_INVALID;
    uint64_t at = 0;
    out->verts3 = at;       at = blob_align(at + 12 * n_piece_verts);
    out->vert_off = at;     at = blob_align(at + 4 * ((uint64_t)n_pieces + 1));
    out->ring_base = at;    at = blob_align(at + 4 * ((uint64_t)n_pieces + 1));
    out->ring_len = at;     at = blob_align(at + n_piece_verts);
    out->ring = at;         at = blob_align(at + ring_entry_bytes * n_piece_ring);
    out->planes4 = at;      at = blob_align(at + 16 * n_planes);
    out->plane_off = at;    at = blob_align(at + 4 * ((uint64_t)n_cells + 1));
    out->cell_verts3 = at;  at = blob_align(at + 12 * n_cell_verts);
    out->cvert_off = at;    at = blob_align(at + 4 * ((uint64_t)n_cells + 1));
    out->ev_piece_off = at; at = blob_align(at + 4 * ((uint64_t)n_events + 1));
    out->ev_cell_off = at;  at = blob_align(at + 4 * ((uint64_t)n_events + 1));
    out->total = at;
    return SURTR_OK;
}

int surtr_upload_blob(surtr_ctx* ctx, const void* blob, uint32_t n_pieces, uint64_t n_piece_verts, uint64_t n_piece_ring,
                      uint32_t n_cells, uint64_t n_planes, uint64_t n_cell_verts, uint32_t n_events, uint32_t ring_entry_bytes)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!blob) return fail(ctx, SURTR_ERR_INVALID, "NULL blob");
    surtr_in_layout L;
    if (surtr_input_blob_layout(n_pieces, n_piece_verts, n_piece_ring, n_cells, n_planes, n_cell_verts, n_events, ring_entry_bytes, &L))
        return fail(ctx, SURTR_ERR_INVALID, "ring_entry_bytes must be 1 or 2");
    const unsigned char* h = static_cast<const unsigned char*>(blob);
    const uint32_t* vert_off = reinterpret_cast<const uint32_t*>(h + L.vert_off);
    const uint32_t* ring_base = reinterpret_cast<const uint32_t*>(h + L.ring_base);
    const uint32_t* plane_off = reinterpret_cast<const uint32_t*>(h + L.plane_off);
    const uint32_t* cvert_off = reinterpret_cast<const uint32_t*>(h + L.cvert_off);
    // everything is validated against the host copy BEFORE the context is touched (per piece, not per vertex: a ring
    // length that contradicts its piece's ring range is caught by the clipper, which reports the pair as failed)
    std::vector<uint32_t> lp, lc;
    if (!make_layout(n_events ? reinterpret_cast<const uint32_t*>(h + L.ev_piece_off) : nullptr, n_events, n_pieces, lp) ||
        !make_layout(n_events ? reinterpret_cast<const uint32_t*>(h + L.ev_cell_off) : nullptr, n_events, n_cells, lc))
        return fail(ctx, SURTR_ERR_INVALID, "event offsets must start at 0, be non-decreasing and end at the piece / cell count");
    if (vert_off[n_pieces] != n_piece_verts || ring_base[0] != 0u || ring_base[n_pieces] != n_piece_ring ||
        plane_off[n_cells] != n_planes || (n_cell_verts && cvert_off[n_cells] != n_cell_verts))
        return fail(ctx, SURTR_ERR_INVALID, "blob offsets do not match the stated sizes");
    uint32_t max_verts = 0;
    for (uint32_t i = 0; i < n_pieces; i++)
    {
        if (vert_off[i + 1] < vert_off[i] || vert_off[i + 1] - vert_off[i] > 65520u)
            return fail(ctx, SURTR_ERR_INVALID, "vert_off must be non-decreasing and a piece may have at most 65520 vertices");
        if (ring_base[i + 1] < ring_base[i]) return fail(ctx, SURTR_ERR_INVALID, "ring_base must be non-decreasing");
        max_verts = std::max(max_verts, vert_off[i + 1] - vert_off[i]);
    }
    if (ring_entry_bytes == 1 && max_verts > 256u) return fail(ctx, SURTR_ERR_INVALID, "one-byte ring entries need pieces of at most 256 vertices");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->in_blob.reserve(std::max<uint64_t>(L.total, 256)));
    CK(ctx->p_verts.reserve(std::max<size_t>(16 * n_piece_verts, 16)));
    CK(ctx->p_ring_off.reserve(4 * (n_piece_verts + 1)));
    if (ring_entry_bytes == 1) CK(ctx->p_ring.reserve(std::max<size_t>(2 * n_piece_ring, 16)));
    if (n_cell_verts) CK(ctx->c_verts.reserve(16 * n_cell_verts));
    // ONE host -> device copy, then one kernel that widens / rebuilds the compact sections into the resident arrays; the
    // remaining index arrays are used in place
    CK(cudaMemcpyAsync(ctx->in_blob.p, blob, L.total, cudaMemcpyHostToDevice, ctx->stream));
    unsigned char* d = ctx->in_blob.as<unsigned char>();
    if (ring_entry_bytes == 2) ctx->p_ring.set_view(d + L.ring, 2 * n_piece_ring);
    const uint64_t work = std::max<uint64_t>(n_piece_verts + n_cell_verts, std::max<uint64_t>((uint64_t)n_pieces * 8, n_piece_ring / 16));
    const unsigned blocks = (unsigned)std::min<uint64_t>((work + 255) / 256 + 1, (uint64_t)ctx->num_sm * 8);
    auto kern = ring_entry_bytes == 1 ? expand_blob_kernel<1> : expand_blob_kernel<2>;
    kern<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const float*>(d + L.verts3), ctx->p_verts.as<float4>(), n_piece_verts,
                                          reinterpret_cast<const float*>(d + L.cell_verts3), ctx->c_verts.as<float4>(), n_cell_verts,
                                          reinterpret_cast<const uint32_t*>(d + L.vert_off), reinterpret_cast<const uint32_t*>(d + L.ring_base),
                                          d + L.ring_len, n_pieces, ctx->p_ring_off.as<uint32_t>(), d + L.ring, ctx->p_ring.as<uint16_t>(), n_piece_ring);
    CK(cudaGetLastError());
    ctx->p_vert_off.set_view(d + L.vert_off, 4 * ((size_t)n_pieces + 1));
    ctx->c_planes.set_view(d + L.planes4, 16 * n_planes);
    ctx->c_plane_off.set_view(d + L.plane_off, 4 * ((size_t)n_cells + 1));
    ctx->c_vert_off.set_view(d + L.cvert_off, 4 * ((size_t)n_cells + 1));
    ctx->n_pieces = n_pieces;
    ctx->n_pverts = n_piece_verts;
    ctx->n_pring = n_piece_ring;
    ctx->max_piece_verts = max_verts;
    ctx->n_cells = n_cells;
    ctx->n_planes = n_planes;
    ctx->n_cverts = n_cell_verts;
    ctx->cells_bounded = n_cell_verts != 0;
    if (lp != ctx->h_ev_piece_off) { ctx->h_ev_piece_off.swap(lp); ctx->tables_dirty = true; }
    if (lc != ctx->h_ev_cell_off) { ctx->h_ev_cell_off.swap(lc); ctx->tables_dirty = true; }
    ctx->n_events_p = (uint32_t)ctx->h_ev_piece_off.size() - 1;
    ctx->n_events_c = (uint32_t)ctx->h_ev_cell_off.size() - 1;
    ctx->have_pieces = ctx->have_cells = true;
    ctx->event_launched = false;
    return SURTR_OK;
}

int surtr_download_blob_async(surtr_ctx* ctx, void* host_blob, uint64_t capacity, surtr_out_layout* out)
{
    if (!ctx || !out) return SURTR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->event_resolved && ctx->copy_pending) ctx->copy_pending = false;
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    const surtr_counts& c = ctx->last;
    uint64_t at = 0;
    out->n_fragments = c.n_fragments; out->n_verts = c.n_verts; out->n_ring = c.n_ring;
    out->fragments = at; at = blob_align(at + sizeof(surtr_fragment) * c.n_fragments);
    out->verts3 = at;    at = blob_align(at + 12 * c.n_verts);
    out->ring_len = at;  at = blob_align(at + c.n_verts);
    out->ring_entry_bytes = ctx->last_ring_bytes;   // 1 when no fragment of the event has more than 256 vertices
    out->ring = at;      at = blob_align(at + out->ring_entry_bytes * c.n_ring);
    out->total = at;
    if (!host_blob || capacity < at) return fail(ctx, SURTR_ERR_INVALID, "host blob too small: " + std::to_string(at) + " bytes needed");
    if (!at) return SURTR_OK;
    cudaStream_t cs = ctx->copy_stream ? ctx->copy_stream : ctx->stream;
    CK(ctx->out_blob.reserve(at));
    unsigned char* d = ctx->out_blob.as<unsigned char>();
    const uint64_t work = std::max<uint64_t>(c.n_verts, std::max<uint64_t>(c.n_fragments * 4, c.n_ring / 8));
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((work + 255) / 256, (uint64_t)ctx->num_sm * 8));
    auto kern = out->ring_entry_bytes == 1 ? pack_blob_kernel<1> : pack_blob_kernel<2>;
    kern<<<blocks, 256, 0, cs>>>(ctx->f_rec.as<uint4>(), c.n_fragments * (sizeof(surtr_fragment) / 16), ctx->f_verts.as<float4>(),
                                 ctx->f_ring_off.as<uint32_t>(), c.n_verts, ctx->f_ring.as<uint16_t>(), c.n_ring,
                                 reinterpret_cast<uint4*>(d + out->fragments), reinterpret_cast<float*>(d + out->verts3),
                                 d + out->ring_len, d + out->ring);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host_blob, d, at, cudaMemcpyDefault, cs));   // ONE copy (the destination may also be device memory: the multi-GPU gather packs into a device tensor)
    if (ctx->copy_stream)
    {
        CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
        ctx->copy_pending = true;
    }
    return SURTR_OK;
}

int surtr_sync(surtr_ctx* ctx)
{
    if (!ctx) return SURTR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->copy_pending = false;
    return SURTR_OK;
}

int surtr_download_fragments(surtr_ctx* ctx, surtr_fragment* fragments, float* verts4, uint32_t* ring_off, uint16_t* ring)
{
    const int rc = surtr_download_fragments_async(ctx, fragments, verts4, ring_off, ring);
    return rc ? rc : surtr_sync(ctx);
}

int surtr_device_fragments(surtr_ctx* ctx, surtr_device_view* out)
{
    if (!ctx || !out) return SURTR_ERR_INVALID;
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    out->fragments = ctx->f_rec.as<surtr_fragment>();
    out->verts4 = ctx->f_verts.as<float>();
    out->ring_off = ctx->f_ring_off.as<uint32_t>();
    out->ring = ctx->f_ring.as<uint16_t>();
    return SURTR_OK;
}

int surtr_upload_pattern(surtr_ctx* ctx, const float* face_verts4, const uint32_t* face_vert_off, uint32_t n_faces,
                         const uint32_t* cell_face_off, uint32_t n_cells)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!face_verts4 || !face_vert_off || !cell_face_off || !n_faces || !n_cells) return fail(ctx, SURTR_ERR_INVALID, "NULL or empty pattern");
    if (cell_face_off[n_cells] != n_faces) return fail(ctx, SURTR_ERR_INVALID, "cell_face_off does not end at n_faces");
    for (uint32_t f = 0; f < n_faces; f++)
        if (face_vert_off[f + 1] < face_vert_off[f] + 3) return fail(ctx, SURTR_ERR_INVALID, "pattern face with fewer than 3 vertices");
    CK(cudaSetDevice(ctx->device));
    const uint32_t nfv = face_vert_off[n_faces];
    int rc;
    if ((rc = upload(ctx, ctx->pat_verts, face_verts4, 16 * (size_t)nfv))) return rc;
    if ((rc = upload(ctx, ctx->pat_face_off, face_vert_off, 4 * ((size_t)n_faces + 1)))) return rc;
    ctx->h_pat_cell_face_off.assign(cell_face_off, cell_face_off + n_cells + 1);
    ctx->h_pat_cell_vert_off.resize(n_cells + 1);
    for (uint32_t c = 0; c <= n_cells; c++) ctx->h_pat_cell_vert_off[c] = face_vert_off[cell_face_off[c]];
    ctx->pat_faces = n_faces;
    ctx->pat_cells = n_cells;
    ctx->pat_fverts = nfv;
    return SURTR_OK;
}

int surtr_place_pattern(surtr_ctx* ctx, const float* scale3, const float* translate3, uint32_t n_place)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!ctx->pat_cells) return fail(ctx, SURTR_ERR_INVALID, "no pattern uploaded");
    if (!scale3 || !translate3 || !n_place) return fail(ctx, SURTR_ERR_INVALID, "NULL or empty placement list");
    const uint64_t nf = (uint64_t)ctx->pat_faces * n_place, nv = (uint64_t)ctx->pat_fverts * n_place, nc = (uint64_t)ctx->pat_cells * n_place;
    if (nf > 0xfffffff0ull || nv > 0xfffffff0ull) return fail(ctx, SURTR_ERR_INVALID, "too many placements for one batch");
    CK(cudaSetDevice(ctx->device));
    // a stream-ordered copy out of these host vectors must have left before they are rewritten (pageable -> staged
    // copies return only after the staging, so this is cheap; it also orders the placement after the previous event)
    ctx->h_xform.resize(6 * (size_t)n_place);
    for (uint32_t p = 0; p < n_place; p++)
    {
        for (int k = 0; k < 3; k++) { ctx->h_xform[6 * p + k] = scale3[3 * p + k]; ctx->h_xform[6 * p + 3 + k] = translate3[3 * p + k]; }
    }
    int rc;
    if ((rc = upload(ctx, ctx->pat_xform, ctx->h_xform.data(), 4 * ctx->h_xform.size()))) return rc;
    CK(ctx->c_planes.reserve(16 * nf));
    CK(ctx->c_verts.reserve(16 * nv));
    // offsets: placement p repeats the pattern's offsets shifted by p * (faces | face vertices)
    std::vector<uint32_t>& off = ctx->h_off_scratch;
    off.resize(2 * (nc + 1));
    for (uint32_t p = 0; p < n_place; p++)
        for (uint32_t c = 0; c < ctx->pat_cells; c++)
        {
            off[(size_t)p * ctx->pat_cells + c] = p * ctx->pat_faces + ctx->h_pat_cell_face_off[c];
            off[nc + 1 + (size_t)p * ctx->pat_cells + c] = p * ctx->pat_fverts + ctx->h_pat_cell_vert_off[c];
        }
    off[nc] = (uint32_t)nf;
    off[2 * nc + 1] = (uint32_t)nv;
    if ((rc = upload(ctx, ctx->c_plane_off, off.data(), 4 * (nc + 1)))) return rc;
    if ((rc = upload(ctx, ctx->c_vert_off, off.data() + nc + 1, 4 * (nc + 1)))) return rc;
    const unsigned blocks = (unsigned)((nf + 255) / 256);
    place_pattern_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->pat_verts.as<float4>(), ctx->pat_face_off.as<uint32_t>(), ctx->pat_faces,
                                                         ctx->pat_fverts, ctx->pat_xform.as<float>(), n_place, ctx->c_planes.as<float4>(),
                                                         ctx->c_verts.as<float4>());
    CK(cudaGetLastError());
    ctx->cells_bounded = true;
    ctx->n_cverts = nv;
    ctx->n_cells = (uint32_t)nc;
    ctx->n_planes = nf;
    std::vector<uint32_t> layout(n_place + 1);
    for (uint32_t p = 0; p <= n_place; p++) layout[p] = p * ctx->pat_cells;
    if (layout != ctx->h_ev_cell_off)
    {
        ctx->h_ev_cell_off.swap(layout);
        ctx->tables_dirty = true;
    }
    ctx->n_events_c = n_place;
    ctx->have_cells = true;
    ctx->event_launched = false;
    return SURTR_OK;
}

int surtr_transform_pieces(surtr_ctx* ctx, const float* matrices16, const uint32_t* piece_matrix, uint32_t n_matrices)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!ctx->have_pieces) return fail(ctx, SURTR_ERR_INVALID, "no pieces resident");
    if (!matrices16 || !n_matrices) return fail(ctx, SURTR_ERR_INVALID, "NULL or empty matrix list");
    if (piece_matrix)
        for (uint32_t i = 0; i < ctx->n_pieces; i++)
            if (piece_matrix[i] >= n_matrices) return fail(ctx, SURTR_ERR_INVALID, "piece_matrix entry out of range");
    if (!ctx->n_pieces) return SURTR_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload(ctx, ctx->xf_mat, matrices16, 64 * (size_t)n_matrices))) return rc;
    if (piece_matrix && (rc = upload(ctx, ctx->xf_idx, piece_matrix, 4 * (size_t)ctx->n_pieces))) return rc;
    const unsigned blocks = (ctx->n_pieces + 7) / 8;
    transform_pieces_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->p_verts.as<float4>(), ctx->p_vert_off.as<uint32_t>(), ctx->n_pieces,
                                                            ctx->xf_mat.as<float>(), piece_matrix ? ctx->xf_idx.as<uint32_t>() : nullptr);
    CK(cudaGetLastError());
    ctx->event_launched = false;
    return SURTR_OK;
}

int surtr_download_pieces(surtr_ctx* ctx, float* verts4)
{
    if (!ctx || !verts4) return SURTR_ERR_INVALID;
    if (!ctx->have_pieces) return fail(ctx, SURTR_ERR_INVALID, "no pieces resident");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_pverts) CK(cudaMemcpyAsync(verts4, ctx->p_verts.p, 16 * ctx->n_pverts, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SURTR_OK;
}

namespace
{
// the last event's fragment arrays become the piece arrays (buffers swapped, vert_off table rebuilt from the records)
int fragments_to_pieces_impl(surtr_ctx* ctx, uint32_t n, std::vector<uint32_t>& layout)
{
    const surtr_counts c = ctx->last;
    CK(ctx->p_vert_off.reserve(4 * ((size_t)n + 1)));
    fragments_vert_off_kernel<<<(n + 1 + 255) / 256, 256, 0, ctx->stream>>>(ctx->f_rec.as<surtr_fragment>(), n, (uint32_t)c.n_verts,
                                                                             ctx->p_vert_off.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->p_ring_off.view) ctx->p_ring_off.release();   // sections of the input blob are never recycled as output arrays
    if (ctx->p_ring.view) ctx->p_ring.release();
    std::swap(ctx->p_verts, ctx->f_verts);
    std::swap(ctx->p_ring_off, ctx->f_ring_off);
    std::swap(ctx->p_ring, ctx->f_ring);
    ctx->cap_fverts = ctx->f_verts.cap / 16;
    ctx->cap_fring = ctx->f_ring.cap / 2;
    if (ctx->f_ring_off.cap < 4 * (ctx->cap_fverts + 1)) ctx->cap_fverts = ctx->f_ring_off.cap >= 8 ? ctx->f_ring_off.cap / 4 - 1 : 0;
    ctx->max_piece_verts = std::min<uint32_t>(32000u, 2u * std::max(ctx->max_piece_verts, 64u));   // fragments can outgrow their piece
    ctx->n_pieces = n;
    ctx->n_pverts = c.n_verts;
    ctx->n_pring = c.n_ring;
    ctx->h_ev_piece_off.swap(layout);
    ctx->n_events_p = (uint32_t)ctx->h_ev_piece_off.size() - 1;
    ctx->have_pieces = true;
    ctx->tables_dirty = true;
    ctx->event_launched = false;
    return SURTR_OK;
}
} // namespace

int surtr_fragments_to_pieces(surtr_ctx* ctx, const uint32_t* ev_piece_off, uint32_t n_events)
{
    if (!ctx) return SURTR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    if (ctx->last.n_fragments > 0xfffffff0ull) return fail(ctx, SURTR_ERR_INVALID, "too many fragments");
    const uint32_t n = (uint32_t)ctx->last.n_fragments;
    std::vector<uint32_t> layout;
    if (!make_layout(ev_piece_off, n_events, n, layout))
        return fail(ctx, SURTR_ERR_INVALID, "ev_piece_off must start at 0, be non-decreasing and end at the fragment count");
    return fragments_to_pieces_impl(ctx, n, layout);
}

int surtr_fragments_to_pieces_per_event(surtr_ctx* ctx)
{
    if (!ctx) return SURTR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    if (ctx->last.n_fragments > 0xfffffff0ull) return fail(ctx, SURTR_ERR_INVALID, "too many fragments");
    const uint32_t n = (uint32_t)ctx->last.n_fragments;
    if (ctx->h_ev_cell_off.size() < 2) return fail(ctx, SURTR_ERR_INVALID, "no cell set: there is no event layout to regroup by");
    const uint32_t ne = (uint32_t)ctx->h_ev_cell_off.size() - 1;   // the event layout the fragments were cut under
    // event boundaries in the fragment list, found on the device (the records never travel): n_events + 1 words come back
    CK(ctx->d_ev_frag_off.reserve(8 * ((size_t)ne + 1)));
    uint32_t* d_cell_off = ctx->d_ev_frag_off.as<uint32_t>();
    uint32_t* d_frag_off = d_cell_off + ne + 1;
    CK(cudaMemcpyAsync(d_cell_off, ctx->h_ev_cell_off.data(), 4 * ((size_t)ne + 1), cudaMemcpyHostToDevice, ctx->stream));
    event_fragment_off_kernel<<<(ne + 1 + 127) / 128, 128, 0, ctx->stream>>>(ctx->f_rec.as<surtr_fragment>(), n, d_cell_off, ne, d_frag_off);
    CK(cudaGetLastError());
    std::vector<uint32_t> layout((size_t)ne + 1);
    CK(cudaMemcpyAsync(layout.data(), d_frag_off, 4 * ((size_t)ne + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> checked;
    if (!make_layout(layout.data(), ne, n, checked))
        return fail(ctx, SURTR_ERR_INVALID, "fragment records are not grouped by event");
    return fragments_to_pieces_impl(ctx, n, checked);
}

int surtr_kdop_calc(surtr_ctx* ctx, const float* verts4, uint32_t n_verts, const float* normals3, uint32_t k, float* dist,
                    int32_t* arg, float* planes8)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!verts4 || !normals3 || !dist || !arg || !k) return fail(ctx, SURTR_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    DevBuf dv, dn, dd, da, dp;
    int rc = SURTR_OK;
    auto done = [&](int code) { dv.release(); dn.release(); dd.release(); da.release(); dp.release(); return code; };
    if (dv.reserve(16 * std::max(1u, n_verts)) || dn.reserve(12 * k) || dd.reserve(8 * k) || da.reserve(8 * k) || dp.reserve(32 * k))
        return done(fail(ctx, SURTR_ERR_NOMEM, "cudaMalloc failed"));
    cudaMemcpyAsync(dv.p, verts4, 16 * (size_t)n_verts, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dn.p, normals3, 12 * (size_t)k, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(dp.p, 0, 32 * (size_t)k, ctx->stream);
    kdop_arg_kernel<<<k, 256, 0, ctx->stream>>>(dv.as<float4>(), n_verts, dn.as<float>(), dd.as<float>(), da.as<int32_t>(), dp.as<float4>());
    cudaMemcpyAsync(dist, dd.p, 8 * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(arg, da.p, 8 * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream);
    if (planes8) cudaMemcpyAsync(planes8, dp.p, 32 * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, SURTR_ERR_CUDA, cudaGetErrorString(e));
    return done(rc);
}

int surtr_kdop_calc_batch(surtr_ctx* ctx, const float* verts4, const uint32_t* vert_off, uint32_t n_objects, const float* normals3,
                          const uint32_t* normal_off, float* dist, int32_t* arg, float* planes8)
{
    if (!ctx) return SURTR_ERR_INVALID;
    if (!vert_off || !normal_off || !dist || !arg) return fail(ctx, SURTR_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    const uint64_t nv = vert_off[n_objects], nn = normal_off[n_objects];
    if (!nn) return SURTR_OK;
    if (!normals3 || (nv && !verts4)) return fail(ctx, SURTR_ERR_INVALID, "NULL argument");
    std::vector<uint32_t> obj(nn);
    for (uint32_t o = 0; o < n_objects; o++)
        for (uint32_t e = normal_off[o]; e < normal_off[o + 1]; e++) obj[e] = o;
    DevBuf dv, dvo, dn, dobj, dd, da, dp;
    int rc = SURTR_OK;
    auto done = [&](int code) { dv.release(); dvo.release(); dn.release(); dobj.release(); dd.release(); da.release(); dp.release(); return code; };
    if (dv.reserve(16 * std::max<uint64_t>(1, nv)) || dvo.reserve(4 * ((size_t)n_objects + 1)) || dn.reserve(12 * nn) || dobj.reserve(4 * nn) ||
        dd.reserve(8 * nn) || da.reserve(8 * nn) || dp.reserve(32 * nn))
        return done(fail(ctx, SURTR_ERR_NOMEM, "cudaMalloc failed"));
    cudaMemcpyAsync(dv.p, verts4, 16 * nv, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dvo.p, vert_off, 4 * ((size_t)n_objects + 1), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dn.p, normals3, 12 * nn, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dobj.p, obj.data(), 4 * nn, cudaMemcpyHostToDevice, ctx->stream);
    const int blocks = (int)std::min<uint64_t>((nn * 32 + 255) / 256, (uint64_t)ctx->num_sm * 8);
    kdop_arg_batch_kernel<<<blocks, 256, 0, ctx->stream>>>(dv.as<float4>(), dvo.as<uint32_t>(), dn.as<float>(), dobj.as<uint32_t>(),
                                                           (uint32_t)nn, dd.as<float>(), da.as<int32_t>(), dp.as<float4>());
    cudaMemcpyAsync(dist, dd.p, 8 * nn, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(arg, da.p, 8 * nn, cudaMemcpyDeviceToHost, ctx->stream);
    if (planes8) cudaMemcpyAsync(planes8, dp.p, 32 * nn, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, SURTR_ERR_CUDA, cudaGetErrorString(e));
    return done(rc);
}

int surtr_last_event_ms(surtr_ctx* ctx, float* total_ms, float* clip_ms)
{
    if (!ctx) return SURTR_ERR_INVALID;
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    if (total_ms) CK(cudaEventElapsedTime(total_ms, ctx->ev[0], ctx->ev[surtr_ctx::NPH]));
    if (clip_ms)
    {
        *clip_ms = 0.f;
        if (ctx->profiled_last) CK(cudaEventElapsedTime(clip_ms, ctx->ev[3], ctx->ev[4]));
    }
    return SURTR_OK;
}

int surtr_last_event_phases(surtr_ctx* ctx, float* ms8)
{
    if (!ctx || !ms8) return SURTR_ERR_INVALID;
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    if (!ctx->profiled_last) return fail(ctx, SURTR_ERR_INVALID, "the last event ran without profiling (surtr_set_profiling)");
    for (int i = 0; i < surtr_ctx::NPH; i++) CK(cudaEventElapsedTime(ms8 + i, ctx->ev[i], ctx->ev[i + 1]));
    return SURTR_OK;
}

int surtr_measure_fp32_peak(surtr_ctx* ctx, float* tflops)
{
    if (!ctx || !tflops) return SURTR_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->dbg.reserve(4));
    const int iters = 8192, blocks = ctx->num_sm * 16, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    fma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->dbg.as<float>(), 64, 0.999f, 1.0e-6f);   // warm-up
    CK(cudaEventRecord(e0, ctx->stream));
    fma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->dbg.as<float>(), iters, 0.999f, 1.0e-6f);
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = (float)(2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    return SURTR_OK;
}

int surtr_last_event_launches(const surtr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int surtr_set_profiling(surtr_ctx* ctx, int on)
{
    if (!ctx) return SURTR_ERR_INVALID;
    ctx->profile = on != 0;
    return SURTR_OK;
}

// Development aids (csrc/surtr_debug.h, not part of the drop-in ABI): per-candidate cycle counters of K3.
int surtr_debug_enable(surtr_ctx* ctx, int on)
{
    if (!ctx) return SURTR_ERR_INVALID;
    ctx->debug = on != 0;
    return SURTR_OK;
}

void* surtr_debug_copy_stream(surtr_ctx* ctx) { return ctx ? (void*)ctx->copy_stream : nullptr; }

int surtr_debug_read(surtr_ctx* ctx, uint32_t* out, uint64_t n_cand)
{
    if (!ctx || !ctx->debug) return SURTR_ERR_INVALID;
    const int rc = resolve_event(ctx);
    if (rc) return rc;
    CK(cudaMemcpy(out, ctx->dbg.p, 32 * std::min<uint64_t>(n_cand, ctx->cap_cand), cudaMemcpyDeviceToHost));
    return SURTR_OK;
}
}
