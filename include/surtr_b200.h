/* surtr_b200.h -- C ABI of the B200-native fracture engine (libsurtr_b200.so).
 *
 * Drop-in boundary for ONE path of W298/Surtr: the per-fracture-event cutting of an object's convex
 * pieces by a Voronoi cell set.  The reference exposes no FFI for this path; the seams this ABI
 * replaces are (all citations relative to the reference tree):
 *
 *   m_fractureTask (convex branch)      Inc/Surtr.h:272, Src/Surtr.cpp:1457-1468, 1497-1503
 *   Surtr::ApplyFracture                Src/Surtr.cpp:2098-2149   (all cells x all pieces, bind order)
 *   Surtr::SetExtract                   Src/Surtr.cpp:2151-2155   (face count per fragment)
 *   Poly::ClipPolyhedron (both)         Inc/Poly.h:40-41, Src/Poly.cpp:265-566
 *   Poly::ExtractFaces / Poly::Moments  Inc/Poly.h:36-37, Src/Poly.cpp:55-126
 *   Kdop::KdopContainer::Calc           Inc/Kdop.h:35-37, Src/Kdop.cpp:15-115
 *   Kdop::KdopContainer::ClipWithPolyhedron  Src/Kdop.cpp:166-179 (plane list [Min0,Max0,Min1,...] -> clip)
 *   mass properties for InitCompound    Src/Surtr.cpp:2499-2529 (PhysX updateMassAndInertia, :2520)
 *
 * The calls are batch-granular (a per-pair call cannot feed a GPU): upload a piece set and a cell set
 * (optionally many independent fracture events at once), run the event, read fragments back.
 * INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the context owns device memory;
 *   - every function returns a SURTR_* status, never throws; surtr_last_error() has the text;
 *   - one context per host thread / GPU / stream; calls on one context are ordered on its stream;
 *   - there is NO CPU fallback: without a usable CUDA device surtr_ctx_create fails with SURTR_ERR_NO_DEVICE.
 *
 * Flat polyhedron-set layout ("pieces", "fragments")
 *   verts4    float[4*NV]   x y z 0, all polyhedra back to back
 *   vert_off  u32[n+1]      first vertex of polyhedron i
 *   ring_off  u32[NV+1]     first ring entry of (global) vertex v
 *   ring      u16[NE]       neighbour rings (Poly::Vertex::NeighborVertexVec, Inc/Poly.h:18), indices LOCAL
 *                           to the polyhedron, same cyclic order as the reference
 * Cell-set layout
 *   planes4   float[4*NP]   (nx ny nz d) = VMACH::PolygonFace::FacePlane (Inc/VMACH.h:21), outward normals
 *   plane_off u32[n_cells+1]
 *   cell_verts4 / cvert_off optional: every vertex of every face loop of the cell (duplicates allowed); used
 *                           only for the broad-phase bounds.  NULL = cells are unbounded (no culling by cell).
 * Events
 *   ev_piece_off / ev_cell_off  u32[n_events+1]: event e cuts pieces [ev_piece_off[e], ev_piece_off[e+1]) by
 *   cells [ev_cell_off[e], ev_cell_off[e+1]).  NULL = one event covering everything.
 * Fragment order: event-major, then cell-major, then piece-minor -- the order in which ApplyFracture
 *   consumes its futures (Src/Surtr.cpp:2133-2146); empty results are dropped exactly as there (:1467).
 */
#ifndef SURTR_B200_H
#define SURTR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURTR_OK 0
#define SURTR_ERR_CUDA 1        /* a CUDA runtime call failed */
#define SURTR_ERR_INVALID 2     /* bad argument / call order */
#define SURTR_ERR_NO_DEVICE 3   /* no usable sm_100 device: the product path has no CPU fallback */
#define SURTR_ERR_OVERFLOW 4    /* some (piece, cell) pairs could not be cut: malformed rings or beyond the 16-bit index range */
#define SURTR_ERR_NOMEM 5       /* device or pinned-host allocation failed */

typedef struct surtr_ctx surtr_ctx;

/* One record per non-empty fragment (64 bytes). */
typedef struct surtr_fragment {
    uint32_t cell;        /* global cell index (into the uploaded cell set) */
    uint32_t piece;       /* global piece index (into the uploaded piece set) */
    uint32_t vert_off;    /* first vertex in the fragment vertex array */
    uint16_t n_verts;     /* Poly::Polyhedron::size() of the result */
    uint16_t n_faces;     /* Poly::ExtractFaces(result)->size()  (Src/Poly.cpp:89-126) */
    double volume;        /* Poly::Moments zerothMoment (Src/Poly.cpp:55-87) */
    float centroid[3];    /* Poly::Moments firstMoment */
    float inertia[6];     /* Ixx Iyy Izz Ixy Ixz Iyz about the centroid, unit density (replaces Surtr.cpp:2520) */
    uint32_t n_ring;      /* sum of ring lengths (directed edges) */
} surtr_fragment;

typedef struct surtr_counts {
    uint64_t n_pairs;       /* (piece, cell) pairs of all events */
    uint64_t n_candidates;  /* pairs surviving the k-DOP broad phase */
    uint64_t n_fragments;   /* non-empty clip results */
    uint64_t n_verts;       /* total fragment vertices */
    uint64_t n_ring;        /* total fragment ring entries */
    uint64_t n_seq_cuts;    /* cuts that took the sequential in-plane/anomaly path (diagnostic) */
    uint64_t n_tier2;       /* pairs cut in the large on-chip tier: 65..256 vertex slots or ring degree 9..16 (diagnostic) */
    uint64_t n_tier3;       /* pairs cut in the global-memory tier: more than 256 vertex slots (diagnostic) */
    uint64_t n_tier1b;      /* pairs re-run in the 128-slot warp-per-pair tier (diagnostic) */
    uint64_t n_failed;      /* pairs that could not be cut (malformed rings, > 65520 vertex slots): see surtr_failed_pairs */
} surtr_counts;

/* Device-side view of the last event's fragments (pointers stay valid until the next event / upload). */
typedef struct surtr_device_view {
    const surtr_fragment* fragments;
    const float* verts4;
    const uint32_t* ring_off;   /* n_verts + 1 */
    const uint16_t* ring;
} surtr_device_view;

/* --- context ------------------------------------------------------------------------------------------ */
/* `stream` is a cudaStream_t to order all work on (e.g. torch's current stream), or NULL for a private one. */
int surtr_ctx_create(int device, void* stream, surtr_ctx** out);
void surtr_ctx_destroy(surtr_ctx* ctx);
const char* surtr_last_error(const surtr_ctx* ctx);   /* ctx may be NULL: last create error */
const char* surtr_version(void);

/* Broad-phase direction set: k in {3, 7, 13} (AABB, 14-DOP, 26-DOP).  Default 13: the extra slabs cost a few
 * microseconds in K1 / K2 and keep 40 % of the AABB's dead candidates away from the clipper.  The set never changes
 * the fragments, only how many pairs reach K3. */
int surtr_set_kdop_directions(surtr_ctx* ctx, int k);
/* Which build of the small-tier clipper (K3) an event launches: 0 (default) = by the size of the context's previous event --
 * the latency build (32 warps per SM, 64 registers, plane prefetch) when its candidate pairs fit one wave of warps, the
 * throughput build (40 warps per SM, 48 registers) otherwise; 1 = always throughput (a caller that keeps several small
 * events in flight on different streams), 2 = always latency.  Never changes a result, only the event time. */
int surtr_set_clip_build(surtr_ctx* ctx, int mode);

/* --- inputs (host -> device; replaces the per-task deep copies of Src/Poly.cpp:562, Surtr.cpp:2129-2131) - */
int surtr_upload_pieces(surtr_ctx* ctx, const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off,
                        const uint16_t* ring, uint32_t n_pieces, const uint32_t* ev_piece_off, uint32_t n_events);
int surtr_upload_cells(surtr_ctx* ctx, const float* planes4, const uint32_t* plane_off, const float* cell_verts4,
                       const uint32_t* cvert_off, uint32_t n_cells, const uint32_t* ev_cell_off, uint32_t n_events);
/* PCIe wire format.  The same two calls with tightly packed float3 vertex streams (xyz, 12 bytes per vertex, the
 * size of the reference's Poly::Vertex::Position / VMACH vertex, Inc/Poly.h:15-21) instead of the resident float4
 * layout: a quarter fewer bytes cross the bus, and one small kernel on the context stream widens them to float4
 * (w = 0) in HBM, so every kernel of the event still reads aligned float4 streams.  Results are identical. */
int surtr_upload_pieces3(surtr_ctx* ctx, const float* verts3, const uint32_t* vert_off, const uint32_t* ring_off,
                         const uint16_t* ring, uint32_t n_pieces, const uint32_t* ev_piece_off, uint32_t n_events);
int surtr_upload_cells3(surtr_ctx* ctx, const float* planes4, const uint32_t* plane_off, const float* cell_verts3,
                        const uint32_t* cvert_off, uint32_t n_cells, const uint32_t* ev_cell_off, uint32_t n_events);
/* Recursive re-fracture: make the last event's fragments the piece set of the next event (device side, no
 * copy through the host).  Every fragment becomes one piece; ev_piece_off (host, n_events+1) regroups them, or
 * NULL to keep one event. */
int surtr_fragments_to_pieces(surtr_ctx* ctx, const uint32_t* ev_piece_off, uint32_t n_events);
/* The same with the grouping the recursion of the reference keeps (a compound's fragments stay with the compound,
 * Src/Surtr.cpp:2133-2146 binds every result to the object that was hit): the fragments of event e become the pieces
 * of event e.  The event boundaries are found on the device from the records' cell ids (fragments come in (event,
 * cell, piece) order); n_events + 1 words cross the bus instead of every fragment record. */
int surtr_fragments_to_pieces_per_event(surtr_ctx* ctx);

/* World transform of the resident pieces, in place: replaces the host loop of Surtr::ExecuteFractureRoutine
 * (Src/Surtr.cpp:1846-1852) that runs Poly::Transform (Src/Poly.cpp:580-585) over every piece before DoFracture.
 * matrices16: n_matrices row-major 4x4 matrices exactly as the reference hands them to Poly::Transform (it transposes
 * them itself); piece_matrix[n_pieces] selects the matrix of each piece, NULL = matrix 0 for every piece.  Together
 * with surtr_fragments_to_pieces this keeps a compound on the device across events: only 64 bytes per rigid body
 * travel per event. */
int surtr_transform_pieces(surtr_ctx* ctx, const float* matrices16, const uint32_t* piece_matrix, uint32_t n_matrices);
/* Vertex positions of the resident pieces (after transforms), n_verts x float4, synchronous. */
int surtr_download_pieces(surtr_ctx* ctx, float* verts4);

/* A fracture pattern kept on the device in its own frame and placed per event.  Replaces the host loop
 * Surtr::DoFracture runs over a copy of the pattern before every event (Src/Surtr.cpp:1887-1896): Polygon3D::Scale
 * + Polygon3D::Translate (Src/VMACH.cpp:506-534), which move every face vertex and re-derive every face plane from
 * the face's first three vertices (PolygonFace::ConstructFacePlane, VMACH.cpp:303-310).
 * face_verts4: the VertexVec of every face of every cell, back to back (xyz + pad); face_vert_off[n_faces+1];
 * cell_face_off[n_cells+1].  Every face needs >= 3 vertices. */
int surtr_upload_pattern(surtr_ctx* ctx, const float* face_verts4, const uint32_t* face_vert_off, uint32_t n_faces,
                         const uint32_t* cell_face_off, uint32_t n_cells);
/* Places the resident pattern n_placements times: vertex' = (vertex * scale) + translate per component, planes
 * rebuilt on the device.  The result becomes the cell set of the next event(s), one independent event per placement
 * (pair it with surtr_upload_pieces(..., ev_piece_off, n_placements)); scale3 / translate3 hold 3 floats per
 * placement.  Asynchronous on the context stream. */
int surtr_place_pattern(surtr_ctx* ctx, const float* scale3, const float* translate3, uint32_t n_placements);

/* --- the hot path (replaces Surtr::ApplyFracture + m_fractureTask + SetExtract + mass properties) ------- */
/* Asynchronous on the context stream: K1 k-DOP extents -> K2 broad phase + ordered compaction ->
 * K3 one-warp-per-pair half-space clipping -> K4 fragment assembly with moments. */
int surtr_fracture_event(surtr_ctx* ctx);
/* Waits for the event; grows internal buffers (candidate list, clip tiers, the global tier's vertex and ring slots,
 * fragment arrays) and re-runs it until nothing more is asked for.  The reference's Poly::ClipPolyhedron has no limit on
 * vertex count or valence; here a piece may have up to 65520 vertices (ring entries are 16-bit) and ANY valence.
 * Returns SURTR_ERR_OVERFLOW -- with *out filled in and the event complete -- when some pairs could not be cut
 * (malformed rings, more than 65520 vertex slots during a cut, more than 65535 faces): they are listed by
 * surtr_failed_pairs and produce no fragment; every other fragment of the event is valid and the downloads succeed. */
int surtr_event_counts(surtr_ctx* ctx, surtr_counts* out);
/* (piece, cell) of every failed pair of the last event, 2 x uint32 per pair; *n_pairs = how many there are. */
int surtr_failed_pairs(surtr_ctx* ctx, uint32_t* piece_cell, uint64_t capacity_pairs, uint64_t* n_pairs);

/* --- outputs ------------------------------------------------------------------------------------------ */
/* Any pointer may be NULL to skip that array.  Sizes come from surtr_event_counts. */
int surtr_download_fragments(surtr_ctx* ctx, surtr_fragment* fragments, float* verts4, uint32_t* ring_off,
                             uint16_t* ring);
/* Same, but returns once the copies are enqueued (after waiting for the event itself, which sizes them).  They run
 * on a copy stream of the context: the next event may be uploaded and launched right away -- its uploads and its
 * K1-K3 overlap the copies, only K4 (which rewrites the fragment arrays) is ordered behind them.  The host buffers
 * (pinned, for the copies to be asynchronous) are complete when surtr_sync has returned, or any call that waits for an
 * event launched after this one (its surtr_event_counts / download).  Lets one host thread keep several contexts busy
 * without ever blocking on PCIe. */
int surtr_download_fragments_async(surtr_ctx* ctx, surtr_fragment* fragments, float* verts4, uint32_t* ring_off,
                                   uint16_t* ring);
/* PCIe wire format of the fragments: float3 positions (n_verts x 3 floats) and ONE BYTE per vertex for its ring
 * length (ring_off is their exclusive prefix sum; a Poly::Vertex::NeighborVertexVec never exceeds 16 entries on
 * this path) instead of float4 + a 32-bit offset: 13 instead of 20 bytes per vertex.  A small kernel on the
 * context's copy stream packs the resident arrays before the copies; records and ring entries travel unchanged.
 * Same completion rules as surtr_download_fragments(_async). */
int surtr_download_fragments_packed(surtr_ctx* ctx, surtr_fragment* fragments, float* verts3, uint8_t* ring_len,
                                    uint16_t* ring);
int surtr_download_fragments_packed_async(surtr_ctx* ctx, surtr_fragment* fragments, float* verts3, uint8_t* ring_len,
                                          uint16_t* ring);
/* One-copy transfers.  Host<->device copy rates on PCIe depend strongly on the size of each copy (measured on the B200
 * boxes: 4 MB copies reach a third of the rate of 64 MB copies when both directions are busy, profiles/r2_e2e_sweep.md), so
 * a caller that streams event batches moves ONE blob per direction instead of 8 + 4 arrays.
 *
 * Both blobs are COMPACT wire formats -- the loop is bound by the bytes that cross PCIe (bench.py: e2e.achieved_gbs
 * against dma_ceiling), so nothing travels wider than it has to: float3 positions, ONE length byte per vertex instead
 * of a 32-bit ring offset, and one-byte ring entries whenever no polyhedron of the batch has more than 256 vertices
 * (ring_entry_bytes = 1; a Poly::Vertex::NeighborVertexVec entry is a vertex index local to its polyhedron).  One small
 * kernel per direction converts between the wire format and the resident arrays (float4, 32-bit offsets, 16-bit entries).
 *
 * Input blob, sections at the byte offsets surtr_input_blob_layout returns (every section 256-byte aligned):
 *   verts3       float[3 * n_piece_verts]   piece vertices
 *   vert_off     u32[n_pieces + 1]          first vertex of every piece
 *   ring_base    u32[n_pieces + 1]          first ring entry of every piece (ring_base[n_pieces] = n_piece_ring)
 *   ring_len     u8[n_piece_verts]          neighbours of every vertex
 *   ring         u8 | u16 [n_piece_ring]    neighbour ids, local to the piece (ring_entry_bytes = 1 needs pieces of <= 256 vertices)
 *   planes4, plane_off, cell_verts3, cvert_off   as surtr_upload_cells3 (n_cell_verts == 0 = unbounded cells)
 *   ev_piece_off, ev_cell_off   u32[n_events + 1] each, ignored when n_events == 0 = one event
 * The blob is copied with one cudaMemcpyAsync on the context stream (pinned memory for it to be asynchronous; the per-piece
 * and per-event offsets are also read on the host for validation before the context is touched). */
typedef struct surtr_in_layout {
    uint64_t verts3, vert_off, ring_base, ring_len, ring, planes4, plane_off, cell_verts3, cvert_off, ev_piece_off, ev_cell_off, total;
} surtr_in_layout;
int surtr_input_blob_layout(uint32_t n_pieces, uint64_t n_piece_verts, uint64_t n_piece_ring, uint32_t n_cells, uint64_t n_planes,
                            uint64_t n_cell_verts, uint32_t n_events, uint32_t ring_entry_bytes, surtr_in_layout* out);
int surtr_upload_blob(surtr_ctx* ctx, const void* blob, uint32_t n_pieces, uint64_t n_piece_verts, uint64_t n_piece_ring,
                      uint32_t n_cells, uint64_t n_planes, uint64_t n_cell_verts, uint32_t n_events, uint32_t ring_entry_bytes);
/* Output blob = records | float3 positions | one byte of ring length per vertex | ring entries (one byte each when no
 * fragment of the event has more than 256 vertices, else two: out->ring_entry_bytes says which) at the byte offsets
 * returned in *out, assembled on the device and moved with ONE copy on the context's copy stream.  Waits for the event
 * (which sizes the blob); SURTR_ERR_INVALID with *out filled in when `capacity` is too small (64 n_fragments + 13 n_verts
 * + 2 n_ring + 1024 always suffices).  Completion rules as surtr_download_fragments_async.  `host_blob` may also point
 * to DEVICE memory (the copy is issued with cudaMemcpyDefault): that is how the multi-GPU gather gets one contiguous
 * blob per rank. */
typedef struct surtr_out_layout {
    uint64_t fragments, verts3, ring_len, ring, total;   /* byte offsets, total size */
    uint64_t n_fragments, n_verts, n_ring;
    uint64_t ring_entry_bytes;                           /* 1 or 2 */
} surtr_out_layout;
int surtr_download_blob_async(surtr_ctx* ctx, void* host_blob, uint64_t capacity, surtr_out_layout* out);
int surtr_sync(surtr_ctx* ctx);
int surtr_device_fragments(surtr_ctx* ctx, surtr_device_view* out);

/* --- k-DOP (replaces Kdop::KdopContainer::Calc(const Poly::Polyhedron&), Src/Kdop.cpp:92-115) ----------- */
/* For each of k normals: dist[2i] = MinDist, dist[2i+1] = MaxDist (float values), arg[2i], arg[2i+1] = index of
 * the first extremal vertex (strict </> as the reference), planes8[8i] = MinPlane, MaxPlane
 * (Plane(v,-n), Plane(v,n)).  Host buffers in and out. */
int surtr_kdop_calc(surtr_ctx* ctx, const float* verts4, uint32_t n_verts, const float* normals3, uint32_t k,
                    float* dist, int32_t* arg, float* planes8);

/* Batched form for many small objects (the refitting pass, Surtr::Refitting, Src/Surtr.cpp:2405-2413 over
 * m_refittingTask :1449-1455): object i has vertices [vert_off[i], vert_off[i+1]) and its own normals
 * [normal_off[i], normal_off[i+1]); outputs are indexed by global normal index, arg is local to the object. */
int surtr_kdop_calc_batch(surtr_ctx* ctx, const float* verts4, const uint32_t* vert_off, uint32_t n_objects,
                          const float* normals3, const uint32_t* normal_off, float* dist, int32_t* arg, float* planes8);

/* Timing of the last surtr_fracture_event in milliseconds (CUDA events on the context stream).  clip_ms (the K3
 * small-tier kernel alone) is only measured while profiling is on: the extra events sit between the kernels of an
 * event and serialise their programmatic dependent launches, so they are off by default. */
int surtr_last_event_ms(surtr_ctx* ctx, float* total_ms, float* clip_ms);
int surtr_set_profiling(surtr_ctx* ctx, int on);
/* Per-kernel durations of the last event in milliseconds (profiling must have been on; measurement only, SURVEY.md
 * section 8d): ms8[0] K1 k-DOP extents, [1] K2a broad-phase masks, [2] K2b pair compaction, [3] K3 small tier
 * (clip_fast_kernel<2>), [4] the 128-slot, large and global tiers (0 when not launched), [5] K4 scan, [6] K4 gather, [7] counters to
 * the host + reset.  CUDA events on the context stream between the launches. */
int surtr_last_event_phases(surtr_ctx* ctx, float* ms8);
/* Number of kernels the last surtr_fracture_event launched. */
/* Measurement helper: FP32 FMA throughput of the device (TFLOP/s) from a register-resident FFMA kernel on the context
 * stream -- the denominator of the secondary (FP32 pipe) roofline in bench.py. */
int surtr_measure_fp32_peak(surtr_ctx* ctx, float* tflops);
int surtr_last_event_launches(const surtr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
