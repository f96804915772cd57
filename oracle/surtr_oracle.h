/* oracle/surtr_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or executed
 * from the product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it).
 *
 * Plain-C restatement ("port") of the reference's convex-piece cutting path over flat arrays.
 * Parity is PINNED: tests/test_oracle_port.py checks every function below bit-for-bit against
 * oracle/_ref/libsurtr_ref.so (the reference's own Src/Poly.cpp / Src/Kdop.cpp compiled headless)
 * and against the committed fixtures under tests/golden/ generated from that library.
 * Build with -ffp-contract=off (see oracle/Makefile): every float product/sum rounds separately.
 */
#ifndef SURTR_ORACLE_H
#define SURTR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Poly::ComparePlanePoint (Src/Poly.cpp:716-723): +1 keep, 0 in-plane, -1 clipped. */
int so_compare_plane_point(const float plane[4], const float p[3]);
/* Poly::PlaneLineIntersection (Src/Poly.cpp:746-751). */
void so_plane_line_intersection(const float a[3], const float b[3], const float plane[4], float out[3]);
/* SimpleMath Plane(p1,p2,p3) / Plane(point,normal) (SimpleMath.inl:2773-2788). */
void so_plane_from_points(const float a[3], const float b[3], const float c[3], float out[4]);
void so_plane_from_point_normal(const float p[3], const float n[3], float out[4]);

/* A growable polyhedron: vertex i has ring[i*SO_MAXDEG .. +deg[i]). */
#define SO_MAXDEG 64
typedef struct so_poly {
    int nv, cap;
    float* pos;   /* 3 per vertex */
    int* deg;
    int* ring;    /* SO_MAXDEG per vertex */
    int* comp;
    int* id;
} so_poly;

void so_poly_init(so_poly* p);
void so_poly_free(so_poly* p);
void so_poly_load(so_poly* p, const float* verts4, const uint32_t* ring_off, const uint16_t* ring, uint32_t v0,
                  uint32_t v1);

/* Poly::ClipPolyhedron(Polyhedron&, const std::vector<Plane>&) (Src/Poly.cpp:265-500).
 * Returns 0, or -1 if a ring outgrew SO_MAXDEG (never seen on the tested inputs). */
int so_clip(so_poly* p, const float* planes4, int nplanes);

/* Poly::ExtractFaces (Src/Poly.cpp:89-126).  face_off[nf+1], face_idx[sum loop lengths]; either may be NULL
 * to count only.  Returns the number of faces. */
int so_extract_faces(const so_poly* p, uint32_t* face_off, uint16_t* face_idx);

/* Poly::Moments (Src/Poly.cpp:55-87): volume (double) and centroid (float3). */
void so_moments(const so_poly* p, double* volume, float centroid[3]);

/* Inertia tensor about the centroid at unit density, double precision polyhedral integral over the same face
 * fans.  No reference counterpart (PhysX computes it, Surtr.cpp:2520): independent check for kernel K4.
 * out = {Ixx, Iyy, Izz, Ixy, Ixz, Iyz}. */
void so_inertia(const so_poly* p, double out[6]);

/* Kdop::KdopContainer::Calc(const Poly::Polyhedron&) (Src/Kdop.cpp:92-115): per normal min/max t (float values),
 * index of the first extremal vertex, and MinPlane/MaxPlane.  dist[2k], arg[2k], planes[8k]. */
void so_kdop_calc(const float* verts4, uint32_t nv, const float* normals3, uint32_t k, float* dist, int32_t* arg,
                  float* planes);

/* Restatement of Surtr::ApplyFracture + m_fractureTask, convex branch (Src/Surtr.cpp:1457-1468, 2098-2149):
 * every cell (plane list) against every piece, cell-major / piece-minor, non-empty results only.
 * Results are appended to caller-provided flat arrays with the given capacities; returns the number of
 * fragments, or -1 on capacity overflow.  frag_rec[i] = {cell, piece, nv, nf}; moments optional (NULL). */
typedef struct so_out {
    float* verts4;       uint64_t cap_verts;
    uint32_t* vert_off;  /* cap_frags+1 */
    uint32_t* ring_off;  /* cap_verts+1 */
    uint16_t* ring;      uint64_t cap_ring;
    uint32_t* rec;       uint64_t cap_frags; /* 4 per fragment */
    double* volume;      /* cap_frags or NULL */
    float* centroid;     /* 3*cap_frags or NULL */
    double* inertia;     /* 6*cap_frags or NULL */
} so_out;

int64_t so_apply_fracture(const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off,
                          const uint16_t* ring, uint32_t n_pieces, const float* planes4, const uint32_t* plane_off,
                          uint32_t n_cells, so_out* out);

/* Face planes of a polyhedron exactly as a VMACH cell face gets them (Src/VMACH.cpp:289-310): loop vertices
 * are appended with PolygonFace::AddVertex (duplicates closer than 1e-12 dropped) and the plane is
 * Plane(v0, v1, v2) of the first three kept vertices.  Returns the face count; planes4 has 4 floats per face
 * (zeros when fewer than three distinct vertices). */
int so_face_planes(const so_poly* p, float* planes4);

/* NEW derivation shared with the product's host library (replaces the voro++ call sites Surtr.cpp:2007-2067):
 * cell i = Poly::GetBB() clipped by Plane((Si+Sj)*0.5, Sj-Si) for every neighbour j in nb (ascending).
 * Appends the cells to `out` (rec = {i, 0, nv, nf}); plane_off[n+1] / planes4 receive the face planes. */
int64_t so_voronoi_cells(const float* seeds3, uint32_t n, const uint32_t* nb_off, const uint32_t* nb_idx, so_out* out,
                         uint32_t* plane_off, float* planes4, uint64_t cap_planes);

/* Clip polyhedron i by its own plane list [pl_off[i], pl_off[i+1]) -- Poly.cpp:265 in place; EVERY input gets
 * one output entry (possibly with 0 vertices). */
int64_t so_clip_each(const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off, const uint16_t* ring,
                     uint32_t n, const float* planes4, const uint32_t* pl_off, so_out* out);

/* so_face_planes for every polyhedron of a flat set; returns the plane count or -1 on overflow. */
int64_t so_face_planes_set(const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off,
                           const uint16_t* ring, uint32_t n, uint32_t* plane_off, float* planes4, uint64_t cap_planes);

#ifdef __cplusplus
}
#endif
#endif
