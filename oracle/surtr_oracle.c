/* oracle/surtr_oracle.c -- TEST INFRASTRUCTURE ONLY.  See surtr_oracle.h.
 *
 * Plain-C restatement of the reference's clipping path.  Each function cites the reference lines it
 * follows; arithmetic follows SURVEY.md Appendix A (float32, every product/sum rounded separately,
 * Vector3/float == multiply by 1.f/s).  Compile with -ffp-contract=off.
 */
#include "surtr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- SimpleMath subset (ThirdParty/Inc/SimpleMath.inl:918-946, 2773-2788) ---- */
static inline float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline void cross3(const float* a, const float* b, float* o)
{
    const float x = a[1] * b[2] - a[2] * b[1];
    const float y = a[2] * b[0] - a[0] * b[2];
    const float z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline void normalize3(float* v)
{
    const float lsq = dot3(v, v);
    const float len = sqrtf(lsq);
    if (lsq == 0.f) { v[0] = v[1] = v[2] = 0.f; return; }
    if (isinf(lsq)) { v[0] = v[1] = v[2] = NAN; return; }
    v[0] = v[0] / len; v[1] = v[1] / len; v[2] = v[2] / len;
}

void so_plane_from_points(const float a[3], const float b[3], const float c[3], float out[4])
{
    const float d1[3] = { a[0] - b[0], a[1] - b[1], a[2] - b[2] };
    const float d2[3] = { a[0] - c[0], a[1] - c[1], a[2] - c[2] };
    float n[3];
    cross3(d1, d2, n);
    normalize3(n);
    out[0] = n[0]; out[1] = n[1]; out[2] = n[2];
    out[3] = -dot3(n, a);
}

void so_plane_from_point_normal(const float p[3], const float n[3], float out[4])
{
    out[0] = n[0]; out[1] = n[1]; out[2] = n[2];
    out[3] = -dot3(p, n);
}

/* Src/Poly.cpp:716-723 */
int so_compare_plane_point(const float plane[4], const float p[3])
{
    const float sgndist = plane[3] + dot3(plane, p);
    if ((double)fabsf(sgndist) < 1.0e-10)
        return 0;
    const double m = -(double)sgndist;
    return m > 0.0 ? 1 : (m < 0.0 ? -1 : 0);
}

/* Src/Poly.cpp:725-744 */
static int compare_plane_bb(const float plane[4], const float lo[3], const float hi[3])
{
    int cmin = 2, cmax = -2;
    /* corner order of the reference is irrelevant to min/max */
    for (int k = 0; k < 8; k++)
    {
        const float p[3] = { (k & 1) ? hi[0] : lo[0], (k & 2) ? hi[1] : lo[1], (k & 4) ? hi[2] : lo[2] };
        const int c = so_compare_plane_point(plane, p);
        if (c < cmin) cmin = c;
        if (c > cmax) cmax = c;
    }
    if (cmin >= 0) return 1;
    if (cmax <= 0) return -1;
    return 0;
}

/* Src/Poly.cpp:746-751 */
void so_plane_line_intersection(const float a[3], const float b[3], const float plane[4], float out[3])
{
    const float asgn = plane[3] + dot3(plane, a);
    const float bsgn = plane[3] + dot3(plane, b);
    const float r = 1.f / (bsgn - asgn);
    for (int k = 0; k < 3; k++)
        out[k] = (a[k] * bsgn - b[k] * asgn) * r;
}

/* ---- polyhedron container ---- */
void so_poly_init(so_poly* p) { memset(p, 0, sizeof(*p)); }

void so_poly_free(so_poly* p)
{
    free(p->pos); free(p->deg); free(p->ring); free(p->comp); free(p->id);
    memset(p, 0, sizeof(*p));
}

static void reserve(so_poly* p, int n)
{
    if (n <= p->cap) return;
    int cap = p->cap ? p->cap : 64;
    while (cap < n) cap *= 2;
    p->pos = (float*)realloc(p->pos, sizeof(float) * 3 * cap);
    p->deg = (int*)realloc(p->deg, sizeof(int) * cap);
    p->ring = (int*)realloc(p->ring, sizeof(int) * SO_MAXDEG * cap);
    p->comp = (int*)realloc(p->comp, sizeof(int) * cap);
    p->id = (int*)realloc(p->id, sizeof(int) * cap);
    p->cap = cap;
}

void so_poly_load(so_poly* p, const float* verts4, const uint32_t* ring_off, const uint16_t* ring, uint32_t v0,
                  uint32_t v1)
{
    const int n = (int)(v1 - v0);
    reserve(p, n);
    p->nv = n;
    for (int i = 0; i < n; i++)
    {
        const uint32_t v = v0 + i;
        p->pos[3 * i] = verts4[4 * v]; p->pos[3 * i + 1] = verts4[4 * v + 1]; p->pos[3 * i + 2] = verts4[4 * v + 2];
        const int d = (int)(ring_off[v + 1] - ring_off[v]);
        p->deg[i] = d;
        for (int j = 0; j < d && j < SO_MAXDEG; j++)
            p->ring[i * SO_MAXDEG + j] = ring[ring_off[v] + j];
        p->comp[i] = 1;  /* Poly.cpp:11-13 default */
        p->id[i] = -1;
    }
}

#define RING(p, i) ((p)->ring + (size_t)(i) * SO_MAXDEG)

/* FaceLoop (Src/Poly.cpp:34-41): the ring entry just before `vprev`, wrapping; if vprev is absent std::find
 * returns end() and *(end-1) is the last entry. */
static int face_loop(const so_poly* p, int v, int vprev)
{
    const int* r = RING(p, v);
    const int d = p->deg[v];
    int k = 0;
    while (k < d && r[k] != vprev) k++;
    if (k == 0) return r[d - 1];
    return r[k - 1];
}

static int ring_insert(so_poly* p, int* r, int* d, int at, int value)
{
    (void)p;
    if (*d >= SO_MAXDEG) return -1;
    for (int k = *d; k > at; k--) r[k] = r[k - 1];
    r[at] = value;
    (*d)++;
    return 0;
}

static void bbox(const so_poly* p, float lo[3], float hi[3], int only_kept)
{
    lo[0] = lo[1] = lo[2] = FLT_MAX;   /* reference keeps doubles holding float values (Poly.cpp:276-287) */
    hi[0] = hi[1] = hi[2] = -FLT_MAX;
    for (int i = 0; i < p->nv; i++)
    {
        if (only_kept && p->comp[i] < 0) continue;
        for (int k = 0; k < 3; k++)
        {
            const float c = p->pos[3 * i + k];
            if (c < lo[k]) lo[k] = c;
            if (c > hi[k]) hi[k] = c;
        }
    }
}

/* Src/Poly.cpp:265-500 */
int so_clip(so_poly* p, const float* planes4, int nplanes)
{
    float lo[3], hi[3];
    bbox(p, lo, hi, 0);
    int* old_ring = NULL;
    int* old_deg = NULL;
    int old_cap = 0;
    int status = 0;

    for (int kplane = 0; kplane < nplanes && p->nv > 0; kplane++)
    {
        const float* plane = planes4 + 4 * kplane;

        /* :297-299 bounding-box shortcut */
        const int boxcomp = compare_plane_bb(plane, lo, hi);
        int above = boxcomp == 1;
        int below = boxcomp == -1;

        /* :303-319 */
        if (!(above || below))
        {
            above = 1; below = 1;
            for (int i = 0; i < p->nv; i++)
            {
                p->comp[i] = so_compare_plane_point(plane, p->pos + 3 * i);
                if (p->comp[i] == 1) below = 0;
                else if (p->comp[i] == -1) above = 0;
            }
        }

        if (below) { p->nv = 0; break; }   /* :322-327 */
        if (above) continue;               /* :328 */

        /* :332-363 insert new vertices on straddling edges */
        const int nverts0 = p->nv;
        for (int i = 0; i < nverts0; i++)
        {
            if (p->comp[i] != -1) continue;
            const int nneigh = p->deg[i];
            for (int j = 0; j < nneigh; j++)
            {
                const int jn = RING(p, i)[j];
                if (p->comp[jn] > 0)
                {
                    const int inew = p->nv;
                    reserve(p, inew + 1);
                    p->nv = inew + 1;
                    so_plane_line_intersection(p->pos + 3 * i, p->pos + 3 * jn, plane, p->pos + 3 * inew);
                    p->comp[inew] = 2;
                    p->id[inew] = -1;
                    p->deg[inew] = 2;
                    RING(p, inew)[0] = i;
                    RING(p, inew)[1] = jn;
                    int* rj = RING(p, jn);
                    int k = 0;
                    while (k < p->deg[jn] && rj[k] != i) k++;
                    if (k < p->deg[jn]) rj[k] = inew;   /* std::find hit; a miss would be UB in the reference */
                    RING(p, i)[j] = inew;
                }
            }
        }
        const int nverts = p->nv;

        /* :367-369 snapshot of all rings */
        if (old_cap < nverts)
        {
            old_cap = p->cap;
            old_ring = (int*)realloc(old_ring, sizeof(int) * SO_MAXDEG * old_cap);
            old_deg = (int*)realloc(old_deg, sizeof(int) * old_cap);
        }
        memcpy(old_ring, p->ring, sizeof(int) * SO_MAXDEG * nverts);
        memcpy(old_deg, p->deg, sizeof(int) * nverts);

        /* :370-425 patch links to clipped vertices; new vertices first */
        for (int ii = 0; ii < nverts; ii++)
        {
            const int i = (ii + nverts0) % nverts;
            if (!(p->comp[i] == 0 || p->comp[i] == 2)) continue;
            const int nneigh = p->deg[i];
            for (int j = 0; j < nneigh; j++)
            {
                const int jn = RING(p, i)[j];
                if (jn < 0 || p->comp[jn] != -1) continue;
                int iprev = i, inext = jn, itmp;
                int k = 0;
                while (p->comp[inext] == -1 && k++ < nverts)
                {
                    itmp = inext;
                    inext = face_loop(p, inext, iprev);
                    iprev = itmp;
                }
                if (RING(p, i)[(j + 1) % p->deg[i]] == inext || inext == i)
                {
                    RING(p, i)[j] = -1;
                }
                else
                {
                    RING(p, i)[j] = inext;
                    if (p->comp[inext] == 2)
                    {
                        if (ring_insert(p, RING(p, inext), &p->deg[inext], 0, i)) status = -1;
                        int* od = &old_deg[inext];
                        if (ring_insert(p, old_ring + (size_t)inext * SO_MAXDEG, od, 0, -1)) status = -1;
                    }
                    else
                    {
                        int* orr = old_ring + (size_t)inext * SO_MAXDEG;
                        int off = 0;
                        while (off < old_deg[inext] && orr[off] != iprev) off++;
                        if (ring_insert(p, RING(p, inext), &p->deg[inext], off, i)) status = -1;
                        if (ring_insert(p, orr, &old_deg[inext], off, i)) status = -1;
                    }
                }
                if (status) goto done;
            }
        }
        /* :426-431 drop the -1 marks */
        for (int i = 0; i < nverts; i++)
        {
            int* r = RING(p, i);
            int w = 0;
            for (int k = 0; k < p->deg[i]; k++)
                if (r[k] != -1) r[w++] = r[k];
            p->deg[i] = w;
        }

        /* :433-462 bypass kept vertices with exactly two neighbours */
        int updated = 1;
        while (updated)
        {
            updated = 0;
            for (int i = 0; i < nverts; i++)
            {
                if (p->comp[i] >= 0 && p->deg[i] == 2)
                {
                    updated = 1;
                    const int iprev = RING(p, i)[0];
                    const int inext = RING(p, i)[1];
                    int k = 0;
                    while (k < p->deg[iprev] && RING(p, iprev)[k] != i) ++k;
                    if (k < p->deg[iprev]) RING(p, iprev)[k] = inext;
                    k = 0;
                    while (k < p->deg[inext] && RING(p, inext)[k] != i) ++k;
                    if (k < p->deg[inext]) RING(p, inext)[k] = iprev;
                    p->comp[i] = -1;
                }
            }
        }

        /* :464-499 renumber, compress, recompute the box */
        int n = 0;
        for (int i = 0; i < nverts; i++)
            if (p->comp[i] >= 0) p->id[i] = n++;
        bbox(p, lo, hi, 1);
        for (int i = 0; i < nverts; i++)
            if (p->comp[i] >= 0)
                for (int j = 0; j < p->deg[i]; j++)
                    RING(p, i)[j] = p->id[RING(p, i)[j]];
        for (int i = 0; i < nverts; i++)
        {
            if (p->comp[i] < 0) continue;
            const int t = p->id[i];
            if (t != i)
            {
                memcpy(p->pos + 3 * t, p->pos + 3 * i, sizeof(float) * 3);
                memcpy(RING(p, t), RING(p, i), sizeof(int) * p->deg[i]);
                p->deg[t] = p->deg[i];
                p->comp[t] = p->comp[i];
                p->id[t] = p->id[i];
            }
        }
        p->nv = n;
        if (p->nv < 4) p->nv = 0;
    }
done:
    free(old_ring);
    free(old_deg);
    return status;
}

/* Src/Poly.cpp:89-126.  visited[(v, slot)] replaces the std::set of directed edges: a directed edge (a,b) is
 * identified by the slot of b in a's ring. */
int so_extract_faces(const so_poly* p, uint32_t* face_off, uint16_t* face_idx)
{
    const int nv = p->nv;
    unsigned char* visited = (unsigned char*)calloc((size_t)(nv > 0 ? nv : 1) * SO_MAXDEG, 1);
    int nf = 0;
    uint32_t w = 0;
    if (face_off) face_off[0] = 0;
    for (int i = 0; i < nv; i++)
    {
        if (p->comp[i] < 0) continue;
        for (int s = 0; s < p->deg[i]; s++)
        {
            if (visited[(size_t)i * SO_MAXDEG + s]) continue;
            const int adj = RING(p, i)[s];
            if (face_idx) face_idx[w] = (uint16_t)i;
            w++;
            int iprev = i, inext = adj, itmp;
            int guard = 0;
            while (inext != i && guard++ <= nv * SO_MAXDEG)
            {
                /* mark (iprev -> inext) */
                const int* r = RING(p, iprev);
                for (int k = 0; k < p->deg[iprev]; k++)
                    if (r[k] == inext) { visited[(size_t)iprev * SO_MAXDEG + k] = 1; break; }
                if (face_idx) face_idx[w] = (uint16_t)inext;
                w++;
                itmp = inext;
                inext = face_loop(p, inext, iprev);
                iprev = itmp;
            }
            {
                const int* r = RING(p, iprev);
                for (int k = 0; k < p->deg[iprev]; k++)
                    if (r[k] == inext) { visited[(size_t)iprev * SO_MAXDEG + k] = 1; break; }
            }
            nf++;
            if (face_off) face_off[nf] = w;
        }
    }
    free(visited);
    return nf;
}

static int count_face_entries(const so_poly* p, int* nf_out)
{
    /* upper bound: every directed edge belongs to exactly one loop */
    int e = 0;
    for (int i = 0; i < p->nv; i++) e += p->deg[i];
    *nf_out = e;
    return e;
}

/* Src/Poly.cpp:55-87 */
void so_moments(const so_poly* p, double* volume, float centroid[3])
{
    double zeroth = 0.0;
    float first[3] = { 0.f, 0.f, 0.f };
    if (p->nv > 3)
    {
        int maxf;
        const int ne = count_face_entries(p, &maxf);
        uint32_t* foff = (uint32_t*)malloc(sizeof(uint32_t) * (maxf + 2));
        uint16_t* fidx = (uint16_t*)malloc(sizeof(uint16_t) * (ne + 2));
        const int nf = so_extract_faces(p, foff, fidx);
        const float* origin = p->pos;
        for (int f = 0; f < nf; f++)
        {
            const uint16_t* facet = fidx + foff[f];
            const unsigned n = foff[f + 1] - foff[f];
            float p0[3], p1[3], p2[3], c[3];
            for (int k = 0; k < 3; k++) p0[k] = p->pos[3 * facet[0] + k] - origin[k];
            for (unsigned k = 1u; k + 1 < n; ++k)
            {
                const int i = facet[k];
                const int j = facet[(k + 1) % n];
                for (int q = 0; q < 3; q++) p1[q] = p->pos[3 * i + q] - origin[q];
                for (int q = 0; q < 3; q++) p2[q] = p->pos[3 * j + q] - origin[q];
                cross3(p1, p2, c);
                const float dV = dot3(p0, c);
                zeroth += dV;
                for (int q = 0; q < 3; q++)
                    first[q] = first[q] + ((p0[q] + p1[q]) + p2[q]) * dV;
            }
        }
        zeroth /= 6.0;
        /* safeInv (Poly.cpp:33) then Vector3::operator*=(float) */
        const double x = 24.0 * zeroth;
        const double inv = (x >= 0.0 ? 1.0 : -1.0) / fmax(1.0e-30, fabs(x));
        const float s = (float)inv;
        for (int q = 0; q < 3; q++)
            first[q] = first[q] * s + origin[q];
        free(foff);
        free(fidx);
    }
    *volume = zeroth;
    centroid[0] = first[0]; centroid[1] = first[1]; centroid[2] = first[2];
}

void so_inertia(const so_poly* p, double out[6])
{
    for (int k = 0; k < 6; k++) out[k] = 0.0;
    if (p->nv <= 3) return;
    int maxf;
    const int ne = count_face_entries(p, &maxf);
    uint32_t* foff = (uint32_t*)malloc(sizeof(uint32_t) * (maxf + 2));
    uint16_t* fidx = (uint16_t*)malloc(sizeof(uint16_t) * (ne + 2));
    const int nf = so_extract_faces(p, foff, fidx);
    const double o[3] = { p->pos[0], p->pos[1], p->pos[2] };
    double vol6 = 0.0, m1[3] = { 0, 0, 0 }, cov[3][3] = { { 0 } };
    for (int f = 0; f < nf; f++)
    {
        const uint16_t* facet = fidx + foff[f];
        const unsigned n = foff[f + 1] - foff[f];
        double a[3], b[3], c[3];
        for (int q = 0; q < 3; q++) a[q] = p->pos[3 * facet[0] + q] - o[q];
        for (unsigned k = 1u; k + 1 < n; ++k)
        {
            for (int q = 0; q < 3; q++) b[q] = p->pos[3 * facet[k] + q] - o[q];
            for (int q = 0; q < 3; q++) c[q] = p->pos[3 * facet[k + 1] + q] - o[q];
            const double d = a[0] * (b[1] * c[2] - b[2] * c[1]) + a[1] * (b[2] * c[0] - b[0] * c[2]) +
                             a[2] * (b[0] * c[1] - b[1] * c[0]);
            vol6 += d;
            for (int q = 0; q < 3; q++) m1[q] += d * (a[q] + b[q] + c[q]);
            for (int r = 0; r < 3; r++)
                for (int s = 0; s < 3; s++)
                    cov[r][s] += d * ((a[r] + b[r] + c[r]) * (a[s] + b[s] + c[s]) + a[r] * a[s] + b[r] * b[s] +
                                      c[r] * c[s]);
        }
    }
    free(foff);
    free(fidx);
    const double V = vol6 / 6.0;
    if (V == 0.0) return;
    double cm[3];
    for (int q = 0; q < 3; q++) cm[q] = m1[q] / 24.0 / V;
    double C[3][3];
    for (int r = 0; r < 3; r++)
        for (int s = 0; s < 3; s++)
            C[r][s] = cov[r][s] / 120.0 - V * cm[r] * cm[s];
    out[0] = C[1][1] + C[2][2];
    out[1] = C[0][0] + C[2][2];
    out[2] = C[0][0] + C[1][1];
    out[3] = -C[0][1];
    out[4] = -C[0][2];
    out[5] = -C[1][2];
}

/* Src/Kdop.cpp:92-115 */
void so_kdop_calc(const float* verts4, uint32_t nv, const float* normals3, uint32_t k, float* dist, int32_t* arg,
                  float* planes)
{
    for (uint32_t e = 0; e < k; e++)
    {
        const float* n = normals3 + 3 * e;
        double mind = DBL_MAX, maxd = -DBL_MAX;
        int32_t amin = -1, amax = -1;
        for (uint32_t v = 0; v < nv; v++)
        {
            const float t = dot3(verts4 + 4 * v, n);
            if (mind > t) { mind = t; amin = (int32_t)v; }
            if (maxd < t) { maxd = t; amax = (int32_t)v; }
        }
        dist[2 * e] = (float)mind; dist[2 * e + 1] = (float)maxd;
        arg[2 * e] = amin; arg[2 * e + 1] = amax;
        if (planes)
        {
            const float neg[3] = { -n[0], -n[1], -n[2] };
            if (amin >= 0) so_plane_from_point_normal(verts4 + 4 * amin, neg, planes + 8 * e);
            if (amax >= 0) so_plane_from_point_normal(verts4 + 4 * amax, n, planes + 8 * e + 4);
        }
    }
}

/* Src/Surtr.cpp:1457-1468 + 2098-2149 (convex branch, non-partial) */
int64_t so_apply_fracture(const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off,
                          const uint16_t* ring, uint32_t n_pieces, const float* planes4, const uint32_t* plane_off,
                          uint32_t n_cells, so_out* out)
{
    so_poly w;
    so_poly_init(&w);
    uint64_t nf = 0, nvtx = 0, nring = 0;
    out->vert_off[0] = 0;
    out->ring_off[0] = 0;
    for (uint32_t c = 0; c < n_cells; c++)
    {
        const float* pl = planes4 + 4 * (size_t)plane_off[c];
        const int npl = (int)(plane_off[c + 1] - plane_off[c]);
        for (uint32_t pi = 0; pi < n_pieces; pi++)
        {
            so_poly_load(&w, verts4, ring_off, ring, vert_off[pi], vert_off[pi + 1]);
            if (so_clip(&w, pl, npl) != 0) { so_poly_free(&w); return -2; }
            if (w.nv == 0) continue;
            if (nf >= out->cap_frags || nvtx + (uint64_t)w.nv > out->cap_verts) { so_poly_free(&w); return -1; }
            for (int i = 0; i < w.nv; i++)
            {
                float* d = out->verts4 + 4 * (nvtx + i);
                d[0] = w.pos[3 * i]; d[1] = w.pos[3 * i + 1]; d[2] = w.pos[3 * i + 2]; d[3] = 0.f;
                if (nring + (uint64_t)w.deg[i] > out->cap_ring) { so_poly_free(&w); return -1; }
                for (int j = 0; j < w.deg[i]; j++)
                    out->ring[nring++] = (uint16_t)w.ring[(size_t)i * SO_MAXDEG + j];
                out->ring_off[nvtx + i + 1] = (uint32_t)nring;
            }
            out->rec[4 * nf] = c;
            out->rec[4 * nf + 1] = pi;
            out->rec[4 * nf + 2] = (uint32_t)w.nv;
            out->rec[4 * nf + 3] = (uint32_t)so_extract_faces(&w, NULL, NULL);
            if (out->volume) so_moments(&w, out->volume + nf, out->centroid + 3 * nf);
            if (out->inertia) so_inertia(&w, out->inertia + 6 * nf);
            nvtx += (uint64_t)w.nv;
            nf++;
            out->vert_off[nf] = (uint32_t)nvtx;
        }
    }
    so_poly_free(&w);
    return (int64_t)nf;
}

/* Src/VMACH.cpp:289-310 + NearlyEqual (:1205): (v1 - v2).Length() < 1e-12 */
int so_face_planes(const so_poly* p, float* planes4)
{
    int maxf;
    const int ne = count_face_entries(p, &maxf);
    uint32_t* foff = (uint32_t*)malloc(sizeof(uint32_t) * (maxf + 2));
    uint16_t* fidx = (uint16_t*)malloc(sizeof(uint16_t) * (ne + 2));
    const int nf = so_extract_faces(p, foff, fidx);
    for (int f = 0; f < nf; f++)
    {
        float kept[3][3];
        int nk = 0;
        for (uint32_t k = foff[f]; k < foff[f + 1] && nk < 3; k++)
        {
            const float* v = p->pos + 3 * fidx[k];
            int dup = 0;
            for (int q = 0; q < nk; q++)
            {
                const float d[3] = { kept[q][0] - v[0], kept[q][1] - v[1], kept[q][2] - v[2] };
                if ((double)sqrtf(dot3(d, d)) < 1e-12) { dup = 1; break; }
            }
            if (dup) continue;
            kept[nk][0] = v[0]; kept[nk][1] = v[1]; kept[nk][2] = v[2];
            nk++;
        }
        float* o = planes4 + 4 * f;
        if (nk == 3) so_plane_from_points(kept[0], kept[1], kept[2], o);
        else { o[0] = 0.f; o[1] = 1.f; o[2] = 0.f; o[3] = 0.f; }   /* default-constructed Plane */
    }
    free(foff);
    free(fidx);
    return nf;
}

static int append_poly(so_out* out, const so_poly* w, uint64_t* nf, uint64_t* nvtx, uint64_t* nring, uint32_t a,
                       uint32_t b)
{
    if (*nf >= out->cap_frags || *nvtx + (uint64_t)w->nv > out->cap_verts) return -1;
    for (int i = 0; i < w->nv; i++)
    {
        float* d = out->verts4 + 4 * (*nvtx + i);
        d[0] = w->pos[3 * i]; d[1] = w->pos[3 * i + 1]; d[2] = w->pos[3 * i + 2]; d[3] = 0.f;
        if (*nring + (uint64_t)w->deg[i] > out->cap_ring) return -1;
        for (int j = 0; j < w->deg[i]; j++)
            out->ring[(*nring)++] = (uint16_t)w->ring[(size_t)i * SO_MAXDEG + j];
        out->ring_off[*nvtx + i + 1] = (uint32_t)*nring;
    }
    out->rec[4 * *nf] = a;
    out->rec[4 * *nf + 1] = b;
    out->rec[4 * *nf + 2] = (uint32_t)w->nv;
    out->rec[4 * *nf + 3] = w->nv ? (uint32_t)so_extract_faces(w, NULL, NULL) : 0u;
    if (out->volume) so_moments(w, out->volume + *nf, out->centroid + 3 * *nf);
    if (out->inertia) so_inertia(w, out->inertia + 6 * *nf);
    *nvtx += (uint64_t)w->nv;
    (*nf)++;
    out->vert_off[*nf] = (uint32_t)*nvtx;
    return 0;
}

/* Poly::GetBB (Src/Poly.cpp:587-617) */
static void load_unit_cube(so_poly* p)
{
    static const float pts[8][3] = { { -0.5f, -0.5f, -0.5f }, { 0.5f, -0.5f, -0.5f }, { 0.5f, 0.5f, -0.5f },
                                     { -0.5f, 0.5f, -0.5f }, { -0.5f, -0.5f, 0.5f }, { 0.5f, -0.5f, 0.5f },
                                     { 0.5f, 0.5f, 0.5f },   { -0.5f, 0.5f, 0.5f } };
    static const int nb[8][3] = { { 1, 4, 3 }, { 5, 0, 2 }, { 3, 6, 1 }, { 7, 2, 0 },
                                  { 5, 7, 0 }, { 1, 6, 4 }, { 5, 2, 7 }, { 4, 6, 3 } };
    reserve(p, 8);
    p->nv = 8;
    for (int i = 0; i < 8; i++)
    {
        memcpy(p->pos + 3 * i, pts[i], sizeof(float) * 3);
        p->deg[i] = 3;
        for (int j = 0; j < 3; j++) RING(p, i)[j] = nb[i][j];
        p->comp[i] = 1;
        p->id[i] = -1;
    }
}

int64_t so_voronoi_cells(const float* seeds3, uint32_t n, const uint32_t* nb_off, const uint32_t* nb_idx, so_out* out,
                         uint32_t* plane_off, float* planes4, uint64_t cap_planes)
{
    so_poly w;
    so_poly_init(&w);
    uint64_t nf = 0, nvtx = 0, nring = 0, npl = 0;
    out->vert_off[0] = 0;
    out->ring_off[0] = 0;
    plane_off[0] = 0;
    float* bis = NULL;
    uint32_t bis_cap = 0;
    for (uint32_t i = 0; i < n; i++)
    {
        const uint32_t k0 = nb_off[i], k1 = nb_off[i + 1];
        if (k1 - k0 > bis_cap) { bis_cap = 2 * (k1 - k0) + 16; bis = (float*)realloc(bis, sizeof(float) * 4 * bis_cap); }
        const float* si = seeds3 + 3 * i;
        for (uint32_t k = k0; k < k1; k++)
        {
            const float* sj = seeds3 + 3 * nb_idx[k];
            const float mid[3] = { (si[0] + sj[0]) * 0.5f, (si[1] + sj[1]) * 0.5f, (si[2] + sj[2]) * 0.5f };
            const float nrm[3] = { sj[0] - si[0], sj[1] - si[1], sj[2] - si[2] };
            so_plane_from_point_normal(mid, nrm, bis + 4 * (k - k0));
        }
        load_unit_cube(&w);
        if (so_clip(&w, bis, (int)(k1 - k0)) != 0) { free(bis); so_poly_free(&w); return -2; }
        if (append_poly(out, &w, &nf, &nvtx, &nring, i, 0)) { free(bis); so_poly_free(&w); return -1; }
        int maxf;
        count_face_entries(&w, &maxf);
        if (npl + (uint64_t)maxf > cap_planes) { free(bis); so_poly_free(&w); return -1; }
        npl += (uint64_t)(w.nv ? so_face_planes(&w, planes4 + 4 * npl) : 0);
        plane_off[i + 1] = (uint32_t)npl;
    }
    free(bis);
    so_poly_free(&w);
    return (int64_t)nf;
}

int64_t so_clip_each(const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off, const uint16_t* ring,
                     uint32_t n, const float* planes4, const uint32_t* pl_off, so_out* out)
{
    so_poly w;
    so_poly_init(&w);
    uint64_t nf = 0, nvtx = 0, nring = 0;
    out->vert_off[0] = 0;
    out->ring_off[0] = 0;
    for (uint32_t i = 0; i < n; i++)
    {
        so_poly_load(&w, verts4, ring_off, ring, vert_off[i], vert_off[i + 1]);
        if (so_clip(&w, planes4 + 4 * (size_t)pl_off[i], (int)(pl_off[i + 1] - pl_off[i])) != 0) { so_poly_free(&w); return -2; }
        if (append_poly(out, &w, &nf, &nvtx, &nring, i, i)) { so_poly_free(&w); return -1; }
    }
    so_poly_free(&w);
    return (int64_t)nf;
}

/* Face planes of every polyhedron of a flat set (so_face_planes per polyhedron). */
int64_t so_face_planes_set(const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off,
                           const uint16_t* ring, uint32_t n, uint32_t* plane_off, float* planes4, uint64_t cap_planes)
{
    so_poly w;
    so_poly_init(&w);
    uint64_t npl = 0;
    plane_off[0] = 0;
    for (uint32_t i = 0; i < n; i++)
    {
        so_poly_load(&w, verts4, ring_off, ring, vert_off[i], vert_off[i + 1]);
        int maxf;
        count_face_entries(&w, &maxf);
        if (npl + (uint64_t)maxf > cap_planes) { so_poly_free(&w); return -1; }
        npl += (uint64_t)(w.nv ? so_face_planes(&w, planes4 + 4 * npl) : 0);
        plane_off[i + 1] = (uint32_t)npl;
    }
    so_poly_free(&w);
    return (int64_t)npl;
}
