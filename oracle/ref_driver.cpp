// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Headless driver around the REFERENCE's own geometry code.  It is compiled together with
// /root/reference/Src/{Poly,Kdop,VMACH}.cpp and Inc/DT3D.h (from where they lie, see
// oracle/Makefile) into oracle/_ref/libsurtr_ref.so and exposes a flat C interface for
// the tests, the golden-fixture generator and bench.py's CPU-baseline leg.
//
// Src/Surtr.cpp cannot be compiled here (DX12 + PhysX + voro++ + assimp + Win32), so the
// few orchestration routines of the hot path are restated below, each citing the lines it
// follows.  Everything geometric is done by calling the reference functions themselves.
//
// Flat polyhedron-set layout (shared with include/surtr_b200.h):
//   verts    float[4*NV]   xyz + 0 pad, all polyhedra back to back
//   vert_off u32[n+1]      first vertex of polyhedron i
//   ring_off u32[NV+1]     first ring entry of (global) vertex v
//   ring     u16[NE]       neighbour rings, indices LOCAL to the polyhedron
#include "pch.h"

#include "Poly.h"
#include "Kdop.h"
#include "VMACH.h"
#include "DT3D.h" // non-inline definitions: include from exactly this TU

#include "thread_pool.h"

#include <chrono>
#include <future>
#include <set>
#include <unordered_map>

using DirectX::SimpleMath::Plane;
using DirectX::SimpleMath::Vector3;

namespace
{
struct PolySet
{
	std::vector<float> verts;
	std::vector<uint32_t> vert_off{ 0 };
	std::vector<uint32_t> ring_off{ 0 };
	std::vector<uint16_t> ring;
	// per polyhedron
	std::vector<uint32_t> cell, piece, nfaces;
	std::vector<double> volume;
	std::vector<float> centroid; // 3 per polyhedron
	// face loops (ExtractFaces order) and face planes (PolygonFace::AddVertex route)
	std::vector<uint32_t> face_off{ 0 };     // per face: first loop entry
	std::vector<uint16_t> face_idx;          // local vertex ids
	std::vector<uint32_t> poly_face_off{ 0 }; // per polyhedron: first face
	std::vector<float> planes;               // 4 per face
	double seconds = 0.0;

	size_t count() const { return vert_off.size() - 1; }
};

Poly::Polyhedron to_poly(const float* verts, const uint32_t* vert_off, const uint32_t* ring_off, const uint16_t* ring,
						 uint32_t i)
{
	Poly::Polyhedron p;
	const uint32_t v0 = vert_off[i], v1 = vert_off[i + 1];
	std::vector<Vector3> pos;
	std::vector<std::vector<int>> nei;
	pos.reserve(v1 - v0);
	nei.reserve(v1 - v0);
	for (uint32_t v = v0; v < v1; v++)
	{
		pos.emplace_back(verts[4 * v], verts[4 * v + 1], verts[4 * v + 2]);
		nei.emplace_back(ring + ring_off[v], ring + ring_off[v + 1]);
	}
	Poly::InitPolyhedron(p, pos, nei);
	return p;
}

// Face plane exactly as a VMACH cell face gets it: PolygonFace(true) + AddVertex per loop vertex
// (VMACH.cpp:289-310; duplicates closer than 1e-12 dropped; plane from the first three kept vertices).
bool face_plane(const Poly::Polyhedron& p, const std::vector<int>& loop, Plane& out)
{
	VMACH::PolygonFace f(true);
	for (int v : loop)
		f.AddVertex(p[v].Position);
	if (!f.FacePlaneConstructed)
		return false;
	out = f.FacePlane;
	return true;
}

void append(PolySet& s, const Poly::Polyhedron& p, uint32_t cell, uint32_t piece, bool with_moments = true,
			Poly::Extract* precomputed = nullptr, bool with_planes = true)
{
	for (const auto& v : p)
	{
		s.verts.push_back(v.Position.x);
		s.verts.push_back(v.Position.y);
		s.verts.push_back(v.Position.z);
		s.verts.push_back(0.f);
		for (int n : v.NeighborVertexVec)
			s.ring.push_back((uint16_t)n);
		s.ring_off.push_back((uint32_t)s.ring.size());
	}
	s.vert_off.push_back((uint32_t)(s.verts.size() / 4));
	s.cell.push_back(cell);
	s.piece.push_back(piece);

	Poly::Extract* faces = precomputed ? precomputed : Poly::ExtractFaces(p); // Poly.cpp:89-126
	s.nfaces.push_back((uint32_t)faces->size());
	for (const auto& loop : *faces)
	{
		for (int v : loop)
			s.face_idx.push_back((uint16_t)v);
		s.face_off.push_back((uint32_t)s.face_idx.size());
		Plane pl(0, 0, 0, 0);
		if (with_planes)
			face_plane(p, loop, pl);
		s.planes.push_back(pl.x);
		s.planes.push_back(pl.y);
		s.planes.push_back(pl.z);
		s.planes.push_back(pl.w);
	}
	s.poly_face_off.push_back((uint32_t)(s.face_off.size() - 1));
	delete faces;

	double vol = 0.0;
	Vector3 c(0, 0, 0);
	if (with_moments)
		Poly::Moments(vol, c, p); // Poly.cpp:55-87
	s.volume.push_back(vol);
	s.centroid.push_back(c.x);
	s.centroid.push_back(c.y);
	s.centroid.push_back(c.z);
}

VMACH::Polygon3D planes_to_polygon(const float* planes, uint32_t p0, uint32_t p1)
{
	// Poly::ClipPolyhedron(const Polyhedron&, const Polygon3D&) only reads FaceVec[i].FacePlane (Poly.cpp:558-560).
	VMACH::Polygon3D poly(true);
	for (uint32_t k = p0; k < p1; k++)
	{
		VMACH::PolygonFace f(true);
		f.ManuallySetFacePlane(Plane(planes[4 * k], planes[4 * k + 1], planes[4 * k + 2], planes[4 * k + 3]));
		poly.AddFace(f);
	}
	return poly;
}

// Restatement of m_fractureTask, convex branch (Surtr.cpp:1457-1468, 1497-1503): one cell against every piece.
std::vector<std::pair<uint32_t, Poly::Polyhedron>> fracture_task(const VMACH::Polygon3D& voroPoly,
																 const std::vector<Poly::Polyhedron>& pieces)
{
	std::vector<std::pair<uint32_t, Poly::Polyhedron>> local;
	for (uint32_t c = 0; c < pieces.size(); c++)
	{
		Poly::Polyhedron convex = Poly::ClipPolyhedron(pieces[c], voroPoly);
		if (convex.empty())
			continue;
		local.emplace_back(c, std::move(convex));
	}
	return local;
}

std::vector<std::vector<uint32_t>> dt_neighbors(const std::vector<Vector3>& seeds)
{
	const DT3D::Delaunay dt = DT3D::Triangulate(seeds); // DT3D.h:159-267
	// Tets carry point VALUES; map back by exact equality.
	auto key = [](const Vector3& v) {
		uint32_t b[3];
		std::memcpy(b, &v.x, 12);
		return std::make_tuple(b[0], b[1], b[2]);
	};
	std::map<std::tuple<uint32_t, uint32_t, uint32_t>, uint32_t> index;
	for (uint32_t i = 0; i < seeds.size(); i++)
		index[key(seeds[i])] = i;
	std::vector<std::set<uint32_t>> nb(seeds.size());
	for (const auto& tet : dt.TetVec)
	{
		const Vector3* p[4] = { &tet.p0, &tet.p1, &tet.p2, &tet.p3 };
		uint32_t id[4];
		bool ok = true;
		for (int k = 0; k < 4; k++)
		{
			auto it = index.find(key(*p[k]));
			if (it == index.end()) { ok = false; break; }
			id[k] = it->second;
		}
		if (!ok)
			continue;
		for (int a = 0; a < 4; a++)
			for (int b = 0; b < 4; b++)
				if (a != b)
					nb[id[a]].insert(id[b]);
	}
	std::vector<std::vector<uint32_t>> out(seeds.size());
	for (size_t i = 0; i < seeds.size(); i++)
		out[i].assign(nb[i].begin(), nb[i].end()); // ascending seed index
	return out;
}

// NEW derivation (replaces the voro++ call sites Surtr.cpp:2007-2067, dependency absent):
// cell i = container box clipped by the bisector half-spaces towards its neighbours j (ascending j),
// plane = Plane((Si+Sj)*0.5, Sj-Si), outward normal, unnormalised.
Plane bisector(const Vector3& si, const Vector3& sj)
{
	const Vector3 mid = (si + sj) * 0.5f;
	return Plane(mid, sj - si);
}
} // namespace

extern "C"
{
void* ref_polyset_new() { return new PolySet(); }
void ref_polyset_free(void* h) { delete (PolySet*)h; }

// sizes: [n_poly, n_verts, n_ring, n_faces, n_face_idx]
void ref_polyset_sizes(void* h, uint64_t* sizes)
{
	PolySet* s = (PolySet*)h;
	sizes[0] = s->count();
	sizes[1] = s->verts.size() / 4;
	sizes[2] = s->ring.size();
	sizes[3] = s->face_off.size() - 1;
	sizes[4] = s->face_idx.size();
}

double ref_polyset_seconds(void* h) { return ((PolySet*)h)->seconds; }

#define COPY_OUT(dst, vec) \
	if (dst)               \
	std::memcpy(dst, (vec).data(), (vec).size() * sizeof((vec)[0]))

void ref_polyset_export(void* h, float* verts, uint32_t* vert_off, uint32_t* ring_off, uint16_t* ring, uint32_t* cell,
						uint32_t* piece, uint32_t* nfaces, double* volume, float* centroid, uint32_t* poly_face_off,
						uint32_t* face_off, uint16_t* face_idx, float* planes)
{
	PolySet* s = (PolySet*)h;
	COPY_OUT(verts, s->verts);
	COPY_OUT(vert_off, s->vert_off);
	COPY_OUT(ring_off, s->ring_off);
	COPY_OUT(ring, s->ring);
	COPY_OUT(cell, s->cell);
	COPY_OUT(piece, s->piece);
	COPY_OUT(nfaces, s->nfaces);
	COPY_OUT(volume, s->volume);
	COPY_OUT(centroid, s->centroid);
	COPY_OUT(poly_face_off, s->poly_face_off);
	COPY_OUT(face_off, s->face_off);
	COPY_OUT(face_idx, s->face_idx);
	COPY_OUT(planes, s->planes);
}

// Seeds: Surtr.cpp:1988-1998 (mt19937 + uniform_real_distribution<double>(-0.5,0.5), narrowed at emplace_back).
void ref_seeds_uniform(uint32_t seed, uint32_t n, float* out)
{
	std::mt19937 gen(seed);
	std::uniform_real_distribution<double> uniformDist(-0.5, 0.5);
	for (uint32_t i = 0; i < n; i++)
	{
		const double x = uniformDist(gen);
		const double y = uniformDist(gen);
		const double z = uniformDist(gen);
		const Vector3 v(x, y, z);
		out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
	}
}

// Radial pattern seeds: Surtr.cpp:2072-2096.
void ref_seeds_radial(uint32_t seed, uint32_t n, double mean, float* out)
{
	std::mt19937 gen(seed);
	std::uniform_real_distribution<double> directionUniformDist(-1.0, 1.0);
	std::exponential_distribution<double> lengthExpDist(1.0 / mean);
	for (uint32_t i = 0; i < n; i++)
	{
		double len = std::max(std::min(lengthExpDist(gen), 0.5), 1e-12);
		double x = directionUniformDist(gen);
		double y = directionUniformDist(gen);
		double z = directionUniformDist(gen);
		Vector3 v = Vector3(x, y, z);
		v.Normalize();
		v *= len;
		out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
	}
}

// The unit cube of Poly::GetBB() (Poly.cpp:587-617) as a one-element polyset.
void ref_unit_cube(void* out)
{
	append(*(PolySet*)out, Poly::GetBB(), 0, 0);
}

// DT3D neighbour lists (CSR).  Returns total entries; call with idx == nullptr to size.
uint64_t ref_dt3d_neighbors(const float* seeds, uint32_t n, uint32_t* off, uint32_t* idx, double* seconds)
{
	std::vector<Vector3> s;
	for (uint32_t i = 0; i < n; i++)
		s.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
	static thread_local std::vector<std::vector<uint32_t>> cache;
	static thread_local std::vector<float> cache_key;
	std::vector<float> k(seeds, seeds + 3 * n);
	if (k != cache_key)
	{
		const auto t0 = std::chrono::steady_clock::now();
		cache = dt_neighbors(s);
		cache_key = k;
		if (seconds)
			*seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	}
	uint64_t total = 0;
	for (uint32_t i = 0; i < n; i++)
	{
		if (off) off[i] = (uint32_t)total;
		if (idx) std::copy(cache[i].begin(), cache[i].end(), idx + total);
		total += cache[i].size();
	}
	if (off) off[n] = (uint32_t)total;
	return total;
}

// Voronoi cells of `seeds` inside the unit container box, clipped with the REFERENCE clipper.
// nb_off/nb_idx: neighbour CSR (e.g. from ref_dt3d_neighbors); nullptr = brute force (all j != i).
void ref_voronoi_cells(const float* seeds, uint32_t n, const uint32_t* nb_off, const uint32_t* nb_idx, void* out)
{
	PolySet& o = *(PolySet*)out;
	std::vector<Vector3> s;
	for (uint32_t i = 0; i < n; i++)
		s.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
	for (uint32_t i = 0; i < n; i++)
	{
		std::vector<Plane> planes;
		if (nb_off)
			for (uint32_t k = nb_off[i]; k < nb_off[i + 1]; k++)
				planes.push_back(bisector(s[i], s[nb_idx[k]]));
		else
			for (uint32_t j = 0; j < n; j++)
				if (j != i)
					planes.push_back(bisector(s[i], s[j]));
		Poly::Polyhedron cell = Poly::GetBB();
		Poly::ClipPolyhedron(cell, planes); // Poly.cpp:265
		append(o, cell, i, 0);
	}
}

// Clip every polyhedron of a set by its own plane list (pl_off per polyhedron) -- Poly.cpp:265 in place.
void ref_clip_each(const float* verts, const uint32_t* vert_off, const uint32_t* ring_off, const uint16_t* ring,
				   uint32_t n, const float* planes, const uint32_t* pl_off, void* out)
{
	PolySet& o = *(PolySet*)out;
	for (uint32_t i = 0; i < n; i++)
	{
		Poly::Polyhedron p = to_poly(verts, vert_off, ring_off, ring, i);
		std::vector<Plane> pls;
		for (uint32_t k = pl_off[i]; k < pl_off[i + 1]; k++)
			pls.emplace_back(planes[4 * k], planes[4 * k + 1], planes[4 * k + 2], planes[4 * k + 3]);
		Poly::ClipPolyhedron(p, pls);
		append(o, p, i, i);
	}
}

// Restatement of Surtr::ApplyFracture (Surtr.cpp:2098-2149), non-partial, convex branch:
// one task per cell, results consumed in cell order, piece-minor inside a cell.
// nthreads == 0: run the tasks inline on the calling thread; otherwise dp::thread_pool(nthreads)
// (the reference uses 16, Surtr.cpp:28).  `seconds` covers enqueue -> last future, like the
// reference's "ApplyFracture" timer (Surtr.cpp:1917-1924), not the result flattening.
void ref_apply_fracture(const float* verts, const uint32_t* vert_off, const uint32_t* ring_off, const uint16_t* ring,
						uint32_t n_pieces, const float* planes, const uint32_t* plane_off, uint32_t n_cells,
						uint32_t nthreads, int with_moments, void* out)
{
	// `seconds` = fan-out + SetExtract; flattening into the flat arrays (and the optional moments) is not timed.
	PolySet& o = *(PolySet*)out;
	std::vector<Poly::Polyhedron> pieces;
	for (uint32_t i = 0; i < n_pieces; i++)
		pieces.push_back(to_poly(verts, vert_off, ring_off, ring, i));
	std::vector<VMACH::Polygon3D> cells;
	for (uint32_t c = 0; c < n_cells; c++)
		cells.push_back(planes_to_polygon(planes, plane_off[c], plane_off[c + 1]));

	std::vector<std::vector<std::pair<uint32_t, Poly::Polyhedron>>> results(n_cells);
	const auto t0 = std::chrono::steady_clock::now();
	if (nthreads == 0)
	{
		for (uint32_t c = 0; c < n_cells; c++)
			results[c] = fracture_task(cells[c], pieces);
	}
	else
	{
		dp::thread_pool pool(nthreads);
		std::vector<std::future<std::vector<std::pair<uint32_t, Poly::Polyhedron>>>> futures;
		for (uint32_t c = 0; c < n_cells; c++)
			futures.push_back(pool.enqueue(fracture_task, std::cref(cells[c]), std::cref(pieces)));
		for (uint32_t c = 0; c < n_cells; c++)
			results[c] = futures[c].get();
	}
	// Surtr::SetExtract (Surtr.cpp:2151-2155): Poly::ExtractFaces per resulting piece, as DoFracture does right
	// after ApplyFracture (:1921-1922).  Timed together with the fan-out: that pair is the CPU event.
	std::vector<std::vector<Poly::Extract*>> extracts(n_cells);
	for (uint32_t c = 0; c < n_cells; c++)
		for (auto& [piece, poly] : results[c])
			extracts[c].push_back(Poly::ExtractFaces(poly));
	o.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

	for (uint32_t c = 0; c < n_cells; c++)
		for (size_t k = 0; k < results[c].size(); k++)
			append(o, results[c][k].second, c, results[c][k].first, with_moments != 0, extracts[c][k], false);
}

// Restatement of Surtr::_MeshIslandLoop / CheckMeshIsland (Surtr.cpp:2157-2199): connected components of the ring
// graph, each a sorted set, discovered from the lowest vertex not yet in a group.  (The reference recurses; an explicit
// stack visits the same sets.)
static std::vector<std::set<int>> check_mesh_island(const Poly::Polyhedron& polyhedron)
{
	std::vector<std::set<int>> groupVec;
	std::vector<char> seen(polyhedron.size(), 0);
	int start = 0;
	while (true)
	{
		std::set<int> group;
		std::vector<int> stack{ start };
		while (!stack.empty())
		{
			const int v = stack.back();
			stack.pop_back();
			for (const int a : polyhedron[v].NeighborVertexVec)
				if (group.insert(a).second)
					stack.push_back(a);
		}
		groupVec.push_back(group);
		for (const int v : group)
			seen[v] = 1;
		bool remain = false;
		for (int v = 0; v < (int)polyhedron.size(); v++)
			if (!seen[v]) { remain = true; start = v; break; }
		if (!remain)
			break;
	}
	return groupVec;
}

// Restatement of the full m_fractureTask (Surtr.cpp:1457-1504), convex AND mesh branch, cells in order (inline, one
// thread): out_convex / out_mesh receive the Piece::Convex / Piece::Mesh of every resulting piece, in PieceVec order.
void ref_apply_fracture_mesh(const float* cverts, const uint32_t* cvert_off, const uint32_t* cring_off, const uint16_t* cring,
							 const float* mverts, const uint32_t* mvert_off, const uint32_t* mring_off, const uint16_t* mring,
							 uint32_t n_pieces, const float* planes, const uint32_t* plane_off, uint32_t n_cells,
							 void* out_convex, void* out_mesh)
{
	PolySet& oc = *(PolySet*)out_convex;
	PolySet& om = *(PolySet*)out_mesh;
	const auto t0 = std::chrono::steady_clock::now();
	for (uint32_t c = 0; c < n_cells; c++)
	{
		const VMACH::Polygon3D voroPoly = planes_to_polygon(planes, plane_off[c], plane_off[c + 1]);
		for (uint32_t i = 0; i < n_pieces; i++)
		{
			const Poly::Polyhedron convex = Poly::ClipPolyhedron(to_poly(cverts, cvert_off, cring_off, cring, i), voroPoly);
			if (convex.empty())
				continue;
			const Poly::Polyhedron mesh = Poly::ClipPolyhedron(to_poly(mverts, mvert_off, mring_off, mring, i), voroPoly);
			if (mesh.empty())
				continue;
			const auto groupVec = check_mesh_island(mesh);
			if (groupVec.size() >= 2)
			{
				for (const auto& group : groupVec)
				{
					Poly::Polyhedron island;
					std::unordered_map<int, int> mapping;
					for (const int iVert : group)
					{
						mapping[iVert] = (int)island.size();
						island.push_back(mesh[iVert]);
					}
					for (auto& vert : island)
						for (int& iAdj : vert.NeighborVertexVec)
							iAdj = mapping[iAdj];
					append(oc, convex, c, i);
					append(om, island, c, i);
				}
			}
			else
			{
				append(oc, convex, c, i);
				append(om, mesh, c, i);
			}
		}
	}
	oc.seconds = om.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Kdop::KdopContainer::Calc(const Poly::Polyhedron&) (Kdop.cpp:92-115) on raw vertices:
// out_dist[2k] = {MinDist, MaxDist}, out_planes[8k] = {MinPlane, MaxPlane}, out_vtx[6k] = {MinVertex, MaxVertex}.
void ref_kdop_calc_poly(const float* verts, uint32_t nv, const float* normals, uint32_t k, double* out_dist,
						float* out_planes, float* out_vtx)
{
	std::vector<Vector3> nrm;
	for (uint32_t i = 0; i < k; i++)
		nrm.emplace_back(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
	Poly::Polyhedron p(nv);
	for (uint32_t v = 0; v < nv; v++)
		p[v].Position = Vector3(verts[4 * v], verts[4 * v + 1], verts[4 * v + 2]);
	Kdop::KdopContainer kd(nrm);
	kd.Calc(p);
	for (uint32_t i = 0; i < k; i++)
	{
		const auto& e = kd.ElementVec[i];
		out_dist[2 * i] = e.MinDist; out_dist[2 * i + 1] = e.MaxDist;
		const float pl[8] = { e.MinPlane.x, e.MinPlane.y, e.MinPlane.z, e.MinPlane.w,
							  e.MaxPlane.x, e.MaxPlane.y, e.MaxPlane.z, e.MaxPlane.w };
		std::memcpy(out_planes + 8 * i, pl, sizeof(pl));
		const float vx[6] = { e.MinVertex.x, e.MinVertex.y, e.MinVertex.z, e.MaxVertex.x, e.MaxVertex.y, e.MaxVertex.z };
		std::memcpy(out_vtx + 6 * i, vx, sizeof(vx));
	}
}

// Kdop::KdopContainer::Calc(vertices, maxAxisScale, planeGapInv) (Kdop.cpp:15-51), same outputs.
void ref_kdop_calc_gap(const float* verts, uint32_t nv, const float* normals, uint32_t k, double maxAxisScale,
					   float planeGapInv, double* out_dist, float* out_planes, float* out_vtx)
{
	std::vector<Vector3> nrm, vv;
	for (uint32_t i = 0; i < k; i++)
		nrm.emplace_back(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
	for (uint32_t v = 0; v < nv; v++)
		vv.emplace_back(verts[4 * v], verts[4 * v + 1], verts[4 * v + 2]);
	Kdop::KdopContainer kd(nrm);
	kd.Calc(vv, maxAxisScale, planeGapInv);
	for (uint32_t i = 0; i < k; i++)
	{
		const auto& e = kd.ElementVec[i];
		out_dist[2 * i] = e.MinDist; out_dist[2 * i + 1] = e.MaxDist;
		const float pl[8] = { e.MinPlane.x, e.MinPlane.y, e.MinPlane.z, e.MinPlane.w,
							  e.MaxPlane.x, e.MaxPlane.y, e.MaxPlane.z, e.MaxPlane.w };
		std::memcpy(out_planes + 8 * i, pl, sizeof(pl));
		const float vx[6] = { e.MinVertex.x, e.MinVertex.y, e.MinVertex.z, e.MaxVertex.x, e.MaxVertex.y, e.MaxVertex.z };
		std::memcpy(out_vtx + 6 * i, vx, sizeof(vx));
	}
}

// Surtr::GenerateICHNormal (Surtr.cpp:1961-1974) over VMACH::ConvexHull (VMACH.cpp:869-1161).
// Returns the number of normals written (<= cap).
uint32_t ref_ich_normals(const float* verts, uint32_t nv, int limit, float* out, uint32_t cap)
{
	std::vector<Vector3> vv;
	for (uint32_t v = 0; v < nv; v++)
		vv.emplace_back(verts[4 * v], verts[4 * v + 1], verts[4 * v + 2]);
	VMACH::ConvexHull ich(vv, (uint32_t)limit);
	uint32_t n = 0;
	for (const VMACH::ConvexHullFace& f : ich.GetFaces())
	{
		Vector3 normal = (f.Vertices[1] - f.Vertices[0]).Cross(f.Vertices[2] - f.Vertices[0]);
		normal.Normalize();
		if (n < cap)
		{
			out[3 * n] = normal.x; out[3 * n + 1] = normal.y; out[3 * n + 2] = normal.z;
		}
		n++;
	}
	return n;
}

// Restatement of m_refittingTask (Surtr.cpp:1449-1455) for every piece i: ICH normals of the mesh points (limit =
// min(#points, RefittingPointLimit)) -> KdopContainer::Calc(mesh) -> convex = ClipWithPolyhedron(convex).
// Every piece gets one output entry (possibly empty).
void ref_refit(const float* cverts, const uint32_t* cvert_off, const uint32_t* cring_off, const uint16_t* cring, uint32_t n,
			   const float* mverts, const uint32_t* mvert_off, int limit, void* out)
{
	PolySet& o = *(PolySet*)out;
	for (uint32_t i = 0; i < n; i++)
	{
		Poly::Polyhedron convex = to_poly(cverts, cvert_off, cring_off, cring, i);
		std::vector<Vector3> pts;
		Poly::Polyhedron mesh;
		for (uint32_t v = mvert_off[i]; v < mvert_off[i + 1]; v++)
		{
			pts.emplace_back(mverts[4 * v], mverts[4 * v + 1], mverts[4 * v + 2]);
			mesh.push_back(Poly::Vertex(pts.back()));
		}
		// Surtr::GenerateICHNormal (Surtr.cpp:1961-1974)
		VMACH::ConvexHull ich(pts, (uint32_t)std::min((int)pts.size(), limit));
		std::vector<Vector3> normals;
		for (const VMACH::ConvexHullFace& f : ich.GetFaces())
		{
			Vector3 normal = (f.Vertices[1] - f.Vertices[0]).Cross(f.Vertices[2] - f.Vertices[0]);
			normal.Normalize();
			normals.push_back(normal);
		}
		Kdop::KdopContainer kdop(normals);
		kdop.Calc(mesh);
		convex = kdop.ClipWithPolyhedron(convex);
		append(o, convex, i, i);
	}
}

// Config 1, convex branch: restatement of Surtr::PrepareFracture steps 1-6, 8, 10 (Surtr.cpp:1747-1811) on a vertex
// cloud: ICH normals (limit) -> bbox -> k-DOP with gap -> ACH = 2x bbox clipped -> Voronoi cells of `seeds`
// (unit box) scaled by the bbox extent and translated to its centre (Polygon3D::Scale/Translate re-derive every
// plane, VMACH.cpp:506-534) -> ApplyFracture on the single ACH piece.  out_ach gets the ACH, out the fragments.
static void config1(const float* verts4, uint32_t nv, const int32_t* indices, uint32_t n_idx, int ich_limit, float gap_inv, int refit_limit,
					const float* seeds, uint32_t n_seeds, const uint32_t* nb_off, const uint32_t* nb_idx, void* out_ach, void* out, void* out_mesh)
{
	std::vector<Vector3> vertices;
	for (uint32_t v = 0; v < nv; v++)
		vertices.emplace_back(verts4[4 * v], verts4[4 * v + 1], verts4[4 * v + 2]);
	VMACH::ConvexHull ich(vertices, (uint32_t)ich_limit);
	std::vector<Vector3> normals;
	for (const VMACH::ConvexHullFace& f : ich.GetFaces())
	{
		Vector3 normal = (f.Vertices[1] - f.Vertices[0]).Cross(f.Vertices[2] - f.Vertices[0]);
		normal.Normalize();
		normals.push_back(normal);
	}
	double minX, maxX, minY, maxY, minZ, maxZ;
	{
		const auto x = std::minmax_element(vertices.begin(), vertices.end(), [](const Vector3& p1, const Vector3& p2) { return p1.x < p2.x; });
		const auto y = std::minmax_element(vertices.begin(), vertices.end(), [](const Vector3& p1, const Vector3& p2) { return p1.y < p2.y; });
		const auto z = std::minmax_element(vertices.begin(), vertices.end(), [](const Vector3& p1, const Vector3& p2) { return p1.z < p2.z; });
		minX = (*x.first).x; maxX = (*x.second).x;
		minY = (*y.first).y; maxY = (*y.second).y;
		minZ = (*z.first).z; maxZ = (*z.second).z;
	}
	const Vector3 BBCenter((maxX + minX) / 2.0, (maxY + minY) / 2.0, (maxZ + minZ) / 2.0);
	const float MaxAxisScale = std::max(std::max(maxX - minX, maxY - minY), maxZ - minZ);   // stored as float (Surtr.h:154)
	Kdop::KdopContainer achKdop(normals);
	achKdop.Calc(vertices, MaxAxisScale, gap_inv);
	Poly::Polyhedron ach = Poly::GetBB();
	Poly::Scale(ach, Vector3((maxX - minX), (maxY - minY), (maxZ - minZ)));
	Poly::Scale(ach, Vector3(2.0, 2.0, 2.0));
	Poly::Translate(ach, BBCenter);
	ach = achKdop.ClipWithPolyhedron(ach);
	append(*(PolySet*)out_ach, ach, 0, 0);

	// cells: the new derivation (unit box), as VMACH::Polygon3D built through PolygonFace::AddVertex
	std::vector<Vector3> s;
	for (uint32_t i = 0; i < n_seeds; i++)
		s.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
	std::vector<VMACH::Polygon3D> voroPolyVec;
	for (uint32_t i = 0; i < n_seeds; i++)
	{
		std::vector<Plane> planes;
		for (uint32_t k = nb_off[i]; k < nb_off[i + 1]; k++)
			planes.push_back(bisector(s[i], s[nb_idx[k]]));
		Poly::Polyhedron cell = Poly::GetBB();
		Poly::ClipPolyhedron(cell, planes);
		Poly::Extract* faces = Poly::ExtractFaces(cell);
		VMACH::Polygon3D poly(true);
		for (const auto& loop : *faces)
		{
			VMACH::PolygonFace f(true);
			for (int v : loop)
				f.AddVertex(cell[v].Position);
			poly.AddFace(f);
		}
		delete faces;
		voroPolyVec.push_back(poly);
	}
	for (VMACH::Polygon3D& voro : voroPolyVec)
	{
		voro.Scale(Vector3((maxX - minX), (maxY - minY), (maxZ - minZ)));
		voro.Translate(BBCenter);
	}
	PolySet& o = *(PolySet*)out;
	if (!indices)
	{
		// ApplyFracture, convex branch, single piece
		for (uint32_t c = 0; c < n_seeds; c++)
		{
			Poly::Polyhedron convex = Poly::ClipPolyhedron(ach, voroPolyVec[c]);
			if (convex.empty())
				continue;
			append(o, convex, c, 0);
		}
		return;
	}
	// step 7 (Surtr.cpp:1788-1795) + step 10: full m_fractureTask on the (ACH, mesh) piece, then Refitting (:1813)
	Poly::Polyhedron meshPolyhedron;
	{
		std::vector<int> idx(indices, indices + n_idx);
		const std::vector<std::vector<int>> nei = Poly::ExtractNeighborFromMesh(vertices, idx);
		Poly::InitPolyhedron(meshPolyhedron, vertices, nei);
	}
	PolySet pre_convex, pre_mesh;
	{
		PolySet a, m;
		append(a, ach, 0, 0);
		append(m, meshPolyhedron, 0, 0);
		std::vector<float> planes;
		std::vector<uint32_t> plane_off{ 0 };
		for (const VMACH::Polygon3D& voro : voroPolyVec)
		{
			for (const VMACH::PolygonFace& f : voro.FaceVec)
				planes.insert(planes.end(), { f.FacePlane.x, f.FacePlane.y, f.FacePlane.z, f.FacePlane.w });
			plane_off.push_back((uint32_t)(planes.size() / 4));
		}
		ref_apply_fracture_mesh(a.verts.data(), a.vert_off.data(), a.ring_off.data(), a.ring.data(),
								m.verts.data(), m.vert_off.data(), m.ring_off.data(), m.ring.data(), 1,
								planes.data(), plane_off.data(), n_seeds, &pre_convex, &pre_mesh);
	}
	PolySet& om = *(PolySet*)out_mesh;
	const uint32_t n = (uint32_t)pre_convex.vert_off.size() - 1;
	for (uint32_t i = 0; i < n; i++)
	{
		Poly::Polyhedron convex = to_poly(pre_convex.verts.data(), pre_convex.vert_off.data(), pre_convex.ring_off.data(), pre_convex.ring.data(), i);
		const Poly::Polyhedron mesh = to_poly(pre_mesh.verts.data(), pre_mesh.vert_off.data(), pre_mesh.ring_off.data(), pre_mesh.ring.data(), i);
		// m_refittingTask (Surtr.cpp:1449-1455)
		std::vector<Vector3> pts;
		for (const Poly::Vertex& v : mesh)
			pts.push_back(v.Position);
		VMACH::ConvexHull ich(pts, (uint32_t)std::min((int)pts.size(), refit_limit));
		std::vector<Vector3> nrm;
		for (const VMACH::ConvexHullFace& f : ich.GetFaces())
		{
			Vector3 normal = (f.Vertices[1] - f.Vertices[0]).Cross(f.Vertices[2] - f.Vertices[0]);
			normal.Normalize();
			nrm.push_back(normal);
		}
		Kdop::KdopContainer kdop(nrm);
		kdop.Calc(mesh);
		convex = kdop.ClipWithPolyhedron(convex);
		append(o, convex, pre_convex.cell[i], pre_convex.piece[i]);
		append(om, mesh, pre_mesh.cell[i], pre_mesh.piece[i]);
	}
}

void ref_config1_convex(const float* verts4, uint32_t nv, int ich_limit, float gap_inv, const float* seeds, uint32_t n_seeds,
						const uint32_t* nb_off, const uint32_t* nb_idx, void* out_ach, void* out)
{
	config1(verts4, nv, nullptr, 0, ich_limit, gap_inv, 0, seeds, n_seeds, nb_off, nb_idx, out_ach, out, nullptr);
}

// Surtr::PrepareFracture in full (Surtr.cpp:1747-1827): as above plus the mesh polyhedron, the mesh branch of the
// fracture task with its island split, and Refitting.  out = Piece::Convex after the refit, out_mesh = Piece::Mesh.
void ref_config1_full(const float* verts4, uint32_t nv, const int32_t* indices, uint32_t n_idx, int ich_limit, float gap_inv, int refit_limit,
					  const float* seeds, uint32_t n_seeds, const uint32_t* nb_off, const uint32_t* nb_idx, void* out_ach, void* out, void* out_mesh)
{
	config1(verts4, nv, indices, n_idx, ich_limit, gap_inv, refit_limit, seeds, n_seeds, nb_off, nb_idx, out_ach, out, out_mesh);
}

// ------------------------------------------------------------------------------------------------------------
// Restatement of Surtr::DoFracture (Surtr.cpp:1885-1959) and the members it calls, which live in the DX12 application
// class and cannot be compiled here: ApplyFracture with the partial mode (:2098-2149), the full m_fractureTask
// (:1457-1504), SetExtract (:2151-2155), MergeOutOfImpact (:2368-2403), ConvexOutOfSphere (:2415-2458),
// HandleConvexIsland (:2203-2366), m_refittingTask (:1449-1455).  Geometry is done by the reference's own compiled
// Poly / Kdop / VMACH code.
namespace
{
struct RPiece
{
	Poly::Polyhedron Convex, Mesh;
};

bool convex_out_of_sphere(const Poly::Polyhedron& polyhedron, const Poly::Extract* extract, const std::vector<Vector3>& cloud,
						  const Vector3 origin, const float radius)
{
	for (const auto& v : polyhedron)
		if ((origin - v.Position).Length() < radius)
			return false;
	for (const auto& po : cloud)
	{
		bool contain = true;
		for (const auto& f : *extract)
		{
			Vector3 normal = (polyhedron[f[1]].Position - polyhedron[f[0]].Position).Cross(polyhedron[f[2]].Position - polyhedron[f[0]].Position);
			normal.Normalize();
			const float d = -polyhedron[f[0]].Position.Dot(normal);
			const float dist = normal.Dot(po) + d;
			if (dist > 0) { contain = false; break; }
		}
		if (contain)
			return false;
	}
	return true;
}

bool face_point_inside(const std::vector<Vector3>& pts, const std::vector<Vector3>& outline, const Vector3& n)
{
	const int m = (int)outline.size();
	for (const Vector3& p : pts)
	{
		bool included = true;
		for (int v = 0; v < m; v++)
			if (!VMACH::OnYourRight(outline[v], outline[(v + 1) % m], p, n)) { included = false; break; }
		if (included)
			return true;
	}
	return false;
}

void handle_convex_island(std::vector<std::set<int>>& bind, const std::vector<RPiece*>& pieces, const std::vector<Poly::Extract*>& extracts)
{
	struct FaceNode { int CID; double AbsD; Plane FacePlane; std::vector<Vector3> FacePoints; };
	std::vector<std::set<int>> newBind;
	for (auto& localBind : bind)
	{
		if (localBind.size() <= 1)
			continue;
		std::vector<FaceNode> nodes;
		for (const int cid : localBind)
			for (const auto& poly : *extracts[cid])
			{
				std::vector<Vector3> points;
				for (const int v : poly)
					points.push_back(pieces[cid]->Convex[v].Position);
				Plane p(points[0], points[1], points[2]);
				nodes.push_back(FaceNode{ cid, std::abs(p.D()), p, points });
			}
		std::sort(nodes.begin(), nodes.end(), [](const FaceNode& a, const FaceNode& b) { return a.AbsD < b.AbsD; });
		std::unordered_map<int, std::set<int>> nei;
		for (int i = 0; i + 1 < (int)nodes.size(); i++)
		{
			bool lowerBoundFound = false;
			for (int j = i + 1; j < (int)nodes.size(); j++)   // quadratic, exactly as :2232-2318 (no window)
			{
				if (lowerBoundFound && nodes[i].AbsD > nodes[j].AbsD)
					break;
				if (std::abs(nodes[i].AbsD - nodes[j].AbsD) > 1e-3)
					continue;
				lowerBoundFound = true;
				Vector3 in = nodes[i].FacePlane.Normal(), jn = nodes[j].FacePlane.Normal();
				in.Normalize(); jn.Normalize();
				if (!(std::abs(1 + in.Dot(jn)) < 1e-4))
					continue;
				if (face_point_inside(nodes[i].FacePoints, nodes[j].FacePoints, jn) || face_point_inside(nodes[j].FacePoints, nodes[i].FacePoints, in))
				{
					nei[nodes[i].CID].insert(nodes[j].CID);
					nei[nodes[j].CID].insert(nodes[i].CID);
				}
			}
		}
		std::set<int> remain(localBind.begin(), localBind.end());
		std::vector<std::set<int>> splitGroup;
		while (!remain.empty())
		{
			std::set<int> split;
			std::vector<int> queue{ *remain.begin() };
			for (size_t q = 0; q < queue.size(); q++)
			{
				const int curr = queue[q];
				if (remain.count(curr))
				{
					split.insert(curr);
					remain.erase(curr);
					for (const int a : nei[curr])
						queue.push_back(a);
				}
			}
			splitGroup.push_back(split);
		}
		if (splitGroup.size() >= 2)
		{
			localBind = splitGroup[0];
			newBind.insert(newBind.end(), std::next(splitGroup.begin()), splitGroup.end());
		}
	}
	bind.insert(bind.end(), newBind.begin(), newBind.end());
}

std::vector<VMACH::Polygon3D> unit_box_cells(const float* seeds, uint32_t n_seeds, const uint32_t* nb_off, const uint32_t* nb_idx)
{
	// cells of the unit container from neighbour bisectors, as VMACH::Polygon3D through PolygonFace::AddVertex
	std::vector<Vector3> s;
	for (uint32_t i = 0; i < n_seeds; i++)
		s.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
	std::vector<VMACH::Polygon3D> out;
	for (uint32_t i = 0; i < n_seeds; i++)
	{
		std::vector<Plane> planes;
		for (uint32_t k = nb_off[i]; k < nb_off[i + 1]; k++)
			planes.push_back(bisector(s[i], s[nb_idx[k]]));
		Poly::Polyhedron cell = Poly::GetBB();
		Poly::ClipPolyhedron(cell, planes);
		Poly::Extract* faces = Poly::ExtractFaces(cell);
		VMACH::Polygon3D poly(true);
		for (const auto& loop : *faces)
		{
			VMACH::PolygonFace f(true);
			for (int v : loop)
				f.AddVertex(cell[v].Position);
			poly.AddFace(f);
		}
		delete faces;
		out.push_back(poly);
	}
	return out;
}
} // namespace

// pieces: the target compound (convex_i, mesh_i) in world space.  pattern: seeds + neighbour CSR of the pattern cells
// in the unit box.  cloud3: m_spherePointCloud (unit sphere samples, already scaled by 0.5 as at Surtr.cpp:1508).
// Outputs in second.PieceVec order: Piece::Convex after Refitting (cell field = index of the bind set / compound the
// piece ends up in, piece field = 1 when the piece is one of the caller's untouched pieces), Piece::Mesh likewise.
void ref_do_fracture(const float* cverts, const uint32_t* cvert_off, const uint32_t* cring_off, const uint16_t* cring,
					 const float* mverts, const uint32_t* mvert_off, const uint32_t* mring_off, const uint16_t* mring, uint32_t n_pieces,
					 const float* seeds, uint32_t n_seeds, const uint32_t* nb_off, const uint32_t* nb_idx,
					 const float* cloud3, uint32_t n_cloud, const float* impact3, float impact_radius, float max_axis_scale,
					 int partial, int refit_limit, void* out_convex, void* out_mesh, uint32_t* n_compounds)
{
	std::vector<RPiece*> target;
	std::vector<Poly::Extract*> targetExtract;
	for (uint32_t i = 0; i < n_pieces; i++)
	{
		target.push_back(new RPiece{ to_poly(cverts, cvert_off, cring_off, cring, i), to_poly(mverts, mvert_off, mring_off, mring, i) });
		targetExtract.push_back(Poly::ExtractFaces(target.back()->Convex));
	}
	const Vector3 impact(impact3[0], impact3[1], impact3[2]);
	// DoFracture: pattern placement (:1887-1896) and sphere samples (:1911-1916)
	std::vector<VMACH::Polygon3D> pattern = unit_box_cells(seeds, n_seeds, nb_off, nb_idx);
	for (VMACH::Polygon3D& voro : pattern)
		voro.Scale(Vector3(max_axis_scale, max_axis_scale, max_axis_scale) * 2);
	for (VMACH::Polygon3D& voro : pattern)
		voro.Translate(impact);
	std::vector<Vector3> cloud;
	for (uint32_t i = 0; i < n_cloud; i++)
	{
		Vector3 v(cloud3[3 * i], cloud3[3 * i + 1], cloud3[3 * i + 2]);
		v *= impact_radius;
		v += impact;
		cloud.push_back(v);
	}
	// ApplyFracture (:2098-2149)
	std::vector<RPiece*> decompose;
	std::vector<std::set<int>> bind;
	std::set<int> outside, outsideBind;
	if (partial)
		for (int c = 0; c < (int)target.size(); c++)
			if (convex_out_of_sphere(target[c]->Convex, targetExtract[c], cloud, impact, impact_radius))
			{
				outside.insert(c);
				outsideBind.insert((int)decompose.size());
				decompose.push_back(target[c]);
			}
	const size_t n_untouched = decompose.size();
	bind.push_back(outsideBind);
	for (const VMACH::Polygon3D& voroPoly : pattern)
	{
		std::set<int> localBind;
		for (int c = 0; c < (int)target.size(); c++)   // m_fractureTask (:1457-1504)
		{
			if (outside.count(c))
				continue;
			const Poly::Polyhedron convex = Poly::ClipPolyhedron(target[c]->Convex, voroPoly);
			if (convex.empty())
				continue;
			const Poly::Polyhedron mesh = Poly::ClipPolyhedron(target[c]->Mesh, voroPoly);
			if (mesh.empty())
				continue;
			const auto groupVec = check_mesh_island(mesh);
			if (groupVec.size() >= 2)
			{
				for (const auto& group : groupVec)
				{
					Poly::Polyhedron island;
					std::unordered_map<int, int> mapping;
					for (const int iVert : group)
					{
						mapping[iVert] = (int)island.size();
						island.push_back(mesh[iVert]);
					}
					for (auto& vert : island)
						for (int& iAdj : vert.NeighborVertexVec)
							iAdj = mapping[iAdj];
					localBind.insert((int)decompose.size());
					decompose.push_back(new RPiece{ convex, island });
				}
			}
			else
			{
				localBind.insert((int)decompose.size());
				decompose.push_back(new RPiece{ convex, mesh });
			}
		}
		if (!localBind.empty())
			bind.push_back(localBind);
	}
	// SetExtract
	std::vector<Poly::Extract*> extracts;
	for (const RPiece* p : decompose)
		extracts.push_back(Poly::ExtractFaces(p->Convex));
	// MergeOutOfImpact
	if (partial)
	{
		for (size_t i = 1; i < bind.size(); i++)
		{
			std::set<int> out;
			for (const int c : bind[i])
				if (convex_out_of_sphere(decompose[c]->Convex, extracts[c], cloud, impact, impact_radius))
					out.insert(c);
			for (const int c : out)
			{
				bind[i].erase(c);
				bind[0].insert(c);
			}
		}
		bind.erase(std::remove_if(std::next(bind.begin()), bind.end(), [](const std::set<int>& b) { return b.empty(); }), bind.end());
	}
	handle_convex_island(bind, decompose, extracts);
	// Refitting (m_refittingTask) on every piece
	for (RPiece* piece : decompose)
	{
		std::vector<Vector3> pts;
		for (const Poly::Vertex& v : piece->Mesh)
			pts.push_back(v.Position);
		VMACH::ConvexHull ich(pts, (uint32_t)std::min((int)pts.size(), refit_limit));
		std::vector<Vector3> nrm;
		for (const VMACH::ConvexHullFace& f : ich.GetFaces())
		{
			Vector3 normal = (f.Vertices[1] - f.Vertices[0]).Cross(f.Vertices[2] - f.Vertices[0]);
			normal.Normalize();
			nrm.push_back(normal);
		}
		Kdop::KdopContainer kdop(nrm);
		kdop.Calc(piece->Mesh);
		piece->Convex = kdop.ClipWithPolyhedron(piece->Convex);
	}
	std::vector<uint32_t> compound_of(decompose.size(), 0xffffffffu);
	for (size_t b = 0; b < bind.size(); b++)
		for (const int c : bind[b])
			compound_of[c] = (uint32_t)b;
	*n_compounds = (uint32_t)bind.size();
	for (size_t i = 0; i < decompose.size(); i++)
	{
		append(*(PolySet*)out_convex, decompose[i]->Convex, compound_of[i], i < n_untouched ? 1u : 0u);
		append(*(PolySet*)out_mesh, decompose[i]->Mesh, compound_of[i], i < n_untouched ? 1u : 0u);
	}
}

// Triangle mesh -> vertex-ring polyhedron, as PrepareFracture step 7 does (Surtr.cpp:1788-1795):
// Poly::ExtractNeighborFromMesh (Poly.cpp:128-263) + InitPolyhedron.  Returns 0, or 1 if the reference throws
// (asymmetric adjacency, Poly.cpp:253-260).
int ref_mesh_polyhedron(const float* verts4, uint32_t nv, const int32_t* indices, uint32_t n_idx, void* out)
{
	std::vector<Vector3> vertices;
	for (uint32_t v = 0; v < nv; v++)
		vertices.emplace_back(verts4[4 * v], verts4[4 * v + 1], verts4[4 * v + 2]);
	std::vector<int> idx(indices, indices + n_idx);
	try
	{
		const std::vector<std::vector<int>> nei = Poly::ExtractNeighborFromMesh(vertices, idx);
		Poly::Polyhedron mesh;
		Poly::InitPolyhedron(mesh, vertices, nei);
		append(*(PolySet*)out, mesh, 0, 0);
	}
	catch (const std::exception&)
	{
		return 1;
	}
	return 0;
}

// Poly::Transform (Poly.cpp:580-585) on a vertex list with one row-major matrix.
void ref_transform(const float* verts4, uint32_t nv, const float* matrix16, float* out4)
{
	Poly::Polyhedron p(nv);
	for (uint32_t v = 0; v < nv; v++)
		p[v].Position = Vector3(verts4[4 * v], verts4[4 * v + 1], verts4[4 * v + 2]);
	DirectX::XMMATRIX m;
	std::memcpy(m.r, matrix16, 64);
	Poly::Transform(p, m);
	for (uint32_t v = 0; v < nv; v++)
	{
		out4[4 * v] = p[v].Position.x; out4[4 * v + 1] = p[v].Position.y; out4[4 * v + 2] = p[v].Position.z; out4[4 * v + 3] = verts4[4 * v + 3];
	}
}

// Scalar helpers for the unit KATs (Poly.cpp:716-751).
int ref_compare_plane_point(const float* plane, const float* p)
{
	return Poly::ComparePlanePoint(Plane(plane[0], plane[1], plane[2], plane[3]), Vector3(p[0], p[1], p[2]));
}
void ref_plane_line_intersection(const float* a, const float* b, const float* plane, float* out)
{
	const Vector3 r = Poly::PlaneLineIntersection(Vector3(a[0], a[1], a[2]), Vector3(b[0], b[1], b[2]),
												   Plane(plane[0], plane[1], plane[2], plane[3]));
	out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void ref_plane_from_points(const float* a, const float* b, const float* c, float* out)
{
	const Plane p(Vector3(a[0], a[1], a[2]), Vector3(b[0], b[1], b[2]), Vector3(c[0], c[1], c[2]));
	out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = p.w;
}
void ref_plane_from_point_normal(const float* a, const float* n, float* out)
{
	const Plane p(Vector3(a[0], a[1], a[2]), Vector3(n[0], n[1], n[2]));
	out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = p.w;
}
// VMACH::GetBoxPolygon (VMACH.cpp:1207-1226): the six outward planes of the unit cube.
void ref_box_planes(float* out24)
{
	const VMACH::Polygon3D box = VMACH::GetBoxPolygon();
	for (int f = 0; f < 6; f++)
	{
		out24[4 * f] = box.FaceVec[f].FacePlane.x; out24[4 * f + 1] = box.FaceVec[f].FacePlane.y;
		out24[4 * f + 2] = box.FaceVec[f].FacePlane.z; out24[4 * f + 3] = box.FaceVec[f].FacePlane.w;
	}
}
}
