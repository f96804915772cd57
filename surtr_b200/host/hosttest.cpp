// hosttest.cpp -- flat C entry points over the host-side mirror CLASSES, so the Python tests can drive the
// class-level API (Poly / Kdop / VMACH / DT3D / SurtrHost) exactly as a C++ caller would.  Test harness only.
#include <chrono>
#include "ConvexHull.h"
#include "DT3D.h"
#include "Fracture.h"
#include "Kdop.h"

#include <atomic>
#include <cstring>
#include <mutex>
#include <set>
#include <thread>
#include <stdexcept>
#include <string>

using DirectX::SimpleMath::Plane;
using DirectX::SimpleMath::Vector3;

namespace
{
std::string g_err;

Poly::Polyhedron to_poly(const float* verts, const uint32_t* ring_off, const uint16_t* ring, uint32_t v0, uint32_t v1)
{
	std::vector<Vector3> pos;
	std::vector<std::vector<int>> nei;
	for (uint32_t v = v0; v < v1; v++)
	{
		pos.emplace_back(verts[4 * v], verts[4 * v + 1], verts[4 * v + 2]);
		nei.emplace_back(ring + ring_off[v], ring + ring_off[v + 1]);
	}
	Poly::Polyhedron p;
	Poly::InitPolyhedron(p, pos, nei);
	return p;
}

struct Out
{
	std::vector<float> verts;
	std::vector<uint32_t> vert_off{ 0 }, ring_off{ 0 }, cell, piece, nfaces, plane_off{ 0 };
	std::vector<uint16_t> ring;
	std::vector<double> volume;
	std::vector<float> centroid, planes;
	void add(const Poly::Polyhedron& p)
	{
		for (const auto& v : p)
		{
			verts.insert(verts.end(), { v.Position.x, v.Position.y, v.Position.z, 0.f });
			for (int n : v.NeighborVertexVec) ring.push_back((uint16_t)n);
			ring_off.push_back((uint32_t)ring.size());
		}
		vert_off.push_back((uint32_t)(verts.size() / 4));
	}
};
Out g_out, g_mesh;   // g_mesh: Piece::Mesh of the same pieces (mesh-branch entry points)
} // namespace

extern "C"
{
const char* hosttest_error() { return g_err.c_str(); }

// sizes: n_poly, n_verts, n_ring, n_planes
void hosttest_sizes(uint64_t* s)
{
	s[0] = g_out.vert_off.size() - 1; s[1] = g_out.verts.size() / 4; s[2] = g_out.ring.size(); s[3] = g_out.planes.size() / 4;
}

#define CP(dst, vec) if (dst) std::memcpy(dst, (vec).data(), (vec).size() * sizeof((vec)[0]))
void hosttest_export(float* verts, uint32_t* vert_off, uint32_t* ring_off, uint16_t* ring, uint32_t* cell, uint32_t* piece,
					 uint32_t* nfaces, double* volume, float* centroid, float* planes, uint32_t* plane_off)
{
	CP(verts, g_out.verts); CP(vert_off, g_out.vert_off); CP(ring_off, g_out.ring_off); CP(ring, g_out.ring);
	CP(cell, g_out.cell); CP(piece, g_out.piece); CP(nfaces, g_out.nfaces); CP(volume, g_out.volume);
	CP(centroid, g_out.centroid); CP(planes, g_out.planes); CP(plane_off, g_out.plane_off);
}

// SurtrHost::GenerateSeeds + DT3D::Neighbors (CPU only)
int hosttest_seeds(int seed, int n, float* out)
{
	const auto s = SurtrHost::GenerateSeeds(seed, n);
	for (int i = 0; i < n; i++) { out[3 * i] = s[i].x; out[3 * i + 1] = s[i].y; out[3 * i + 2] = s[i].z; }
	return 0;
}

uint64_t hosttest_dt3d_neighbors(const float* seeds, uint32_t n, uint32_t* off, uint32_t* idx, uint64_t cap)
{
	std::vector<Vector3> s;
	for (uint32_t i = 0; i < n; i++) s.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
	std::vector<uint32_t> o, x;
	DT3D::Neighbors(s, o, x);
	std::memcpy(off, o.data(), 4 * o.size());
	if (x.size() <= cap) std::memcpy(idx, x.data(), 4 * x.size());
	return x.size();
}

// Tets (4 point indices each, construction order) of the grid-accelerated (use_grid = 1) or plain-scan triangulation
uint64_t hosttest_dt3d_tets(const float* seeds, uint32_t n, int use_grid, int32_t* out, uint64_t cap)
{
	std::vector<Vector3> s;
	for (uint32_t i = 0; i < n; i++) s.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
	const auto tets = DT3D::detail::TetIndices(s, use_grid != 0);
	if (tets.size() <= cap)
		for (size_t i = 0; i < tets.size(); i++)
			for (int k = 0; k < 4; k++) out[4 * i + k] = tets[i][k];
	return tets.size();
}

// DT3D::Triangulate by value: returns the tet count and the number of tets whose circumsphere contains another seed
int hosttest_dt3d_triangulate(const float* seeds, uint32_t n, uint32_t* n_tets, uint32_t* n_faces, uint32_t* n_voronoi_edges)
{
	std::vector<Vector3> s;
	for (uint32_t i = 0; i < n; i++) s.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
	const DT3D::Delaunay dt = DT3D::Triangulate(s);
	*n_tets = (uint32_t)dt.TetVec.size();
	*n_faces = (uint32_t)dt.FaceVec.size();
	*n_voronoi_edges = (uint32_t)DT3D::Voronoi(dt).size();
	int violations = 0;
	for (const auto& t : dt.TetVec)
		for (const auto& p : s)
			if (!(p == t.p0 || p == t.p1 || p == t.p2 || p == t.p3) && Vector3::Distance(t.sphere.center, p) < t.sphere.radius * (1.f - 1e-4f))
				violations++;
	return violations;
}

// VMACH::GetBoxPolygon planes + PolygonFace semantics (CPU only)
void hosttest_box_planes(float* out24)
{
	const VMACH::Polygon3D box = VMACH::GetBoxPolygon();
	for (int f = 0; f < 6; f++)
		std::memcpy(out24 + 4 * f, &box.FaceVec[f].FacePlane.x, 16);
}

// Poly::ExtractFaces (host bookkeeping) on a flat polyhedron: returns face count, writes loops
int hosttest_extract_faces(const float* verts, const uint32_t* ring_off, const uint16_t* ring, uint32_t nv, uint32_t* face_off, uint16_t* face_idx)
{
	const Poly::Polyhedron p = to_poly(verts, ring_off, ring, 0, nv);
	Poly::Extract* e = Poly::ExtractFaces(p);
	uint32_t w = 0;
	face_off[0] = 0;
	for (size_t f = 0; f < e->size(); f++)
	{
		for (int v : (*e)[f]) face_idx[w++] = (uint16_t)v;
		face_off[f + 1] = w;
	}
	const int n = (int)e->size();
	delete e;
	return n;
}

// Scalar helpers
int hosttest_compare_plane_point(const float* pl, const float* p) { return Poly::ComparePlanePoint(Plane(pl[0], pl[1], pl[2], pl[3]), Vector3(p[0], p[1], p[2])); }
void hosttest_plane_line_intersection(const float* a, const float* b, const float* pl, float* out)
{
	const Vector3 r = Poly::PlaneLineIntersection(Vector3(a[0], a[1], a[2]), Vector3(b[0], b[1], b[2]), Plane(pl[0], pl[1], pl[2], pl[3]));
	out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// VMACH::GenerateICHNormal (host ICH): returns the number of normals
uint32_t hosttest_ich_normals(const float* verts4, uint32_t nv, int limit, float* out, uint32_t cap)
{
	std::vector<Vector3> vv;
	for (uint32_t i = 0; i < nv; i++) vv.emplace_back(verts4[4 * i], verts4[4 * i + 1], verts4[4 * i + 2]);
	const std::vector<Vector3> n = VMACH::GenerateICHNormal(vv, limit);
	for (uint32_t i = 0; i < n.size() && i < cap; i++) { out[3 * i] = n[i].x; out[3 * i + 1] = n[i].y; out[3 * i + 2] = n[i].z; }
	return (uint32_t)n.size();
}

// ---- GPU-backed class API ----
// SurtrHost::Refitting over pieces {Convex = polyset A, Mesh = point sets B}
int hosttest_refit(const float* cverts, const uint32_t* cvert_off, const uint32_t* cring_off, const uint16_t* cring, uint32_t n,
				   const float* mverts, const uint32_t* mvert_off, int limit)
{
	try
	{
		std::vector<SurtrHost::Piece*> pieces;
		for (uint32_t i = 0; i < n; i++)
		{
			const Poly::Polyhedron convex = to_poly(cverts, cring_off, cring, cvert_off[i], cvert_off[i + 1]);
			Poly::Polyhedron mesh(mvert_off[i + 1] - mvert_off[i]);
			for (uint32_t v = mvert_off[i]; v < mvert_off[i + 1]; v++)
				mesh[v - mvert_off[i]].Position = Vector3(mverts[4 * v], mverts[4 * v + 1], mverts[4 * v + 2]);
			pieces.push_back(new SurtrHost::Piece(convex, mesh));
		}
		SurtrHost::FractureArgs args;
		args.RefittingPointLimit = limit;
		SurtrHost::Refitting(pieces, args);
		g_out = Out();
		for (auto* p : pieces) { g_out.add(p->Convex); delete p; }
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// SurtrHost::GenerateVoronoi(seeds) -> cells as polyhedra-free Polygon3D: exports planes + face vertices
int hosttest_voronoi(const float* seeds, uint32_t n)
{
	try
	{
		std::vector<Vector3> s;
		for (uint32_t i = 0; i < n; i++) s.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
		const auto cells = SurtrHost::GenerateVoronoi(s);
		g_out = Out();
		for (const auto& c : cells)
		{
			for (const auto& f : c.FaceVec)
				g_out.planes.insert(g_out.planes.end(), { f.FacePlane.x, f.FacePlane.y, f.FacePlane.z, f.FacePlane.w });
			g_out.plane_off.push_back((uint32_t)(g_out.planes.size() / 4));
		}
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// SurtrHost::ApplyFracture + SetExtract over Compound / Polygon3D objects built from flat inputs.
// Cells are given as planes + face-vertex streams (one PolygonFace per plane, vertices split evenly is not needed:
// the clipper reads only FacePlane; the vertex stream of the whole cell is attached to its first face for the bounds).
int hosttest_apply_fracture(const float* verts, const uint32_t* vert_off, const uint32_t* ring_off, const uint16_t* ring, uint32_t n_pieces,
							const float* planes, const uint32_t* plane_off, const float* cverts, const uint32_t* cvert_off, uint32_t n_cells)
{
	try
	{
		SurtrHost::Compound compound;
		for (uint32_t i = 0; i < n_pieces; i++)
		{
			const Poly::Polyhedron p = to_poly(verts, ring_off, ring, vert_off[i], vert_off[i + 1]);
			compound.PieceVec.push_back(new SurtrHost::Piece(p, p));
		}
		std::vector<VMACH::Polygon3D> cells;
		for (uint32_t c = 0; c < n_cells; c++)
		{
			VMACH::Polygon3D poly(true);
			for (uint32_t k = plane_off[c]; k < plane_off[c + 1]; k++)
			{
				VMACH::PolygonFace f(true);
				if (k == plane_off[c])
					for (uint32_t v = cvert_off[c]; v < cvert_off[c + 1]; v++)
						f.VertexVec.emplace_back(cverts[4 * v], cverts[4 * v + 1], cverts[4 * v + 2]);
				else
					f.VertexVec.emplace_back(cverts[4 * cvert_off[c]], cverts[4 * cvert_off[c] + 1], cverts[4 * cvert_off[c] + 2]);
				f.ManuallySetFacePlane(Plane(planes[4 * k], planes[4 * k + 1], planes[4 * k + 2], planes[4 * k + 3]));
				poly.AddFace(f);
			}
			cells.push_back(poly);
		}
		SurtrHost::CompoundInfo info = SurtrHost::ApplyFracture(compound, cells, false);
		SurtrHost::SetExtract(info);
		g_out = Out();
		for (size_t i = 0; i < info.PieceVec.size(); i++)
		{
			g_out.add(info.PieceVec[i]->Convex);
			g_out.cell.push_back((uint32_t)info.PieceSourceCell[i]);
			g_out.piece.push_back((uint32_t)info.PieceSourcePiece[i]);
			g_out.nfaces.push_back((uint32_t)info.PieceExtractedConvex[i]->size());
			// Poly::Moments through the class API for the first few pieces, K4's record for all
			g_out.volume.push_back(info.PieceMass[i].Volume);
			g_out.centroid.insert(g_out.centroid.end(), { info.PieceMass[i].Centroid.x, info.PieceMass[i].Centroid.y, info.PieceMass[i].Centroid.z });
			if ((int)info.PieceExtractedConvex[i]->size() != info.PieceMass[i].FaceCount) { g_err = "face count mismatch between ExtractFaces and K4"; return 2; }
		}
		// bind sets: consecutive, one per non-empty cell
		size_t expect = 0;
		for (size_t b = 1; b < info.CompoundBind.size(); b++)
			for (int idx : info.CompoundBind[b])
				if ((size_t)idx != expect++) { g_err = "CompoundBind order"; return 3; }
		for (auto* p : compound.PieceVec) delete p;
		for (auto* p : info.PieceVec) delete p;
		for (auto* e : info.PieceExtractedConvex) delete e;
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// select which set hosttest_sizes / hosttest_export read: 0 = convex (default), 1 = mesh
void hosttest_swap_sets() { std::swap(g_out, g_mesh); }

// Poly::ExtractNeighborFromMesh (CPU only): rings into g_out as one polyhedron
int hosttest_mesh_polyhedron(const float* verts4, uint32_t nv, const int32_t* indices, uint32_t n_idx)
{
	try
	{
		std::vector<Vector3> vv;
		for (uint32_t i = 0; i < nv; i++) vv.emplace_back(verts4[4 * i], verts4[4 * i + 1], verts4[4 * i + 2]);
		const std::vector<int> idx(indices, indices + n_idx);
		Poly::Polyhedron mesh;
		Poly::InitPolyhedron(mesh, vv, Poly::ExtractNeighborFromMesh(vv, idx));
		g_out = Out();
		g_out.add(mesh);
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

static void export_pieces(const SurtrHost::CompoundInfo& info)
{
	g_out = Out();
	g_mesh = Out();
	for (size_t i = 0; i < info.PieceVec.size(); i++)
	{
		g_out.add(info.PieceVec[i]->Convex);
		g_mesh.add(info.PieceVec[i]->Mesh);
		for (Out* o : { &g_out, &g_mesh })
		{
			o->cell.push_back((uint32_t)info.PieceSourceCell[i]);
			o->piece.push_back((uint32_t)info.PieceSourcePiece[i]);
			o->nfaces.push_back(i < info.PieceExtractedConvex.size() ? (uint32_t)info.PieceExtractedConvex[i]->size() : 0u);
			o->volume.push_back(info.PieceMass[i].Volume);
			o->centroid.insert(o->centroid.end(), { info.PieceMass[i].Centroid.x, info.PieceMass[i].Centroid.y, info.PieceMass[i].Centroid.z });
		}
	}
}

// SurtrHost::ApplyFracture with the mesh branch (full m_fractureTask): pieces = (convex_i, mesh_i); cells as plane lists
// with their vertex streams.  g_out = Piece::Convex, g_mesh = Piece::Mesh of the result, PieceVec order.
int hosttest_apply_fracture_mesh(const float* cv, const uint32_t* cvo, const uint32_t* cro, const uint16_t* cr,
								 const float* mv, const uint32_t* mvo, const uint32_t* mro, const uint16_t* mr, uint32_t n_pieces,
								 const float* planes, const uint32_t* plane_off, const float* cverts, const uint32_t* cvert_off, uint32_t n_cells)
{
	try
	{
		SurtrHost::Compound compound;
		for (uint32_t i = 0; i < n_pieces; i++)
			compound.PieceVec.push_back(new SurtrHost::Piece(to_poly(cv, cro, cr, cvo[i], cvo[i + 1]), to_poly(mv, mro, mr, mvo[i], mvo[i + 1])));
		std::vector<VMACH::Polygon3D> cells;
		for (uint32_t c = 0; c < n_cells; c++)
		{
			VMACH::Polygon3D poly(true);
			for (uint32_t k = plane_off[c]; k < plane_off[c + 1]; k++)
			{
				VMACH::PolygonFace f(true);
				if (k == plane_off[c])
					for (uint32_t v = cvert_off[c]; v < cvert_off[c + 1]; v++)
						f.VertexVec.emplace_back(cverts[4 * v], cverts[4 * v + 1], cverts[4 * v + 2]);
				else
					f.VertexVec.emplace_back(cverts[4 * cvert_off[c]], cverts[4 * cvert_off[c] + 1], cverts[4 * cvert_off[c] + 2]);
				f.ManuallySetFacePlane(Plane(planes[4 * k], planes[4 * k + 1], planes[4 * k + 2], planes[4 * k + 3]));
				poly.AddFace(f);
			}
			cells.push_back(poly);
		}
		SurtrHost::CompoundInfo info = SurtrHost::ApplyFracture(compound, cells, true);
		SurtrHost::SetExtract(info);
		export_pieces(info);
		for (auto* p : compound.PieceVec) delete p;
		for (auto* p : info.PieceVec) delete p;
		for (auto* e : info.PieceExtractedConvex) delete e;
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// detail::parallel_for (CPU only): every index exactly once, from more than one thread for large n; a throwing task
// surfaces as the call's exception after all other tasks ran.  Returns 0 on success.
int hosttest_parallel_for(uint32_t n, uint32_t throw_at, uint32_t* n_threads_seen)
{
	std::vector<std::atomic<int>> hits(n);
	for (auto& h : hits) h.store(0);
	std::mutex mu;
	std::set<std::thread::id> ids;
	bool thrown = false;
	try
	{
		SurtrHost::detail::parallel_for(n, [&](size_t i) {
			hits[i].fetch_add(1);
			{
				std::lock_guard<std::mutex> lock(mu);
				ids.insert(std::this_thread::get_id());
			}
			volatile double x = 0;
			for (int k = 0; k < 2000; k++) x = x + k * 0.5;
			if (i == throw_at) throw std::runtime_error("task failed");
		});
	}
	catch (const std::runtime_error&) { thrown = true; }
	*n_threads_seen = (uint32_t)ids.size();
	for (uint32_t i = 0; i < n; i++)
		if (hits[i].load() != 1) { g_err = "index not visited exactly once"; return 1; }
	if (thrown != (throw_at < n)) { g_err = "exception not propagated"; return 2; }
	return 0;
}

// Poly::Transform (CPU only)
void hosttest_transform(const float* verts4, uint32_t nv, const float* matrix16, float* out4)
{
	Poly::Polyhedron p(nv);
	for (uint32_t v = 0; v < nv; v++) p[v].Position = Vector3(verts4[4 * v], verts4[4 * v + 1], verts4[4 * v + 2]);
	DirectX::XMMATRIX m;
	std::memcpy(m.r, matrix16, 64);
	Poly::Transform(p, m);
	for (uint32_t v = 0; v < nv; v++)
	{
		out4[4 * v] = p[v].Position.x; out4[4 * v + 1] = p[v].Position.y; out4[4 * v + 2] = p[v].Position.z; out4[4 * v + 3] = verts4[4 * v + 3];
	}
}

// SurtrHost::CombineMass (CPU only): per piece volume, centroid[3], inertia[6] -> {mass, c[3], I[6]}
void hosttest_combine_mass(uint32_t n, const double* volume, const float* centroid3, const float* inertia6, float density, float* out10)
{
	std::vector<SurtrHost::MassProperties> parts(n);
	for (uint32_t i = 0; i < n; i++)
	{
		parts[i].Volume = volume[i];
		parts[i].Centroid = Vector3(centroid3[3 * i], centroid3[3 * i + 1], centroid3[3 * i + 2]);
		std::memcpy(parts[i].Inertia, inertia6 + 6 * i, 6 * sizeof(float));
	}
	const SurtrHost::MassProperties m = SurtrHost::CombineMass(parts, density);
	const float row[10] = { (float)m.Volume, m.Centroid.x, m.Centroid.y, m.Centroid.z, m.Inertia[0], m.Inertia[1], m.Inertia[2],
							m.Inertia[3], m.Inertia[4], m.Inertia[5] };
	std::memcpy(out10, row, sizeof(row));
}

// SurtrHost::DoFracture on one compound (convex_i, mesh_i).  g_out / g_mesh = Piece::Convex / Piece::Mesh in PieceVec order;
// cell = index of the compound (bind set) a piece ends up in, piece = 1 for the caller's untouched pieces.
// mass10 (optional, 10 floats per compound): CombineMass at density 10 = {mass, cx, cy, cz, Ixx, Iyy, Izz, Ixy, Ixz, Iyz}.
// wall time of the SurtrHost::DoFracture call inside the last hosttest_do_fracture (without this wrapper's array <-> class
// conversions and the pattern generation, which the reference does once at start-up, Surtr.cpp:1436-1450)
static double g_do_fracture_ms = 0.0;
double hosttest_last_do_fracture_ms() { return g_do_fracture_ms; }

int hosttest_do_fracture(const float* cv, const uint32_t* cvo, const uint32_t* cro, const uint16_t* cr,
						 const float* mv, const uint32_t* mvo, const uint32_t* mro, const uint16_t* mr, uint32_t n_pieces,
						 const float* seeds, uint32_t n_seeds, const float* cloud3, uint32_t n_cloud, const float* impact3,
						 float impact_radius, float max_axis_scale, int partial, uint32_t* n_compounds, float* mass10, uint32_t mass_cap)
{
	try
	{
		SurtrHost::Compound compound;
		for (uint32_t i = 0; i < n_pieces; i++)
		{
			compound.PieceVec.push_back(new SurtrHost::Piece(to_poly(cv, cro, cr, cvo[i], cvo[i + 1]), to_poly(mv, mro, mr, mvo[i], mvo[i + 1])));
			compound.PieceExtractedConvex.push_back(Poly::ExtractFaces(compound.PieceVec.back()->Convex));
		}
		std::vector<Vector3> ss, cloud;
		for (uint32_t i = 0; i < n_seeds; i++) ss.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
		for (uint32_t i = 0; i < n_cloud; i++) cloud.emplace_back(cloud3[3 * i], cloud3[3 * i + 1], cloud3[3 * i + 2]);
		SurtrHost::FractureStorage storage;
		storage.MaxAxisScale = max_axis_scale;
		storage.PartialFracturePattern = storage.GeneralFracturePattern = SurtrHost::GenerateVoronoi(ss);
		SurtrHost::FractureArgs args;
		args.ImpactPosition = DirectX::XMFLOAT3(impact3[0], impact3[1], impact3[2]);
		args.ImpactRadius = impact_radius;
		args.PartialFracture = partial != 0;
		SurtrHost::CompoundInfo info;
		const auto t0 = std::chrono::steady_clock::now();
		const std::vector<SurtrHost::Compound> result = SurtrHost::DoFracture(compound, storage, cloud, args, &info);
		g_do_fracture_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		*n_compounds = (uint32_t)result.size();
		std::vector<uint32_t> compound_of(info.PieceVec.size(), 0xffffffffu);
		for (size_t b = 0; b < info.CompoundBind.size(); b++)
			for (const int c : info.CompoundBind[b])
				compound_of[c] = (uint32_t)b;
		export_pieces(info);
		for (Out* o : { &g_out, &g_mesh })
			for (size_t i = 0; i < info.PieceVec.size(); i++)
			{
				o->cell[i] = compound_of[i];
				o->piece[i] = info.PieceSourceCell[i] < 0 ? 1u : 0u;
			}
		for (size_t b = 0; b < result.size() && mass10 && b < mass_cap; b++)
		{
			std::vector<SurtrHost::MassProperties> parts;
			for (const int c : info.CompoundBind[b])
				parts.push_back(info.PieceMass[c]);
			const SurtrHost::MassProperties m = SurtrHost::CombineMass(parts, 10.0f);
			const float row[10] = { (float)m.Volume, m.Centroid.x, m.Centroid.y, m.Centroid.z, m.Inertia[0], m.Inertia[1], m.Inertia[2],
									m.Inertia[3], m.Inertia[4], m.Inertia[5] };
			std::memcpy(mass10 + 10 * b, row, sizeof(row));
		}
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// SurtrHost::PrepareFracture in full (mesh polyhedron, mesh branch, Refitting, SetExtract)
int hosttest_config1_full(const float* verts4, uint32_t nv, const int32_t* indices, uint32_t n_idx, const float* seeds, uint32_t n_seeds,
						  uint32_t* ach_nv)
{
	try
	{
		std::vector<Vector3> vv, ss;
		for (uint32_t i = 0; i < nv; i++) vv.emplace_back(verts4[4 * i], verts4[4 * i + 1], verts4[4 * i + 2]);
		for (uint32_t i = 0; i < n_seeds; i++) ss.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
		const std::vector<int> idx(indices, indices + n_idx);
		SurtrHost::PreparedObject r = SurtrHost::PrepareFracture(vv, idx, ss);
		*ach_nv = (uint32_t)r.ACH.size();
		export_pieces(r.Initial);
		for (auto* p : r.Initial.PieceVec) delete p;
		for (auto* e : r.Initial.PieceExtractedConvex) delete e;
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// SurtrHost::PrepareFracture (config 1, convex branch): exports the fragments; ach_out gets {nv, nf-unused}
int hosttest_config1(const float* verts4, uint32_t nv, const float* seeds, uint32_t n_seeds, uint32_t* ach_nv)
{
	try
	{
		std::vector<Vector3> vv, ss;
		for (uint32_t i = 0; i < nv; i++) vv.emplace_back(verts4[4 * i], verts4[4 * i + 1], verts4[4 * i + 2]);
		for (uint32_t i = 0; i < n_seeds; i++) ss.emplace_back(seeds[3 * i], seeds[3 * i + 1], seeds[3 * i + 2]);
		SurtrHost::PreparedObject r = SurtrHost::PrepareFracture(vv, ss);
		*ach_nv = (uint32_t)r.ACH.size();
		g_out = Out();
		for (size_t i = 0; i < r.Initial.PieceVec.size(); i++)
		{
			g_out.add(r.Initial.PieceVec[i]->Convex);
			g_out.cell.push_back((uint32_t)r.Initial.PieceSourceCell[i]);
			g_out.piece.push_back((uint32_t)r.Initial.PieceSourcePiece[i]);
			g_out.nfaces.push_back((uint32_t)r.Initial.PieceMass[i].FaceCount);
			g_out.volume.push_back(r.Initial.PieceMass[i].Volume);
			g_out.centroid.insert(g_out.centroid.end(), { r.Initial.PieceMass[i].Centroid.x, r.Initial.PieceMass[i].Centroid.y, r.Initial.PieceMass[i].Centroid.z });
			delete r.Initial.PieceVec[i];
		}
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// Poly::ClipPolyhedron (in place) + Poly::Moments + Kdop::KdopContainer through the class API
int hosttest_clip_and_moments(const float* verts, const uint32_t* ring_off, const uint16_t* ring, uint32_t nv, const float* planes, uint32_t npl,
							  double* volume, float* centroid)
{
	try
	{
		Poly::Polyhedron p = to_poly(verts, ring_off, ring, 0, nv);
		std::vector<Plane> pls;
		for (uint32_t k = 0; k < npl; k++) pls.emplace_back(planes[4 * k], planes[4 * k + 1], planes[4 * k + 2], planes[4 * k + 3]);
		Poly::ClipPolyhedron(p, pls);
		g_out = Out();
		g_out.add(p);
		Vector3 c;
		Poly::Moments(*volume, c, p);
		centroid[0] = c.x; centroid[1] = c.y; centroid[2] = c.z;
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int hosttest_kdop_ach(const float* verts, uint32_t nv, const float* normals, uint32_t k, double maxAxisScale, float gapInv,
					  const float* boxverts, float* out_planes)
{
	try
	{
		std::vector<Vector3> vv, nn;
		for (uint32_t i = 0; i < nv; i++) vv.emplace_back(verts[4 * i], verts[4 * i + 1], verts[4 * i + 2]);
		for (uint32_t i = 0; i < k; i++) nn.emplace_back(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
		Kdop::KdopContainer kd(nn);
		kd.Calc(vv, maxAxisScale, gapInv);
		for (uint32_t i = 0; i < k; i++)
		{
			std::memcpy(out_planes + 8 * i, &kd.ElementVec[i].MinPlane.x, 16);
			std::memcpy(out_planes + 8 * i + 4, &kd.ElementVec[i].MaxPlane.x, 16);
		}
		Poly::Polyhedron box = Poly::GetBB();
		for (int i = 0; i < 8; i++) box[i].Position = Vector3(boxverts[4 * i], boxverts[4 * i + 1], boxverts[4 * i + 2]);
		const Poly::Polyhedron ach = kd.ClipWithPolyhedron(box);
		g_out = Out();
		g_out.add(ach);
		return 0;
	}
	catch (const std::exception& e) { g_err = e.what(); return 1; }
}
}
