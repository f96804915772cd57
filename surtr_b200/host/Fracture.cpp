#include "Fracture.h"

#include "ConvexHull.h"
#include "DT3D.h"
#include "Engine.h"
#include "Kdop.h"

#include <algorithm>
#include <random>

namespace SurtrHost
{
std::vector<Vector3> GenerateSeeds(int seed, int cellCount)
{
	std::vector<Vector3> cellPointVec;
	std::mt19937 gen(seed);
	std::uniform_real_distribution<double> uniformDist(-0.5, 0.5);
	for (int i = 0; i < cellCount; i++)
	{
		const double x = uniformDist(gen);
		const double y = uniformDist(gen);
		const double z = uniformDist(gen);
		cellPointVec.emplace_back(x, y, z);
	}
	return cellPointVec;
}

std::vector<Vector3> GenerateRadialSeeds(int seed, int cellCount, double mean)
{
	std::vector<Vector3> cellPointVec;
	std::mt19937 gen(seed);
	std::uniform_real_distribution<double> directionUniformDist(-1.0, 1.0);
	std::exponential_distribution<double> lengthExpDist(1.0 / mean);
	for (int i = 0; i < cellCount; i++)
	{
		const double len = std::max(std::min(lengthExpDist(gen), 0.5), 1e-12);
		const double x = directionUniformDist(gen);
		const double y = directionUniformDist(gen);
		const double z = directionUniformDist(gen);
		Vector3 v(x, y, z);
		v.Normalize();
		v *= (float)len;
		cellPointVec.push_back(v);
	}
	return cellPointVec;
}

std::vector<VMACH::Polygon3D> GenerateVoronoi(const std::vector<Vector3>& cellPointVec) { return DT3D::VoronoiCells(cellPointVec); }

PreparedObject PrepareFracture(const std::vector<Vector3>& vertices, const std::vector<Vector3>& seeds, const FractureArgs& args)
{
	PreparedObject r;
	// 1-2. intermediate convex hull with limit count -> face normals
	const std::vector<Vector3> normals = VMACH::GenerateICHNormal(vertices, args.ICHIncludePointLimit);
	r.ICHFaceCnt = (int)normals.size();
	// 3. bounding box (doubles holding float values, as the reference)
	double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
	for (const Vector3& v : vertices)
	{
		const double c[3] = { v.x, v.y, v.z };
		for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], c[k]); hi[k] = std::max(hi[k], c[k]); }
	}
	r.BBCenter = Vector3((hi[0] + lo[0]) / 2.0, (hi[1] + lo[1]) / 2.0, (hi[2] + lo[2]) / 2.0);
	r.MinBB = Vector3(lo[0], lo[1], lo[2]);
	r.MaxBB = Vector3(hi[0], hi[1], hi[2]);
	r.MaxAxisScale = (float)std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
	// 4. min/max planes of the k-DOP (GPU)
	Kdop::KdopContainer achKdop(normals);
	achKdop.Calc(vertices, r.MaxAxisScale, args.ACHPlaneGapInverse);
	// 5-6. ACH seed box, clipped by the k-DOP (GPU)
	Poly::Polyhedron ach = Poly::GetBB();
	Poly::Scale(ach, Vector3(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]));
	Poly::Scale(ach, Vector3(2.0, 2.0, 2.0));
	Poly::Translate(ach, r.BBCenter);
	r.ACH = achKdop.ClipWithPolyhedron(ach);
	// 8. Voronoi cells for the initial decomposition, placed on the object
	r.Cells = GenerateVoronoi(seeds);
	for (VMACH::Polygon3D& voro : r.Cells)
	{
		voro.Scale(Vector3(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]));
		voro.Translate(r.BBCenter);
	}
	// 10. initial pieces
	Compound pre;
	Piece ach_piece(r.ACH, r.ACH);
	pre.PieceVec.push_back(&ach_piece);
	r.Initial = ApplyFracture(pre, r.Cells);
	return r;
}

CompoundInfo ApplyFracture(const Compound& compound, const std::vector<VMACH::Polygon3D>& voroPolyVec)
{
	detail::FlatPolys pieces;
	for (const Piece* p : compound.PieceVec)
		pieces.add(p->Convex);
	detail::FlatCells cells;
	for (const VMACH::Polygon3D& cell : voroPolyVec)
		cells.add(cell);
	detail::Fragments fr;
	detail::run_event(pieces, cells, fr);

	CompoundInfo info;
	info.CompoundBind.push_back(std::set<int>());   // 0-th element is reserved (Surtr.cpp:2126)
	int current_cell = -1;
	for (size_t f = 0; f < fr.rec.size(); f++)
	{
		const surtr_fragment& r = fr.rec[f];
		const Poly::Polyhedron convex = fr.polyhedron(f);
		info.PieceVec.push_back(new Piece(convex, convex));   // mesh branch: "next" row f-1
		if ((int)r.cell != current_cell)   // fragments arrive cell-major: one bind set per non-empty cell (Surtr.cpp:2133-2146)
		{
			info.CompoundBind.push_back(std::set<int>());
			current_cell = (int)r.cell;
		}
		info.CompoundBind.back().insert((int)f);
		MassProperties m;
		m.Volume = r.volume;
		m.Centroid = Vector3(r.centroid[0], r.centroid[1], r.centroid[2]);
		std::copy(r.inertia, r.inertia + 6, m.Inertia);
		m.FaceCount = r.n_faces;
		info.PieceMass.push_back(m);
		info.PieceSourceCell.push_back((int)r.cell);
		info.PieceSourcePiece.push_back((int)r.piece);
	}
	return info;
}

void Refitting(std::vector<Piece*>& targetPieceVec, const FractureArgs& args)
{
	const uint32_t n = (uint32_t)targetPieceVec.size();
	if (!n)
		return;
	// 1. ICH normals per piece (host; <= 4 points by default)
	std::vector<float> verts4, normals3;
	std::vector<uint32_t> vert_off{ 0 }, normal_off{ 0 };
	for (const Piece* p : targetPieceVec)
	{
		std::vector<Vector3> pts;
		for (const Poly::Vertex& v : p->Mesh)
		{
			pts.push_back(v.Position);
			verts4.insert(verts4.end(), { v.Position.x, v.Position.y, v.Position.z, 0.f });
		}
		vert_off.push_back((uint32_t)(verts4.size() / 4));
		const std::vector<Vector3> nrm = VMACH::GenerateICHNormal(pts, std::min((int)pts.size(), args.RefittingPointLimit));
		for (const Vector3& nv : nrm)
			normals3.insert(normals3.end(), { nv.x, nv.y, nv.z });
		normal_off.push_back((uint32_t)(normals3.size() / 3));
	}
	// 2. k-DOP extents of every piece->Mesh on the GPU (Kdop::Calc(Polyhedron), Kdop.cpp:92-115)
	const size_t nn = normals3.size() / 3;
	std::vector<float> dist(2 * nn), planes8(8 * nn);
	std::vector<int32_t> arg(2 * nn);
	detail::check(surtr_kdop_calc_batch(detail::context(), verts4.data(), vert_off.data(), n, normals3.data(), normal_off.data(),
										dist.data(), arg.data(), planes8.data()), "surtr_kdop_calc_batch");
	// 3. clip piece->Convex by its own plane list: n independent events of one piece x one cell
	detail::FlatPolys pieces;
	detail::FlatCells cells;
	std::vector<uint32_t> ev(n + 1);
	for (uint32_t i = 0; i < n; i++)
	{
		pieces.add(targetPieceVec[i]->Convex);
		std::vector<Poly::Plane> pl;
		for (uint32_t e = normal_off[i]; e < normal_off[i + 1]; e++)
		{
			pl.emplace_back(planes8[8 * e], planes8[8 * e + 1], planes8[8 * e + 2], planes8[8 * e + 3]);       // MinPlane
			pl.emplace_back(planes8[8 * e + 4], planes8[8 * e + 5], planes8[8 * e + 6], planes8[8 * e + 7]);   // MaxPlane
		}
		cells.add(pl);
		ev[i] = i;
	}
	ev[n] = n;
	surtr_ctx* c = detail::context();
	detail::check(surtr_upload_pieces(c, pieces.verts4.data(), pieces.vert_off.data(), pieces.ring_off.data(), pieces.ring.data(), n, ev.data(), n),
				  "surtr_upload_pieces");
	detail::check(surtr_upload_cells(c, cells.planes4.data(), cells.plane_off.data(), nullptr, nullptr, n, ev.data(), n), "surtr_upload_cells");
	detail::check(surtr_fracture_event(c), "surtr_fracture_event");
	surtr_counts cnt;
	detail::check(surtr_event_counts(c, &cnt), "surtr_event_counts");
	detail::Fragments fr;
	fr.rec.resize(cnt.n_fragments);
	fr.verts4.resize(4 * cnt.n_verts);
	fr.ring_off.resize(cnt.n_verts + 1);
	fr.ring.resize(cnt.n_ring);
	detail::check(surtr_download_fragments(c, fr.rec.data(), fr.verts4.data(), fr.ring_off.data(), fr.ring.data()), "surtr_download_fragments");
	for (Piece* p : targetPieceVec)
		p->Convex.clear();
	for (size_t f = 0; f < fr.rec.size(); f++)
		targetPieceVec[fr.rec[f].piece]->Convex = fr.polyhedron(f);
}

void SetExtract(CompoundInfo& preResult)
{
	preResult.PieceExtractedConvex.resize(preResult.PieceVec.size(), nullptr);
	std::transform(preResult.PieceVec.begin(), preResult.PieceVec.end(), preResult.PieceExtractedConvex.begin(),
				   [](const Piece* p) { return Poly::ExtractFaces(p->Convex); });
}
} // namespace SurtrHost
