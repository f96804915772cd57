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

static PreparedObject prepare(const std::vector<Vector3>& vertices, const std::vector<int>* indices, const std::vector<Vector3>& seeds,
							  const FractureArgs& args)
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
	// 7. mesh polyhedron
	if (indices)
	{
		const std::vector<std::vector<int>> nei = Poly::ExtractNeighborFromMesh(vertices, *indices);
		Poly::InitPolyhedron(r.Mesh, vertices, nei);
	}
	// 10. initial pieces
	Compound pre;
	Piece first_piece(r.ACH, indices ? r.Mesh : r.ACH);
	pre.PieceVec.push_back(&first_piece);
	r.Initial = ApplyFracture(pre, r.Cells, indices != nullptr);
	if (indices)
	{
		Refitting(r.Initial.PieceVec, args);
		SetExtract(r.Initial);
	}
	return r;
}

PreparedObject PrepareFracture(const std::vector<Vector3>& vertices, const std::vector<Vector3>& seeds, const FractureArgs& args)
{
	return prepare(vertices, nullptr, seeds, args);
}

PreparedObject PrepareFracture(const std::vector<Vector3>& vertices, const std::vector<int>& indices, const std::vector<Vector3>& seeds,
							   const FractureArgs& args)
{
	return prepare(vertices, &indices, seeds, args);
}

std::vector<std::set<int>> CheckMeshIsland(const Poly::Polyhedron& polyhedron)
{
	// Surtr.cpp:2157-2199 recurses from an arbitrary vertex; an explicit stack reaches the same sets.  Groups come
	// out in order of their lowest not-yet-grouped vertex, each as a sorted set.
	std::vector<std::set<int>> groupVec;
	std::vector<char> grouped(polyhedron.size(), 0);
	std::vector<int> stack;
	int start = 0;
	for (;;)
	{
		std::set<int> group;
		stack.assign(1, start);
		while (!stack.empty())
		{
			const int v = stack.back();
			stack.pop_back();
			for (const int a : polyhedron[v].NeighborVertexVec)
				if (group.insert(a).second)
					stack.push_back(a);
		}
		for (const int v : group)
			grouped[v] = 1;
		groupVec.push_back(std::move(group));
		grouped[start] = 1;   // a start vertex without neighbours is in no group (the reference would spin on it)
		const auto rest = std::find(grouped.begin() + start, grouped.end(), 0);
		if (rest == grouped.end())
			break;
		start = (int)(rest - grouped.begin());
	}
	return groupVec;
}

CompoundInfo ApplyFracture(const Compound& compound, const std::vector<VMACH::Polygon3D>& voroPolyVec, bool meshBranch)
{
	detail::FlatPolys pieces;
	for (const Piece* p : compound.PieceVec)
		pieces.add(p->Convex);
	detail::FlatCells cells;
	for (const VMACH::Polygon3D& cell : voroPolyVec)
		cells.add(cell);
	detail::Fragments fr, mfr;
	detail::run_event(pieces, cells, fr);
	if (meshBranch)
	{
		// second clip of m_fractureTask (Surtr.cpp:1470): every Piece::Mesh against the same resident cells, one more
		// GPU event.  The broad phase culls with the mesh's own (tighter) extents; a pair yields pieces only when both
		// its convex and its mesh fragment exist (:1466-1472).
		detail::FlatPolys meshes;
		for (const Piece* p : compound.PieceVec)
			meshes.add(p->Mesh);
		detail::run_event(meshes, cells, mfr, true, false);
	}

	CompoundInfo info;
	info.CompoundBind.push_back(std::set<int>());   // 0-th element is reserved (Surtr.cpp:2126)
	int current_cell = -1;
	size_t m = 0;
	const auto key = [](const surtr_fragment& r) { return ((uint64_t)r.cell << 32) | r.piece; };
	for (size_t f = 0; f < fr.rec.size(); f++)
	{
		const surtr_fragment& r = fr.rec[f];
		const Poly::Polyhedron convex = fr.polyhedron(f);
		std::vector<Poly::Polyhedron> meshes;
		if (meshBranch)
		{
			while (m < mfr.rec.size() && key(mfr.rec[m]) < key(r))   // both lists are cell-major, piece-minor
				m++;
			if (m == mfr.rec.size() || key(mfr.rec[m]) != key(r))
				continue;   // mesh clipped away (Surtr.cpp:1471)
			const Poly::Polyhedron mesh = mfr.polyhedron(m);
			const std::vector<std::set<int>> groupVec = CheckMeshIsland(mesh);
			if (groupVec.size() >= 2)
			{
				std::vector<int> mapping(mesh.size(), -1);
				for (const std::set<int>& group : groupVec)   // Surtr.cpp:1475-1495
				{
					Poly::Polyhedron island;
					for (const int iVert : group)
					{
						mapping[iVert] = (int)island.size();
						island.push_back(mesh[iVert]);
					}
					for (Poly::Vertex& vert : island)
						for (int& iAdj : vert.NeighborVertexVec)
							iAdj = mapping[iAdj];
					meshes.push_back(std::move(island));
				}
			}
			else
				meshes.push_back(mesh);
		}
		else
			meshes.push_back(convex);
		for (Poly::Polyhedron& mesh : meshes)
		{
			if ((int)r.cell != current_cell)   // one bind set per cell that produced pieces, cell order (Surtr.cpp:2133-2146)
			{
				info.CompoundBind.push_back(std::set<int>());
				current_cell = (int)r.cell;
			}
			info.CompoundBind.back().insert((int)info.PieceVec.size());
			info.PieceVec.push_back(new Piece(convex, mesh));
			MassProperties mp;
			mp.Volume = r.volume;
			mp.Centroid = Vector3(r.centroid[0], r.centroid[1], r.centroid[2]);
			std::copy(r.inertia, r.inertia + 6, mp.Inertia);
			mp.FaceCount = r.n_faces;
			info.PieceMass.push_back(mp);
			info.PieceSourceCell.push_back((int)r.cell);
			info.PieceSourcePiece.push_back((int)r.piece);
		}
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
