#include "Fracture.h"

#include "ConvexHull.h"
#include "DT3D.h"
#include "Engine.h"
#include "Kdop.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <random>

namespace SurtrHost
{
namespace
{
// SURTR_TRACE=1 prints the wall time of every orchestration phase to stderr (the reference's TIMER_START_NAME /
// TIMER_STOP_PRINT around the same phases, Surtr.cpp:1917-1944).
struct Phase
{
	const char* name;
	std::chrono::steady_clock::time_point t0;
	static bool enabled()
	{
		static const bool on = std::getenv("SURTR_TRACE") != nullptr;
		return on;
	}
	explicit Phase(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
	~Phase()
	{
		if (enabled())
			std::fprintf(stderr, "[surtr] %-28s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
};
} // namespace

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
	std::unique_ptr<Phase> ph(new Phase("ICH normals"));
	const std::vector<Vector3> normals = VMACH::GenerateICHNormal(vertices, args.ICHIncludePointLimit);
	r.ICHFaceCnt = (int)normals.size();
	ph.reset(new Phase("bbox + k-DOP + ACH"));
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
	ph.reset(new Phase("Voronoi cells (DT3D + clip)"));
	// 8. Voronoi cells for the initial decomposition, placed on the object
	r.Cells = GenerateVoronoi(seeds);
	for (VMACH::Polygon3D& voro : r.Cells)
	{
		voro.Scale(Vector3(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]));
		voro.Translate(r.BBCenter);
	}
	ph.reset(new Phase("mesh polyhedron"));
	// 7. mesh polyhedron
	if (indices)
	{
		const std::vector<std::vector<int>> nei = Poly::ExtractNeighborFromMesh(vertices, *indices);
		Poly::InitPolyhedron(r.Mesh, vertices, nei);
	}
	ph.reset(new Phase("ApplyFracture (initial)"));
	// 10. initial pieces
	Compound pre;
	Piece first_piece(r.ACH, indices ? r.Mesh : r.ACH);
	pre.PieceVec.push_back(&first_piece);
	r.Initial = ApplyFracture(pre, r.Cells, indices != nullptr);
	ph.reset(new Phase("Refitting + SetExtract"));
	if (indices)
	{
		Refitting(r.Initial.PieceVec, args, &r.Initial.PieceMass);
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

namespace
{
// The islands of CheckMeshIsland as sorted vertex lists (no node-based sets: one flag array, one stack).
std::vector<std::vector<int>> mesh_islands(const Poly::Polyhedron& polyhedron)
{
	std::vector<std::vector<int>> groups;
	const int n = (int)polyhedron.size();
	std::vector<char> grouped(n, 0);
	std::vector<int> in_group(n, 0), stack;   // in_group[v] = stamp of the group that holds v (every group starts from an empty set)
	int start = 0, stamp = 0;
	while (n)
	{
		std::vector<int> group;
		stamp++;
		stack.assign(1, start);
		while (!stack.empty())
		{
			const int v = stack.back();
			stack.pop_back();
			for (const int a : polyhedron[v].NeighborVertexVec)
				if (a >= 0 && a < n && in_group[a] != stamp)
				{
					in_group[a] = stamp;
					group.push_back(a);
					stack.push_back(a);
				}
		}
		std::sort(group.begin(), group.end());
		for (const int v : group)
			grouped[v] = 1;
		groups.push_back(std::move(group));
		grouped[start] = 1;   // a start vertex without neighbours is in no group (the reference would spin on it)
		const auto rest = std::find(grouped.begin() + start, grouped.end(), 0);
		if (rest == grouped.end())
			break;
		start = (int)(rest - grouped.begin());
	}
	return groups;
}
} // namespace

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

bool ConvexOutOfSphere(const Poly::Polyhedron& polyhedron, const Extract* extract, const std::vector<Vector3>& spherePointCloud,
					   const Vector3 origin, const float radius)
{
	// approximate test of Surtr.cpp:2415-2458: no vertex inside the sphere, and no sample point of the sphere's
	// surface inside the convex
	for (const Poly::Vertex& v : polyhedron)
		if ((origin - v.Position).Length() < radius)
			return false;
	std::vector<Vector3> normals;
	std::vector<float> offsets;
	for (const std::vector<int>& f : *extract)   // hoisted out of the point loop; same values, same order per point
	{
		Vector3 normal = (polyhedron[f[1]].Position - polyhedron[f[0]].Position).Cross(polyhedron[f[2]].Position - polyhedron[f[0]].Position);
		normal.Normalize();
		normals.push_back(normal);
		offsets.push_back(-polyhedron[f[0]].Position.Dot(normal));
	}
	for (const Vector3& po : spherePointCloud)
	{
		bool contain = true;
		for (size_t k = 0; k < normals.size(); k++)
		{
			const float dist = normals[k].Dot(po) + offsets[k];
			if (dist > 0)
			{
				contain = false;
				break;
			}
		}
		if (contain)
			return false;
	}
	return true;
}

namespace
{
// where the cells of an event come from: a host-side Polygon3D list (packed and uploaded), or the context's resident
// pattern placed on the device
struct CellSource
{
	const std::vector<VMACH::Polygon3D>* polys = nullptr;
	const detail::FlatPattern* pattern = nullptr;
	Vector3 scale, translate;
	uint32_t count() const { return polys ? (uint32_t)polys->size() : pattern->count(); }
};

CompoundInfo apply_fracture(const Compound& compound, const CellSource& source, const std::vector<Vector3>& spherePointCloud, bool partial,
							const FractureArgs& args, bool meshBranch);
bool refit_begin(const std::vector<Piece*>& targetPieceVec, const FractureArgs& args);
void refit_end(std::vector<Piece*>& targetPieceVec, std::vector<MassProperties>* mass);
} // namespace

CompoundInfo ApplyFracture(const Compound& compound, const std::vector<VMACH::Polygon3D>& voroPolyVec, bool meshBranch)
{
	return ApplyFracture(compound, voroPolyVec, std::vector<Vector3>(), false, FractureArgs(), meshBranch);
}

CompoundInfo ApplyFracture(const Compound& compound, const std::vector<VMACH::Polygon3D>& voroPolyVec,
						   const std::vector<Vector3>& spherePointCloud, bool partial, const FractureArgs& args, bool meshBranch)
{
	CellSource source;
	source.polys = &voroPolyVec;
	return apply_fracture(compound, source, spherePointCloud, partial, args, meshBranch);
}

namespace
{
CompoundInfo apply_fracture(const Compound& compound, const CellSource& source, const std::vector<Vector3>& spherePointCloud, bool partial,
							const FractureArgs& args, bool meshBranch)
{
	const std::vector<Piece*>& targetPieceVec = compound.PieceVec;
	// pieces that lie outside the impact sphere are not cut (Surtr.cpp:2109-2124)
	std::vector<int> inside, outside;
	for (int c = 0; c < (int)targetPieceVec.size(); c++)
	{
		bool out = false;
		if (partial)
		{
			const Extract* extract = c < (int)compound.PieceExtractedConvex.size() ? compound.PieceExtractedConvex[c] : nullptr;
			Extract* own = extract ? nullptr : Poly::ExtractFaces(targetPieceVec[c]->Convex);
			out = ConvexOutOfSphere(targetPieceVec[c]->Convex, extract ? extract : own, spherePointCloud, args.ImpactPosition, args.ImpactRadius);
			delete own;
		}
		(out ? outside : inside).push_back(c);
	}
	const uint32_t n_in = (uint32_t)inside.size(), n_out = (uint32_t)outside.size(), n_cells = source.count();

	// event 1: the pieces inside the sphere x the cells (one pool task per cell in the reference, :2129-2131)
	detail::Fragments fr, mfr, ofr;
	detail::FlatCells cells;
	std::vector<Poly::Polyhedron> convexPoly;   // mesh branch: fragment f of the convex event, unpacked early
	if (n_in)
	{
		// The convex and the mesh clip are independent given the cells: they run on the thread's two contexts, so the
		// host packs the meshes while the GPU cuts the convex pieces, and both events share the device afterwards.
		detail::FlatPolys pieces, meshes;
		const auto install_cells = [&](int slot) {
			if (!source.polys)
				detail::place_pattern(*source.pattern, source.scale, source.translate, slot);
		};
		{
			Phase ph("  pack convex + cells");
			for (const int c : inside)
				pieces.add(targetPieceVec[c]->Convex);
			if (source.polys)
				for (const VMACH::Polygon3D& cell : *source.polys)
					cells.add(cell);
			install_cells(0);
		}
		detail::begin_event(pieces, cells, source.polys != nullptr, 0);
		if (meshBranch)
		{
			// event 2, the second clip of m_fractureTask (Surtr.cpp:1470): every Piece::Mesh against the same cells.  The
			// broad phase culls with the mesh's own (tighter) extents; a pair yields pieces only when both its convex and
			// its mesh fragment exist (:1466-1472).
			{
				Phase ph2("  pack meshes");
				for (const int c : inside)
					meshes.add(targetPieceVec[c]->Mesh);
				install_cells(1);
			}
			detail::begin_event(meshes, cells, source.polys != nullptr, 1);
		}
		{
			Phase ph("  convex event (rest)");
			detail::end_event(fr, true, 0);
		}
		if (meshBranch)
		{
			{
				// the convex fragments become Poly::Polyhedron objects while the GPU is still cutting the meshes
				Phase ph("  unpack convex");
				convexPoly.resize(fr.rec.size());
				detail::parallel_for(fr.rec.size(), [&](size_t f) { convexPoly[f] = fr.polyhedron(f); });
			}
			Phase ph("  mesh event (rest)");
			detail::end_event(mfr, true, 1);
		}
	}
	if (n_out)
	{
		// event 3: the untouched pieces x one cell without planes, which hands them back uncut -- only so that K3/K4
		// compute their face counts and mass properties like everybody else's
		detail::FlatPolys pieces;
		for (const int c : outside)
			pieces.add(targetPieceVec[c]->Convex);
		detail::FlatCells keep;
		keep.add_keep_all();
		detail::run_event(pieces, keep, ofr, false);
	}

	Phase ph_unpack("  unpack + islands");
	CompoundInfo info;
	const auto mass_of = [](const surtr_fragment& r) {
		MassProperties mp;
		mp.Volume = r.volume;
		mp.Centroid = Vector3(r.centroid[0], r.centroid[1], r.centroid[2]);
		std::copy(r.inertia, r.inertia + 6, mp.Inertia);
		mp.FaceCount = r.n_faces;
		return mp;
	};
	// untouched pieces first, bound together in the reserved 0-th set (Surtr.cpp:2116-2126)
	info.CompoundBind.push_back(std::set<int>());
	for (uint32_t k = 0; k < n_out; k++)
	{
		info.CompoundBind[0].insert((int)info.PieceVec.size());
		info.PieceVec.push_back(targetPieceVec[outside[k]]);
		info.PieceMass.push_back(MassProperties());
		info.PieceSourceCell.push_back(-1);
		info.PieceSourcePiece.push_back(outside[k]);
	}
	for (const surtr_fragment& r : ofr.rec)
		info.PieceMass[r.piece] = mass_of(r);
	(void)n_cells;
	// match every convex fragment with its mesh fragment (both lists are cell-major, piece-minor) ...
	const auto key = [](const surtr_fragment& r) { return ((uint64_t)r.cell << 32) | r.piece; };
	std::vector<std::pair<size_t, size_t>> matched;   // (convex fragment, mesh fragment)
	size_t m = 0;
	for (size_t f = 0; f < fr.rec.size(); f++)
	{
		if (!meshBranch)
		{
			matched.emplace_back(f, f);
			continue;
		}
		while (m < mfr.rec.size() && key(mfr.rec[m]) < key(fr.rec[f]))
			m++;
		if (m < mfr.rec.size() && key(mfr.rec[m]) == key(fr.rec[f]))   // otherwise: mesh clipped away (Surtr.cpp:1471)
			matched.emplace_back(f, m);
	}
	// ... build the pieces of every pair on the worker pool (AoS conversion + island split, Surtr.cpp:1474-1500) ...
	std::vector<std::vector<Piece*>> built(matched.size());
	detail::parallel_for(matched.size(), [&](size_t k) {
		Poly::Polyhedron convex = convexPoly.empty() ? fr.polyhedron(matched[k].first) : std::move(convexPoly[matched[k].first]);
		if (!meshBranch)
		{
			Poly::Polyhedron same = convex;
			built[k].push_back(new Piece(std::move(convex), std::move(same)));
			return;
		}
		Poly::Polyhedron mesh = mfr.polyhedron(matched[k].second);
		const std::vector<std::vector<int>> groupVec = mesh_islands(mesh);   // = CheckMeshIsland(mesh), as sorted lists
		if (groupVec.size() >= 2)
		{
			std::vector<int> mapping(mesh.size(), -1);
			for (const std::vector<int>& group : groupVec)
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
				built[k].push_back(new Piece(convex, std::move(island)));
			}
		}
		else
			built[k].push_back(new Piece(std::move(convex), std::move(mesh)));
	});
	// ... and bind them in order: one set per cell that produced pieces, cell order (Surtr.cpp:2133-2146)
	int current_cell = -1;
	for (size_t k = 0; k < matched.size(); k++)
	{
		const surtr_fragment& r = fr.rec[matched[k].first];
		for (Piece* piece : built[k])
		{
			if ((int)r.cell != current_cell)
			{
				info.CompoundBind.push_back(std::set<int>());
				current_cell = (int)r.cell;
			}
			info.CompoundBind.back().insert((int)info.PieceVec.size());
			info.PieceVec.push_back(piece);
			info.PieceMass.push_back(mass_of(r));
			info.PieceSourceCell.push_back((int)r.cell);
			info.PieceSourcePiece.push_back(inside[r.piece]);
		}
	}
	return info;
}
} // namespace

void SetExtract(CompoundInfo& preResult)
{
	for (Extract* e : preResult.PieceExtractedConvex)
		delete e;   // the reference leaks the previous lists (Surtr.cpp:2151-2155 overwrites the pointers)
	preResult.PieceExtractedConvex.assign(preResult.PieceVec.size(), nullptr);
	detail::parallel_for(preResult.PieceVec.size(), [&](size_t i) { preResult.PieceExtractedConvex[i] = Poly::ExtractFaces(preResult.PieceVec[i]->Convex); });
}

void MergeOutOfImpact(CompoundInfo& compoundInfo, const std::vector<Vector3>& spherePointCloud, const FractureArgs& args)
{
	// Surtr.cpp:2368-2403; the 0-th set is skipped, emptied sets are dropped
	for (size_t i = 1; i < compoundInfo.CompoundBind.size(); i++)
	{
		std::set<int>& local = compoundInfo.CompoundBind[i];
		std::set<int> outside;
		for (const int c : local)
			if (ConvexOutOfSphere(compoundInfo.PieceVec[c]->Convex, compoundInfo.PieceExtractedConvex[c], spherePointCloud,
								  args.ImpactPosition, args.ImpactRadius))
				outside.insert(c);
		for (const int c : outside)
		{
			local.erase(c);
			compoundInfo.CompoundBind[0].insert(c);
		}
	}
	compoundInfo.CompoundBind.erase(std::remove_if(std::next(compoundInfo.CompoundBind.begin()), compoundInfo.CompoundBind.end(),
												   [](const std::set<int>& local) { return local.empty(); }),
									compoundInfo.CompoundBind.end());
}

void HandleConvexIsland(CompoundInfo& compoundInfo)
{
	// Surtr.cpp:2203-2366.  Two pieces of a bind set are neighbours when they own a pair of faces with (almost) the same
	// plane offset, opposite normals and overlapping outlines; every bind set is split into its connected groups.
	struct FaceNode
	{
		int CID;
		double AbsD;
		Vector3 Normal;   // FacePlane.Normal(), normalised once (the reference renormalises per pair: same value)
		std::vector<Vector3> FacePoints;
	};
	const auto any_point_inside = [](const FaceNode& of, const FaceNode& in) {
		const int n = (int)in.FacePoints.size();
		for (const Vector3& p : of.FacePoints)
		{
			bool included = true;
			for (int v = 0; v < n && included; v++)
				included = VMACH::OnYourRight(in.FacePoints[v], in.FacePoints[(v + 1) % n], p, in.Normal);
			if (included)
				return true;
		}
		return false;
	};

	// every bind set is split on its own: one pool task per set, applied in set order afterwards
	const size_t nBind = compoundInfo.CompoundBind.size();
	std::vector<std::vector<std::set<int>>> splitOf(nBind);
	detail::parallel_for(nBind, [&](size_t iBind) {
		const std::set<int>& localBind = compoundInfo.CompoundBind[iBind];
		if (localBind.size() <= 1)
			return;
		std::vector<FaceNode> nodes;
		for (const int cid : localBind)
			for (const std::vector<int>& poly : *compoundInfo.PieceExtractedConvex[cid])
			{
				FaceNode node;
				node.CID = cid;
				for (const int v : poly)
					node.FacePoints.push_back(compoundInfo.PieceVec[cid]->Convex[v].Position);
				const DirectX::SimpleMath::Plane p(node.FacePoints[0], node.FacePoints[1], node.FacePoints[2]);
				node.AbsD = std::abs(p.D());
				node.Normal = p.Normal();
				node.Normal.Normalize();
				nodes.push_back(std::move(node));
			}
		// sorted by offset, so that the candidates of a face are one contiguous window instead of every later face
		// (the reference's early exit at :2237 never fires on a sorted list, its scan is quadratic)
		std::sort(nodes.begin(), nodes.end(), [](const FaceNode& a, const FaceNode& b) { return a.AbsD < b.AbsD; });
		std::map<int, std::set<int>> nei;
		for (size_t i = 0; i + 1 < nodes.size(); i++)
			for (size_t j = i + 1; j < nodes.size(); j++)
			{
				if (std::abs(nodes[i].AbsD - nodes[j].AbsD) > 1e-3)
					break;   // later faces are farther still
				if (!(std::abs(1 + nodes[i].Normal.Dot(nodes[j].Normal)) < 1e-4))
					continue;
				if (any_point_inside(nodes[i], nodes[j]) || any_point_inside(nodes[j], nodes[i]))
				{
					nei[nodes[i].CID].insert(nodes[j].CID);
					nei[nodes[j].CID].insert(nodes[i].CID);
				}
			}
		// flood fill (:2321-2350): groups come out by their lowest remaining piece id
		std::set<int> remain(localBind.begin(), localBind.end());
		std::vector<std::set<int>> splitGroup;
		while (!remain.empty())
		{
			std::set<int> split;
			std::vector<int> stack{ *remain.begin() };
			while (!stack.empty())
			{
				const int curr = stack.back();
				stack.pop_back();
				if (remain.erase(curr))
				{
					split.insert(curr);
					for (const int iAdj : nei[curr])
						stack.push_back(iAdj);
				}
			}
			splitGroup.push_back(std::move(split));
		}
		splitOf[iBind] = std::move(splitGroup);
	});
	std::vector<std::set<int>> newBind;
	for (size_t iBind = 0; iBind < nBind; iBind++)
		if (splitOf[iBind].size() >= 2)
		{
			compoundInfo.CompoundBind[iBind] = splitOf[iBind][0];
			newBind.insert(newBind.end(), std::next(splitOf[iBind].begin()), splitOf[iBind].end());
		}
	compoundInfo.CompoundBind.insert(compoundInfo.CompoundBind.end(), newBind.begin(), newBind.end());
}

const detail::FlatPattern& FractureStorage::Resident(bool partial) const
{
	detail::FlatPattern& flat = m_flat[partial ? 0 : 1];
	if (!flat.id)
		flat.build(partial ? PartialFracturePattern : GeneralFracturePattern);
	return flat;
}

void FractureStorage::PatternsChanged()
{
	m_flat[0] = detail::FlatPattern();
	m_flat[1] = detail::FlatPattern();
}

std::vector<VMACH::Polygon3D> GenerateFracturePattern(int seed, int cellCount, double mean)
{
	return GenerateVoronoi(GenerateRadialSeeds(seed, cellCount, mean));
}

std::vector<Compound> DoFracture(const Compound& targetCompound, const FractureStorage& storage, const std::vector<Vector3>& spherePointCloud,
								 const FractureArgs& args, CompoundInfo* out_info)
{
	// Surtr.cpp:1885-1959
	// The pattern stays on the device in its own frame; scale (:1889-1891) and alignment (:1893-1896) happen there,
	// together with the re-derivation of every face plane that Polygon3D::Scale / Translate do on the host.
	const detail::FlatPattern& pattern = storage.Resident(args.PartialFracture);
	std::vector<Vector3> localSpherePointCloud = spherePointCloud;
	for (Vector3& v : localSpherePointCloud)
	{
		v *= args.ImpactRadius;
		v += args.ImpactPosition;
	}
	CellSource source;
	source.pattern = &pattern;
	source.scale = Vector3(storage.MaxAxisScale, storage.MaxAxisScale, storage.MaxAxisScale) * 2;
	source.translate = args.ImpactPosition;
	CompoundInfo second;
	{
		Phase ph("ApplyFracture");
		second = apply_fracture(targetCompound, source, localSpherePointCloud, args.PartialFracture, args, true);
		SetExtract(second);
	}
	if (args.PartialFracture)
	{
		Phase ph("MergeOutOfImpact");
		MergeOutOfImpact(second, localSpherePointCloud, args);
	}
	{
		// The refit's inputs (every piece's Mesh and Convex) are final after the merge, and HandleConvexIsland only regroups
		// CompoundBind from the PRE-refit convexes (the reference's order, Surtr.cpp:1931-1944): the refit's GPU event is
		// launched first and runs while the host does the grouping; its results replace the convexes afterwards.
		Phase ph("HandleConvexIsland + Refitting");
		const bool refit = refit_begin(second.PieceVec, args);
		{
			Phase ph2("  HandleConvexIsland");
			HandleConvexIsland(second);
		}
		if (refit)
			refit_end(second.PieceVec, &second.PieceMass);
		SetExtract(second);
	}

	std::vector<Compound> result;
	for (const std::set<int>& iComp : second.CompoundBind)
	{
		Compound c;
		for (const int iPiece : iComp)
		{
			c.PieceVec.push_back(second.PieceVec[iPiece]);
			c.PieceExtractedConvex.push_back(second.PieceExtractedConvex[iPiece]);
		}
		result.push_back(std::move(c));
	}
	if (out_info)
		*out_info = second;
	return result;
}

MassProperties CombineMass(const std::vector<MassProperties>& pieces, float density)
{
	// what PxRigidBodyExt::updateMassAndInertia(body, density) (Surtr.cpp:2520) derives from the compound's shapes:
	// total mass, centre of mass, and the inertia tensor about it by the parallel-axis theorem.  `Volume` of the result
	// holds the MASS; `Inertia` is scaled by the density.
	MassProperties total;
	double mass = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
	for (const MassProperties& p : pieces)
	{
		const double m = density * p.Volume;
		mass += m;
		cx += m * p.Centroid.x; cy += m * p.Centroid.y; cz += m * p.Centroid.z;
	}
	if (mass == 0.0)
		return total;
	cx /= mass; cy /= mass; cz /= mass;
	double I[6] = { 0, 0, 0, 0, 0, 0 };
	for (const MassProperties& p : pieces)
	{
		const double m = density * p.Volume;
		const double dx = p.Centroid.x - cx, dy = p.Centroid.y - cy, dz = p.Centroid.z - cz;
		I[0] += density * p.Inertia[0] + m * (dy * dy + dz * dz);
		I[1] += density * p.Inertia[1] + m * (dx * dx + dz * dz);
		I[2] += density * p.Inertia[2] + m * (dx * dx + dy * dy);
		I[3] += density * p.Inertia[3] - m * dx * dy;
		I[4] += density * p.Inertia[4] - m * dx * dz;
		I[5] += density * p.Inertia[5] - m * dy * dz;
		total.FaceCount += p.FaceCount;
	}
	total.Volume = mass;
	total.Centroid = Vector3((float)cx, (float)cy, (float)cz);
	for (int k = 0; k < 6; k++)
		total.Inertia[k] = (float)I[k];
	return total;
}

namespace
{
// Refitting in two halves, so that a caller can do host work (HandleConvexIsland) while the GPU clips:
// refit_begin = ICH normals, k-DOP extents, upload + launch of the clip event; refit_end = wait, download, unpack.
bool refit_begin(const std::vector<Piece*>& targetPieceVec, const FractureArgs& args)
{
	const uint32_t n = (uint32_t)targetPieceVec.size();
	if (!n)
		return false;
	// 1. ICH normals per piece (host; <= 4 points by default: the seed tetrahedron)
	Phase* ph = new Phase("  refit: ICH normals");
	std::vector<std::vector<Vector3>> piece_normals(n);
	detail::parallel_for(n, [&](size_t i) {   // one task per piece, like the reference's pool (Surtr.cpp:2405-2413)
		std::vector<Vector3> pts;
		pts.reserve(targetPieceVec[i]->Mesh.size());
		for (const Poly::Vertex& v : targetPieceVec[i]->Mesh)
			pts.push_back(v.Position);
		piece_normals[i] = VMACH::GenerateICHNormal(pts, std::min((int)pts.size(), args.RefittingPointLimit));
	});
	std::vector<float> verts4, normals3;
	std::vector<uint32_t> vert_off{ 0 }, normal_off{ 0 };
	for (uint32_t i = 0; i < n; i++)
	{
		for (const Poly::Vertex& v : targetPieceVec[i]->Mesh)
			verts4.insert(verts4.end(), { v.Position.x, v.Position.y, v.Position.z, 0.f });
		vert_off.push_back((uint32_t)(verts4.size() / 4));
		for (const Vector3& nv : piece_normals[i])
			normals3.insert(normals3.end(), { nv.x, nv.y, nv.z });
		normal_off.push_back((uint32_t)(normals3.size() / 3));
	}
	delete ph;
	ph = new Phase("  refit: k-DOP batch");
	// 2. k-DOP extents of every piece->Mesh on the GPU (Kdop::Calc(Polyhedron), Kdop.cpp:92-115)
	const size_t nn = normals3.size() / 3;
	std::vector<float> dist(2 * nn), planes8(8 * nn);
	std::vector<int32_t> arg(2 * nn);
	detail::check(surtr_kdop_calc_batch(detail::context(), verts4.data(), vert_off.data(), n, normals3.data(), normal_off.data(),
										dist.data(), arg.data(), planes8.data()), "surtr_kdop_calc_batch");
	delete ph;
	Phase ph3("  refit: pack + launch");
	// 3. clip piece->Convex by its own plane list: n independent events of one piece x one cell
	detail::FlatPolys pieces;
	detail::FlatCells cells;
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
		pieces.ev_off.push_back(i);
	}
	pieces.ev_off.push_back(n);
	cells.ev_off = pieces.ev_off;
	detail::begin_event(pieces, cells, true, 0);   // (pageable uploads are staged before the call returns)
	return true;
}

void refit_end(std::vector<Piece*>& targetPieceVec, std::vector<MassProperties>* mass)
{
	Phase ph("  refit: wait + unpack");
	const uint32_t n = (uint32_t)targetPieceVec.size();
	detail::Fragments fr;
	detail::end_event(fr, true, 0);
	for (Piece* p : targetPieceVec)
		p->Convex.clear();
	if (mass)
		mass->assign(n, MassProperties());   // a convex that the refit clips away has no mass left
	detail::parallel_for(fr.rec.size(), [&](size_t f) { targetPieceVec[fr.rec[f].piece]->Convex = fr.polyhedron(f); });   // one fragment per piece
	for (size_t f = 0; f < fr.rec.size(); f++)
	{
		const surtr_fragment& r = fr.rec[f];
		if (mass)
		{
			MassProperties& mp = (*mass)[r.piece];
			mp.Volume = r.volume;
			mp.Centroid = Vector3(r.centroid[0], r.centroid[1], r.centroid[2]);
			std::copy(r.inertia, r.inertia + 6, mp.Inertia);
			mp.FaceCount = r.n_faces;
		}
	}
}
} // namespace

void Refitting(std::vector<Piece*>& targetPieceVec, const FractureArgs& args, std::vector<MassProperties>* mass)
{
	if (refit_begin(targetPieceVec, args))
		refit_end(targetPieceVec, mass);
}

} // namespace SurtrHost
