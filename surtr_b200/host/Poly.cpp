#include "Poly.h"

#include "Engine.h"
#include "VMACH.h"

#include <algorithm>
#include <stdexcept>

using SurtrHost::detail::FlatCells;
using SurtrHost::detail::FlatPolys;
using SurtrHost::detail::Fragments;

namespace Poly
{
void InitPolyhedron(Polyhedron& polyhedron, const std::vector<Vector3>& positionVec, const std::vector<std::vector<int>>& neighborVec)
{
	polyhedron.resize(positionVec.size());
	for (size_t i = 0; i < positionVec.size(); i++)
	{
		polyhedron[i].Position = positionVec[i];
		polyhedron[i].NeighborVertexVec = neighborVec[i];
	}
}

std::vector<std::vector<int>> ExtractNeighborFromMesh(const std::vector<Vector3>& vertices, const std::vector<int>& indices)
{
	// The reference builds per-triangle adjacency lists and walks each vertex's triangle fan through them
	// (Poly.cpp:131-212).  Same walk here over a CSR of incident triangles: the fan's next triangle is the first one,
	// in the order "edges (0,1), (1,2), (2,0) of the current triangle, ascending triangle index across an edge", that
	// touches the vertex and is not yet part of the fan -- which is what candidate[0] of Poly.cpp:183-210 selects.
	const int n_tri = (int)indices.size() / 3;
	const int n_vert = (int)vertices.size();
	std::vector<int> inc_off(n_vert + 1, 0);
	for (int i = 0; i < 3 * n_tri; i++)
		inc_off[indices[i] + 1]++;
	for (int v = 0; v < n_vert; v++)
		inc_off[v + 1] += inc_off[v];
	std::vector<int> inc(inc_off[n_vert]), fill(inc_off.begin(), inc_off.end() - 1);
	for (int t = 0; t < n_tri; t++)   // ascending triangle index per vertex, repeats kept (as push_back does at :145-147)
		for (int k = 0; k < 3; k++)
			inc[fill[indices[3 * t + k]]++] = t;
	const auto touches = [&](const int t, const int v) { return indices[3 * t] == v || indices[3 * t + 1] == v || indices[3 * t + 2] == v; };

	// per-triangle adjacency, duplicates removed in first-seen order (:150-170); triangles are independent -> worker pool
	constexpr int CHUNK = 256;
	std::vector<std::vector<int>> adj(n_tri);
	SurtrHost::detail::parallel_for((size_t)(n_tri + CHUNK - 1) / CHUNK, [&](size_t chunk) {
		for (int curr = (int)chunk * CHUNK; curr < std::min(n_tri, ((int)chunk + 1) * CHUNK); curr++)
		{
			std::vector<int>& out = adj[curr];
			for (int e = 0; e < 3; e++)
			{
				const int a = indices[3 * curr + e], b = indices[3 * curr + (e + 1) % 3];
				const int* pa = &inc[inc_off[a]], * ea = &inc[inc_off[a + 1]];
				const int* pb = &inc[inc_off[b]], * eb = &inc[inc_off[b + 1]];
				while (pa != ea && pb != eb)   // std::set_intersection of two ascending lists (:156-158)
				{
					if (*pa < *pb) ++pa;
					else if (*pb < *pa) ++pb;
					else
					{
						const int t = *pa;
						++pa; ++pb;
						if (t != curr && out.end() == std::find(out.begin(), out.end(), t))
							out.push_back(t);
					}
				}
			}
		}
	});

	// one fan walk per vertex (:172-251), independent of each other -> worker pool
	std::vector<std::vector<int>> nei(n_vert);
	SurtrHost::detail::parallel_for((size_t)(n_vert + CHUNK - 1) / CHUNK, [&](size_t chunk) {
		std::vector<int> fan;
		for (int iVert = (int)chunk * CHUNK; iVert < std::min(n_vert, ((int)chunk + 1) * CHUNK); iVert++)
		{
			if (inc_off[iVert] == inc_off[iVert + 1])
				continue;   // a vertex no triangle uses keeps an empty ring
			fan.assign(1, inc[inc_off[iVert]]);
			for (int curr = fan[0];;)
			{
				int next = -1, n_candidates = 0;
				for (const int t : adj[curr])
					if (fan.end() == std::find(fan.begin(), fan.end(), t) && touches(t, iVert))
					{
						if (n_candidates++ == 0)
							next = t;
					}
				if (n_candidates == 0)
					break;
				if (n_candidates > 2)
					throw std::runtime_error("ExtractNeighborFromMesh: non-manifold fan (the reference does not terminate here)");
				fan.push_back(next);
				curr = next;
			}

			std::vector<int> collection;
			for (const int t : fan)
			{
				int start = 0;
				for (int k = 0; k < 3; k++)
					if (indices[3 * t + k] == iVert) { start = k; break; }
				collection.push_back(indices[3 * t + (start + 1) % 3]);
				collection.push_back(indices[3 * t + (start + 2) % 3]);
			}
			if (collection.size() >= 3)
			{
				const bool isCCW = collection[1] != collection[2];   // :236
				if (isCCW)
					for (size_t i = 0; i + 1 < collection.size(); i += 2)
						std::swap(collection[i], collection[i + 1]);
				std::vector<int> unique;
				for (const int v : collection)
					if (unique.end() == std::find(unique.begin(), unique.end(), v))
						unique.push_back(v);
				if (isCCW)
					std::reverse(unique.begin(), unique.end());
				collection.swap(unique);
			}
			nei[iVert].swap(collection);
		}
	});
	for (int v = 0; v < n_vert; v++)   // :253-260
		for (const int a : nei[v])
			if (nei[a].end() == std::find(nei[a].begin(), nei[a].end(), v))
				throw std::runtime_error("ExtractNeighborFromMesh: asymmetric adjacency");
	return nei;
}

void Moments(double& zerothMoment, Vector3& firstMoment, const Polyhedron& polyhedron)
{
	zerothMoment = 0.0;
	firstMoment = Vector3(0.0, 0.0, 0.0);
	if (polyhedron.size() <= 3)   // Poly.cpp:61
		return;
	FlatPolys pieces;
	pieces.add(polyhedron);
	FlatCells cells;
	cells.add(std::vector<Plane>());   // an empty plane list keeps the piece whole; K4 still integrates it
	Fragments fr;
	SurtrHost::detail::run_event(pieces, cells, fr, false);
	if (fr.rec.size() == 1)
	{
		zerothMoment = fr.rec[0].volume;
		firstMoment = Vector3(fr.rec[0].centroid[0], fr.rec[0].centroid[1], fr.rec[0].centroid[2]);
	}
}

Extract* ExtractFaces(const Polyhedron& polyhedron)
{
	// One loop per face, started at the lowest-numbered vertex's first unvisited outgoing edge and following
	// "the ring entry before the one we came from" -- the traversal rule of the reference (Poly.cpp:34-41, 94-122).
	Extract* faces = new Extract();
	const int nv = (int)polyhedron.size();
	std::vector<std::vector<char>> seen(nv);
	for (int v = 0; v < nv; v++)
		seen[v].assign(polyhedron[v].NeighborVertexVec.size(), 0);
	auto slot_of = [&](int from, int to) {
		const std::vector<int>& r = polyhedron[from].NeighborVertexVec;
		return (int)(std::find(r.begin(), r.end(), to) - r.begin());
	};
	for (int v = 0; v < nv; v++)
	{
		if (polyhedron[v].comp < 0)
			continue;
		for (size_t s = 0; s < polyhedron[v].NeighborVertexVec.size(); s++)
		{
			if (seen[v][s])
				continue;
			std::vector<int> loop(1, v);
			int prev = v, cur = polyhedron[v].NeighborVertexVec[s];
			seen[v][s] = 1;
			size_t guard = 0;
			while (cur != v && guard++ < (size_t)nv * 64)
			{
				loop.push_back(cur);
				const std::vector<int>& r = polyhedron[cur].NeighborVertexVec;
				const int k = slot_of(cur, prev);
				const int next = r[(k == 0 || k >= (int)r.size()) ? r.size() - 1 : k - 1];
				const int ks = slot_of(cur, next);
				if (ks < (int)seen[cur].size())
					seen[cur][ks] = 1;
				prev = cur;
				cur = next;
			}
			faces->push_back(loop);
		}
	}
	return faces;
}

void ClipPolyhedron(Polyhedron& polyhedron, const std::vector<Plane>& planes)
{
	if (polyhedron.empty())
		return;
	FlatPolys pieces;
	pieces.add(polyhedron);
	FlatCells cells;
	cells.add(planes);
	Fragments fr;
	SurtrHost::detail::run_event(pieces, cells, fr);
	if (fr.rec.empty())
		polyhedron.clear();
	else
		polyhedron = fr.polyhedron(0);
}

Polyhedron ClipPolyhedron(const Polyhedron& polyhedron, const VMACH::Polygon3D& polygon3D)
{
	std::vector<Plane> planes;
	for (const auto& f : polygon3D.FaceVec)
		planes.push_back(f.FacePlane);
	Polyhedron res = polyhedron;
	ClipPolyhedron(res, planes);
	return res;
}

void Translate(Polyhedron& polyhedron, const Vector3& v)
{
	for (auto& i : polyhedron)
		i.Position += v;
}

void Scale(Polyhedron& polyhedron, const Vector3& v)
{
	for (auto& i : polyhedron)
		i.Position *= v;
}

void Transform(Polyhedron& polyhedron, const DirectX::XMMATRIX& matrix)
{
	const DirectX::XMMATRIX mat = DirectX::XMMatrixTranspose(matrix);
	for (Vertex& vert : polyhedron)
		vert.Position = DirectX::XMVector3TransformCoord(vert.Position, mat);
}

Polyhedron GetBB()
{
	static const float pts[8][3] = { { -0.5f, -0.5f, -0.5f }, { +0.5f, -0.5f, -0.5f }, { +0.5f, +0.5f, -0.5f }, { -0.5f, +0.5f, -0.5f },
									 { -0.5f, -0.5f, +0.5f }, { +0.5f, -0.5f, +0.5f }, { +0.5f, +0.5f, +0.5f }, { -0.5f, +0.5f, +0.5f } };
	static const int nb[8][3] = { { 1, 4, 3 }, { 5, 0, 2 }, { 3, 6, 1 }, { 7, 2, 0 }, { 5, 7, 0 }, { 1, 6, 4 }, { 5, 2, 7 }, { 4, 6, 3 } };
	Polyhedron poly(8);
	for (int i = 0; i < 8; i++)
	{
		poly[i].Position = Vector3(pts[i][0], pts[i][1], pts[i][2]);
		poly[i].NeighborVertexVec.assign(nb[i], nb[i] + 3);
	}
	return poly;
}

int ComparePlanePoint(const Plane& plane, const Vector3& point)
{
	const float s = plane.D() + plane.Normal().Dot(point);
	if (std::abs(s) < 1.0e-10)
		return 0;
	return s < 0.f ? 1 : (s > 0.f ? -1 : 0);
}

int ComparePlaneBB(const Plane& plane, const double xmin, const double ymin, const double zmin, const double xmax, const double ymax, const double zmax)
{
	int cmin = 2, cmax = -2;
	for (int k = 0; k < 8; k++)
	{
		const int c = ComparePlanePoint(plane, Vector3((k & 1) ? xmax : xmin, (k & 2) ? ymax : ymin, (k & 4) ? zmax : zmin));
		cmin = std::min(cmin, c);
		cmax = std::max(cmax, c);
	}
	return cmin >= 0 ? 1 : (cmax <= 0 ? -1 : 0);
}

Vector3 PlaneLineIntersection(const Vector3& a, const Vector3& b, const Plane& plane)
{
	const float sa = plane.D() + plane.Normal().Dot(a);
	const float sb = plane.D() + plane.Normal().Dot(b);
	return ((a * sb) - (b * sa)) / (sb - sa);
}
} // namespace Poly
