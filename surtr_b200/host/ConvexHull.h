// ConvexHull.h -- host-side mirror of VMACH::ConvexHull (Inc/VMACH.h:88-165, Src/VMACH.cpp:869-1203): the greedy
// incremental convex hull ("ICH") whose face normals seed the k-DOPs (Surtr::GenerateICHNormal, Surtr.cpp:1961-1982).
// Tiny and inherently sequential (<= 20 points at ACH time, <= 4 at refit time): it stays on the host, exactly as
// SURVEY.md section 2 row 4 scopes it.
//
// The public types keep the reference's names (drop-in surface); the hull itself is an index-based structure of its
// own: triangles and edges are records in two append-only arrays that refer to points by INDEX, removal is a flag
// (array order = the reference's list order, which decides the order of the output faces and so of the k-DOP planes),
// and the edge table is keyed by a pair of small integers instead of a hash of printed coordinates.  The reference
// keys an edge by hash(to_string(p1)) ^ hash(to_string(p2)) (VMACH.cpp:940-947): end points that print alike (closer
// than 1e-6) share a key, and an edge whose two end points print alike has key 0 whatever they are.  Those collisions
// change which faces get built on near-degenerate clouds, so they are reproduced -- through a "print class" per point
// (points with the same six-decimal print) -- and only there: distinct classes never collide.
#pragma once

#include "SimpleMath.h"

#include <array>
#include <cstdint>
#include <list>
#include <unordered_map>
#include <map>
#include <string>
#include <vector>

namespace VMACH
{
using DirectX::SimpleMath::Vector3;

struct ConvexHullVertex : public Vector3
{
	bool Processed;
	ConvexHullVertex() : Vector3(), Processed(false) {}
	ConvexHullVertex(float ix, float iy, float iz) : Vector3(ix, iy, iz), Processed(false) {}
	ConvexHullVertex(const Vector3& v3) : Vector3(v3), Processed(false) {}
};

struct ConvexHullFace
{
	bool Visible;
	ConvexHullVertex Vertices[3];
	ConvexHullFace(const ConvexHullVertex& p1, const ConvexHullVertex& p2, const ConvexHullVertex& p3) : Visible(false)
	{
		Vertices[0] = p1; Vertices[1] = p2; Vertices[2] = p3;
	}
	void Rewind() { std::swap(Vertices[0], Vertices[2]); }
	float CalcArea();
};

struct ConvexHullEdge
{
	bool Remove;
	ConvexHullFace* Face1;   // (not populated by GetEdges(): the hull keeps face indices, not pointers)
	ConvexHullFace* Face2;
	ConvexHullVertex EndPoints[2];
	ConvexHullEdge(const ConvexHullVertex& p1, const ConvexHullVertex& p2) : Remove(false), Face1(nullptr), Face2(nullptr)
	{
		EndPoints[0] = p1; EndPoints[1] = p2;
	}
};

class ConvexHull
{
public:
	ConvexHull(const std::vector<ConvexHullVertex>& pointCloud, uint32_t limitCnt);
	ConvexHull(const std::vector<Vector3>& pointCloud, uint32_t limitCnt);

	bool Contains(const ConvexHullVertex& point) const;
	const std::list<ConvexHullFace> GetFaces() const;   // live triangles, creation order
	const std::list<ConvexHullEdge> GetEdges() const;   // live edges, creation order

	static bool Colinear(const ConvexHullVertex& p1, const ConvexHullVertex& p2, const ConvexHullVertex& p3);
	static float Volume(const ConvexHullFace& face, const ConvexHullVertex& point);

	// the live triangles as point indices into the input cloud (creation order): what GenerateICHNormal consumes
	std::vector<std::array<int, 3>> FaceIndices() const;
	const std::vector<Vector3>& Points() const { return m_pos; }

private:
	struct Tri { int v[3]; bool visible = false, alive = true; };
	struct Rim { int a, b; int f1 = -1, f2 = -1; bool remove = false, alive = true; };   // an edge with its (up to) two triangles

	void Build(uint32_t limitCnt);
	bool SeedTetrahedron();
	float TetVolume(const Tri& t, int p) const;
	uint64_t EdgeKey(int a, int b) const;
	void AttachEdge(int a, int b, int tri);
	void AddTriangle(int a, int b, int c, int inner);
	void Absorb(int p);
	void Sweep();

	std::vector<Vector3> m_pos;
	std::vector<char> m_used;
	std::vector<float> m_outside;               // summed outside volume of every unused point against the current hull
	static constexpr uint32_t UNCLASSED = 0xffffffffu;
	mutable std::vector<uint32_t> m_printClass; // points with the same six-decimal print share a class (lazily, hull points only)
	mutable std::map<std::string, uint32_t> m_byPrint;
	uint32_t PrintClass(int i) const;
	std::vector<Tri> m_tris;
	std::vector<Rim> m_rims;
	std::unordered_map<uint64_t, int> m_rimOf;  // edge key -> index into m_rims
	std::vector<int> m_lit, m_fresh;            // triangles the last absorbed point saw / added
	uint32_t m_usedCnt = 0;
	bool m_seedOnly = false;      // limitCnt <= 4: the hull is its seed tetrahedron, nothing else is computed
};

// Surtr::GenerateICHNormal (Surtr.cpp:1961-1982): normalised (v1-v0) x (v2-v0) of every hull face, list order.
std::vector<Vector3> GenerateICHNormal(const std::vector<Vector3>& vertices, int ichIncludePointLimit);
} // namespace VMACH
