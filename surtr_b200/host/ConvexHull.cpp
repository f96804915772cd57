#include <cstring>
#include "ConvexHull.h"

#include "Engine.h"

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <map>
#include <string>
#include <tuple>

namespace VMACH
{
float ConvexHullFace::CalcArea()
{
	const Vector3 d1 = Vertices[1] - Vertices[0];
	const Vector3 d2 = Vertices[2] - Vertices[0];
	return 0.5f * d1.Cross(d2).Length();
}

bool ConvexHull::Colinear(const ConvexHullVertex& p1, const ConvexHullVertex& p2, const ConvexHullVertex& p3)
{
	return ((p3.z - p1.z) * (p2.y - p1.y) - (p2.z - p1.z) * (p3.y - p1.y)) == 0 &&
		   ((p2.z - p1.z) * (p3.x - p1.x) - (p2.x - p1.x) * (p3.z - p1.z)) == 0 &&
		   ((p2.x - p1.x) * (p3.y - p1.y) - (p2.y - p1.y) * (p3.x - p1.x)) == 0;
}

// Signed volume of the tetrahedron (triangle, point) in float, the reference's term order (VMACH.cpp:919-938):
// negative = the point sees the triangle from outside.
static inline float SignedVolume(const Vector3& a, const Vector3& b, const Vector3& c, const Vector3& p)
{
	const float ax = a.x - p.x, ay = a.y - p.y, az = a.z - p.z;
	const float bx = b.x - p.x, by = b.y - p.y, bz = b.z - p.z;
	const float cx = c.x - p.x, cy = c.y - p.y, cz = c.z - p.z;
	return ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
}

float ConvexHull::Volume(const ConvexHullFace& face, const ConvexHullVertex& point)
{
	return SignedVolume(face.Vertices[0], face.Vertices[1], face.Vertices[2], point);
}

float ConvexHull::TetVolume(const Tri& t, int p) const { return SignedVolume(m_pos[t.v[0]], m_pos[t.v[1]], m_pos[t.v[2]], m_pos[p]); }

ConvexHull::ConvexHull(const std::vector<ConvexHullVertex>& pointCloud, uint32_t limitCnt)
{
	m_pos.assign(pointCloud.begin(), pointCloud.end());
	Build(limitCnt);
}

ConvexHull::ConvexHull(const std::vector<Vector3>& pointCloud, uint32_t limitCnt) : m_pos(pointCloud) { Build(limitCnt); }

// Key of the edge table.  Reference: hash(print(p1)) ^ hash(print(p2)) -- symmetric, equal for end points that print
// alike, and 0 for ANY edge whose two end points print alike.  Same partition with two class ids.
// Print class of a point, computed when the point first becomes an end point of a hull edge (only the few points that
// join the hull ever need one: formatting all 2503 bunny vertices was two thirds of the hull's time).
uint32_t ConvexHull::PrintClass(int i) const
{
	if (m_printClass[i] == UNCLASSED)
	{
		char buf[192];
		std::snprintf(buf, sizeof buf, "%f%f%f", m_pos[i].x, m_pos[i].y, m_pos[i].z);   // std::to_string(float) prints "%f"
		m_printClass[i] = m_byPrint.emplace(buf, (uint32_t)m_byPrint.size()).first->second;
	}
	return m_printClass[i];
}

uint64_t ConvexHull::EdgeKey(int a, int b) const
{
	const uint32_t ca = PrintClass(a), cb = PrintClass(b);
	if (ca == cb)
		return ~0ull;
	return ((uint64_t)std::min(ca, cb) << 32) | std::max(ca, cb);
}

void ConvexHull::AttachEdge(int a, int b, int tri)
{
	const uint64_t key = EdgeKey(a, b);
	auto it = m_rimOf.find(key);
	if (it == m_rimOf.end())
	{
		m_rims.push_back(Rim{ a, b });
		it = m_rimOf.emplace(key, (int)m_rims.size() - 1).first;
	}
	Rim& r = m_rims[it->second];
	if (r.f1 >= 0 && r.f2 >= 0)
		return;                          // both sides taken: the triangle stays unlinked, as in the reference
	(r.f1 < 0 ? r.f1 : r.f2) = tri;
}

// New triangle (a, b, c), wound so that `inner` lies on its positive side; its three edges join the edge table.
void ConvexHull::AddTriangle(int a, int b, int c, int inner)
{
	Tri t;
	t.v[0] = a; t.v[1] = b; t.v[2] = c;
	if (TetVolume(t, inner) < 0)
		std::swap(t.v[0], t.v[2]);
	m_tris.push_back(t);
	const int id = (int)m_tris.size() - 1;
	m_fresh.push_back(id);
	AttachEdge(a, b, id);
	AttachEdge(a, c, id);
	AttachEdge(b, c, id);
}

// The point p joins the hull: triangles that see it are marked, every horizon edge (one marked, one unmarked triangle)
// gets a new triangle to p.  Edges appended by those new triangles are visited by the same pass (their second side is
// still open, so they are skipped), which is the reference's iteration over a list it appends to.
void ConvexHull::Absorb(int p)
{
	for (int t = 0; t < (int)m_tris.size(); t++)
		if (m_tris[t].alive && TetVolume(m_tris[t], p) < 0)
		{
			m_tris[t].visible = true;
			m_lit.push_back(t);
		}
	if (m_lit.empty())
		return;
	for (size_t e = 0; e < m_rims.size(); e++)
	{
		if (!m_rims[e].alive || m_rims[e].f1 < 0 || m_rims[e].f2 < 0)
			continue;
		const bool v1 = m_tris[m_rims[e].f1].visible, v2 = m_tris[m_rims[e].f2].visible;
		if (v1 && v2)
		{
			m_rims[e].remove = true;
			continue;
		}
		if (!v1 && !v2)
			continue;
		if (v1)
			std::swap(m_rims[e].f1, m_rims[e].f2);       // f1 = the hidden side, f2 = the side p sees
		const int a = m_rims[e].a, b = m_rims[e].b, seen = m_rims[e].f2;
		// orientation probe: the marked triangle's corner off the edge, compared by VALUE like the reference's operator==
		int inner = m_tris[seen].v[0];
		for (int i = 0; i < 3; i++)
		{
			const int v = m_tris[seen].v[i];
			if (m_pos[v] == m_pos[a] || m_pos[v] == m_pos[b])   // (exact coordinates: the reference's operator==)
				continue;
			inner = v;
			break;
		}
		// the marked side lets go of the edge (first matching slot, as ConvexHullEdge::EraseFace)
		if (m_rims[e].f1 == seen) m_rims[e].f1 = -1; else m_rims[e].f2 = -1;
		AddTriangle(a, b, p, inner);                     // (may grow m_rims: index-based access only)
	}
}

// End of a step: edges between two marked triangles and the marked triangles themselves leave the hull.
void ConvexHull::Sweep()
{
	m_lit.clear();
	m_fresh.clear();
	for (Rim& r : m_rims)
		if (r.alive && r.remove)
		{
			m_rimOf.erase(EdgeKey(r.a, r.b));
			r.alive = false;
		}
	for (Tri& t : m_tris)
		if (t.alive && t.visible)
			t.alive = false;
}

// First tetrahedron: the point of largest x, the point farthest from it, the point spanning the largest triangle with
// those two, the point of largest signed volume over that triangle -- first maximum wins in each scan (std::max_element).
bool ConvexHull::SeedTetrahedron()
{
	const int n = (int)m_pos.size();
	if (n <= 3)
		return false;
	int i1 = 0;
	for (int i = 1; i < n; i++)
		if (m_pos[i1].x < m_pos[i].x) i1 = i;
	auto far = [&](int i) {
		const double dx = m_pos[i].x - m_pos[i1].x, dy = m_pos[i].y - m_pos[i1].y, dz = m_pos[i].z - m_pos[i1].z;   // float differences, squared in double
		return std::sqrt(dx * dx + dy * dy + dz * dz);
	};
	int i2 = 0;
	for (int i = 1; i < n; i++)
		if (far(i2) < far(i)) i2 = i;
	auto area = [&](int i) { return ConvexHullFace(m_pos[i1], m_pos[i2], m_pos[i]).CalcArea(); };
	int i3 = 0;
	for (int i = 1; i < n; i++)
		if (area(i3) < area(i)) i3 = i;
	auto vol = [&](int i) { return SignedVolume(m_pos[i1], m_pos[i2], m_pos[i3], m_pos[i]); };
	int i4 = 0;
	for (int i = 1; i < n; i++)
		if (vol(i4) < vol(i)) i4 = i;
	m_used[i1] = m_used[i2] = m_used[i3] = m_used[i4] = 1;
	m_usedCnt = 4;
	AddTriangle(i1, i2, i3, i4);
	AddTriangle(i1, i2, i4, i3);
	AddTriangle(i1, i3, i4, i2);
	AddTriangle(i2, i3, i4, i1);
	return true;
}

void ConvexHull::Build(uint32_t limitCnt)
{
	const size_t n = m_pos.size();
	m_used.assign(n, 0);
	m_printClass.assign(n, UNCLASSED);   // six-decimal print classes (the reference's edge keys), filled in on demand
	m_byPrint.clear();
	if (limitCnt != 0 && limitCnt <= 4)
	{
		// A hull limited to four points is its seed tetrahedron (the greedy loop below never runs): the refit of every
		// piece after every fracture (Surtr::Refitting, RefittingPointLimit = 4, Surtr.cpp:2405-2413) takes this path.
		// No outside volumes: the four faces, their winding and their order depend on the four scans of SeedTetrahedron
		// alone.
		m_seedOnly = true;
		SeedTetrahedron();
		return;
	}
	m_outside.assign(n, 0.0f);
	if (!SeedTetrahedron())
		return;
	// a point's outside volume is its own sequential sum, so the points spread over the worker pool in chunks
	constexpr size_t CHUNK = 128;
	const size_t nChunks = (n + CHUNK - 1) / CHUNK;
	auto forUnused = [&](const std::function<void(size_t)>& fn) {
		SurtrHost::detail::parallel_for(nChunks, [&](size_t c) {
			for (size_t i = c * CHUNK; i < std::min(n, (c + 1) * CHUNK); i++)
				if (!m_used[i]) fn(i);
		});
	};
	forUnused([&](size_t i) {
		for (const Tri& t : m_tris)
			if (t.alive) m_outside[i] += std::max(0.0f, TetVolume(t, (int)i));
	});
	// (m_fresh still lists the four seed triangles here: the reference clears its added-face list only at the end of a
	// step, VMACH.cpp:1144-1146, so the first step adds their volumes a second time -- kept, it decides the greedy order)
	const uint32_t limit = limitCnt == 0 ? (uint32_t)n : limitCnt;
	while (m_usedCnt < limit)
	{
		// greedy: the unused point with the largest outside volume (first maximum) joins the hull
		const int k = (int)(std::max_element(m_outside.begin(), m_outside.end()) - m_outside.begin());
		Absorb(k);
		m_used[k] = 1;
		m_outside[k] = -FLT_MAX;
		m_usedCnt++;
		forUnused([&](size_t i) {
			float gone = 0.0f, came = 0.0f;
			for (int t : m_lit) gone += std::max(0.0f, TetVolume(m_tris[t], (int)i));
			for (int t : m_fresh) came += std::max(0.0f, TetVolume(m_tris[t], (int)i));
			m_outside[i] -= gone;
			m_outside[i] += came;
		});
		Sweep();
	}
}

bool ConvexHull::Contains(const ConvexHullVertex& point) const
{
	for (const Tri& t : m_tris)
		if (t.alive && SignedVolume(m_pos[t.v[0]], m_pos[t.v[1]], m_pos[t.v[2]], point) <= 0)
			return false;
	return true;
}

std::vector<std::array<int, 3>> ConvexHull::FaceIndices() const
{
	std::vector<std::array<int, 3>> out;
	for (const Tri& t : m_tris)
		if (t.alive) out.push_back({ t.v[0], t.v[1], t.v[2] });
	return out;
}

const std::list<ConvexHullFace> ConvexHull::GetFaces() const
{
	std::list<ConvexHullFace> out;
	for (const Tri& t : m_tris)
		if (t.alive) out.emplace_back(m_pos[t.v[0]], m_pos[t.v[1]], m_pos[t.v[2]]);
	return out;
}

const std::list<ConvexHullEdge> ConvexHull::GetEdges() const
{
	std::list<ConvexHullEdge> out;
	for (const Rim& r : m_rims)
		if (r.alive) out.emplace_back(m_pos[r.a], m_pos[r.b]);
	return out;
}

std::vector<Vector3> GenerateICHNormal(const std::vector<Vector3>& vertices, int ichIncludePointLimit)
{
	const ConvexHull ich(vertices, (uint32_t)ichIncludePointLimit);
	std::vector<Vector3> normals;
	for (const std::array<int, 3>& f : ich.FaceIndices())
	{
		Vector3 n = (vertices[f[1]] - vertices[f[0]]).Cross(vertices[f[2]] - vertices[f[0]]);
		n.Normalize();
		normals.push_back(n);
	}
	return normals;
}
} // namespace VMACH
