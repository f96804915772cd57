#include "ConvexHull.h"

#include "Engine.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <string>

namespace VMACH
{
float ConvexHullFace::CalcArea()
{
	const Vector3 d1 = Vertices[1] - Vertices[0];
	const Vector3 d2 = Vertices[2] - Vertices[0];
	return 0.5f * d1.Cross(d2).Length();
}

void ConvexHullEdge::LinkFace(ConvexHullFace* face)
{
	if (Face1 != nullptr && Face2 != nullptr)
		return;
	(Face1 == nullptr ? Face1 : Face2) = face;
}

void ConvexHullEdge::EraseFace(ConvexHullFace* face)
{
	if (Face1 != face && Face2 != face)
		return;
	(Face1 == face ? Face1 : Face2) = nullptr;
}

ConvexHull::ConvexHull(const std::vector<ConvexHullVertex>& pointCloud, uint32_t limitCnt) : m_limitCnt(limitCnt), m_pointCloud(pointCloud)
{
	m_pointVolume.assign(m_pointCloud.size(), 0.0f);
	CreateConvexHull();
}

ConvexHull::ConvexHull(const std::vector<Vector3>& pointCloud, uint32_t limitCnt) : m_limitCnt(limitCnt)
{
	for (const Vector3& v : pointCloud)
		m_pointCloud.emplace_back(v);
	m_pointVolume.assign(m_pointCloud.size(), 0.0f);
	CreateConvexHull();
}

bool ConvexHull::Contains(const ConvexHullVertex& point) const
{
	for (const ConvexHullFace& f : m_faceList)
		if (Volume(f, point) <= 0)
			return false;
	return true;
}

bool ConvexHull::Colinear(const ConvexHullVertex& p1, const ConvexHullVertex& p2, const ConvexHullVertex& p3)
{
	return ((p3.z - p1.z) * (p2.y - p1.y) - (p2.z - p1.z) * (p3.y - p1.y)) == 0 &&
		   ((p2.z - p1.z) * (p3.x - p1.x) - (p2.x - p1.x) * (p3.z - p1.z)) == 0 &&
		   ((p2.x - p1.x) * (p3.y - p1.y) - (p2.y - p1.y) * (p3.x - p1.x)) == 0;
}

// Signed volume of the tetrahedron (face, point), float arithmetic in the reference's term order (VMACH.cpp:919-938).
float ConvexHull::Volume(const ConvexHullFace& face, const ConvexHullVertex& point)
{
	const float ax = face.Vertices[0].x - point.x, ay = face.Vertices[0].y - point.y, az = face.Vertices[0].z - point.z;
	const float bx = face.Vertices[1].x - point.x, by = face.Vertices[1].y - point.y, bz = face.Vertices[1].z - point.z;
	const float cx = face.Vertices[2].x - point.x, cy = face.Vertices[2].y - point.y, cz = face.Vertices[2].z - point.z;
	return ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
}

// The reference keys an edge by the XOR of string hashes of its end points printed with std::to_string (six
// decimals, VMACH.cpp:940-947).  End points closer than the printing resolution therefore share a key; keeping the
// same key keeps the same hull on such inputs.
size_t ConvexHull::Key2Edge(const ConvexHullVertex& p1, const ConvexHullVertex& p2)
{
	std::hash<std::string> h;
	return h(std::to_string(p1.x) + std::to_string(p1.y) + std::to_string(p1.z)) ^
		   h(std::to_string(p2.x) + std::to_string(p2.y) + std::to_string(p2.z));
}

void ConvexHull::CreateEdge(const ConvexHullVertex& p1, const ConvexHullVertex& p2, ConvexHullFace& newFace)
{
	const size_t key = Key2Edge(p1, p2);
	auto it = m_edgeMap.find(key);
	if (it == m_edgeMap.end())
	{
		m_edgeList.emplace_back(p1, p2);
		it = m_edgeMap.insert({ key, &m_edgeList.back() }).first;
	}
	it->second->LinkFace(&newFace);
}

void ConvexHull::CreateFace(const ConvexHullVertex& p1, const ConvexHullVertex& p2, const ConvexHullVertex& p3, const ConvexHullVertex& innerPoint)
{
	m_faceList.emplace_back(p1, p2, p3);
	ConvexHullFace& face = m_faceList.back();
	m_addedFaceVec.push_back(&face);
	if (Volume(face, innerPoint) < 0)   // orient so that the inner point is on the positive side
		face.Rewind();
	CreateEdge(p1, p2, face);
	CreateEdge(p1, p3, face);
	CreateEdge(p2, p3, face);
}

void ConvexHull::AddPointToHull(const ConvexHullVertex& point)
{
	bool any = false;
	for (ConvexHullFace& face : m_faceList)
		if (Volume(face, point) < 0)
		{
			face.Visible = true;
			m_visibleFaceVec.push_back(&face);
			any = true;
		}
	if (!any)
		return;
	// horizon edges (one visible, one hidden face) get a new face to the point; edges appended meanwhile are visited too
	for (auto it = m_edgeList.begin(); it != m_edgeList.end(); ++it)
	{
		ConvexHullEdge& edge = *it;
		if (edge.Face1 == nullptr || edge.Face2 == nullptr)
			continue;
		if (edge.Face1->Visible && edge.Face2->Visible)
			edge.Remove = true;
		else if (edge.Face1->Visible || edge.Face2->Visible)
		{
			if (edge.Face1->Visible)
				std::swap(edge.Face1, edge.Face2);
			// now Face1 is hidden and Face2 visible; the orientation probe is the visible face's vertex off the edge
			// (FindInnerPoint(face2, edge), VMACH.cpp:950-962, 1028)
			ConvexHullVertex inner = edge.Face2->Vertices[0];
			bool found = false;
			for (int i = 0; i < 3 && !found; i++)
			{
				const ConvexHullVertex& v = edge.Face2->Vertices[i];
				if (v == edge.EndPoints[0] || v == edge.EndPoints[1])
					continue;
				inner = v;
				found = true;
			}
			edge.EraseFace(edge.Face2);
			CreateFace(edge.EndPoints[0], edge.EndPoints[1], point, inner);
		}
	}
}

bool ConvexHull::BuildFirstHull()
{
	if (m_pointCloud.size() <= 3)
		return false;
	auto& P = m_pointCloud;
	const auto v1 = std::max_element(P.begin(), P.end(), [](const ConvexHullVertex& a, const ConvexHullVertex& b) { return a.x < b.x; });
	auto dist1 = [&](const ConvexHullVertex& a) {
		return std::sqrt(std::pow(a.x - v1->x, 2) + std::pow(a.y - v1->y, 2) + std::pow(a.z - v1->z, 2));
	};
	const auto v2 = std::max_element(P.begin(), P.end(), [&](const ConvexHullVertex& a, const ConvexHullVertex& b) { return dist1(a) < dist1(b); });
	const auto v3 = std::max_element(P.begin(), P.end(), [&](const ConvexHullVertex& a, const ConvexHullVertex& b) {
		ConvexHullFace f1(*v1, *v2, a), f2(*v1, *v2, b);
		return f1.CalcArea() < f2.CalcArea();
	});
	const auto v4 = std::max_element(P.begin(), P.end(), [&](const ConvexHullVertex& a, const ConvexHullVertex& b) {
		const ConvexHullFace f(*v1, *v2, *v3);
		return Volume(f, a) < Volume(f, b);
	});
	v1->Processed = v2->Processed = v3->Processed = v4->Processed = true;
	m_processedPointCnt = 4;
	CreateFace(*v1, *v2, *v3, *v4);
	CreateFace(*v1, *v2, *v4, *v3);
	CreateFace(*v1, *v3, *v4, *v2);
	CreateFace(*v2, *v3, *v4, *v1);
	return true;
}

void ConvexHull::CreateConvexHull()
{
	if (!BuildFirstHull())
		return;
	// every point's outside volume is its own sequential sum: the points are spread over the worker pool, the order
	// of the additions per point is the reference's
	constexpr size_t CHUNK = 256;
	const size_t n_chunks = (m_pointCloud.size() + CHUNK - 1) / CHUNK;
	SurtrHost::detail::parallel_for(n_chunks, [&](size_t c) {
		for (size_t i = c * CHUNK; i < std::min(m_pointCloud.size(), (c + 1) * CHUNK); i++)
		{
			if (m_pointCloud[i].Processed)
				continue;
			for (const ConvexHullFace& f : m_faceList)
				m_pointVolume[i] += std::max(0.0f, Volume(f, m_pointCloud[i]));
		}
	});
	if (m_limitCnt == 0)
		m_limitCnt = (uint32_t)m_pointCloud.size();
	while (m_processedPointCnt < m_limitCnt)
	{
		// greedy: the point with the largest outside volume joins the hull
		const int k = (int)std::distance(m_pointVolume.begin(), std::max_element(m_pointVolume.begin(), m_pointVolume.end()));
		AddPointToHull(m_pointCloud[k]);
		m_pointCloud[k].Processed = true;
		m_pointVolume[k] = -FLT_MAX;
		m_processedPointCnt++;
		SurtrHost::detail::parallel_for(n_chunks, [&](size_t c) {
			for (size_t i = c * CHUNK; i < std::min(m_pointCloud.size(), (c + 1) * CHUNK); i++)
			{
				if (m_pointCloud[i].Processed)
					continue;
				float removed = 0.0f, added = 0.0f;
				for (ConvexHullFace* f : m_visibleFaceVec)
					removed += std::max(0.0f, Volume(*f, m_pointCloud[i]));
				for (ConvexHullFace* f : m_addedFaceVec)
					added += std::max(0.0f, Volume(*f, m_pointCloud[i]));
				m_pointVolume[i] -= removed;
				m_pointVolume[i] += added;
			}
		});
		CleanUp();
	}
}

void ConvexHull::CleanUp()
{
	m_visibleFaceVec.clear();
	m_addedFaceVec.clear();
	for (auto it = m_edgeList.begin(); it != m_edgeList.end();)
	{
		if (it->Remove)
		{
			m_edgeMap.erase(Key2Edge(it->EndPoints[0], it->EndPoints[1]));
			it = m_edgeList.erase(it);
		}
		else
			++it;
	}
	m_faceList.remove_if([](const ConvexHullFace& f) { return f.Visible; });
}

std::vector<Vector3> GenerateICHNormal(const std::vector<Vector3>& vertices, int ichIncludePointLimit)
{
	ConvexHull ich(vertices, (uint32_t)ichIncludePointLimit);
	std::vector<Vector3> normals;
	for (const ConvexHullFace& f : ich.GetFaces())
	{
		Vector3 n = (f.Vertices[1] - f.Vertices[0]).Cross(f.Vertices[2] - f.Vertices[0]);
		n.Normalize();
		normals.push_back(n);
	}
	return normals;
}
} // namespace VMACH
