// ConvexHull.h -- host-side mirror of VMACH::ConvexHull (Inc/VMACH.h:88-165, Src/VMACH.cpp:869-1203): the greedy
// incremental convex hull ("ICH") whose face normals seed the k-DOPs (Surtr::GenerateICHNormal, Surtr.cpp:1961-1982).
// Tiny and inherently sequential (<= 20 points at ACH time, <= 4 at refit time): it stays on the host, exactly as
// SURVEY.md section 2 row 4 scopes it.  Own implementation; the visiting orders that decide the order of the output
// faces (edge list order, first-maximum selection) follow the reference so the k-DOP plane order is the same.
#pragma once

#include "SimpleMath.h"

#include <cstdint>
#include <list>
#include <unordered_map>
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
	ConvexHullFace* Face1;
	ConvexHullFace* Face2;
	ConvexHullVertex EndPoints[2];
	ConvexHullEdge(const ConvexHullVertex& p1, const ConvexHullVertex& p2) : Remove(false), Face1(nullptr), Face2(nullptr)
	{
		EndPoints[0] = p1; EndPoints[1] = p2;
	}
	void LinkFace(ConvexHullFace* face);
	void EraseFace(ConvexHullFace* face);
};

class ConvexHull
{
public:
	ConvexHull(const std::vector<ConvexHullVertex>& pointCloud, uint32_t limitCnt);
	ConvexHull(const std::vector<Vector3>& pointCloud, uint32_t limitCnt);

	bool Contains(const ConvexHullVertex& point) const;
	const std::list<ConvexHullFace> GetFaces() const { return m_faceList; }
	const std::list<ConvexHullEdge> GetEdges() const { return m_edgeList; }

	static bool Colinear(const ConvexHullVertex& p1, const ConvexHullVertex& p2, const ConvexHullVertex& p3);
	static float Volume(const ConvexHullFace& face, const ConvexHullVertex& point);

private:
	static size_t Key2Edge(const ConvexHullVertex& p1, const ConvexHullVertex& p2);
	void CreateFace(const ConvexHullVertex& p1, const ConvexHullVertex& p2, const ConvexHullVertex& p3, const ConvexHullVertex& innerPoint);
	void CreateEdge(const ConvexHullVertex& p1, const ConvexHullVertex& p2, ConvexHullFace& newFace);
	void AddPointToHull(const ConvexHullVertex& point);
	bool BuildFirstHull();
	void CreateConvexHull();
	void CleanUp();

	std::vector<ConvexHullFace*> m_visibleFaceVec, m_addedFaceVec;
	uint32_t m_limitCnt = 0, m_processedPointCnt = 0;
	std::vector<ConvexHullVertex> m_pointCloud;
	std::vector<float> m_pointVolume;
	std::list<ConvexHullFace> m_faceList;
	std::list<ConvexHullEdge> m_edgeList;
	std::unordered_map<size_t, ConvexHullEdge*> m_edgeMap;
};

// Surtr::GenerateICHNormal (Surtr.cpp:1961-1982): normalised (v1-v0) x (v2-v0) of every hull face, list order.
std::vector<Vector3> GenerateICHNormal(const std::vector<Vector3>& vertices, int ichIncludePointLimit);
} // namespace VMACH
