// Poly.h -- host-side mirror of the reference's namespace Poly (Inc/Poly.h:9-77), headless.
//
// Same names, signatures and conventions as the reference, so a caller written against Inc/Poly.h recompiles
// unchanged.  The geometric work is NOT done here: ClipPolyhedron / Moments route through the C ABI
// (include/surtr_b200.h) as a batch of one, and the batch entry points in Fracture.h are what a hot caller uses.
// Render helpers (Inc/Poly.h:50-68), EarClipping / IsCCW (:75-76) and ExtractNeighborFromMesh (:38) are outside
// the accelerated path (renderer stubbed out, mesh clip = "next" row f-1) and not declared.
#pragma once

#include "SimpleMath.h"

#include <vector>

namespace VMACH { struct Polygon3D; }

namespace Poly
{
using DirectX::SimpleMath::Plane;
using DirectX::SimpleMath::Vector3;

struct Vertex   // Inc/Poly.h:15-29
{
	Vector3 Position;
	std::vector<int> NeighborVertexVec;
	int comp;
	mutable int ID;

	Vertex() : Position(0, 0, 0), comp(1), ID(-1) {}                                   // Poly.cpp:11
	Vertex(const Vector3& pos) : Position(pos), comp(1), ID(-1) {}                     // Poly.cpp:12
	Vertex(const Vector3& pos, const int c) : Position(pos), comp(c), ID(-1) {}        // Poly.cpp:13
	Vertex(const Vertex& rhs) = default;
	Vertex& operator=(const Vertex& rhs) = default;
	bool operator==(const Vertex& rhs) const   // compares all four fields (Poly.cpp:25-28)
	{
		return Position == rhs.Position && NeighborVertexVec == rhs.NeighborVertexVec && comp == rhs.comp && ID == rhs.ID;
	}
};

typedef std::vector<Poly::Vertex> Polyhedron;   // empty vector <=> "no polyhedron"
typedef std::vector<std::vector<int>> Extract;

void InitPolyhedron(Polyhedron& polyhedron, const std::vector<Vector3>& positionVec, const std::vector<std::vector<int>>& neighborVec);
// Triangle list -> CCW neighbour ring per vertex (Poly.cpp:128-263), the input of InitPolyhedron for Piece::Mesh
// (Surtr.cpp:1788-1795).  Index bookkeeping only (once per object, host).  Throws std::runtime_error where the reference
// throws (asymmetric adjacency) or would not terminate (an edge fan with three or more continuations).
std::vector<std::vector<int>> ExtractNeighborFromMesh(const std::vector<Vector3>& vertices, const std::vector<int>& indices);
// Volume and centroid (Poly.cpp:55-87), computed by kernel K4 on the GPU.
void Moments(double& zerothMoment, Vector3& firstMoment, const Polyhedron& polyhedron);
// Face loops (Poly.cpp:89-126); caller owns the result.  Pure index bookkeeping over the rings (no arithmetic):
// the face COUNT of every fragment comes from K4, the loops are only materialised on request for host consumers.
Extract* ExtractFaces(const Polyhedron& polyhedron);

// Successive half-space clipping (Poly.cpp:265-566), on the GPU through the C ABI.  Result empty when culled.
void ClipPolyhedron(Polyhedron& polyhedron, const std::vector<Plane>& planes);
Polyhedron ClipPolyhedron(const Polyhedron& polyhedron, const VMACH::Polygon3D& polygon3D);

void Translate(Polyhedron& polyhedron, const Vector3& v);
void Scale(Polyhedron& polyhedron, const Vector3& v);
void Transform(Polyhedron& polyhedron, const DirectX::XMMATRIX& matrix);   // transposes its argument first (Poly.cpp:582)

Polyhedron GetBB();   // unit cube centred at the origin (Poly.cpp:587-617)

// Scalar helpers of the interface (Poly.cpp:716-751); +1 keep / 0 in-plane / -1 clipped.
int ComparePlanePoint(const Plane& plane, const Vector3& point);
int ComparePlaneBB(const Plane& plane, const double xmin, const double ymin, const double zmin, const double xmax, const double ymax, const double zmax);
Vector3 PlaneLineIntersection(const Vector3& a, const Vector3& b, const Plane& plane);
} // namespace Poly
