// VMACH.h -- host-side mirror of the cell container of the reference (Inc/VMACH.h:17-86, 190-195), headless.
//
// Only the data interface of the hot path is kept: PolygonFace / Polygon3D hold the face loops of a Voronoi cell
// and the face plane the clipper reads (Poly.cpp:558-560).  The legacy face-loop clipper (ClipWithPlane/Face/
// Polygon, VMACH.cpp:550-867), EarClipping, Render, VisualMesh are dead or debug code in the reference (SURVEY.md
// section 2 rows 3 and 5) and are not declared.  ConvexHull (ICH, VMACH.cpp:869-1203) belongs to the "next" row f-2.
#pragma once

#include "SimpleMath.h"

#include <vector>

namespace VMACH
{
using DirectX::SimpleMath::Plane;
using DirectX::SimpleMath::Vector3;

bool NearlyEqual(const Vector3& v1, const Vector3& v2);   // (v1 - v2).Length() < 1e-12 (VMACH.cpp:1205)

struct PolygonFace   // Inc/VMACH.h:17-58
{
	bool GuaranteeConvex;
	std::vector<Vector3> VertexVec;
	Plane FacePlane;
	bool FacePlaneConstructed;
	bool ForceColor;

	PolygonFace(bool _guranteeConvex) : GuaranteeConvex(_guranteeConvex), FacePlaneConstructed(false), ForceColor(false) {}
	PolygonFace(bool _guranteeConvex, std::vector<Vector3> _vertexVec)
		: GuaranteeConvex(_guranteeConvex), VertexVec(_vertexVec), FacePlaneConstructed(false), ForceColor(false)
	{
		ConstructFacePlane();
	}

	bool IsEmpty() const { return VertexVec.empty(); }
	Vector3 GetNormal() const;                  // throws if the plane was never constructed (VMACH.cpp:88-97)
	void AddVertex(const Vector3& newVertex);   // drops near-duplicates; plane appears with the 3rd vertex (VMACH.cpp:289-300)
	void ConstructFacePlane();                  // Plane(v0, v1, v2), only if GuaranteeConvex (VMACH.cpp:302-310)
	void ManuallySetFacePlane(const Plane& plane);
	void Rewind();                              // reverse the loop and rebuild the plane (VMACH.cpp:318-322)
};

struct Polygon3D   // Inc/VMACH.h:60-86
{
	bool GuaranteeConvex;
	std::vector<PolygonFace> FaceVec;

	Polygon3D(bool _guranteeConvex) : GuaranteeConvex(_guranteeConvex) {}
	Polygon3D(bool _guranteeConvex, std::vector<PolygonFace> _faceVec) : GuaranteeConvex(_guranteeConvex), FaceVec(_faceVec) {}

	void AddFace(const PolygonFace& newFace);
	// each call re-derives every face plane from the first three face vertices (VMACH.cpp:506-534)
	void Translate(const Vector3& vector);
	void Scale(const float& scalar);
	void Scale(const Vector3& vector);
};

Polygon3D GetBoxPolygon();   // the six outward planes of the unit cube (VMACH.cpp:1207-1226)
// (b - a) x (c - a) . n > 0 (VMACH.cpp:1240-1243): c lies to the left of a->b seen against n
bool OnYourRight(const Vector3& a, const Vector3& b, const Vector3& c, const Vector3& n);
} // namespace VMACH
