// Kdop.h -- host-side mirror of the reference's k-DOP container (Inc/Kdop.h:16-41), headless.
// Calc = kernel kdop_arg_kernel through surtr_kdop_calc; ClipWithPolyhedron = the GPU clipper with the plane list
// [Min0, Max0, Min1, Max1, ...] (Kdop.cpp:166-179).  ClipWithPolygon (legacy, no callers) and Render are not declared.
#pragma once

#include "SimpleMath.h"

#include <cfloat>
#include <vector>

namespace VMACH { struct Polygon3D; }
namespace Poly { struct Vertex; typedef std::vector<Poly::Vertex> Polyhedron; }

namespace Kdop
{
using DirectX::SimpleMath::Plane;
using DirectX::SimpleMath::Vector3;

struct KdopElement   // Inc/Kdop.h:16-27; extents are doubles holding float values
{
	Vector3 Normal;
	Vector3 MinVertex;
	Vector3 MaxVertex;
	double MinDist = DBL_MAX;
	double MaxDist = -DBL_MAX;
	Plane MinPlane;
	Plane MaxPlane;

	KdopElement(const Vector3& _normal) : Normal(_normal) {}
};

struct KdopContainer
{
	std::vector<KdopElement> ElementVec;

	KdopContainer(const std::vector<Vector3>& normalVec);

	// Calc accumulates like the reference (no reset): construct a fresh container per use (Surtr.cpp:1451, 1775).
	void Calc(const std::vector<Vector3>& vertices, const double& maxAxisScale, const float& planeGapInv);   // Kdop.cpp:15-51
	void Calc(const VMACH::Polygon3D& mesh);                                                                 // Kdop.cpp:53-90
	void Calc(const Poly::Polyhedron& mesh);                                                                 // Kdop.cpp:92-115
	Poly::Polyhedron ClipWithPolyhedron(const Poly::Polyhedron& polyhedron);                                 // Kdop.cpp:166-179

private:
	void Accumulate(const std::vector<Vector3>& vertices);
};
} // namespace Kdop
