#include "VMACH.h"

#include <algorithm>
#include <stdexcept>

namespace VMACH
{
bool NearlyEqual(const Vector3& v1, const Vector3& v2) { return (v1 - v2).Length() < 1e-12; }

Vector3 PolygonFace::GetNormal() const
{
	if (!FacePlaneConstructed)
		throw std::runtime_error("PolygonFace::GetNormal: face plane was never constructed");
	Vector3 n = FacePlane.Normal();
	n.Normalize();
	return n;
}

void PolygonFace::AddVertex(const Vector3& newVertex)
{
	for (const Vector3& v : VertexVec)
		if (NearlyEqual(v, newVertex))
			return;
	VertexVec.push_back(newVertex);
	if (GuaranteeConvex && VertexVec.size() == 3)
		ConstructFacePlane();
}

void PolygonFace::ConstructFacePlane()
{
	if (GuaranteeConvex && VertexVec.size() >= 3)
	{
		FacePlane = Plane(VertexVec[0], VertexVec[1], VertexVec[2]);
		FacePlaneConstructed = true;
	}
}

void PolygonFace::ManuallySetFacePlane(const Plane& plane)
{
	FacePlane = plane;
	FacePlaneConstructed = true;
}

void PolygonFace::Rewind()
{
	std::reverse(VertexVec.begin(), VertexVec.end());
	ConstructFacePlane();
}

void Polygon3D::AddFace(const PolygonFace& newFace)
{
	if (GuaranteeConvex && !newFace.GuaranteeConvex)
		GuaranteeConvex = false;
	FaceVec.push_back(newFace);
}

void Polygon3D::Translate(const Vector3& vector)
{
	for (PolygonFace& f : FaceVec)
	{
		for (Vector3& v : f.VertexVec)
			v += vector;
		f.ConstructFacePlane();
	}
}

void Polygon3D::Scale(const float& scalar)
{
	for (PolygonFace& f : FaceVec)
	{
		for (Vector3& v : f.VertexVec)
			v *= scalar;
		f.ConstructFacePlane();
	}
}

void Polygon3D::Scale(const Vector3& vector)
{
	for (PolygonFace& f : FaceVec)
	{
		for (Vector3& v : f.VertexVec)
			v *= vector;
		f.ConstructFacePlane();
	}
}

Polygon3D GetBoxPolygon()
{
	const float h = 0.5f;
	const Vector3 c[8] = { Vector3(-h, -h, -h), Vector3(+h, -h, -h), Vector3(+h, +h, -h), Vector3(-h, +h, -h),
						   Vector3(-h, -h, +h), Vector3(+h, -h, +h), Vector3(+h, +h, +h), Vector3(-h, +h, +h) };
	// loops listed so that Rewind() leaves outward planes: -z, +x, +z, -x, +y, -y
	const int loops[6][4] = { { 1, 2, 3, 0 }, { 5, 6, 2, 1 }, { 4, 7, 6, 5 }, { 0, 3, 7, 4 }, { 6, 7, 3, 2 }, { 4, 5, 1, 0 } };
	Polygon3D box(true);
	for (int f = 0; f < 6; f++)
	{
		PolygonFace face(true, { c[loops[f][0]], c[loops[f][1]], c[loops[f][2]], c[loops[f][3]] });
		face.Rewind();
		box.FaceVec.push_back(face);
	}
	return box;
}
bool OnYourRight(const Vector3& a, const Vector3& b, const Vector3& c, const Vector3& n) { return (b - a).Cross(c - a).Dot(n) > 0; }
} // namespace VMACH
