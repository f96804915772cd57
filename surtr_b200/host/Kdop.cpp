#include "Kdop.h"

#include "Engine.h"
#include "Poly.h"
#include "VMACH.h"

namespace Kdop
{
KdopContainer::KdopContainer(const std::vector<Vector3>& normalVec)
{
	for (const Vector3& n : normalVec)
		ElementVec.push_back(KdopElement(n));
}

// Extents of a vertex stream on the GPU (kdop_arg_kernel), merged into the elements with the reference's strict
// compares so that repeated Calc calls accumulate exactly like Kdop.cpp:21-35.
void KdopContainer::Accumulate(const std::vector<Vector3>& vertices)
{
	if (vertices.empty() || ElementVec.empty())
		return;
	std::vector<float> v4(4 * vertices.size(), 0.f), n3(3 * ElementVec.size());
	for (size_t i = 0; i < vertices.size(); i++)
	{
		v4[4 * i] = vertices[i].x; v4[4 * i + 1] = vertices[i].y; v4[4 * i + 2] = vertices[i].z;
	}
	for (size_t e = 0; e < ElementVec.size(); e++)
	{
		n3[3 * e] = ElementVec[e].Normal.x; n3[3 * e + 1] = ElementVec[e].Normal.y; n3[3 * e + 2] = ElementVec[e].Normal.z;
	}
	std::vector<float> dist(2 * ElementVec.size()), planes(8 * ElementVec.size());
	std::vector<int32_t> arg(2 * ElementVec.size());
	SurtrHost::detail::check(surtr_kdop_calc(SurtrHost::detail::context(), v4.data(), (uint32_t)vertices.size(), n3.data(),
											 (uint32_t)ElementVec.size(), dist.data(), arg.data(), planes.data()),
							 "surtr_kdop_calc");
	for (size_t e = 0; e < ElementVec.size(); e++)
	{
		KdopElement& el = ElementVec[e];
		if (el.MinDist > dist[2 * e])
		{
			el.MinDist = dist[2 * e];
			el.MinVertex = vertices[arg[2 * e]];
			el.MinPlane = Plane(planes[8 * e], planes[8 * e + 1], planes[8 * e + 2], planes[8 * e + 3]);
		}
		if (el.MaxDist < dist[2 * e + 1])
		{
			el.MaxDist = dist[2 * e + 1];
			el.MaxVertex = vertices[arg[2 * e + 1]];
			el.MaxPlane = Plane(planes[8 * e + 4], planes[8 * e + 5], planes[8 * e + 6], planes[8 * e + 7]);
		}
	}
}

static void offset_planes(std::vector<KdopElement>& elements, float gap)
{
	// Kdop.cpp:39-50 / 81-89: normalise the plane normals and push both planes outward by `gap`
	for (KdopElement& el : elements)
	{
		Vector3 mn = el.MinPlane.Normal(), mx = el.MaxPlane.Normal();
		mn.Normalize();
		mx.Normalize();
		el.MinPlane = Plane(el.MinVertex + mn * gap, mn);
		el.MaxPlane = Plane(el.MaxVertex + mx * gap, mx);
	}
}

void KdopContainer::Calc(const std::vector<Vector3>& vertices, const double& maxAxisScale, const float& planeGapInv)
{
	Accumulate(vertices);
	offset_planes(ElementVec, (float)(maxAxisScale / planeGapInv));   // Vector3 * double narrows to float at the call
}

void KdopContainer::Calc(const VMACH::Polygon3D& mesh)
{
	std::vector<Vector3> vertices;
	for (const VMACH::PolygonFace& face : mesh.FaceVec)
		vertices.insert(vertices.end(), face.VertexVec.begin(), face.VertexVec.end());
	Accumulate(vertices);
	offset_planes(ElementVec, (float)0.001);
}

void KdopContainer::Calc(const Poly::Polyhedron& mesh)
{
	std::vector<Vector3> vertices;
	for (const Poly::Vertex& v : mesh)
		vertices.push_back(v.Position);
	Accumulate(vertices);
}

Poly::Polyhedron KdopContainer::ClipWithPolyhedron(const Poly::Polyhedron& polyhedron)
{
	std::vector<Plane> planes;
	for (const KdopElement& el : ElementVec)
	{
		planes.push_back(el.MinPlane);
		planes.push_back(el.MaxPlane);
	}
	Poly::Polyhedron res = polyhedron;
	Poly::ClipPolyhedron(res, planes);
	return res;
}
} // namespace Kdop
