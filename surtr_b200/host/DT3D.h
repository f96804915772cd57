// DT3D.h -- host-side mirror of the reference's 3-D Delaunay interface (Inc/DT3D.h:10-315) plus the Voronoi
// cell builder the fracture path needs.
//
// The reference header is dead code there (no translation unit includes it) and its Voronoi() only emits edges;
// the live cells came from voro++, which is not vendored.  Here Triangulate keeps the reference's value-based
// interface (tets carry point VALUES; map back with Vector3::operator==) but is an index-based Bowyer-Watson with
// an exact-in-double in-sphere predicate, and VoronoiCells derives the cells from the Delaunay neighbours:
// cell i = container box clipped (on the GPU) by the bisector half-spaces Plane((Si+Sj)*0.5, Sj-Si), j ascending.
#pragma once

#include "VMACH.h"

#include <array>
#include <vector>

namespace DT3D
{
using DirectX::SimpleMath::Vector3;

void tetrahedron_circumcenter(const double a[3], const double b[3], const double c[3], const double d[3],
							  double circumcenter[3], double* xi, double* eta, double* zeta);

struct Sphere { Vector3 center; float radius; };

struct Edge
{
	Vector3 p0, p1;
	Edge(const Vector3& _p0, const Vector3& _p1) : p0(_p0), p1(_p1) {}
	bool operator==(const Edge& o) const { return (o.p0 == p0 && o.p1 == p1) || (o.p0 == p1 && o.p1 == p0); }
};

struct Triangle
{
	Vector3 p0, p1, p2;
	Triangle() = default;
	Triangle(const Vector3& _p0, const Vector3& _p1, const Vector3& _p2) : p0(_p0), p1(_p1), p2(_p2) {}
	bool operator==(const Triangle& o) const;
};

struct Tetrahedron
{
	Vector3 p0, p1, p2, p3;
	Triangle t0, t1, t2, t3;
	Sphere sphere;
	Tetrahedron(Vector3 _p0, Vector3 _p1, Vector3 _p2, Vector3 _p3);
};

struct Delaunay
{
	std::vector<Tetrahedron> TetVec;
	std::vector<Triangle> FaceVec;
};

Delaunay Triangulate(const std::vector<Vector3>& points);   // returns empty for < 3 points (DT3D.h:161-162)
std::vector<Edge> Voronoi(const Delaunay& dt);             // unique circumcentre edges (DT3D.h:269-315)

// Delaunay neighbour lists (CSR, ascending seed index) without materialising tets by value.
void Neighbors(const std::vector<Vector3>& points, std::vector<uint32_t>& off, std::vector<uint32_t>& idx);

namespace detail
{
// Test hook: tets as point indices, in construction order; use_grid = false is the quadratic scan of every live tet per
// inserted point (the reference's search, Inc/DT3D.h:198-246), true the grid-accelerated search -- same output.
std::vector<std::array<int, 4>> TetIndices(const std::vector<Vector3>& points, bool use_grid);
}

// NEW (replaces Surtr::GenerateVoronoi, Surtr.cpp:2003-2070): cells in seed order, faces outward, bounded by the
// unit container box [-0.5, 0.5]^3.  The clipping runs on the GPU.
std::vector<VMACH::Polygon3D> VoronoiCells(const std::vector<Vector3>& seeds);
} // namespace DT3D
