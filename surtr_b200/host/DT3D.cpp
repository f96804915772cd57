#include "DT3D.h"

#include "Engine.h"
#include "Poly.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <unordered_map>

namespace DT3D
{
// Circumcentre relative to `a` (Inc/DT3D.h:10-87): Cramer's rule on the edge vectors, all in double.
void tetrahedron_circumcenter(const double a[3], const double b[3], const double c[3], const double d[3],
							  double circumcenter[3], double* xi, double* eta, double* zeta)
{
	const double ba[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] };
	const double ca[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };
	const double da[3] = { d[0] - a[0], d[1] - a[1], d[2] - a[2] };
	const double lb = ba[0] * ba[0] + ba[1] * ba[1] + ba[2] * ba[2];
	const double lc = ca[0] * ca[0] + ca[1] * ca[1] + ca[2] * ca[2];
	const double ld = da[0] * da[0] + da[1] * da[1] + da[2] * da[2];
	const double cd[3] = { ca[1] * da[2] - da[1] * ca[2], ca[2] * da[0] - da[2] * ca[0], ca[0] * da[1] - da[0] * ca[1] };
	const double db[3] = { da[1] * ba[2] - ba[1] * da[2], da[2] * ba[0] - ba[2] * da[0], da[0] * ba[1] - ba[0] * da[1] };
	const double bc[3] = { ba[1] * ca[2] - ca[1] * ba[2], ba[2] * ca[0] - ca[2] * ba[0], ba[0] * ca[1] - ca[0] * ba[1] };
	double denom = 0.5 / (ba[0] * cd[0] + ba[1] * cd[1] + ba[2] * cd[2]);
	for (int k = 0; k < 3; k++)
		circumcenter[k] = (lb * cd[k] + lc * db[k] + ld * bc[k]) * denom;
	if (xi)
	{
		denom *= 2.0;
		*xi = (circumcenter[0] * cd[0] + circumcenter[1] * cd[1] + circumcenter[2] * cd[2]) * denom;
		*eta = (circumcenter[0] * db[0] + circumcenter[1] * db[1] + circumcenter[2] * db[2]) * denom;
		*zeta = (circumcenter[0] * bc[0] + circumcenter[1] * bc[1] + circumcenter[2] * bc[2]) * denom;
	}
}

bool Triangle::operator==(const Triangle& o) const
{
	const Vector3* a[3] = { &p0, &p1, &p2 };
	const Vector3* b[3] = { &o.p0, &o.p1, &o.p2 };
	static const int perm[6][3] = { { 0, 1, 2 }, { 0, 2, 1 }, { 1, 0, 2 }, { 1, 2, 0 }, { 2, 0, 1 }, { 2, 1, 0 } };
	for (const auto& p : perm)
		if (*a[0] == *b[p[0]] && *a[1] == *b[p[1]] && *a[2] == *b[p[2]])
			return true;
	return false;
}

Tetrahedron::Tetrahedron(Vector3 _p0, Vector3 _p1, Vector3 _p2, Vector3 _p3) : p0(_p0), p1(_p1), p2(_p2), p3(_p3)
{
	t0 = Triangle(p0, p1, p2);
	t1 = Triangle(p0, p1, p3);
	t2 = Triangle(p1, p2, p3);
	t3 = Triangle(p2, p0, p3);
	const double a[] = { p0.x, p0.y, p0.z }, b[] = { p1.x, p1.y, p1.z }, c[] = { p2.x, p2.y, p2.z }, d[] = { p3.x, p3.y, p3.z };
	double cc[3];
	tetrahedron_circumcenter(a, b, c, d, cc, nullptr, nullptr, nullptr);
	sphere.center = Vector3(a[0] + cc[0], a[1] + cc[1], a[2] + cc[2]);
	sphere.radius = Vector3::Distance(p0, sphere.center);
}

namespace
{
struct Tet
{
	int v[4];
	double c[3], r2;   // circumsphere
	bool alive;
};

// Uniform grid over the points' bounding box that finds the tets whose circumsphere may contain a query point: a tet
// is listed in every grid cell its circumsphere's bounding box touches (in `big` when that is more than a few dozen
// cells, or when the sphere is not finite -- those are tested for every point).  A point inside a sphere is inside the
// sphere's box, so the list of the point's cell plus `big` is a superset of the tets the full scan would find, and the
// SAME predicate then picks exactly the same ones.  Dead tets are dropped from a list when it is next scanned.
struct SphereGrid
{
	int G = 1;
	double lo[3] = { 0, 0, 0 }, inv[3] = { 0, 0, 0 };
	std::vector<std::vector<int>> cell;
	std::vector<int> big;
	static constexpr long BIG_CELLS = 128;

	void init(const double lo_[3], const double hi_[3], int n)
	{
		G = std::max(1, (int)std::cbrt((double)n / 32.0));
		for (int k = 0; k < 3; k++)
		{
			lo[k] = lo_[k];
			inv[k] = hi_[k] > lo_[k] ? G / (hi_[k] - lo_[k]) : 0.0;
		}
		cell.assign((size_t)G * G * G, {});
		big.clear();
	}
	int coord(double x, int k) const   // monotone in x; NaN and anything below the box -> 0, above -> G - 1
	{
		const double f = (x - lo[k]) * inv[k];
		if (!(f > 0.0)) return 0;
		if (f >= (double)G) return G - 1;
		return (int)f;
	}
	void add(int id, const Tet& t)
	{
		const bool finite = std::isfinite(t.c[0]) && std::isfinite(t.c[1]) && std::isfinite(t.c[2]) && std::isfinite(t.r2);
		if (!finite) { big.push_back(id); return; }
		const double r = std::sqrt(t.r2 * (1.0 + 1e-12)) * (1.0 + 1e-9) + 1e-300;
		int a[3], b[3];
		for (int k = 0; k < 3; k++) { a[k] = coord(t.c[k] - r, k); b[k] = coord(t.c[k] + r, k); }
		if ((long)(b[0] - a[0] + 1) * (b[1] - a[1] + 1) * (b[2] - a[2] + 1) > BIG_CELLS) { big.push_back(id); return; }
		for (int z = a[2]; z <= b[2]; z++)
			for (int y = a[1]; y <= b[1]; y++)
				for (int x = a[0]; x <= b[0]; x++)
					cell[((size_t)z * G + y) * G + x].push_back(id);
	}
	std::vector<int>& at(const double p[3]) { return cell[((size_t)coord(p[2], 2) * G + coord(p[1], 1)) * G + coord(p[0], 0)]; }
};

// Index-based Bowyer-Watson.  Points 0..n-1 are the input, n..n+3 a far super-tetrahedron.  `use_grid` = false is the
// plain scan over every live tet per inserted point (quadratic, as the reference's Inc/DT3D.h:198-246); true finds the
// same tets through the SphereGrid -- identical output, tet for tet (tests/test_host.py), 1.6 s -> tens of ms at 4096 seeds.
std::vector<Tet> bowyer_watson(const std::vector<Vector3>& points, bool use_grid = true)
{
	const int n = (int)points.size();
	std::vector<std::array<double, 3>> P(n + 4);
	double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
	for (int i = 0; i < n; i++)
	{
		P[i] = { points[i].x, points[i].y, points[i].z };
		for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], P[i][k]); hi[k] = std::max(hi[k], P[i][k]); }
	}
	const double dmax = std::max({ hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-30 });
	const double mid[3] = { (lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, (lo[2] + hi[2]) / 2 };
	const double S = 1000.0 * dmax;   // far enough that hull-adjacent Delaunay edges are not lost
	P[n + 0] = { mid[0] - S, mid[1] - S, mid[2] - S };
	P[n + 1] = { mid[0] + S, mid[1] - S, mid[2] - S * 0.5 };
	P[n + 2] = { mid[0], mid[1] + S, mid[2] - S * 0.75 };
	P[n + 3] = { mid[0], mid[1], mid[2] + S };

	auto make = [&](int a, int b, int c, int d) {
		Tet t;
		t.v[0] = a; t.v[1] = b; t.v[2] = c; t.v[3] = d;
		double cc[3];
		tetrahedron_circumcenter(P[a].data(), P[b].data(), P[c].data(), P[d].data(), cc, nullptr, nullptr, nullptr);
		for (int k = 0; k < 3; k++) t.c[k] = P[a][k] + cc[k];
		t.r2 = cc[0] * cc[0] + cc[1] * cc[1] + cc[2] * cc[2];
		t.alive = true;
		return t;
	};
	SphereGrid grid;
	if (use_grid) grid.init(lo, hi, n);
	std::vector<Tet> tets;
	tets.push_back(make(n, n + 1, n + 2, n + 3));
	if (use_grid) grid.add(0, tets[0]);
	std::vector<int> bad;
	std::vector<std::array<int, 3>> faces;   // the cavity's faces; sorted, a face met once is on its boundary
	size_t n_alive = 1;
	for (int i = 0; i < n; i++)
	{
		bad.clear();
		auto inside = [&](int t) {
			const double dx = P[i][0] - tets[t].c[0], dy = P[i][1] - tets[t].c[1], dz = P[i][2] - tets[t].c[2];
			return dx * dx + dy * dy + dz * dz <= tets[t].r2 * (1.0 + 1e-12);
		};
		if (use_grid)
		{
			for (std::vector<int>* list : { &grid.at(P[i].data()), &grid.big })
			{
				size_t w = 0;
				for (size_t r = 0; r < list->size(); r++)
				{
					const int t = (*list)[r];
					if (!tets[t].alive) continue;   // dropped from the list
					(*list)[w++] = t;
					if (inside(t)) bad.push_back(t);
				}
				list->resize(w);
			}
			std::sort(bad.begin(), bad.end());      // the order of the scan
		}
		else
		{
			for (int t = 0; t < (int)tets.size(); t++)
				if (tets[t].alive && inside(t))
					bad.push_back(t);
		}
		faces.clear();
		for (int t : bad)
		{
			const int* v = tets[t].v;
			const int f[4][3] = { { v[0], v[1], v[2] }, { v[0], v[1], v[3] }, { v[1], v[2], v[3] }, { v[2], v[0], v[3] } };
			for (const auto& tri : f)
			{
				std::array<int, 3> key = { tri[0], tri[1], tri[2] };
				std::sort(key.begin(), key.end());
				faces.push_back(key);
			}
			tets[t].alive = false;
		}
		n_alive -= bad.size();
		std::sort(faces.begin(), faces.end());   // new tets are made in ascending face order
		for (size_t a = 0; a < faces.size();)
		{
			size_t b = a + 1;
			while (b < faces.size() && faces[b] == faces[a]) b++;
			if (b - a == 1)   // boundary of the cavity
			{
				tets.push_back(make(faces[a][0], faces[a][1], faces[a][2], i));
				if (use_grid) grid.add((int)tets.size() - 1, tets.back());
				n_alive++;
			}
			a = b;
		}
		if (tets.size() > 4096 && tets.size() > 8 * n_alive)
		{
			tets.erase(std::remove_if(tets.begin(), tets.end(), [](const Tet& t) { return !t.alive; }), tets.end());
			if (use_grid)   // the ids changed: list the survivors again
			{
				grid.init(lo, hi, n);
				for (int t = 0; t < (int)tets.size(); t++) grid.add(t, tets[t]);
			}
		}
	}
	tets.erase(std::remove_if(tets.begin(), tets.end(),
							  [&](const Tet& t) { return !t.alive || t.v[0] >= n || t.v[1] >= n || t.v[2] >= n || t.v[3] >= n; }),
			   tets.end());
	return tets;
}
} // namespace

namespace detail
{
// Test hook: the tets (point indices, in construction order) of the accelerated or of the plain-scan triangulation.
std::vector<std::array<int, 4>> TetIndices(const std::vector<Vector3>& points, bool use_grid)
{
	std::vector<std::array<int, 4>> out;
	if (points.size() >= 4)
		for (const Tet& t : bowyer_watson(points, use_grid))
			out.push_back({ t.v[0], t.v[1], t.v[2], t.v[3] });
	return out;
}
} // namespace detail

Delaunay Triangulate(const std::vector<Vector3>& points)
{
	Delaunay dt;
	if (points.size() < 3)
		return dt;
	for (const Tet& t : bowyer_watson(points))
		dt.TetVec.emplace_back(points[t.v[0]], points[t.v[1]], points[t.v[2]], points[t.v[3]]);
	for (const Tetrahedron& tet : dt.TetVec)   // DT3D.h:259-264
	{
		dt.FaceVec.push_back(tet.t0);
		dt.FaceVec.push_back(tet.t1);
		dt.FaceVec.push_back(tet.t2);
	}
	return dt;
}

std::vector<Edge> Voronoi(const Delaunay& dt)
{
	// one edge per pair of tets sharing a face (by point values), first occurrence kept (DT3D.h:269-315)
	auto key3 = [](const Vector3& a, const Vector3& b, const Vector3& c) {
		std::array<std::array<uint32_t, 3>, 3> k;
		const Vector3* p[3] = { &a, &b, &c };
		for (int i = 0; i < 3; i++) std::memcpy(k[i].data(), &p[i]->x, 12);
		std::sort(k.begin(), k.end());
		return k;
	};
	std::map<std::array<std::array<uint32_t, 3>, 3>, std::vector<int>> owners;
	for (int i = 0; i < (int)dt.TetVec.size(); i++)
	{
		const Tetrahedron& t = dt.TetVec[i];
		for (const Triangle* f : { &t.t0, &t.t1, &t.t2, &t.t3 })
			owners[key3(f->p0, f->p1, f->p2)].push_back(i);
	}
	std::vector<Edge> edges;
	std::set<std::pair<int, int>> seen;
	for (int i = 0; i < (int)dt.TetVec.size(); i++)
	{
		const Tetrahedron& t = dt.TetVec[i];
		for (const Triangle* f : { &t.t0, &t.t1, &t.t2, &t.t3 })
			for (int j : owners[key3(f->p0, f->p1, f->p2)])
				if (j != i && seen.insert({ std::min(i, j), std::max(i, j) }).second)
					edges.emplace_back(dt.TetVec[i].sphere.center, dt.TetVec[j].sphere.center);
	}
	return edges;
}

void Neighbors(const std::vector<Vector3>& points, std::vector<uint32_t>& off, std::vector<uint32_t>& idx)
{
	const size_t n = points.size();
	std::vector<std::vector<uint32_t>> nb(n);
	if (n >= 4)
		for (const Tet& t : bowyer_watson(points))
			for (int a = 0; a < 4; a++)
				for (int b = 0; b < 4; b++)
					if (a != b) nb[t.v[a]].push_back((uint32_t)t.v[b]);
	off.assign(1, 0);
	idx.clear();
	for (size_t i = 0; i < n; i++)
	{
		std::sort(nb[i].begin(), nb[i].end());
		nb[i].erase(std::unique(nb[i].begin(), nb[i].end()), nb[i].end());
		idx.insert(idx.end(), nb[i].begin(), nb[i].end());
		off.push_back((uint32_t)idx.size());
	}
}

std::vector<VMACH::Polygon3D> VoronoiCells(const std::vector<Vector3>& seeds)
{
	std::vector<uint32_t> off, idx;
	Neighbors(seeds, off, idx);
	SurtrHost::detail::FlatPolys box;
	box.add(Poly::GetBB());
	SurtrHost::detail::FlatCells cells;
	for (size_t i = 0; i < seeds.size(); i++)
	{
		std::vector<Poly::Plane> planes;
		for (uint32_t k = off[i]; k < off[i + 1]; k++)
		{
			const Vector3& sj = seeds[idx[k]];
			planes.emplace_back((seeds[i] + sj) * 0.5f, sj - seeds[i]);   // Plane(point, normal): outward, unnormalised
		}
		cells.add(planes);
	}
	SurtrHost::detail::Fragments fr;
	SurtrHost::detail::run_event(box, cells, fr);
	std::vector<VMACH::Polygon3D> out(seeds.size(), VMACH::Polygon3D(true));
	for (size_t f = 0; f < fr.rec.size(); f++)
	{
		const Poly::Polyhedron cell = fr.polyhedron(f);
		Poly::Extract* loops = Poly::ExtractFaces(cell);
		VMACH::Polygon3D poly(true);
		for (const std::vector<int>& loop : *loops)
		{
			VMACH::PolygonFace face(true);
			for (int v : loop)
				face.AddVertex(cell[v].Position);   // Plane(v0, v1, v2) appears with the third vertex: outward
			poly.AddFace(face);
		}
		delete loops;
		out[fr.rec[f].cell] = poly;
	}
	return out;
}
} // namespace DT3D
