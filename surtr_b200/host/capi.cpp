// capi.cpp -- flat C entry points of libsurtr_host.so for callers that are not C++ (bench.py, surtr_b200/hostlib.py).
//
// Pattern generation at batch scale: the reference builds ONE pattern per call
// (Surtr::GenerateVoronoi, Src/Surtr.cpp:1984-2070); BASELINE config 4 needs 4096 x (1000 + 64) cells, so the two
// host-side steps of the derivation run over many seed sets at once on the host worker pool:
//   surtr_host_dt3d_neighbors_batch   DT3D::Neighbors (Delaunay neighbour lists, ascending) per seed set
//   surtr_host_face_planes            the FacePlane of every face of every polyhedron, the way a VMACH::Polygon3D built
//                                     from Poly::ExtractFaces loops holds them (PolygonFace::AddVertex, VMACH.cpp:289-310)
// The cutting between the two (container box x bisector half-spaces) is one GPU event through the C ABI.
#include "DT3D.h"
#include "Engine.h"
#include "Poly.h"
#include "VMACH.h"

#include <cstring>
#include <string>
#include <vector>

using DirectX::SimpleMath::Plane;
using DirectX::SimpleMath::Vector3;

namespace
{
thread_local std::string g_error;
}

extern "C"
{
const char* surtr_host_last_error() { return g_error.c_str(); }

// seeds3: all seed sets back to back (xyz); set_off[n_sets + 1].  nb_off[total_seeds + 1] is ONE CSR over all seeds,
// nb_idx holds neighbour indices LOCAL to the seed's set, ascending.  Returns the number of neighbour entries; when it
// exceeds cap_idx nothing was written to nb_idx (call again with a larger buffer).  0 with a message on failure.
uint64_t surtr_host_dt3d_neighbors_batch(const float* seeds3, const uint32_t* set_off, uint32_t n_sets, uint32_t* nb_off,
										 uint32_t* nb_idx, uint64_t cap_idx)
{
	try
	{
		std::vector<std::vector<uint32_t>> off(n_sets), idx(n_sets);
		SurtrHost::detail::parallel_for(n_sets, [&](size_t s) {
			const uint32_t a = set_off[s], b = set_off[s + 1];
			std::vector<Vector3> pts(b - a);
			for (uint32_t i = a; i < b; i++)
				pts[i - a] = Vector3(seeds3[3 * (size_t)i], seeds3[3 * (size_t)i + 1], seeds3[3 * (size_t)i + 2]);
			DT3D::Neighbors(pts, off[s], idx[s]);
		});
		uint64_t total = 0;
		for (uint32_t s = 0; s < n_sets; s++)
		{
			const uint32_t a = set_off[s], n = set_off[s + 1] - a;
			for (uint32_t i = 0; i < n; i++)
				nb_off[a + i] = (uint32_t)(total + off[s][i]);
			total += idx[s].size();
		}
		nb_off[set_off[n_sets]] = (uint32_t)total;
		if (total > cap_idx)
			return total;
		uint64_t at = 0;
		for (uint32_t s = 0; s < n_sets; s++)
		{
			if (!idx[s].empty())
				std::memcpy(nb_idx + at, idx[s].data(), 4 * idx[s].size());
			at += idx[s].size();
		}
		return total;
	}
	catch (const std::exception& e)
	{
		g_error = e.what();
		return 0;
	}
}

// Polyhedra in the flat layout of include/surtr_b200.h.  plane_off[n_poly + 1] and planes4 (4 floats per face) are
// written; faces come in Poly::ExtractFaces order (Poly.cpp:89-126), each plane = Plane(v0, v1, v2) of the first three
// loop vertices AddVertex keeps (it drops a vertex closer than 1e-12 to one already in the face); a face left with fewer
// than three vertices gets the never-constructed default plane (0, 1, 0, 0).  Returns the number of faces; when it
// exceeds cap_planes only plane_off is valid.
uint64_t surtr_host_face_planes(const float* verts4, const uint32_t* vert_off, const uint32_t* ring_off, const uint16_t* ring,
								uint32_t n_poly, float* planes4, uint32_t* plane_off, uint64_t cap_planes)
{
	try
	{
		// pass 1: faces per polyhedron (Euler would do for closed genus-0 rings; counting loops is what ExtractFaces does)
		std::vector<std::vector<float>> out(n_poly);
		SurtrHost::detail::parallel_for(n_poly, [&](size_t p) {
			const uint32_t v0 = vert_off[p], nv = vert_off[p + 1] - v0;
			std::vector<uint8_t> seen(ring_off[v0 + nv] - ring_off[v0], 0);
			const uint32_t e0 = ring_off[v0];
			auto deg = [&](uint32_t v) { return ring_off[v0 + v + 1] - ring_off[v0 + v]; };
			auto nb = [&](uint32_t v, uint32_t k) { return (uint32_t)ring[ring_off[v0 + v] + k]; };
			auto slot_of = [&](uint32_t from, uint32_t to) {
				const uint32_t d = deg(from);
				for (uint32_t k = 0; k < d; k++)
					if (nb(from, k) == to)
						return k;
				return d;
			};
			std::vector<float>& planes = out[p];
			std::vector<Vector3> kept;
			for (uint32_t v = 0; v < nv; v++)
				for (uint32_t s = 0; s < deg(v); s++)
				{
					if (seen[ring_off[v0 + v] - e0 + s])
						continue;
					seen[ring_off[v0 + v] - e0 + s] = 1;
					kept.clear();
					auto add = [&](uint32_t u) {
						if (kept.size() >= 3)
							return;
						const Vector3 q(verts4[4 * (size_t)(v0 + u)], verts4[4 * (size_t)(v0 + u) + 1], verts4[4 * (size_t)(v0 + u) + 2]);
						for (const Vector3& k : kept)
							if (VMACH::NearlyEqual(k, q))
								return;
						kept.push_back(q);
					};
					add(v);
					uint32_t prev = v, cur = nb(v, s);
					size_t guard = 0;
					while (cur != v && guard++ < (size_t)nv * 64)
					{
						add(cur);
						const uint32_t d = deg(cur), k = slot_of(cur, prev);
						const uint32_t ks = (k == 0 || k >= d) ? d - 1 : k - 1;
						seen[ring_off[v0 + cur] - e0 + ks] = 1;
						prev = cur;
						cur = nb(cur, ks);
					}
					Plane pl(0.f, 1.f, 0.f, 0.f);
					if (kept.size() >= 3)
						pl = Plane(kept[0], kept[1], kept[2]);
					planes.insert(planes.end(), { pl.x, pl.y, pl.z, pl.w });
				}
		});
		uint64_t total = 0;
		for (uint32_t p = 0; p < n_poly; p++)
		{
			plane_off[p] = (uint32_t)total;
			total += out[p].size() / 4;
		}
		plane_off[n_poly] = (uint32_t)total;
		if (total > cap_planes)
			return total;
		for (uint32_t p = 0; p < n_poly; p++)
			if (!out[p].empty())
				std::memcpy(planes4 + 4 * (size_t)plane_off[p], out[p].data(), 4 * out[p].size());
		return total;
	}
	catch (const std::exception& e)
	{
		g_error = e.what();
		return 0;
	}
}
}
