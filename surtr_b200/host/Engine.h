// Engine.h -- internal glue between the host-side mirror classes and the C ABI (include/surtr_b200.h):
// AoS (Poly::Polyhedron, VMACH::Polygon3D) <-> the flat layout, and one lazily created context per host thread.
#pragma once

#include "../../include/surtr_b200.h"
#include "Poly.h"
#include "VMACH.h"

#include <cstdint>
#include <functional>
#include <vector>

namespace SurtrHost
{
namespace detail
{
struct FlatPolys
{
	std::vector<float> verts4;
	std::vector<uint32_t> vert_off{ 0 }, ring_off{ 0 };
	std::vector<uint16_t> ring;
	std::vector<uint32_t> ev_off;     // optional: independent events in one batch (empty = one event)
	void add(const Poly::Polyhedron& p);
	uint32_t count() const { return (uint32_t)vert_off.size() - 1; }
};

struct FlatCells
{
	std::vector<float> planes4;
	std::vector<uint32_t> plane_off{ 0 };
	std::vector<float> cverts4;       // every vertex of every face loop (bounds for the broad phase)
	std::vector<uint32_t> cvert_off{ 0 };
	bool bounded = true;
	std::vector<uint32_t> ev_off;     // optional, as in FlatPolys
	void add(const VMACH::Polygon3D& cell);
	void add_keep_all();              // a cell without planes whose bounds contain everything: returns pieces uncut
	void add(const std::vector<Poly::Plane>& planes);   // unbounded cell (plain plane list)
	uint32_t count() const { return (uint32_t)plane_off.size() - 1; }
};

// A cell set in its own frame, kept on the device and placed per event (surtr_upload_pattern / surtr_place_pattern):
// the VertexVec of every face of every cell.
struct FlatPattern
{
	std::vector<float> face_verts4;
	std::vector<uint32_t> face_vert_off{ 0 }, cell_face_off{ 0 };
	uint64_t id = 0;                  // 0 = empty; a fresh id per build, so a context knows which pattern it holds
	void build(const std::vector<VMACH::Polygon3D>& cells);
	uint32_t count() const { return (uint32_t)cell_face_off.size() - 1; }
};

struct Fragments
{
	std::vector<surtr_fragment> rec;
	std::vector<float> verts4;
	std::vector<uint32_t> ring_off;
	std::vector<uint16_t> ring;
	Poly::Polyhedron polyhedron(size_t i) const;
};

// Host worker pool for the per-piece bookkeeping around an event (the reference's g_threadPool, Surtr.cpp:28: one task
// per piece for refitting, :2405-2413).  Calls fn(i) for every i in [0, n), the caller's thread included; returns when
// all are done; the first exception thrown by a task is rethrown here.  Small n runs inline.
void parallel_for(size_t n, const std::function<void(size_t)>& fn);

// One context per host thread and slot, created on first use (throws std::runtime_error without a B200).  Slot 1 exists so
// that two independent events of one orchestration step (the convex and the mesh clip of ApplyFracture, Surtr.cpp:1461
// and :1470) are in flight at the same time, on two streams.
surtr_ctx* context(int slot = 0);
void check(int rc, const char* what);                            // throws std::runtime_error with surtr_last_error
// Makes `pattern` the resident pattern of the slot's context (no-op when it already is) and installs it, scaled and
// translated on the device, as the cell set of the next event (Polygon3D::Scale + Translate, VMACH.cpp:506-534).
void place_pattern(const FlatPattern& pattern, const DirectX::SimpleMath::Vector3& scale, const DirectX::SimpleMath::Vector3& translate, int slot = 0);
// An event in two halves: begin_event uploads and launches (returns while the GPU works), end_event waits and downloads.
// upload_cells = false: the cells of the previous event on the slot's context are still resident and are reused.
void begin_event(const FlatPolys& pieces, const FlatCells& cells, bool upload_cells = true, int slot = 0);
void end_event(Fragments& out, bool geometry = true, int slot = 0);
inline void run_event(const FlatPolys& pieces, const FlatCells& cells, Fragments& out, bool geometry = true, bool upload_cells = true)
{
	begin_event(pieces, cells, upload_cells, 0);
	end_event(out, geometry, 0);
}
} // namespace detail
} // namespace SurtrHost
