#include <chrono>
#include <memory>
#include "Engine.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>

#include <stdexcept>
#include <string>

namespace SurtrHost
{
namespace detail
{
using DirectX::SimpleMath::Vector3;

void FlatPolys::add(const Poly::Polyhedron& p)
{
	for (const Poly::Vertex& v : p)
	{
		verts4.push_back(v.Position.x);
		verts4.push_back(v.Position.y);
		verts4.push_back(v.Position.z);
		verts4.push_back(0.f);
		for (int n : v.NeighborVertexVec)
			ring.push_back((uint16_t)n);
		ring_off.push_back((uint32_t)ring.size());
	}
	vert_off.push_back((uint32_t)(verts4.size() / 4));
}

void FlatCells::add(const VMACH::Polygon3D& cell)
{
	for (const VMACH::PolygonFace& f : cell.FaceVec)
	{
		// Poly::ClipPolyhedron(const Polyhedron&, const Polygon3D&) reads exactly FaceVec[i].FacePlane (Poly.cpp:558-560)
		planes4.push_back(f.FacePlane.x);
		planes4.push_back(f.FacePlane.y);
		planes4.push_back(f.FacePlane.z);
		planes4.push_back(f.FacePlane.w);
		for (const Vector3& v : f.VertexVec)
		{
			cverts4.push_back(v.x);
			cverts4.push_back(v.y);
			cverts4.push_back(v.z);
			cverts4.push_back(0.f);
		}
		if (f.VertexVec.empty())
			bounded = false;   // a face given only by its plane: no bound can be derived for this set
	}
	plane_off.push_back((uint32_t)(planes4.size() / 4));
	cvert_off.push_back((uint32_t)(cverts4.size() / 4));
}

void FlatCells::add_keep_all()
{
	// bounds only (the clipper never reads cell vertices): the eight corners of a box that contains any finite piece,
	// so that every slab direction of the broad phase sees the full range
	for (int k = 0; k < 8; k++)
	{
		cverts4.push_back((k & 1) ? 1.0e30f : -1.0e30f);
		cverts4.push_back((k & 2) ? 1.0e30f : -1.0e30f);
		cverts4.push_back((k & 4) ? 1.0e30f : -1.0e30f);
		cverts4.push_back(0.f);
	}
	plane_off.push_back((uint32_t)(planes4.size() / 4));
	cvert_off.push_back((uint32_t)(cverts4.size() / 4));
}

void FlatCells::add(const std::vector<Poly::Plane>& planes)
{
	for (const Poly::Plane& p : planes)
	{
		planes4.push_back(p.x);
		planes4.push_back(p.y);
		planes4.push_back(p.z);
		planes4.push_back(p.w);
	}
	plane_off.push_back((uint32_t)(planes4.size() / 4));
	cvert_off.push_back((uint32_t)(cverts4.size() / 4));
	bounded = false;
}

void FlatPattern::build(const std::vector<VMACH::Polygon3D>& cells)
{
	static std::atomic<uint64_t> next_id{ 1 };
	face_verts4.clear();
	face_vert_off.assign(1, 0u);
	cell_face_off.assign(1, 0u);
	for (const VMACH::Polygon3D& cell : cells)
	{
		for (const VMACH::PolygonFace& f : cell.FaceVec)
		{
			// The device placement re-derives every face plane from the face's first three moved vertices, which is what
			// Polygon3D::Scale / Translate do through ConstructFacePlane -- but only for GuaranteeConvex faces
			// (VMACH.cpp:303-310): a face that carries a hand-set plane keeps it in the reference, so it cannot take this
			// route (use the explicit cell list of ApplyFracture, which ships FacePlane as it is).
			if (!f.GuaranteeConvex || f.VertexVec.size() < 3)
				throw std::invalid_argument("FlatPattern: a resident pattern needs GuaranteeConvex faces of at least three vertices "
											"(their planes are re-derived on the device after every placement)");
			for (const Vector3& v : f.VertexVec)
				face_verts4.insert(face_verts4.end(), { v.x, v.y, v.z, 0.f });
			face_vert_off.push_back((uint32_t)(face_verts4.size() / 4));
		}
		cell_face_off.push_back((uint32_t)face_vert_off.size() - 1);
	}
	id = next_id++;
}

Poly::Polyhedron Fragments::polyhedron(size_t i) const
{
	const surtr_fragment& f = rec[i];
	Poly::Polyhedron p(f.n_verts);
	for (uint32_t k = 0; k < f.n_verts; k++)
	{
		const uint32_t v = f.vert_off + k;
		p[k].Position = Vector3(verts4[4 * v], verts4[4 * v + 1], verts4[4 * v + 2]);
		p[k].NeighborVertexVec.assign(ring.begin() + ring_off[v], ring.begin() + ring_off[v + 1]);
	}
	return p;
}

void check(int rc, const char* what)
{
	if (rc == SURTR_OK)
		return;
	surtr_ctx* c = nullptr;
	try { c = context(); } catch (...) {}
	throw std::runtime_error(std::string(what) + ": " + surtr_last_error(c));
}

namespace
{
void check_on(surtr_ctx* c, int rc, const char* what)
{
	if (rc != SURTR_OK)
		throw std::runtime_error(std::string(what) + ": " + surtr_last_error(c));
}
} // namespace

namespace
{
class WorkerPool
{
	// one object per parallel_for call: a worker that wakes up late finds either a finished job (nothing left to take)
	// or the current one, never a mixture of the two
	struct Job
	{
		const std::function<void(size_t)>* fn;
		size_t n;
		std::atomic<size_t> next{ 0 };
		size_t done = 0;              // guarded by the pool mutex
		std::exception_ptr error;     // likewise
	};

public:
	WorkerPool()
	{
		const unsigned hw = std::thread::hardware_concurrency();
		const unsigned n = std::min(16u, hw ? hw : 1u);   // the reference's pool has 16 threads
		for (unsigned i = 1; i < n; i++)
			m_threads.emplace_back([this] { loop(); });
	}
	~WorkerPool()
	{
		{
			std::lock_guard<std::mutex> lock(m_mutex);
			m_quit = true;
		}
		m_wake.notify_all();
		for (std::thread& t : m_threads)
			t.join();
	}
	void run(size_t n, const std::function<void(size_t)>& fn)
	{
		auto job = std::make_shared<Job>();
		job->fn = &fn;
		job->n = n;
		{
			std::lock_guard<std::mutex> lock(m_mutex);
			m_job = job;
			m_generation++;
		}
		m_wake.notify_all();
		work(*job);
		std::unique_lock<std::mutex> lock(m_mutex);
		m_done.wait(lock, [&] { return job->done == n; });   // every index was taken AND finished; fn stays alive until here
		if (m_job == job)
			m_job.reset();
		if (job->error)
			std::rethrow_exception(job->error);
	}

private:
	void work(Job& job)
	{
		size_t finished = 0;
		std::exception_ptr err;
		for (;;)
		{
			const size_t i = job.next.fetch_add(1);
			if (i >= job.n)
				break;
			try { (*job.fn)(i); }
			catch (...) { if (!err) err = std::current_exception(); }
			finished++;
		}
		if (!finished)
			return;
		{
			std::lock_guard<std::mutex> lock(m_mutex);
			job.done += finished;
			if (err && !job.error)
				job.error = err;
		}
		m_done.notify_all();
	}
	void loop()
	{
		uint64_t seen = 0;
		for (;;)
		{
			std::shared_ptr<Job> job;
			{
				std::unique_lock<std::mutex> lock(m_mutex);
				m_wake.wait(lock, [&] { return m_quit || m_generation != seen; });
				if (m_quit)
					return;
				seen = m_generation;
				job = m_job;
			}
			if (job)
				work(*job);
		}
	}
	std::vector<std::thread> m_threads;
	std::mutex m_mutex;
	std::condition_variable m_wake, m_done;
	std::shared_ptr<Job> m_job;
	uint64_t m_generation = 0;
	bool m_quit = false;
};
} // namespace

void parallel_for(size_t n, const std::function<void(size_t)>& fn)
{
	if (n < 8)
	{
		for (size_t i = 0; i < n; i++)
			fn(i);
		return;
	}
	static WorkerPool pool;
	pool.run(n, fn);
}

surtr_ctx* context(int slot)
{
	struct Holder
	{
		surtr_ctx* ctx[2] = { nullptr, nullptr };
		~Holder() { surtr_ctx_destroy(ctx[0]); surtr_ctx_destroy(ctx[1]); }
	};
	static thread_local Holder h;
	slot = slot ? 1 : 0;
	if (!h.ctx[slot])
	{
		const int rc = surtr_ctx_create(0, nullptr, &h.ctx[slot]);
		if (rc != SURTR_OK)
			throw std::runtime_error(std::string("surtr_ctx_create: ") + surtr_last_error(nullptr));   // no CPU fallback
	}
	return h.ctx[slot];
}

void place_pattern(const FlatPattern& pattern, const Vector3& scale, const Vector3& translate, int slot)
{
	static thread_local uint64_t resident_id[2] = { 0, 0 };
	slot = slot ? 1 : 0;
	surtr_ctx* c = context(slot);
	if (resident_id[slot] != pattern.id)
	{
		check(surtr_upload_pattern(c, pattern.face_verts4.data(), pattern.face_vert_off.data(), (uint32_t)pattern.face_vert_off.size() - 1,
								   pattern.cell_face_off.data(), pattern.count()), "surtr_upload_pattern");
		resident_id[slot] = pattern.id;
	}
	const float s3[3] = { scale.x, scale.y, scale.z }, t3[3] = { translate.x, translate.y, translate.z };
	check(surtr_place_pattern(c, s3, t3, 1), "surtr_place_pattern");
}

namespace
{
// SURTR_TRACE=1: wall time of the legs of an event (stderr), next to the orchestration phases of Fracture.cpp
struct Leg
{
	const char* name;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	explicit Leg(const char* n) : name(n) {}
	~Leg()
	{
		static const bool on = std::getenv("SURTR_TRACE") != nullptr;
		if (on)
			std::fprintf(stderr, "[surtr]       %-22s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
};
} // namespace

void begin_event(const FlatPolys& pieces, const FlatCells& cells, bool upload_cells, int slot)
{
	surtr_ctx* c = context(slot);
	Leg leg("upload + launch");
	check_on(c, surtr_upload_pieces(c, pieces.verts4.data(), pieces.vert_off.data(), pieces.ring_off.data(), pieces.ring.data(),
							  pieces.count(), pieces.ev_off.empty() ? nullptr : pieces.ev_off.data(),
							  pieces.ev_off.empty() ? 0u : (uint32_t)pieces.ev_off.size() - 1), "surtr_upload_pieces");
	if (upload_cells)
		check_on(c, surtr_upload_cells(c, cells.planes4.data(), cells.plane_off.data(), cells.bounded ? cells.cverts4.data() : nullptr,
								 cells.bounded ? cells.cvert_off.data() : nullptr, cells.count(),
								 cells.ev_off.empty() ? nullptr : cells.ev_off.data(),
								 cells.ev_off.empty() ? 0u : (uint32_t)cells.ev_off.size() - 1), "surtr_upload_cells");
	check_on(c, surtr_fracture_event(c), "surtr_fracture_event");
}

void end_event(Fragments& out, bool geometry, int slot)
{
	surtr_ctx* c = context(slot);
	std::unique_ptr<Leg> leg(new Leg("wait"));
	surtr_counts n;
	check_on(c, surtr_event_counts(c, &n), "surtr_event_counts");
	leg.reset(new Leg("download"));
	static const bool trace = std::getenv("SURTR_TRACE") != nullptr;
	if (trace)
	{
		float total_ms = 0.f;
		surtr_last_event_ms(c, &total_ms, nullptr);
		std::fprintf(stderr, "[surtr]     event: %llu pairs, %llu candidates (large tier %llu, global tier %llu), %llu fragments, %.3f ms on the device\n",
					 (unsigned long long)n.n_pairs, (unsigned long long)n.n_candidates, (unsigned long long)n.n_tier2, (unsigned long long)n.n_tier3,
					 (unsigned long long)n.n_fragments, total_ms);
	}
	out.rec.resize(n.n_fragments);
	if (geometry)
	{
		out.verts4.resize(4 * n.n_verts);
		out.ring_off.resize(n.n_verts + 1);
		out.ring.resize(n.n_ring);
	}
	check_on(c, surtr_download_fragments(c, out.rec.data(), geometry ? out.verts4.data() : nullptr,
								   geometry ? out.ring_off.data() : nullptr, geometry ? out.ring.data() : nullptr),
		  "surtr_download_fragments");
}
} // namespace detail
} // namespace SurtrHost
