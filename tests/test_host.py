"""Host-side mirror of the reference interfaces (surtr_b200/host: Poly / Kdop / VMACH / DT3D / SurtrHost).
CPU tests cover the pure host logic; GPU tests drive the class-level API end to end against the oracle."""
import os

import numpy as np
import pytest

import common
import hostapi as H
from common import GOLDEN, bits
from oracle import portapi as P
from test_oracle_port import load_polyset


def test_host_seeds_match_reference_recipe():
    assert np.array_equal(H.seeds(46354, 4096), common.seeds_uniform(46354, 4096))


def test_host_scalar_helpers_match_reference():
    d = np.load(os.path.join(GOLDEN, "kat_scalar.npz"))
    n = len(d["planes"])
    assert np.array_equal(np.array([H.compare_plane_point(d["planes"][i], d["pts"][i]) for i in range(n)], np.int32), d["comp"])
    inter = np.stack([H.plane_line_intersection(d["a"][i], d["b"][i], d["planes"][i]) for i in range(n)])
    assert np.array_equal(bits(inter), bits(d["inter"]))
    assert np.array_equal(bits(H.box_planes()), bits(d["box_planes"]))     # VMACH::GetBoxPolygon, incl. -0.0 signs


def test_host_dt3d_against_reference_and_qhull():
    d = np.load(os.path.join(GOLDEN, "cube_x64.npz"))
    off, idx = H.dt3d_neighbors(d["seeds"])
    ref = [set(d["nb_idx"][d["nb_off"][i]:d["nb_off"][i + 1]]) for i in range(64)]      # the reference's own DT3D
    mine = [set(idx[off[i]:off[i + 1]]) for i in range(64)]
    assert all(r <= m for r, m in zip(ref, mine))          # a far super-tetrahedron only ADDS hull-adjacent edges
    for n in (64, 256, 2000):
        s = common.seeds_uniform(46354, n)
        off, idx = H.dt3d_neighbors(s)
        o2, i2 = common.scipy_neighbors(s)
        assert np.array_equal(off, o2) and np.array_equal(idx, i2)
    nt, nf, ne, viol = H.dt3d_triangulate(common.seeds_uniform(46354, 256))
    assert viol == 0 and nf == 3 * nt and nt > 1400 and ne > nt
    assert H.dt3d_triangulate(common.seeds_uniform(1, 2))[0] == 0     # < 3 points -> empty (DT3D.h:161-162)


def test_host_dt3d_grid_search_equals_the_full_scan():
    """The grid-accelerated cavity search of the host DT3D (surtr_b200/host/DT3D.cpp, SphereGrid) must produce the
    triangulation of the plain scan over every live tet (the reference's search, Inc/DT3D.h:198-246) TET FOR TET, in
    the same order -- on uniform seeds, clustered seeds, a lattice (co-spherical points), coplanar points, duplicates
    and the exponential radial pattern of GenerateFracturePattern."""
    rng = np.random.RandomState(7)
    cases = {
        "uniform_5": common.seeds_uniform(3, 5),
        "uniform_64": common.seeds_uniform(46354, 64),
        "uniform_1000": common.seeds_uniform(1234, 1000),
        "uniform_4096": common.seeds_uniform(46354, 4096),
        "clustered": (rng.normal(0, 0.02, (600, 3)) + rng.choice([-0.3, 0.0, 0.3], (600, 1))).astype(np.float32),
        "lattice_6x6x6": (np.stack(np.meshgrid(*[np.arange(6)] * 3, indexing="ij"), -1).reshape(-1, 3) / 5.0 - 0.5).astype(np.float32),
        "coplanar": np.concatenate([rng.uniform(-0.5, 0.5, (200, 2)), np.zeros((200, 1))], 1).astype(np.float32),
        "duplicates": np.repeat(common.seeds_uniform(5, 100), 2, axis=0),
        "radial": (rng.standard_normal((800, 3)) * rng.exponential(0.05, (800, 1))).astype(np.float32),
        "far_from_origin": common.seeds_uniform(9, 300) * np.float32(1e-3) + np.float32(1000.0),
    }
    for name, s in cases.items():
        a, b = H.dt3d_tets(s, True), H.dt3d_tets(s, False)
        assert a.shape == b.shape and np.array_equal(a, b), name
    assert len(H.dt3d_tets(common.seeds_uniform(46354, 4096), True)) > 26000


def test_host_extract_faces_matches_reference_loops():
    d = np.load(os.path.join(GOLDEN, "cube_x64.npz"))
    fr = load_polyset(d, "frag_")
    foff, fidx = d["frag_face_off"], d["frag_face_idx"]
    f0 = 0
    for i in range(fr.n):
        v0, v1 = int(fr.vert_off[i]), int(fr.vert_off[i + 1])
        r0, r1 = int(fr.ring_off[v0]), int(fr.ring_off[v1])
        got = H.extract_faces(fr.verts[v0:v1], fr.ring_off[v0:v1 + 1], fr.ring[r0:r1])
        nf = int(fr.nfaces[i])
        want = [fidx[foff[f]:foff[f + 1]].tolist() for f in range(f0, f0 + nf)]
        assert got == want
        f0 += nf


@pytest.mark.gpu
def test_host_apply_fracture_class_api():
    for name in ("cube_x64", "pieces200_x32"):
        d = np.load(os.path.join(GOLDEN, name + ".npz"))
        cells, pieces, want = load_polyset(d, "cells_"), load_polyset(d, "pieces_"), load_polyset(d, "frag_")
        got = H.apply_fracture(pieces, cells)
        for f in ("verts", "vert_off", "ring_off", "ring", "cell", "piece", "nfaces", "volume", "centroid"):
            assert np.array_equal(bits(getattr(got, f)), bits(getattr(want, f))), (name, f)


@pytest.mark.gpu
def test_host_voronoi_cells_and_clip_moments_kdop():
    s = common.seeds_uniform(46354, 300)
    planes, off = H.voronoi_planes(s)
    o2, i2 = common.scipy_neighbors(s)
    want = P.voronoi_cells(s, o2, i2)
    assert np.array_equal(off, want.plane_off) and np.array_equal(bits(planes), bits(want.planes))
    # Poly::ClipPolyhedron (in place) + Poly::Moments
    cube = common.unit_cube()
    pl = P.plane_from_point_normal([0.1, 0.05, 0.0], [1, 1, 1])
    got, vol, cen = H.clip_and_moments(cube, pl[None])
    w = P.clip_each(cube, pl[None], [0, 1])
    assert np.array_equal(bits(got.verts), bits(w.verts)) and np.array_equal(got.ring, w.ring)
    assert vol == w.volume[0] and np.array_equal(bits(cen), bits(w.centroid[0]))
    # Kdop::KdopContainer::Calc(vertices, maxAxisScale, gapInv) + ClipWithPolyhedron = the ACH of config 1
    d = np.load(os.path.join(GOLDEN, "config1_kdop.npz"))
    for key in ("bunny", "cube"):
        v4 = d[key + "_verts"]
        ext = v4[:, :3].max(0).astype(np.float64) - v4[:, :3].min(0).astype(np.float64)
        planes, ach = H.kdop_ach(v4, d[key + "_normals"], float(ext.max()), 2000.0, d[key + "_seedbox_verts"])
        assert np.array_equal(bits(planes), bits(d[key + "_gap_planes"]))
        wa = load_polyset(d, key + "_ach_")
        assert np.array_equal(bits(ach.verts), bits(wa.verts)) and np.array_equal(ach.ring, wa.ring)


def test_host_ich_normals_match_reference():
    """VMACH::ConvexHull mirror (greedy incremental hull): same faces in the same order as the reference build."""
    d = np.load(os.path.join(GOLDEN, "config1_kdop.npz"))
    for key, n in (("bunny", 28), ("cube", 12), ("sphere", 36)):
        got = H.ich_normals(d[key + "_verts"], 20)
        assert len(got) == n and np.array_equal(bits(got), bits(d[key + "_normals"])), key
    r = np.load(os.path.join(GOLDEN, "refit96.npz"))
    off = r["mesh_vert_off"]
    for i in range(len(off) - 1):
        pts = r["mesh_verts"][off[i]:off[i + 1]]
        got = H.ich_normals(pts, min(len(pts), 4))
        want = r["ich_normals"][r["ich_normal_off"][i]:r["ich_normal_off"][i + 1]]
        assert np.array_equal(bits(got), bits(want)), i


@pytest.mark.skipif(not common.have_ref(), reason="needs the reference build (oracle/_ref)")
def test_host_ich_on_near_degenerate_clouds_matches_reference_live():
    """The index-based hull against the reference build, live, where the reference's edge keys collide: points closer
    than the six decimals std::to_string prints (same key for different edges, key 0 for an edge between two such
    points), exact duplicates, coplanar and collinear runs, and plain random clouds at several limits."""
    from oracle import refapi as R
    rng = np.random.RandomState(11)
    clouds = []
    for n in (5, 8, 30, 200):
        clouds.append(rng.uniform(-1, 1, (n, 3)))
    base = rng.uniform(-1, 1, (12, 3))
    clouds.append(np.concatenate([base, base + 3e-7, base[:5] - 2e-7]))                      # near-duplicates: print alike
    clouds.append(np.concatenate([base, base[:6]]))                                            # exact duplicates
    clouds.append(np.concatenate([base, np.c_[rng.uniform(-1, 1, (10, 2)), np.zeros(10)]]))    # a coplanar run
    clouds.append(np.concatenate([base, np.outer(np.linspace(-1, 1, 7), [1, 2, 3]) * 0.3]))    # a collinear run
    clouds.append(np.round(rng.uniform(-1, 1, (40, 3)), 1))                                    # lattice points: many ties
    for ci, c in enumerate(clouds):
        v = np.zeros((len(c), 4), np.float32)
        v[:, :3] = c
        for limit in (4, 6, 20, 0):
            lim = min(limit, len(v)) if limit else 0
            want = R.ich_normals(v, lim)
            got = H.ich_normals(v, lim)
            assert got.shape == want.shape and np.array_equal(bits(got), bits(want)), (ci, limit)


@pytest.mark.gpu
def test_host_refitting_matches_reference():
    """SurtrHost::Refitting (m_refittingTask, Surtr.cpp:1449-1455) batched on the GPU == the reference per piece."""
    r = np.load(os.path.join(GOLDEN, "refit96.npz"))
    convex, want = load_polyset(r, "convex_"), load_polyset(r, "out_")
    got = H.refit(convex, r["mesh_verts"], r["mesh_vert_off"], 4)
    assert np.array_equal(got.vert_off, want.vert_off)
    assert np.array_equal(bits(got.verts), bits(want.verts)) and np.array_equal(got.ring, want.ring)


@pytest.mark.gpu
def test_kdop_calc_batch_matches_oracle(ctx):
    r = np.load(os.path.join(GOLDEN, "refit96.npz"))
    dist, arg, planes = ctx.kdop_calc_batch(r["mesh_verts"], r["mesh_vert_off"], r["ich_normals"], r["ich_normal_off"])
    off, noff = r["mesh_vert_off"], r["ich_normal_off"]
    for i in range(len(off) - 1):
        d, a, p = P.kdop_calc(r["mesh_verts"][off[i]:off[i + 1]], r["ich_normals"][noff[i]:noff[i + 1]])
        sl = slice(noff[i], noff[i + 1])
        assert np.array_equal(bits(dist[sl]), bits(d)) and np.array_equal(arg[sl], a) and np.array_equal(bits(planes[sl]), bits(p))


@pytest.mark.gpu
def test_config1_bunny_convex_branch():
    """BASELINE config 1 (bundled bunny, 32 seeds), convex branch of PrepareFracture through the host classes:
    ICH -> k-DOP -> ACH -> DT3D cells scaled/translated onto the object -> ApplyFracture == the reference build."""
    d = np.load(os.path.join(GOLDEN, "config1_bunny32.npz"))
    got = H.config1(d["verts"], d["seeds"])
    want = load_polyset(d, "frag_")
    assert got.ach_nv == int(load_polyset(d, "ach_").nverts[0]) == 107
    assert got.n == want.n == 28
    for f in ("verts", "vert_off", "ring_off", "ring", "cell", "piece", "nfaces", "volume", "centroid"):
        assert np.array_equal(bits(getattr(got, f)), bits(getattr(want, f))), f


def test_host_mesh_polyhedron_matches_reference():
    """Poly::ExtractNeighborFromMesh mirror (host, index bookkeeping) == the reference build's rings for the bunny
    (2503 vertices, 4968 triangles), order included; then CheckMeshIsland finds one island."""
    d = np.load(os.path.join(GOLDEN, "config1_full_bunny32.npz"))
    want = load_polyset(np.load(os.path.join(GOLDEN, "bunny_mesh_x32.npz")), "mesh_")
    got = H.mesh_polyhedron(d["verts"], d["indices"])
    assert np.array_equal(got.ring_off, want.ring_off) and np.array_equal(got.ring, want.ring)
    assert np.array_equal(bits(got.verts), bits(want.verts))
    # a triangle list whose fan orientation is inconsistent must throw like the reference does
    bad = d["indices"].copy().reshape(-1, 3)
    bad = bad[: len(bad) // 2]          # half a surface: open fans still give symmetric rings or throw, never crash
    try:
        H.mesh_polyhedron(d["verts"], bad)
    except RuntimeError as e:
        assert "ExtractNeighborFromMesh" in str(e)


@pytest.mark.gpu
def test_config1_bunny_full_prepare_fracture():
    """BASELINE config 1 in full through the host classes: mesh polyhedron -> ApplyFracture with the mesh branch (two GPU
    events, the 2503-vertex mesh in the global-memory tier, island split on the host) -> Refitting -> SetExtract ==
    the reference build (ref_config1_full), piece by piece."""
    d = np.load(os.path.join(GOLDEN, "config1_full_bunny32.npz"))
    convex, mesh, ach_nv = H.config1_full(d["verts"], d["indices"], d["seeds"])
    want_c, want_m = load_polyset(d, "convex_"), load_polyset(d, "mesh_")
    assert ach_nv == 107 and convex.n == want_c.n == 27
    for f in ("verts", "vert_off", "ring_off", "ring", "cell", "piece"):
        assert np.array_equal(bits(getattr(convex, f)), bits(getattr(want_c, f))), "convex " + f
        assert np.array_equal(bits(getattr(mesh, f)), bits(getattr(want_m, f))), "mesh " + f
    assert np.array_equal(convex.nfaces, want_c.nfaces)     # SetExtract after the refit
    keys = list(zip(want_c.cell.tolist(), want_c.piece.tolist()))
    assert len(keys) - len(set(keys)) == 2                   # two (cell, piece) pairs split into islands


def test_host_combine_mass_parallel_axis():
    """CombineMass = what PxRigidBodyExt::updateMassAndInertia(body, 10) derives (Surtr.cpp:2520): eight octant boxes of
    the unit cube, each with its analytic inertia about its own centre, must combine to the unit cube's mass 10,
    centre 0 and inertia 10/6 * identity."""
    cen = np.array([[x, y, z] for x in (-.25, .25) for y in (-.25, .25) for z in (-.25, .25)], np.float32)
    vol = np.full(8, 0.125)
    i_box = 0.125 * (0.25 + 0.25) / 12.0                      # V (b^2 + c^2) / 12 of a 0.5-cube at unit density
    inertia = np.tile(np.array([i_box, i_box, i_box, 0, 0, 0], np.float32), (8, 1))
    out = H.combine_mass(vol, cen, inertia, 10.0)
    assert abs(out[0] - 10.0) < 1e-6 and np.abs(out[1:4]).max() < 1e-7
    assert np.allclose(out[4:7], 10.0 / 6.0, rtol=1e-6) and np.abs(out[7:]).max() < 1e-7
    # an off-centre pair: products of inertia follow the sign convention Ixy = -sum m x y
    out = H.combine_mass([1.0, 1.0], [[1, 1, 0], [-1, -1, 0]], np.zeros((2, 6), np.float32), 1.0)
    assert np.allclose(out[4:], [2, 2, 4, -2, 0, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["general", "partial"])
def test_do_fracture_bunny(mode):
    """Row f-3: SurtrHost::DoFracture (pattern placement, ApplyFracture incl. partial mode and mesh branch, SetExtract,
    MergeOutOfImpact, HandleConvexIsland, Refitting) on the 27-piece bunny compound == the restatement over the
    reference build (ref_do_fracture): same pieces in the same order, bit for bit, and the same compounds."""
    d = np.load(os.path.join(GOLDEN, "do_fracture_bunny.npz"))
    d0 = np.load(os.path.join(GOLDEN, "config1_full_bunny32.npz"))
    convex, mesh = load_polyset(d0, "convex_"), load_polyset(d0, "mesh_")
    want_c, want_m = load_polyset(d, mode + "_convex_"), load_polyset(d, mode + "_mesh_")
    got_c, got_m, ncomp, mass = H.do_fracture(convex, mesh, d[mode + "_seeds"], d["cloud"], d["impact"], float(d[mode + "_radius"]),
                                              float(d["max_axis_scale"]), mode == "partial")
    assert ncomp == int(d[mode + "_ncomp"]) and got_c.n == want_c.n
    for f in ("verts", "vert_off", "ring_off", "ring", "cell", "piece"):
        assert np.array_equal(bits(getattr(got_c, f)), bits(getattr(want_c, f))), "convex " + f
        assert np.array_equal(bits(getattr(got_m, f)), bits(getattr(want_m, f))), "mesh " + f
    assert np.array_equal(got_c.nfaces, want_c.nfaces)
    if mode == "partial":
        assert int(want_c.piece.sum()) == 20 and np.all(want_c.cell[want_c.piece == 1] == 0)   # untouched pieces stay in bind[0]
    # compound mass properties at density 10 from K4's per-piece records (after the refit) vs the reference's Moments
    for b in range(ncomp):
        sel = want_c.cell == b
        m_ref = 10.0 * want_c.volume[sel].sum()
        assert abs(mass[b, 0] - m_ref) <= 1e-5 * max(1.0, m_ref)
        if m_ref > 0:
            c_ref = (want_c.volume[sel, None] * want_c.centroid[sel]).sum(0) / want_c.volume[sel].sum()
            assert np.abs(mass[b, 1:4] - c_ref).max() < 1e-4
            assert mass[b, 4:7].min() > 0
    # ... and the compound INERTIA (PhysX's updateMassAndInertia, Surtr.cpp:2520, is not in the reference tree, so the
    # yardstick is the oracle's double-precision polyhedral integral of every piece, combined over the compound in double
    # with the parallel-axis theorem): all six terms of every compound within 1e-4 of the tensor's largest term
    far = np.array([[1.0, 0.0, 0.0, -1.0e6]], np.float32)                      # a plane nothing reaches: pieces pass uncut
    exact = P.apply_fracture(want_c, far, np.array([0, 1], np.uint32), inertia=True, cap_frags=want_c.n + 8, cap_verts=len(want_c.verts) + 64)
    assert exact.n == want_c.n and np.array_equal(bits(exact.volume), bits(want_c.volume))
    worst = 0.0
    for b in range(ncomp):
        sel = np.nonzero(want_c.cell == b)[0]
        V, c, I = exact.volume[sel], exact.centroid[sel].astype(np.float64), exact.inertia[sel]
        if V.sum() <= 0:
            continue
        com = (V[:, None] * c).sum(0) / V.sum()
        d = c - com
        want_I = np.array([
            (I[:, 0] + V * (d[:, 1] ** 2 + d[:, 2] ** 2)).sum(), (I[:, 1] + V * (d[:, 0] ** 2 + d[:, 2] ** 2)).sum(),
            (I[:, 2] + V * (d[:, 0] ** 2 + d[:, 1] ** 2)).sum(), (I[:, 3] - V * d[:, 0] * d[:, 1]).sum(),
            (I[:, 4] - V * d[:, 0] * d[:, 2]).sum(), (I[:, 5] - V * d[:, 1] * d[:, 2]).sum()]) * 10.0
        scale = np.abs(want_I[:3]).max()
        worst = max(worst, float(np.abs(mass[b, 4:10].astype(np.float64) - want_I).max() / scale))
    assert worst < 1e-4, worst


def test_host_worker_pool():
    """detail::parallel_for: every index exactly once, repeated jobs of different sizes, exceptions propagate."""
    import ctypes as C
    L = H.lib()
    seen = C.c_uint32(0)
    for n, throw_at in ((3, 99), (64, 99), (1000, 99), (17, 5), (1000, 999), (8, 99), (5000, 99)):
        assert L.hosttest_parallel_for(n, throw_at, C.byref(seen)) == 0, H._err()
    assert seen.value >= 1


def test_host_transform_matches_reference():
    """Poly::Transform mirror (world matrix, transposed inside, TransformCoord with the w division) == reference build."""
    d = np.load(os.path.join(GOLDEN, "transform_kat.npz"))
    pieces = load_polyset(d, "pieces_")
    for i in range(3):
        sl = slice(int(pieces.vert_off[i]), int(pieces.vert_off[i + 1]))
        assert np.array_equal(bits(H.transform(pieces.verts[sl], d["matrices"][i])), bits(d["out"][sl]))


@pytest.mark.gpu
def test_cpp_example_runs_the_reference_flow(tmp_path):
    """examples/fracture_demo.cpp: PrepareFracture + DoFracture written against the host classes like a C++ caller would;
    on the bunny it must report the reference's 27 initial pieces (config1_full fixture)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "examples")], check=True, capture_output=True)
    d = np.load(os.path.join(GOLDEN, "config1_full_bunny32.npz"))
    obj = tmp_path / "bunny.obj"
    with open(obj, "w") as f:
        for v in d["verts"]:
            f.write(f"v {-float(v[0])!r} {float(v[1])!r} {float(v[2])!r}\n")      # the loader negates x again
        for t in d["indices"].reshape(-1, 3):
            f.write(f"f {t[0] + 1} {t[2] + 1} {t[1] + 1}\n")                        # and flips the winding back
    r = subprocess.run([os.path.join(root, "examples", "fracture_demo"), str(obj), "1.0", "32"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "-> 27 pieces" in r.stdout and "DoFracture(partial)" in r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(not common.have_ref(), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("case", list(range(8)))
def test_do_fracture_random_impacts_match_reference(case):
    """DoFracture with random impact points, radii, pattern seeds and modes on the bunny compound: the host classes (GPU
    events + host orchestration) against the restatement over the reference build, live (oracle/_ref travels to the
    GPU box): pieces, order and compounds bit for bit."""
    from oracle import refapi as R
    d = np.load(os.path.join(GOLDEN, "do_fracture_bunny.npz"))
    d0 = np.load(os.path.join(GOLDEN, "config1_full_bunny32.npz"))
    convex, mesh = load_polyset(d0, "convex_"), load_polyset(d0, "mesh_")
    rng = np.random.RandomState(300 + case)
    v = d0["verts"][:, :3]
    impact = v[rng.randint(len(v))].copy()
    partial = bool(case % 2)
    radius = float(rng.uniform(1.0, 5.0))
    n_seeds = int(rng.choice([16, 32, 48]))
    seeds = R.seeds_radial(int(rng.randint(1, 10 ** 6)), n_seeds, float(rng.choice([0.02, 0.05, 0.3, 1.0])))
    off, idx = H.dt3d_neighbors(seeds)
    max_axis = float(d["max_axis_scale"])
    want_c, want_m, want_n = R.do_fracture(convex, mesh, seeds, off, idx, d["cloud"], impact, radius, max_axis, partial)
    got_c, got_m, got_n, _ = H.do_fracture(convex, mesh, seeds, d["cloud"], impact, radius, max_axis, partial)
    assert got_n == want_n and got_c.n == want_c.n
    for f in ("verts", "vert_off", "ring_off", "ring", "cell", "piece"):
        assert np.array_equal(bits(getattr(got_c, f)), bits(getattr(want_c, f))), "convex " + f
        assert np.array_equal(bits(getattr(got_m, f)), bits(getattr(want_m, f))), "mesh " + f


@pytest.mark.gpu
@pytest.mark.skipif(not common.have_ref(), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("case", [0, 1, 2])
def test_prepare_fracture_random_seeds_match_reference(case):
    """Full PrepareFracture on the bunny with other seed sets (16 / 48 / 64 cells) against the live reference restatement."""
    from oracle import refapi as R
    d = np.load(os.path.join(GOLDEN, "config1_full_bunny32.npz"))
    seeds = R.seeds_uniform(7000 + case, (16, 48, 64)[case])
    off, idx = H.dt3d_neighbors(seeds)
    _, want_c, want_m = R.config1_full(d["verts"], d["indices"], seeds, off, idx)
    got_c, got_m, _ = H.config1_full(d["verts"], d["indices"], seeds)
    assert got_c.n == want_c.n
    for f in ("verts", "vert_off", "ring_off", "ring", "cell", "piece"):
        assert np.array_equal(bits(getattr(got_c, f)), bits(getattr(want_c, f))), "convex " + f
        assert np.array_equal(bits(getattr(got_m, f)), bits(getattr(want_m, f))), "mesh " + f
