"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle on identical inputs.

Bar (BASELINE.json north_star): piece-to-cell assignments and per-fragment vertex/face counts bit-exact; vertex
positions within 1e-5 relative; volumes closing on the parent.  The kernels reproduce the reference's vertex
numbering, ring order and accumulation order, so everything below is asserted BITWISE, which implies the bar."""
import json
import os

import numpy as np
import pytest

import common
from common import GOLDEN, bits
from oracle import portapi as P
from test_oracle_port import load_polyset

pytestmark = pytest.mark.gpu


def _summaries():
    return json.load(open(os.path.join(GOLDEN, "summaries.json")))


@pytest.mark.parametrize("name", ["cube_x64", "pieces200_x32"])
def test_golden_fixture_events(ctx, name):
    """Committed outputs of the REFERENCE build (tests/golden/make_golden.py)."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    cells, pieces, want = load_polyset(d, "cells_"), load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    for k in (3, 7, 13):      # every broad-phase direction set must keep exactly the reference's fragments
        ctx.set_kdop_directions(k)
        got = common.run_gpu(ctx, pieces, cells)
        common.assert_fragments_equal(got, want)
    ctx.set_kdop_directions(13)    # the default
    got = common.run_gpu(ctx, pieces, cells, bounded=False)     # no cell bounds: every pair reaches the clipper
    common.assert_fragments_equal(got, want)
    assert ctx.counts().n_candidates == pieces.n * cells.n


def test_config2_cube_x4096(ctx):
    cells = common.voronoi(46354, 4096)
    cube = common.unit_cube()
    got = common.run_gpu(ctx, cube, cells)
    want = P.apply_fracture(cube, cells.planes, cells.plane_off)
    common.assert_fragments_equal(got, want)
    assert common.summary_of_fragments(got) == _summaries()["config2_cube_x4096"]
    # closure: fragment volumes sum to the parent as closely as the reference's own do
    assert abs(got.rec["volume"].sum() - 1.0) <= abs(want.volume.sum() - 1.0) + 1e-12
    assert abs(got.rec["volume"].sum() - 1.0) < 5e-6


def test_config3_10000x256(ctx):
    cells = common.voronoi(46354, 256)
    pieces = common.voronoi(1234, 10000)
    got = common.run_gpu(ctx, pieces, cells)
    want = P.apply_fracture(pieces, cells.planes, cells.plane_off)
    common.assert_fragments_equal(got, want)
    assert got.n == 23864 and common.summary_of_fragments(got) == _summaries()["config3_10000x256"]
    c = ctx.counts()
    assert c.n_pairs == 2_560_000 and c.n_candidates < 3 * c.n_fragments
    assert abs(got.rec["volume"].sum() - 1.0) < 1e-6


def test_config4_batched_events(ctx):
    """Independent events in one batch: event e = 1000-seed pieces (mt19937(1234+e)) x 64-seed cells
    (mt19937(46354+e)); result = the per-event oracle results back to back (event-major order)."""
    n_ev = 4
    psets = [common.voronoi(1234 + e, 1000) for e in range(n_ev)]
    csets = [common.voronoi(46354 + e, 64) for e in range(n_ev)]
    pieces, ev_p = common.concat(psets)
    cells, ev_c = common.concat(csets)
    got = common.run_gpu(ctx, pieces, cells, ev_p, ev_c)
    f0 = 0
    for e in range(n_ev):
        want = P.apply_fracture(psets[e], csets[e].planes, csets[e].plane_off)
        sl = slice(f0, f0 + want.n)
        assert np.array_equal(got.rec["cell"][sl], want.cell + ev_c[e])
        assert np.array_equal(got.rec["piece"][sl], want.piece + ev_p[e])
        assert np.array_equal(got.rec["n_verts"][sl], want.nverts)
        assert np.array_equal(got.rec["n_faces"][sl], want.nfaces)
        v0 = int(got.rec["vert_off"][f0])
        assert np.array_equal(bits(got.verts[v0:v0 + len(want.verts)]), bits(want.verts))
        assert np.array_equal(bits(got.rec["volume"][sl]), bits(want.volume))
        if e == 0:
            assert want.n == 2841
        f0 += want.n
    assert f0 == got.n


def test_config5_recursive_refracture(ctx):
    """Depth-3 re-fracture with fragments staying on the device as the next level's pieces."""
    pieces = common.unit_cube()
    levels = common.recursion_levels()
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    want_pieces = pieces
    for lvl, cells in enumerate(levels):
        ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
        ctx.fracture_event()
        got = ctx.download()
        want = P.apply_fracture(want_pieces, cells.planes, cells.plane_off)
        common.assert_fragments_equal(got, want)
        assert common.summary_of_fragments(got) == _summaries()[f"config5_level{lvl}"]
        want_pieces = want
        ctx.fragments_to_pieces()
    assert got.n == 1620


def test_fragments_to_pieces_per_event_matches_host_regrouping(ctx):
    """surtr_fragments_to_pieces_per_event (event boundaries found on the device) against surtr_fragments_to_pieces with
    the boundaries the host derives from the downloaded records: three events of different size -- the middle one's
    piece lies outside every cell, so it has no fragments at all -- then one more level on both piece sets; every array
    of the second level must be identical, and equal the oracle's."""
    from surtr_b200 import FractureContext
    cube = common.unit_cube()
    far = common.PolySet(cube.verts + np.array([10, 0, 0, 0], np.float32), cube.vert_off, cube.ring_off, cube.ring)
    pieces, ev_p = common.concat([cube, far, cube])
    c0, c1, c2 = common.voronoi(46354, 64), common.voronoi(777, 8), common.voronoi(1001, 24)
    cells, ev_c = common.concat([c0, c1, c2])
    nxt, ev_n = common.concat([common.voronoi(1002, 16), common.voronoi(1003, 5), common.voronoi(1004, 40)])
    other = FractureContext(0)
    try:
        results = []
        for cx, per_event in ((ctx, True), (other, False)):
            cx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring, ev_p)
            cx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off, ev_c)
            cx.fracture_event()
            rec = cx.download(geometry=False).rec
            ev_of = np.searchsorted(np.asarray(ev_c, np.int64), rec["cell"].astype(np.int64), side="right") - 1
            per_ev = np.bincount(ev_of, minlength=3)
            assert per_ev[0] == 64 and per_ev[1] == 0 and per_ev[2] == 24
            if per_event:
                cx.fragments_to_pieces_per_event()
            else:
                cx.fragments_to_pieces(np.concatenate([[0], np.cumsum(per_ev)]).astype(np.uint32))
            cx.upload_cells(nxt.planes, nxt.plane_off, nxt.verts, nxt.vert_off, ev_n)
            cx.fracture_event()
            results.append(cx.download())
        a, b = results
        assert a.n == b.n and a.n > 0
        assert a.rec.tobytes() == b.rec.tobytes()
        assert a.verts.tobytes() == b.verts.tobytes() and a.ring_off.tobytes() == b.ring_off.tobytes() and a.ring.tobytes() == b.ring.tobytes()
        # and against the oracle: level 1 of events 0 and 2, then level 2
        f0 = 0
        for cube_cells, level2 in ((c0, common.voronoi(1002, 16)), (c2, common.voronoi(1004, 40))):
            lvl1 = P.apply_fracture(cube, cube_cells.planes, cube_cells.plane_off)
            want = P.apply_fracture(lvl1, level2.planes, level2.plane_off)
            sl = slice(f0, f0 + want.n)
            assert np.array_equal(a.rec["n_verts"][sl], want.nverts) and np.array_equal(a.rec["n_faces"][sl], want.nfaces)
            assert a.rec["volume"][sl].tobytes() == np.asarray(want.volume, np.float64).tobytes()
            f0 += want.n
        assert f0 == a.n
    finally:
        other.close()


@pytest.mark.parametrize("mode", [1, 2], ids=["throughput", "latency"])
def test_both_builds_of_the_small_tier_match_the_oracle(ctx, mode):
    """surtr_set_clip_build: the throughput build of K3's small tier (40 warps per SM, positions from shared memory, no plane
    prefetch) and the latency build (32 warps, register copies, prefetch) cut the same fragments, bit for bit the oracle's."""
    pieces, cells = common.voronoi(1234, 300), common.voronoi(46354, 40)
    want = P.apply_fracture(pieces, cells.planes, cells.plane_off)
    ctx.set_clip_build(mode)
    try:
        for _ in range(2):      # (the second event has a previous event to size itself by; the pinned build ignores it)
            got = common.run_gpu(ctx, pieces, cells)
            common.assert_fragments_equal(got, want)
    finally:
        ctx.set_clip_build(0)


def test_edge_cases(ctx):
    cube = common.unit_cube()
    cells = common.voronoi(46354, 64)
    # no cells / no pieces -> no fragments
    empty_cells_off = np.zeros(1, np.uint32)
    ctx.upload_pieces(cube.verts, cube.vert_off, cube.ring_off, cube.ring)
    ctx.upload_cells(np.zeros((0, 4), np.float32), empty_cells_off)
    ctx.fracture_event()
    assert ctx.download().n == 0
    ctx.upload_pieces(np.zeros((0, 4), np.float32), np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros(0, np.uint16))
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
    ctx.fracture_event()
    assert ctx.download().n == 0
    # a cell with an empty plane list keeps every piece whole; a far-away half-space removes it
    planes = np.array([[1, 0, 0, 2.0]], np.float32)        # x + 2 <= 0 : everything clipped
    ctx.upload_pieces(cube.verts, cube.vert_off, cube.ring_off, cube.ring)
    ctx.upload_cells(planes, np.array([0, 0, 1], np.uint32))
    ctx.fracture_event()
    fr = ctx.download()
    assert fr.n == 1 and int(fr.rec["cell"][0]) == 0 and int(fr.rec["n_verts"][0]) == 8 and fr.rec["volume"][0] == 1.0
    # ragged batch: events with zero pieces or zero cells in the middle
    p2, c2 = common.voronoi(7, 40), common.voronoi(8, 12)
    pieces, _ = common.concat([p2, p2])
    cells2, _ = common.concat([c2, c2])
    ev_p = np.array([0, 40, 40, 80, 80], np.uint32)
    ev_c = np.array([0, 12, 24, 24, 24], np.uint32)
    got = common.run_gpu(ctx, pieces, cells2, ev_p, ev_c)
    want = P.apply_fracture(p2, c2.planes, c2.plane_off)
    assert got.n == want.n and np.array_equal(got.rec["n_verts"], want.nverts)
    assert np.array_equal(bits(got.verts), bits(want.verts))


def test_large_tier_pieces_and_kdop_clip(ctx):
    """Config-1 front end: ACH = 2x bbox clipped by the 2k k-DOP planes (Kdop.cpp:166-179).  The bunny / sphere
    ACHs (107 / 135 vertices) then exceed the 64-vertex tier and are cut in the 256-vertex tier."""
    d = np.load(os.path.join(GOLDEN, "config1_kdop.npz"))
    for key in ("bunny", "cube", "sphere"):
        want = load_polyset(d, key + "_ach_")
        box = common.unit_cube()
        box.verts = d[key + "_seedbox_verts"]
        planes = d[key + "_gap_planes"].reshape(-1, 4)
        ctx.upload_pieces(box.verts, box.vert_off, box.ring_off, box.ring)
        ctx.upload_cells(planes, np.array([0, len(planes)], np.uint32))
        ctx.fracture_event()
        got = ctx.download()
        assert got.n == 1
        assert np.array_equal(bits(got.verts), bits(want.verts)) and np.array_equal(got.ring, want.ring)
        assert int(got.rec["n_faces"][0]) == int(want.nfaces[0])
        assert np.array_equal(bits(got.rec["volume"]), bits(want.volume))
        # now cut the ACH itself by a cell set scaled onto it (pieces beyond the small tier)
        if key != "cube":
            ach = want
            lo, hi = ach.verts[:, :3].min(0), ach.verts[:, :3].max(0)
            cells = common.voronoi(46354, 32)
            cp = cells.subset(range(cells.n))
            cp.verts = cells.verts.copy()
            cp.verts[:, :3] = cells.verts[:, :3] * (hi - lo) + (hi + lo) / 2
            # planes of the scaled cells, rebuilt from their vertices like Polygon3D::Scale/Translate does
            # (VMACH.cpp:506-534): PolygonFace::AddVertex route of the oracle
            scaled = cp
            planes2, off2 = P.face_planes(scaled)
            want_fr = P.apply_fracture(ach, planes2, off2)
            ctx.upload_pieces(ach.verts, ach.vert_off, ach.ring_off, ach.ring)
            ctx.upload_cells(planes2, off2, scaled.verts, scaled.vert_off)
            ctx.fracture_event()
            got = ctx.download()
            common.assert_fragments_equal(got, want_fr)
            assert ctx.counts().n_tier1b + ctx.counts().n_tier2 > 0


def test_global_tier_mesh_polyhedron(ctx):
    """Row f-1 clip: the bunny as a 2503-vertex, non-convex vertex-ring polyhedron (ExtractNeighborFromMesh,
    Poly.cpp:128-263) cut by 32 cells (Surtr.cpp:1470) -- beyond both shared-memory tiers, so every pair runs in
    the global-memory tier.  Expected fragments are the REFERENCE build's (tests/golden/make_golden.py)."""
    d = np.load(os.path.join(GOLDEN, "bunny_mesh_x32.npz"))
    mesh, want = load_polyset(d, "mesh_"), load_polyset(d, "frag_")
    ctx.upload_pieces(mesh.verts, mesh.vert_off, mesh.ring_off, mesh.ring)
    ctx.upload_cells(d["planes"], d["plane_off"], d["cell_verts"], d["cell_vert_off"])
    for rep in range(2):                      # second pass: tier already enabled, no rerun
        ctx.fracture_event()
        got = ctx.download()
        common.assert_fragments_equal(got, want)
        c = ctx.counts()
        assert c.n_tier3 == c.n_candidates > 0
    # a mixed event: the mesh next to small convex pieces, every tier in one launch sequence
    small = common.voronoi(7, 40)
    lo, hi = mesh.verts[:, :3].min(0), mesh.verts[:, :3].max(0)
    placed = small.subset(range(small.n))
    placed.verts = small.verts.copy()
    placed.verts[:, :3] = small.verts[:, :3] * (hi - lo) + (hi + lo) / 2
    both, _ = common.concat([placed.subset(range(20)), mesh, placed.subset(range(20, 40))])
    want = P.apply_fracture(both, d["planes"], d["plane_off"], cap_frags=4096, cap_verts=400000)
    ctx.upload_pieces(both.verts, both.vert_off, both.ring_off, both.ring)
    ctx.fracture_event()
    common.assert_fragments_equal(ctx.download(), want)
    c = ctx.counts()
    assert 0 < c.n_tier3 < c.n_candidates


def test_kdop_calc(ctx):
    d = np.load(os.path.join(GOLDEN, "config1_kdop.npz"))
    for key in ("bunny", "cube", "sphere"):
        v4, normals = d[key + "_verts"], d[key + "_normals"]
        dist, arg, planes = ctx.kdop_calc(v4, normals)
        pd, pa, pp = P.kdop_calc(v4, normals)
        assert np.array_equal(bits(dist), bits(pd)) and np.array_equal(arg, pa) and np.array_equal(bits(planes), bits(pp))
        assert np.array_equal(dist.astype(np.float64), d[key + "_poly_dist"])
        assert np.array_equal(bits(planes), bits(d[key + "_poly_planes"]))


def test_inertia_against_double_precision_oracle(ctx):
    """No in-repo reference for inertia (PhysX, Surtr.cpp:2520): compare with the oracle's independent double
    precision polyhedral integral, tolerance 1e-4 relative to the tensor's scale (float32 vertex data)."""
    cells = common.voronoi(46354, 64)
    pieces = common.voronoi(1234, 300)
    got = common.run_gpu(ctx, pieces, cells)
    want = P.apply_fracture(pieces, cells.planes, cells.plane_off, inertia=True)
    assert got.n == want.n
    scale = np.abs(want.inertia[:, :3]).max(axis=1, keepdims=True)
    big = want.volume > 1e-9                      # slivers have no meaningful tensor
    err = np.abs(got.rec["inertia"].astype(np.float64) - want.inertia) / np.maximum(scale, 1e-300)
    assert err[big].max() < 1e-4
    cube = common.unit_cube()
    ctx.upload_pieces(cube.verts, cube.vert_off, cube.ring_off, cube.ring)
    ctx.upload_cells(np.zeros((0, 4), np.float32), np.array([0, 0], np.uint32))
    ctx.fracture_event()
    fr = ctx.download()
    assert np.allclose(fr.rec["inertia"][0], [1 / 6, 1 / 6, 1 / 6, 0, 0, 0], atol=1e-6)


def test_determinism_and_rerun_growth(ctx):
    """Same event twice -> identical bytes; a fresh context with tiny buffers grows and still matches."""
    from surtr_b200 import FractureContext
    cells = common.voronoi(46354, 256)
    pieces = common.voronoi(1234, 3000)
    a = common.run_gpu(ctx, pieces, cells)
    b = common.run_gpu(ctx, pieces, cells)
    assert a.rec.tobytes() == b.rec.tobytes() and a.verts.tobytes() == b.verts.tobytes()
    assert a.ring.tobytes() == b.ring.tobytes()
    c2 = FractureContext(0)
    small = common.unit_cube()
    c2.upload_pieces(small.verts, small.vert_off, small.ring_off, small.ring)
    c2.upload_cells(common.voronoi(5, 8).planes, common.voronoi(5, 8).plane_off)
    c2.fracture_event()
    c2.download()
    c = common.run_gpu(c2, pieces, cells)
    assert a.rec.tobytes() == c.rec.tobytes() and a.verts.tobytes() == c.verts.tobytes()
    c2.close()


def test_product_voronoi_builder_matches_oracle(ctx):
    """surtr_b200.synth.voronoi_cells (GPU clipper + host face planes) == the oracle's cell builder, bitwise."""
    from surtr_b200 import synth
    s = common.seeds_uniform(46354, 512)
    assert np.array_equal(s, synth.seeds_uniform(46354, 512))
    off, idx = common.scipy_neighbors(s)
    got = synth.voronoi_cells(ctx, s, off, idx)
    want = P.voronoi_cells(s, off, idx)
    assert np.array_equal(bits(got.verts), bits(want.verts)) and np.array_equal(got.ring, want.ring)
    assert np.array_equal(got.plane_off, want.plane_off) and np.array_equal(bits(got.planes), bits(want.planes))


def test_malformed_and_oversize_inputs_fail_loudly(ctx):
    """Invalid rings are reported (SURTR_ERR_OVERFLOW), never read out of bounds."""
    from surtr_b200 import SurtrError
    cube = common.unit_cube()
    cells = common.voronoi(46354, 8)
    bad = cube.subset([0])
    bad.ring = cube.ring.copy()
    bad.ring[5] = 200                      # neighbour index beyond the piece's vertex count
    ctx.upload_pieces(bad.verts, bad.vert_off, bad.ring_off, bad.ring)
    ctx.upload_cells(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
    ctx.fracture_event()
    with pytest.raises(SurtrError) as e:
        ctx.counts()
    assert e.value.code == 4
    # a 300-vertex "piece" whose rings are not a polyhedron at all goes to the global-memory tier: it may be
    # rejected or produce garbage, but it must neither crash nor hang, and the context stays usable afterwards
    n = 300
    verts = np.zeros((n, 4), np.float32)
    verts[:, :3] = np.random.RandomState(0).uniform(-0.4, 0.4, (n, 3))
    ring = np.stack([(np.arange(n) + 1) % n, (np.arange(n) + 2) % n, (np.arange(n) + n - 1) % n], 1).astype(np.uint16).reshape(-1)
    ctx.upload_pieces(verts, np.array([0, n], np.uint32), np.arange(0, 3 * n + 1, 3, dtype=np.uint32), ring)
    ctx.fracture_event()
    try:
        ctx.counts()
    except SurtrError as err:
        assert err.code == 4
    got = common.run_gpu(ctx, cube, cells)
    want = P.apply_fracture(cube, cells.planes, cells.plane_off)
    common.assert_fragments_equal(got, want)


def _cone_mesh(n, height=1.3, radius=0.9):
    """A triangulated cone: apex and base centre both have n neighbours (fan poles, the valence ADVICE.md warns about)."""
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    verts = np.zeros((n + 2, 4), np.float32)
    verts[0, :3] = (0.05, height, -0.02)                        # apex
    verts[1, :3] = (0.0, 0.0, 0.0)                              # base centre
    verts[2:, 0], verts[2:, 2] = radius * np.cos(ang), radius * np.sin(ang)
    verts[2:, 1] = 0.01 * np.sin(3 * ang)                       # (not exactly coplanar: no degenerate planes)
    tri = []
    for i in range(n):
        a, b = 2 + i, 2 + (i + 1) % n
        tri += [0, b, a, 1, a, b]
    return verts, np.asarray(tri, np.int32)


@pytest.mark.parametrize("n", [40, 150])
def test_high_valence_pieces_have_no_ring_limit(ctx, n):
    """Poly::ClipPolyhedron has no limit on the number of neighbours of a vertex; the global-memory tier widens its ring
    slots on demand (16 -> 32 -> ...).  A cone whose apex and base centre have n neighbours, as a mesh polyhedron
    (ExtractNeighborFromMesh rings), cut by a 24-cell pattern: n = 40 against the oracle port (rings of up to 64),
    n = 150 against the reference build itself."""
    import hostapi as H
    from oracle import refapi as R
    if n > 60 and not common.have_ref():
        pytest.skip("needs the reference build (oracle/_ref)")
    verts, tri = _cone_mesh(n)
    mesh = (R.mesh_polyhedron if common.have_ref() else H.mesh_polyhedron)(verts, tri)
    assert int(np.diff(mesh.ring_off).max()) == n
    cells = common.voronoi(46354, 24)
    cp = cells.subset(range(cells.n))
    cp.verts = cells.verts.copy()
    cp.verts[:, :3] = cells.verts[:, :3] * np.float32(2.2) + np.array([0.0, 0.6, 0.0], np.float32)
    planes, off = P.face_planes(cp)
    want = (R.apply_fracture(mesh, planes, off, 8) if n > 60 else P.apply_fracture(mesh, planes, off, cap_frags=256, cap_verts=20000))
    ctx.upload_pieces(mesh.verts, mesh.vert_off, mesh.ring_off, mesh.ring)
    ctx.upload_cells(planes, off, cp.verts, cp.vert_off)
    ctx.fracture_event()
    got = ctx.download()
    common.assert_fragments_equal(got, want, moments=(n <= 60))
    assert want.n >= 8 and ctx.counts().n_tier3 > 0 and ctx.counts().n_failed == 0


def test_failed_pairs_are_reported_per_pair(ctx):
    """A malformed piece next to healthy ones: its pairs are listed by surtr_failed_pairs, surtr_event_counts says
    SURTR_ERR_OVERFLOW, and every other fragment of the event is there and correct."""
    from surtr_b200 import SurtrError
    good = common.voronoi(1234, 30)
    cells = common.voronoi(46354, 8)
    bad = common.unit_cube()
    bad.ring = bad.ring.copy()
    bad.ring[5] = 200
    pieces, _ = common.concat([good, bad, good])
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(cells.planes, cells.plane_off)          # unbounded: every pair reaches the clipper
    ctx.fracture_event()
    with pytest.raises(SurtrError) as e:
        ctx.counts()
    assert e.value.code == 4
    c = ctx.counts(allow_failed=True)
    failed = ctx.failed_pairs()
    assert c.n_failed == 8 and sorted(map(tuple, failed.tolist())) == [(30, k) for k in range(8)]
    got = ctx.download(allow_failed=True)
    want = P.apply_fracture(good, cells.planes, cells.plane_off)
    sel = got.rec["piece"] < 30
    assert sel.sum() == want.n and np.array_equal(got.rec["n_verts"][sel], want.nverts)
    assert np.array_equal(bits(got.rec["volume"][sel]), bits(want.volume))
    assert not np.any(got.rec["piece"] == 30) and (got.rec["piece"] > 30).sum() == want.n


def test_resident_pattern_placement(ctx):
    """Row f-4: a pattern uploaded once in its own frame and placed on the device (surtr_place_pattern = Polygon3D::Scale +
    Translate with the face planes re-derived from the moved vertices, VMACH.cpp:303-310, 506-534), three placements in
    one batch of independent events.  Expected: the oracle on cells moved in numpy float32 (one rounding per operation)
    with planes from the PolygonFace::AddVertex route."""
    from surtr_b200 import synth
    cells = common.voronoi(46354, 64)
    ctx.upload_pattern(*synth.pattern_arrays(cells.verts, cells.vert_off, cells.ring_off, cells.ring))
    scale = np.array([[2, 2, 2], [1.5, 0.7, 3.0], [21.72226, 21.72226, 21.72226]], np.float32)
    trans = np.array([[0, 0, 0], [0.3, -1.25, 4.0], [-1.1729, 12.41, 3.53]], np.float32)
    cube = common.unit_cube()
    psets, wants = [], []
    for s, t in zip(scale, trans):
        piece = cube.subset([0])
        piece.verts = cube.verts.copy()
        piece.verts[:, :3] = (cube.verts[:, :3] * (s * np.float32(0.6))).astype(np.float32) + (t + s * np.float32(0.1)).astype(np.float32)
        placed = cells.subset(range(cells.n))
        placed.verts = cells.verts.copy()
        placed.verts[:, :3] = (cells.verts[:, :3] * s).astype(np.float32) + t
        planes, off = P.face_planes(placed)
        psets.append(piece)
        wants.append(P.apply_fracture(piece, planes, off))
    pieces, ev_p = common.concat(psets)
    for rep in range(2):                          # the second pass re-places the resident pattern, nothing is re-uploaded
        ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring, ev_p)
        ctx.place_pattern(scale, trans)
        ctx.fracture_event()
        got = ctx.download()
        f0 = 0
        for e, want in enumerate(wants):
            assert 10 < want.n < 64               # the piece covers part of the pattern: both culling and cutting happen
            sl = slice(f0, f0 + want.n)
            assert np.array_equal(got.rec["cell"][sl], want.cell + 64 * e) and np.array_equal(got.rec["piece"][sl], want.piece + e)
            assert np.array_equal(got.rec["n_verts"][sl], want.nverts) and np.array_equal(got.rec["n_faces"][sl], want.nfaces)
            v0 = int(got.rec["vert_off"][f0])
            assert np.array_equal(bits(got.verts[v0:v0 + len(want.verts)]), bits(want.verts))
            assert np.array_equal(bits(got.rec["volume"][sl]), bits(want.volume))
            f0 += want.n
        assert f0 == got.n


def test_degenerate_cuts(ctx):
    """Planes exactly through vertices, along edges and coincident with faces of the unit cube and of a truncated cube
    (400 random sequences): the in-plane band and the sequential patch path.  Expected = the reference build's output."""
    d = np.load(os.path.join(GOLDEN, "degenerate_x400.npz"))
    pieces, want = load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(d["planes"], d["plane_off"])
    ctx.fracture_event()
    got = ctx.download()
    common.assert_fragments_equal(got, want)
    assert ctx.counts().n_seq_cuts > 100          # the in-plane cases really took the sequential replay


def test_degenerate_cuts_large_tiers(ctx):
    """The same in-plane situations on the 107-vertex ACH (shared-memory large tier) and the 2503-vertex mesh
    (global-memory tier), whose sequential patch paths are separate code: fingerprint of the reference build."""
    pieces, planes, off = common.degenerate_large_inputs()
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(planes, off)
    ctx.fracture_event()
    got = ctx.download()
    s = common.summary_of_fragments(got)
    s["ring"] = common.digest(np.asarray(got.ring, np.uint16))
    assert s == _summaries()["degenerate_large"]
    c = ctx.counts()
    # ACH pairs: the 128-slot warp tier (results beyond 64 vertices go on to the large tier); mesh pairs: straight to the global tier
    assert c.n_tier1b == 48 and c.n_tier3 == 48 and c.n_seq_cuts > 50
    # and with the 128-slot tier switched off (test hook) the large tier's own sequential patch code cuts all 48 ACH pairs
    from surtr_b200 import FractureContext
    os.environ["SURTR_DEBUG_NO_TIER1B"] = "1"
    try:
        cx = FractureContext(0)
    finally:
        del os.environ["SURTR_DEBUG_NO_TIER1B"]
    try:
        cx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
        cx.upload_cells(planes, off)
        cx.fracture_event()
        got2 = cx.download()
        s2 = common.summary_of_fragments(got2)
        s2["ring"] = common.digest(np.asarray(got2.ring, np.uint16))
        assert s2 == _summaries()["degenerate_large"]
        assert cx.counts().n_tier2 == 48 and cx.counts().n_tier1b == 0
    finally:
        cx.close()


def test_transform_pieces_on_device(ctx):
    """Row f-4: world transform of resident pieces (surtr_transform_pieces = Poly::Transform, Poly.cpp:580-585), one matrix
    per piece; then an event on the moved pieces equals the oracle on the reference-transformed pieces."""
    d = np.load(os.path.join(GOLDEN, "transform_kat.npz"))
    pieces = load_polyset(d, "pieces_")
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.transform_pieces(d["matrices"], np.array([0, 1, 2], np.uint32))
    got = ctx.download_pieces(len(pieces.verts))
    assert np.array_equal(bits(got), bits(d["out"]))
    moved = pieces.subset(range(3))
    moved.verts = d["out"]
    lo, hi = d["out"][:, :3].min(0), d["out"][:, :3].max(0)
    cells = common.voronoi(46354, 64)
    placed = cells.subset(range(cells.n))
    placed.verts = cells.verts.copy()
    placed.verts[:, :3] = (cells.verts[:, :3] * (hi - lo)).astype(np.float32) + ((hi + lo) / 2).astype(np.float32)
    planes, off = P.face_planes(placed)
    ctx.upload_cells(planes, off, placed.verts, placed.vert_off)
    ctx.fracture_event()
    common.assert_fragments_equal(ctx.download(), P.apply_fracture(moved, planes, off))


def test_async_download_and_contexts_in_flight():
    """surtr_download_fragments_async + surtr_sync: three contexts driven round-robin from one host thread, each
    launching its next event right behind the enqueued copies of the previous one; every result equals the oracle."""
    import torch
    from surtr_b200 import FractureContext, FRAGMENT_DTYPE
    cube = common.unit_cube()
    sets = [common.voronoi(46354 + k, 64) for k in range(3)]
    wants = [P.apply_fracture(cube, c.planes, c.plane_off) for c in sets]
    pipes = []
    for k in range(3):
        st = torch.cuda.Stream()
        cx = FractureContext(0, st.cuda_stream)
        want = wants[k]
        bufs = dict(rec=torch.zeros(want.n * FRAGMENT_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
                    verts=torch.zeros(len(want.verts) * 4, dtype=torch.float32).pin_memory(),
                    ring_off=torch.zeros(len(want.verts) + 1, dtype=torch.int32).pin_memory(),
                    ring=torch.zeros(len(want.ring), dtype=torch.int16).pin_memory())
        pipes.append((cx, st, bufs))
    try:
        for step in range(9):
            k = step % 3
            cx, st, b = pipes[k]
            if step >= 3:
                cx.download_into_async(b["rec"].data_ptr(), b["verts"].data_ptr(), b["ring_off"].data_ptr(), b["ring"].data_ptr())
            cx.upload_pieces(cube.verts, cube.vert_off, cube.ring_off, cube.ring)
            cx.upload_cells(sets[k].planes, sets[k].plane_off, sets[k].verts, sets[k].vert_off)
            cx.fracture_event()
        for k, (cx, st, b) in enumerate(pipes):
            cx.download_into_async(b["rec"].data_ptr(), b["verts"].data_ptr(), b["ring_off"].data_ptr(), b["ring"].data_ptr())
            cx.sync()
            rec = np.frombuffer(b["rec"].numpy().tobytes(), dtype=FRAGMENT_DTYPE)
            assert len(rec) == wants[k].n and np.array_equal(rec["n_verts"], wants[k].nverts)
            assert np.array_equal(bits(b["verts"].numpy().reshape(-1, 4)), bits(wants[k].verts))
            assert np.array_equal(b["ring"].numpy().view(np.uint16), wants[k].ring)
            assert np.array_equal(bits(rec["volume"]), bits(wants[k].volume))
    finally:
        for cx, st, b in pipes:
            cx.close()


def test_packed_wire_format(ctx):
    """surtr_upload_pieces3 / surtr_upload_cells3 / surtr_download_fragments_packed: float3 streams up, float3 + one
    byte of ring length per vertex down -- the same fragments, bit for bit, as the float4 calls and as the oracle
    (config 4's first event: 1000 pieces x 64 cells, and the unit cube x 256 cells)."""
    for pieces, cells in ((common.voronoi(1234, 1000), common.voronoi(46354, 64)), (common.unit_cube(), common.voronoi(46354, 256))):
        ref = common.run_gpu(ctx, pieces, cells)
        want = P.apply_fracture(pieces, cells.planes, cells.plane_off)
        common.assert_fragments_equal(ref, want)
        ctx.upload_pieces3(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
        ctx.upload_cells3(cells.planes, cells.plane_off, cells.verts, cells.vert_off)
        ctx.fracture_event()
        got = ctx.download_packed()
        common.assert_fragments_equal(got, want)
        assert got.rec.tobytes() == ref.rec.tobytes()
        assert np.array_equal(bits(got.verts[:, :3]), bits(ref.verts[:, :3]))
        assert np.array_equal(got.ring_off, ref.ring_off) and np.array_equal(got.ring, ref.ring)
        # mixed use: a packed download after a float4 upload, and the plain download after a packed one
        common.run_gpu(ctx, pieces, cells)
        assert ctx.download_packed().rec.tobytes() == ref.rec.tobytes()
        again = ctx.download()
        assert np.array_equal(bits(again.verts), bits(ref.verts)) and np.array_equal(again.ring_off, ref.ring_off)


def test_one_copy_blob_transfers(ctx):
    """surtr_upload_blob / surtr_download_blob_async: one blob per direction.  Same fragments, bit for bit, as the plain
    calls and the oracle, for a batch of independent events (config 4 shape), a single event, unbounded cells, and a
    plain upload after a blob upload (the index arrays were views into the blob)."""
    import ctypes as C
    from surtr_b200 import FractureContext
    psets = [common.voronoi(1234 + e, 1000) for e in range(3)]
    csets = [common.voronoi(46354 + e, 64) for e in range(3)]
    pieces, ev_p = common.concat(psets)
    cells, ev_c = common.concat(csets)
    for kw in (dict(ev_piece_off=ev_p, ev_cell_off=ev_c), dict(), dict(bounded=False)):
        if "ev_piece_off" in kw:
            ref = common.run_gpu(ctx, pieces, cells, ev_p, ev_c)
        else:
            ref = common.run_gpu(ctx, pieces, cells, bounded=kw.get("bounded", True))
        sizes, total = FractureContext.fill_input_blob(None, pieces, cells, **kw)
        buf = np.zeros(total, np.uint8)
        FractureContext.fill_input_blob(buf, pieces, cells, **kw)
        ctx.upload_blob_ptr(buf.ctypes.data, sizes)
        ctx.fracture_event()
        c = ctx.counts()
        out = np.zeros(64 * c.n_fragments + 13 * c.n_verts + 2 * c.n_ring + 4 * 256, np.uint8)
        L = ctx.download_blob_into_async(out.ctypes.data, len(out))
        ctx.sync()
        got = FractureContext.unpack_output_blob(out, L)
        assert got.rec.tobytes() == ref.rec.tobytes()
        assert np.array_equal(bits(got.verts), bits(ref.verts))
        assert np.array_equal(got.ring_off, ref.ring_off) and np.array_equal(got.ring, ref.ring)
        # too small a host blob is refused with the needed size reported
        with pytest.raises(Exception):
            ctx.download_blob_into_async(out.ctypes.data, 16)
    # a piece and fragments of more than 256 vertices (the bunny mesh in the global tier): two-byte ring entries both ways
    d = np.load(os.path.join(GOLDEN, "bunny_mesh_x32.npz"))
    mesh = load_polyset(d, "mesh_")
    mcells = common.PolySet(d["cell_verts"], d["cell_vert_off"], None, None)
    mcells.planes, mcells.poly_face_off = d["planes"], d["plane_off"]
    ctx.upload_pieces(mesh.verts, mesh.vert_off, mesh.ring_off, mesh.ring)
    ctx.upload_cells(mcells.planes, mcells.plane_off, mcells.verts, mcells.vert_off)
    ctx.fracture_event()
    ref = ctx.download()
    sizes, total = FractureContext.fill_input_blob(None, mesh, mcells)
    assert sizes[-1] == 2
    buf = np.zeros(total, np.uint8)
    FractureContext.fill_input_blob(buf, mesh, mcells)
    ctx.upload_blob_ptr(buf.ctypes.data, sizes)
    ctx.fracture_event()
    c = ctx.counts()
    out = np.zeros(64 * c.n_fragments + 13 * c.n_verts + 2 * c.n_ring + 4 * 256, np.uint8)
    L = ctx.download_blob_into_async(out.ctypes.data, len(out))
    ctx.sync()
    assert int(L.ring_entry_bytes) == 2 and int(ref.rec["n_verts"].max()) > 256
    got = FractureContext.unpack_output_blob(out, L)
    assert got.rec.tobytes() == ref.rec.tobytes() and np.array_equal(bits(got.verts), bits(ref.verts))
    assert np.array_equal(got.ring_off, ref.ring_off) and np.array_equal(got.ring, ref.ring)

    want = P.apply_fracture(psets[0], csets[0].planes, csets[0].plane_off)
    again = common.run_gpu(ctx, psets[0], csets[0])          # plain upload over the blob's views
    common.assert_fragments_equal(again, want)
    # recursion after a blob upload: the blob's index arrays are not recycled as output arrays
    sizes, total = FractureContext.fill_input_blob(None, psets[0], csets[0])
    buf = np.zeros(total, np.uint8)
    FractureContext.fill_input_blob(buf, psets[0], csets[0])
    ctx.upload_blob_ptr(buf.ctypes.data, sizes)
    ctx.fracture_event()
    lvl1 = ctx.download()
    ctx.fragments_to_pieces()
    ctx.upload_cells(csets[1].planes, csets[1].poly_face_off, csets[1].verts, csets[1].vert_off)
    ctx.fracture_event()
    lvl2 = ctx.download()
    want2 = P.apply_fracture(common.fragments_as_polyset(lvl1), csets[1].planes, csets[1].poly_face_off)
    common.assert_fragments_equal(lvl2, want2)


def test_global_tier_workspace_grows(monkeypatch):
    """A global-tier workspace that runs out of vertex slots is doubled and the event re-run (never a failed pair): the
    bunny mesh (2503 vertices) starting from a 1024-slot workspace (test hook SURTR_DEBUG_CAP3)."""
    from surtr_b200 import FractureContext
    monkeypatch.setenv("SURTR_DEBUG_CAP3", "1024")
    d = np.load(os.path.join(GOLDEN, "bunny_mesh_x32.npz"))
    mesh, want = load_polyset(d, "mesh_"), load_polyset(d, "frag_")
    cx = FractureContext(0)
    try:
        cx.upload_pieces(mesh.verts, mesh.vert_off, mesh.ring_off, mesh.ring)
        cx.upload_cells(d["planes"], d["plane_off"], d["cell_verts"], d["cell_vert_off"])
        cx.fracture_event()
        common.assert_fragments_equal(cx.download(), want)
    finally:
        cx.close()
