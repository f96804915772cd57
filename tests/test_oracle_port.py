"""CPU tests: the oracle's C port (oracle/surtr_oracle.c) against the fixtures generated from the REFERENCE build
(tests/golden/, see make_golden.py) and, when oracle/_ref is present, against the reference build directly.
This is what pins the oracle before it is trusted as the checker of the CUDA path."""
import json
import os

import numpy as np
import pytest

import common
from common import GOLDEN, bits
from oracle import portapi as P
from oracle import refapi as R
from oracle.refapi import PolySet


def load_polyset(d, prefix) -> PolySet:
    ps = PolySet(d[prefix + "verts"], d[prefix + "vert_off"], d[prefix + "ring_off"], d[prefix + "ring"])
    for k in ("cell", "piece", "nfaces", "volume", "centroid"):
        if prefix + k in d:
            setattr(ps, k, d[prefix + k])
    if prefix + "planes" in d:
        ps.planes = d[prefix + "planes"]
        ps.poly_face_off = d[prefix + "plane_off"]
    return ps


def assert_polysets_equal(a: PolySet, b: PolySet, moments=True):
    assert a.n == b.n
    for f in ("verts", "vert_off", "ring_off", "ring", "cell", "piece", "nfaces"):
        assert np.array_equal(bits(getattr(a, f)), bits(getattr(b, f))), f
    if moments:
        assert np.array_equal(bits(a.volume), bits(b.volume)), "volume"
        assert np.array_equal(bits(a.centroid), bits(b.centroid)), "centroid"


def test_scalar_kats_match_reference():
    d = np.load(os.path.join(GOLDEN, "kat_scalar.npz"))
    n = len(d["planes"])
    comp = np.array([P.compare_plane_point(d["planes"][i], d["pts"][i]) for i in range(n)], np.int32)
    assert np.array_equal(comp, d["comp"])
    assert set(np.unique(d["comp"])) == {-1, 0, 1}     # the in-plane band is exercised
    inter = np.stack([P.plane_line_intersection(d["a"][i], d["b"][i], d["planes"][i]) for i in range(n)])
    assert np.array_equal(bits(inter), bits(d["inter"]))
    p3 = np.stack([P.plane_from_points(d["a"][i], d["b"][i], d["c"][i]) for i in range(n)])
    assert np.array_equal(bits(p3), bits(d["plane3"]))
    pn = np.stack([P.plane_from_point_normal(d["a"][i], d["b"][i]) for i in range(n)])
    assert np.array_equal(bits(pn), bits(d["plane_pn"]))


def test_known_answers_unit_cube():
    """unit cube cut by x+y+z <= 0.15 through (0.1, 0.05, 0) -> 10 verts, 7 faces,
    vol 0.611374984; unit cube against VMACH::GetBoxPolygon planes -> 8 verts, vol 1."""
    cube = common.unit_cube()
    pl = P.plane_from_point_normal([0.1, 0.05, 0.0], [1, 1, 1])
    r = P.clip_each(cube, pl[None], [0, 1])
    assert (int(r.nverts[0]), int(r.nfaces[0])) == (10, 7)
    assert abs(r.volume[0] - 0.611374984) < 1e-9
    box = np.load(os.path.join(GOLDEN, "kat_scalar.npz"))["box_planes"]
    r = P.clip_each(cube, box, [0, 6])
    assert (int(r.nverts[0]), int(r.nfaces[0])) == (8, 6) and r.volume[0] == 1.0
    # fully outside -> empty, and an empty plane list leaves the piece untouched
    out = P.plane_from_point_normal([-2.0, 0, 0], [1, 0, 0])
    assert int(P.clip_each(cube, out[None], [0, 1]).nverts[0]) == 0
    assert int(P.clip_each(cube, np.zeros((0, 4), np.float32), [0, 0]).nverts[0]) == 8


@pytest.mark.parametrize("name", ["cube_x64", "pieces200_x32"])
def test_golden_events_bit_exact(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    cells = load_polyset(d, "cells_")
    pieces = load_polyset(d, "pieces_")
    want = load_polyset(d, "frag_")
    got = P.apply_fracture(pieces, cells.planes, cells.plane_off)
    assert_polysets_equal(got, want)
    # the cell builder (new derivation) reproduces the reference-clipper cells from the same neighbour lists
    if name == "cube_x64":
        c2 = P.voronoi_cells(d["seeds"], d["nb_off"], d["nb_idx"])
    else:
        c2 = P.voronoi_cells(d["cell_seeds"], d["cell_nb_off"], d["cell_nb_idx"])
    assert np.array_equal(bits(c2.verts), bits(cells.verts)) and np.array_equal(c2.ring, cells.ring)
    assert np.array_equal(bits(c2.planes), bits(cells.planes)) and np.array_equal(c2.plane_off, cells.plane_off)


def _summaries():
    return json.load(open(os.path.join(GOLDEN, "summaries.json")))


def test_summary_config2_cube_x4096():
    cells = common.voronoi(46354, 4096)
    got = P.apply_fracture(common.unit_cube(), cells.planes, cells.plane_off)
    assert common.summary_of_polyset(got) == _summaries()["config2_cube_x4096"]
    assert got.n == 4096


def test_summary_config4_event0():
    cells = common.voronoi(46354, 64)
    got = P.apply_fracture(common.voronoi(1234, 1000), cells.planes, cells.plane_off)
    assert common.summary_of_polyset(got) == _summaries()["config4_e0_1000x64"]
    assert got.n == 2841     # SURVEY.md section 4 integration count


def test_summary_config3_10000x256():
    cells = common.voronoi(46354, 256)
    got = P.apply_fracture(common.voronoi(1234, 10000), cells.planes, cells.plane_off)
    assert common.summary_of_polyset(got) == _summaries()["config3_10000x256"]
    assert got.n == 23864    # SURVEY.md section 4 integration count


def test_summary_config5_recursion():
    pieces = common.unit_cube()
    for lvl, cells in enumerate(common.recursion_levels()):
        fr = P.apply_fracture(pieces, cells.planes, cells.plane_off)
        assert common.summary_of_polyset(fr) == _summaries()[f"config5_level{lvl}"]
        pieces = fr


def test_invariants_euler_and_ring_symmetry():
    cells = common.voronoi(46354, 64)
    fr = P.apply_fracture(common.voronoi(1234, 300), cells.planes, cells.plane_off)
    for i in range(fr.n):
        _, rings = fr.poly(i)
        e2 = sum(len(r) for r in rings)
        assert len(rings) - e2 // 2 + int(fr.nfaces[i]) == 2       # V - E + F = 2
        for v, r in enumerate(rings):                               # the reference's own check, Poly.cpp:253-260
            for u in r:
                assert v in rings[u]


def test_kdop_and_ach_against_reference_fixture():
    d = np.load(os.path.join(GOLDEN, "config1_kdop.npz"))
    for key, nich in (("bunny", 28), ("cube", 12), ("sphere", 36)):
        v4, normals = d[key + "_verts"], d[key + "_normals"]
        assert len(normals) == nich
        dist, arg, planes = P.kdop_calc(v4, normals)
        assert np.array_equal(dist.astype(np.float64), d[key + "_poly_dist"])
        assert np.array_equal(bits(planes), bits(d[key + "_poly_planes"]))
        assert np.array_equal(bits(v4[arg[:, 0], :3]), bits(d[key + "_poly_vtx"][:, 0]))
        assert np.array_equal(bits(v4[arg[:, 1], :3]), bits(d[key + "_poly_vtx"][:, 1]))
        # ACH: the 2x bounding box clipped by [Min0, Max0, Min1, ...] (Kdop.cpp:166-179)
        box = common.unit_cube()
        box.verts = d[key + "_seedbox_verts"]
        ach = P.clip_each(box, d[key + "_gap_planes"].reshape(-1, 4), [0, 2 * nich])
        want = load_polyset(d, key + "_ach_")
        assert np.array_equal(bits(ach.verts), bits(want.verts)) and np.array_equal(ach.ring, want.ring)
        assert np.array_equal(ach.nfaces, want.nfaces) and np.array_equal(bits(ach.volume), bits(want.volume))
    a = load_polyset(d, "cube_ach_")
    assert (int(a.nverts[0]), int(a.nfaces[0])) == (8, 7) and abs(a.volume[0] - 216.648651) < 1e-5   # SURVEY KAT


def test_inertia_port_sanity():
    """so_inertia on the unit cube: I = m/6 = 1/6 on the diagonal, zero products."""
    cube = common.unit_cube()
    box = np.load(os.path.join(GOLDEN, "kat_scalar.npz"))["box_planes"]
    fr = P.apply_fracture(cube, box, [0, 6], inertia=True)
    assert np.allclose(fr.inertia[0], [1 / 6, 1 / 6, 1 / 6, 0, 0, 0], atol=1e-12)


@pytest.mark.skipif(not common.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_matches_reference_build_directly():
    assert np.array_equal(common.seeds_uniform(46354, 4096), R.seeds_uniform(46354, 4096))
    assert np.array_equal(common.seeds_uniform(1234, 777), R.seeds_uniform(1234, 777))
    pieces, cells = common.voronoi(1234, 400), common.voronoi(46354, 48)
    a = R.apply_fracture(pieces, cells.planes, cells.plane_off, 16)
    b = P.apply_fracture(pieces, cells.planes, cells.plane_off)
    assert_polysets_equal(b, a)
    # qhull neighbour lists contain DT3D::Triangulate's and give the same cells in (V, F)
    s = common.seeds_uniform(46354, 256)
    off, idx, _ = R.dt3d_neighbors(s)
    o2, i2 = common.scipy_neighbors(s)
    for i in range(len(s)):
        assert set(idx[off[i]:off[i + 1]]) <= set(i2[o2[i]:o2[i + 1]])
    c1, c2 = R.voronoi_cells(s, off, idx), P.voronoi_cells(s, o2, i2)
    assert np.array_equal(c1.nverts, c2.nverts) and np.array_equal(c1.nfaces, c2.nfaces)
    # ... and wherever the two neighbour sets coincide (247 of these 256 seeds) the cell is the same bit for bit: vertex
    # positions and order, rings, face planes.  (The extra qhull neighbours of the other 9 only add bisectors that do not
    # cut, so on this seed set every cell is identical.)
    n_same = 0
    for i in range(len(s)):
        v = slice(int(c1.vert_off[i]), int(c1.vert_off[i + 1]))
        f = slice(int(c1.poly_face_off[i]), int(c1.poly_face_off[i + 1]))
        r = slice(int(c1.ring_off[v.start]), int(c1.ring_off[v.stop]))
        same = (c1.verts[v].tobytes() == c2.verts[v].tobytes() and c1.planes[f].tobytes() == c2.planes[f].tobytes() and
                np.array_equal(c1.ring[r], c2.ring[r]))
        if set(idx[off[i]:off[i + 1]]) == set(i2[o2[i]:o2[i + 1]]):
            n_same += 1
            assert same, f"cell {i}: same Delaunay neighbours but different cell"
    assert n_same > 200
    assert c1.verts.tobytes() == c2.verts.tobytes() and c1.planes.tobytes() == c2.planes.tobytes()


def test_port_degenerate_cuts_match_reference_fixture():
    """Planes through vertices, edges and faces (comp == 0 band, Poly.cpp:303-319, 365-462): port == reference build."""
    d = np.load(os.path.join(GOLDEN, "degenerate_x400.npz"))
    pieces, want = load_polyset(d, "pieces_"), load_polyset(d, "frag_")
    got = P.apply_fracture(pieces, d["planes"], d["plane_off"])
    assert_polysets_equal(got, want)


def test_port_degenerate_cuts_on_large_pieces_match_reference_summary():
    import json
    pieces, planes, off = common.degenerate_large_inputs()
    got = P.apply_fracture(pieces, planes, off, cap_frags=256, cap_verts=400000)
    want = json.load(open(os.path.join(GOLDEN, "summaries.json")))["degenerate_large"]
    s = common.summary_of_polyset(got)
    s["ring"] = common.digest(np.asarray(got.ring, np.uint16))
    assert s == want


@pytest.mark.skipif(not common.have_ref(), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("seed", [1, 9])
def test_port_matches_reference_on_random_pairs(seed):
    """The randomised sweep of tests/test_gpu_fuzz.py checks the GPU against the PORT; here the port is checked against
    the reference build on inputs of the same generator (rotated pieces, in-plane planes, long plane lists)."""
    import test_gpu_fuzz as F
    rng = np.random.RandomState(1000 + seed)
    pieces = F.random_pieces(rng, seed)
    planes, off = F.random_cells(rng, pieces, 120)
    want = R.apply_fracture(pieces, planes, off, 8)
    got = P.apply_fracture(pieces, planes, off, cap_frags=pieces.n * 120 + 16)
    assert_polysets_equal(got, want)
    assert want.n > 3000
