"""Generates the committed fixtures under tests/golden/ from the REFERENCE build (oracle/_ref/libsurtr_ref.so =
the reference's own Src/Poly.cpp, Src/Kdop.cpp, Src/VMACH.cpp, Inc/DT3D.h compiled headless by oracle/Makefile).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own; these are outputs of the reference
itself run here, which is what pins the oracle port and the CUDA path.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
from oracle import refapi as R  # noqa: E402

REF_MODELS = "/root/reference/Resources/Models"


def save_polyset(d, prefix, ps, full=True):
    d[prefix + "verts"] = ps.verts
    d[prefix + "vert_off"] = ps.vert_off
    d[prefix + "ring_off"] = ps.ring_off
    d[prefix + "ring"] = ps.ring
    if full:
        for k in ("cell", "piece", "nfaces", "volume", "centroid"):
            d[prefix + k] = getattr(ps, k)
    if ps.planes is not None:
        d[prefix + "planes"] = ps.planes
        d[prefix + "plane_off"] = ps.poly_face_off
    if ps.face_idx is not None:     # Poly::ExtractFaces loops of the reference build
        d[prefix + "face_off"] = ps.face_off
        d[prefix + "face_idx"] = ps.face_idx


def scalar_kats():
    rng = np.random.RandomState(7)
    planes = rng.uniform(-1, 1, (400, 4)).astype(np.float32)
    pts = rng.uniform(-1, 1, (400, 3)).astype(np.float32)
    # force exact in-plane / near-plane cases (the |s| < 1e-10 band, Poly.cpp:719)
    for i in range(0, 100):
        n = planes[i, :3]
        p = pts[i]
        planes[i, 3] = -np.float32((np.float32(n[0] * p[0]) + np.float32(n[1] * p[1])) + np.float32(n[2] * p[2]))
    planes[100:110] = [[1, 0, 0, -0.5]] * 10
    pts[100:110, 0] = [0.5, np.nextafter(np.float32(0.5), np.float32(1)), np.nextafter(np.float32(0.5), np.float32(0)),
                       0.5 + 1e-10, 0.5 - 1e-10, 0.5, 0.5, -0.5, 0.0, 1.0]
    comp = np.array([R.compare_plane_point(planes[i], pts[i]) for i in range(400)], np.int32)
    a = rng.uniform(-1, 1, (400, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (400, 3)).astype(np.float32)
    inter = np.stack([R.plane_line_intersection(a[i], b[i], planes[i]) for i in range(400)])
    c = rng.uniform(-1, 1, (400, 3)).astype(np.float32)
    p3 = np.stack([R.plane_from_points(a[i], b[i], c[i]) for i in range(400)])
    pn = np.stack([R.plane_from_point_normal(a[i], b[i]) for i in range(400)])
    np.savez_compressed(os.path.join(HERE, "kat_scalar.npz"), planes=planes, pts=pts, comp=comp, a=a, b=b, c=c,
                        inter=inter, plane3=p3, plane_pn=pn, box_planes=R.box_planes())


def small_events():
    d = {}
    cube = R.unit_cube()
    s = R.seeds_uniform(46354, 64)
    off, idx, _ = R.dt3d_neighbors(s)           # the reference's own DT3D::Triangulate
    cells = R.voronoi_cells(s, off, idx)
    d["seeds"] = s
    d["nb_off"] = off
    d["nb_idx"] = idx
    save_polyset(d, "cells_", cells)
    save_polyset(d, "pieces_", cube, full=False)
    save_polyset(d, "frag_", R.apply_fracture(cube, cells.planes, cells.plane_off, 0))
    np.savez_compressed(os.path.join(HERE, "cube_x64.npz"), **d)

    d = {}
    sp = R.seeds_uniform(1234, 200)
    po, pi, _ = R.dt3d_neighbors(sp)
    pieces = R.voronoi_cells(sp, po, pi)
    sc = R.seeds_uniform(46354, 32)
    co, ci, _ = R.dt3d_neighbors(sc)
    cells = R.voronoi_cells(sc, co, ci)
    d["piece_seeds"] = sp
    d["cell_seeds"] = sc
    d["piece_nb_off"], d["piece_nb_idx"], d["cell_nb_off"], d["cell_nb_idx"] = po, pi, co, ci
    save_polyset(d, "cells_", cells)
    save_polyset(d, "pieces_", pieces)
    save_polyset(d, "frag_", R.apply_fracture(pieces, cells.planes, cells.plane_off, 16))   # dp::thread_pool(16)
    np.savez_compressed(os.path.join(HERE, "pieces200_x32.npz"), **d)


def load_obj(path, scale):
    """Surtr::LoadModelData (Surtr.cpp:2683-2727): v/f only, x negated (:2714), winding flipped."""
    v, f = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([-float(t[1]) * scale, float(t[2]) * scale, float(t[3]) * scale])
        elif t[0] == "f":
            idx = [int(x.split("/")[0]) - 1 for x in t[1:]]
            for k in range(1, len(idx) - 1):
                f.append([idx[0], idx[k + 1], idx[k]])
    return np.asarray(v, np.float32), np.asarray(f, np.int32)


def config1_kdop():
    """Config 1 front end on the bundled bunny (scale 70, Surtr.cpp:1400): ICH normals -> k-DOP with gap ->
    ACH = 2x bbox clipped by [Min0, Max0, Min1, ...] (PrepareFracture steps 1-6, Surtr.cpp:1747-1785)."""
    d = {}
    for name, scale in (("lowpoly-bunny-closed", 70.0), ("cube", 3.0), ("highpoly-sphere", 5.0)):
        path = os.path.join(REF_MODELS, name + ".obj")
        v, _ = load_obj(path, scale)
        v4 = np.zeros((len(v), 4), np.float32)
        v4[:, :3] = v
        normals = R.ich_normals(v4, 20)
        lo, hi = v.min(0).astype(np.float64), v.max(0).astype(np.float64)
        max_axis = float(np.max(hi - lo))
        dist, planes, vtx = R.kdop_calc_gap(v4, normals, max_axis, 2000.0)
        dist2, planes2, vtx2 = R.kdop_calc_poly(v4, normals)
        # ACH seed: GetBB scaled by extent, by 2, translated to the centre (float ops of Poly::Scale/Translate)
        cube = common.unit_cube()
        ext = (hi - lo).astype(np.float32)
        ctr = ((hi + lo) / 2.0).astype(np.float32)
        cv = cube.verts.copy()
        cv[:, :3] = (cv[:, :3] * ext) * np.float32(2.0) + ctr
        cube.verts = cv
        ach = R.clip_each(cube, planes.reshape(-1, 4), np.array([0, 2 * len(normals)], np.uint32))
        key = name.split("-")[-2] if "-" in name else name
        key = {"lowpoly-bunny-closed": "bunny", "cube": "cube", "highpoly-sphere": "sphere"}[name]
        d[key + "_verts"] = v4
        d[key + "_normals"] = normals
        d[key + "_gap_dist"], d[key + "_gap_planes"], d[key + "_gap_vtx"] = dist, planes, vtx
        d[key + "_poly_dist"], d[key + "_poly_planes"], d[key + "_poly_vtx"] = dist2, planes2, vtx2
        d[key + "_seedbox_verts"] = cv
        save_polyset(d, key + "_ach_", ach)
        print(key, "ICH faces", len(normals), "ACH", ach.nverts, ach.nfaces, ach.volume)
    np.savez_compressed(os.path.join(HERE, "config1_kdop.npz"), **d)


def summaries():
    out = {}
    cube = common.unit_cube()
    # config 2: unit cube x 4096 cells
    cells = common.voronoi(46354, 4096)
    out["config2_cube_x4096"] = common.summary_of_polyset(R.apply_fracture(cube, cells.planes, cells.plane_off, 16))
    # config 4, event 0: 1000 pieces x 64 cells
    out["config4_e0_1000x64"] = common.summary_of_polyset(
        R.apply_fracture(common.voronoi(1234, 1000), common.voronoi(46354, 64).planes,
                         common.voronoi(46354, 64).plane_off, 16))
    # config 3: 10000 pieces x 256 cells
    c3 = common.voronoi(46354, 256)
    out["config3_10000x256"] = common.summary_of_polyset(
        R.apply_fracture(common.voronoi(1234, 10000), c3.planes, c3.plane_off, 16))
    # config 5: depth-3 recursion, 64 seeds per level
    pieces = cube
    for lvl, cells in enumerate(common.recursion_levels()):
        fr = R.apply_fracture(pieces, cells.planes, cells.plane_off, 16)
        out[f"config5_level{lvl}"] = common.summary_of_polyset(fr)
        pieces = fr
    # in-plane cuts on the large tiers (bunny ACH + bunny mesh), see common.degenerate_large_inputs
    from oracle import portapi as P
    pieces, planes, off = common.degenerate_large_inputs()
    want = R.apply_fracture(pieces, planes, off, 16)
    port = P.apply_fracture(pieces, planes, off, cap_frags=256, cap_verts=400000)
    assert common.summary_of_polyset(port) == common.summary_of_polyset(want) and np.array_equal(port.ring, want.ring)
    out["degenerate_large"] = common.summary_of_polyset(want)
    out["degenerate_large"]["ring"] = common.digest(np.asarray(want.ring, np.uint16))
    json.dump(out, open(os.path.join(HERE, "summaries.json"), "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, v["n"], v["n_verts"], v["sum_volume"])


def refit_fixture():
    """Refitting (m_refittingTask, Surtr.cpp:1449-1455) on 96 fragments of the 200x32 event; the "mesh" of a piece
    is its convex shrunk towards its centroid by a per-piece factor (the real mesh branch is the next row f-1)."""
    d0 = np.load(os.path.join(HERE, "pieces200_x32.npz"))
    from test_oracle_port import load_polyset
    fr = load_polyset(d0, "frag_")
    sel = [i for i in range(fr.n) if fr.nverts[i] >= 6][:96]
    convex = fr.subset(sel)
    rng = np.random.RandomState(3)
    mverts = convex.verts.copy()
    for i in range(convex.n):
        v0, v1 = int(convex.vert_off[i]), int(convex.vert_off[i + 1])
        c = convex.verts[v0:v1, :3].mean(0, dtype=np.float32)
        f = np.float32(rng.uniform(0.6, 0.98))
        mverts[v0:v1, :3] = ((convex.verts[v0:v1, :3] - c) * f + c).astype(np.float32)
    out = R.refit(convex, mverts, convex.vert_off, 4)
    d = {"mesh_verts": mverts, "mesh_vert_off": convex.vert_off}
    save_polyset(d, "convex_", convex, full=False)
    save_polyset(d, "out_", out)
    # the ICH normals of the first pieces, for the host ConvexHull mirror
    nrm, noff = [], [0]
    for i in range(convex.n):
        v0, v1 = int(convex.vert_off[i]), int(convex.vert_off[i + 1])
        n = R.ich_normals(mverts[v0:v1], min(v1 - v0, 4))
        nrm.append(n)
        noff.append(noff[-1] + len(n))
    d["ich_normals"] = np.concatenate(nrm)
    d["ich_normal_off"] = np.asarray(noff, np.uint32)
    np.savez_compressed(os.path.join(HERE, "refit96.npz"), **d)
    print("refit: out verts", out.nverts[:10], "empty", int((out.nverts == 0).sum()))


def config1_fixture():
    """BASELINE config 1, convex branch: bundled bunny (scale 70), 32 seeds mt19937(46354), one ApplyFracture on the
    ACH piece (see ref_config1_convex in oracle/ref_driver.cpp).  Neighbour lists = the product's host DT3D."""
    import hostapi
    d0 = np.load(os.path.join(HERE, "config1_kdop.npz"))
    s = R.seeds_uniform(46354, 32)
    off, idx = hostapi.dt3d_neighbors(s)
    ach, fr = R.config1_convex(d0["bunny_verts"], s, off, idx)
    d = {"verts": d0["bunny_verts"], "seeds": s, "nb_off": off, "nb_idx": idx}
    save_polyset(d, "ach_", ach)
    save_polyset(d, "frag_", fr)
    np.savez_compressed(os.path.join(HERE, "config1_bunny32.npz"), **d)
    print("config1: ACH", ach.nverts, ach.nfaces, "fragments", fr.n, "sum vol", fr.volume.sum())


def mesh_fixture():
    """Row f-1 input: the bundled bunny as a vertex-ring polyhedron (ExtractNeighborFromMesh) cut by 32 Voronoi cells
    placed on it -- the second clip of m_fractureTask (Surtr.cpp:1470).  2503 vertices, ring degree up to 13: the
    global-memory tier of K3."""
    from oracle import portapi as P
    v, f = load_obj(os.path.join(REF_MODELS, "lowpoly-bunny-closed.obj"), 70.0)
    v4 = np.zeros((len(v), 4), np.float32)
    v4[:, :3] = v
    mesh = R.mesh_polyhedron(v4, f)
    lo, hi = v.min(0), v.max(0)
    cells = common.voronoi(46354, 32)
    placed = cells.subset(range(cells.n))
    placed.verts = cells.verts.copy()
    placed.verts[:, :3] = (cells.verts[:, :3] * (hi - lo).astype(np.float32) + ((hi + lo) / 2).astype(np.float32)).astype(np.float32)
    planes, plane_off = P.face_planes(placed)
    want = R.apply_fracture(mesh, planes, plane_off, 16)
    port = P.apply_fracture(mesh, planes, plane_off, cap_frags=64, cap_verts=200000)
    assert port.n == want.n and np.array_equal(port.verts.view(np.uint32), want.verts.view(np.uint32)) and np.array_equal(port.ring, want.ring)
    d = {"planes": planes, "plane_off": plane_off, "cell_verts": placed.verts, "cell_vert_off": placed.vert_off}
    save_polyset(d, "mesh_", mesh, full=False)
    save_polyset(d, "frag_", want)
    for k in list(d):
        if k.endswith("face_off") or k.endswith("face_idx"):
            del d[k]
    np.savez_compressed(os.path.join(HERE, "bunny_mesh_x32.npz"), **d)
    deg = np.diff(mesh.ring_off)
    print("mesh: verts", mesh.nverts, "max degree", int(deg.max()), "fragments", want.n, "verts per fragment", want.nverts[:8], "max", int(want.nverts.max()))


def config1_full_fixture():
    """BASELINE config 1 in full: Surtr::PrepareFracture (Surtr.cpp:1747-1827) on the bundled bunny with its triangle
    list -- ACH, mesh polyhedron, 32 cells, convex + mesh clip with island split, Refitting (ref_config1_full)."""
    import hostapi
    v, f = load_obj(os.path.join(REF_MODELS, "lowpoly-bunny-closed.obj"), 70.0)
    v4 = np.zeros((len(v), 4), np.float32)
    v4[:, :3] = v
    s = R.seeds_uniform(46354, 32)
    off, idx = hostapi.dt3d_neighbors(s)
    ach, convex, mesh = R.config1_full(v4, f, s, off, idx)
    d = {"verts": v4, "indices": np.asarray(f, np.int32).reshape(-1), "seeds": s}
    save_polyset(d, "convex_", convex)
    save_polyset(d, "mesh_", mesh)
    for k in list(d):
        if k.endswith("face_off") or k.endswith("face_idx") or k.endswith("planes") or k.endswith("plane_off"):
            del d[k]
    np.savez_compressed(os.path.join(HERE, "config1_full_bunny32.npz"), **d)
    keys = list(zip(convex.cell.tolist(), convex.piece.tolist()))
    print("config1 full: pieces", convex.n, "pairs with islands", len(keys) - len(set(keys)), "convex verts", convex.nverts[:6],
          "mesh verts", mesh.nverts[:6], "empty convex after refit", int((convex.nverts == 0).sum()))


def degenerate_planes():
    """Cutting planes that hit the unit cube's vertices, edges and faces exactly (all values exact in float32, normals
    deliberately not normalised): the comp == 0 / in-plane band of Poly.cpp:303-319 and the patch cases around it."""
    menu = []
    for ax in range(3):
        for d in (0.5, -0.5, 0.0, 0.25):
            for sgn in (1.0, -1.0):
                n = [0.0, 0.0, 0.0]
                n[ax] = sgn
                menu.append(n + [d])                       # coincident with a face / through the centre / generic
    for a, b in ((0, 1), (0, 2), (1, 2)):
        for sa in (1.0, -1.0):
            for sb in (1.0, -1.0):
                for d in (0.0, -1.0, 1.0, 0.5):            # through 4 vertices / touching an edge / generic
                    n = [0.0, 0.0, 0.0]
                    n[a], n[b] = sa, sb
                    menu.append(n + [d])
    for sx in (1.0, -1.0):
        for sy in (1.0, -1.0):
            for sz in (1.0, -1.0):
                for d in (-0.5, 0.5, -1.5, 1.5, 0.0):      # through 3 vertices / touching one vertex / hexagonal cut
                    menu.append([sx, sy, sz, d])
    return np.asarray(menu, np.float32)


def degenerate_fixture():
    """In-plane and touching cuts: the unit cube and a truncated cube against 400 random sequences of 1-6 planes from
    degenerate_planes(); expected fragments from the reference build."""
    from oracle import portapi as P
    menu = degenerate_planes()
    rng = np.random.RandomState(11)
    planes, off = [], [0]
    for _ in range(400):
        k = rng.randint(1, 7)
        planes.append(menu[rng.randint(0, len(menu), k)])
        off.append(off[-1] + k)
    planes, off = np.concatenate(planes), np.asarray(off, np.uint32)
    cube = common.unit_cube()
    corners = np.asarray([[sx, sy, sz, -1.0] for sx in (1, -1) for sy in (1, -1) for sz in (1, -1)], np.float32)
    trunc = R.clip_each(cube, corners, np.array([0, 8], np.uint32))           # cuboctahedron-like: vertices on edge midpoints
    pieces, _ = common.concat([cube, trunc.subset([0])])
    want = R.apply_fracture(pieces, planes, off, 0)
    port = P.apply_fracture(pieces, planes, off)
    assert port.n == want.n and np.array_equal(port.verts.view(np.uint32), want.verts.view(np.uint32)) and np.array_equal(port.ring, want.ring)
    assert np.array_equal(port.volume.view(np.uint64), want.volume.view(np.uint64)) and np.array_equal(port.nfaces, want.nfaces)
    d = {"planes": planes, "plane_off": off}
    save_polyset(d, "pieces_", pieces, full=False)
    save_polyset(d, "frag_", want)
    for k in list(d):
        if k.endswith("face_off") or k.endswith("face_idx") or k.startswith("frag_plane"):
            del d[k]
    np.savez_compressed(os.path.join(HERE, "degenerate_x400.npz"), **d)
    print("degenerate: pieces", pieces.nverts, "fragments", want.n, "of", 2 * 400, "pairs; verts", int(want.nverts.min()), "-", int(want.nverts.max()))


def transform_fixture():
    """Poly::Transform (Poly.cpp:580-585) KAT: the vertices of three Voronoi pieces under a rigid world matrix, a
    scale + shear, and a projective matrix (w != 1), row-major as the caller holds them."""
    pieces = common.voronoi(7, 40).subset([3, 11, 29])
    rng = np.random.RandomState(9)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    rigid = np.eye(4)
    rigid[:3, :3] = q
    rigid[:3, 3] = [1.25, -3.5, 0.75]
    shear = np.eye(4)
    shear[:3, :3] = [[2.0, 0.3, 0.0], [0.0, 0.5, -0.2], [0.1, 0.0, 1.5]]
    shear[:3, 3] = [0.0, 10.0, -2.0]
    proj = rigid.copy()
    proj[3, :] = [0.05, -0.02, 0.01, 1.5]
    mats = np.stack([rigid, shear, proj]).astype(np.float32)
    out = np.concatenate([R.transform(pieces.verts[pieces.vert_off[i]:pieces.vert_off[i + 1]], mats[i]) for i in range(3)])
    d = {"matrices": mats, "out": out}
    save_polyset(d, "pieces_", pieces, full=False)
    for k in list(d):
        if k.endswith("face_off") or k.endswith("face_idx") or k.endswith("planes") or k.endswith("plane_off"):
            del d[k]
    np.savez_compressed(os.path.join(HERE, "transform_kat.npz"), **d)
    print("transform: verts", len(out), "max |delta|", float(np.abs(out[:, :3] - pieces.verts[:, :3]).max()))


def do_fracture_fixture():
    """Row f-3: Surtr::DoFracture (Surtr.cpp:1885-1959) on the compound PrepareFracture produced for the bunny (the 27
    pieces of config1_full_bunny32.npz), with a 32-cell radial pattern (GenerateFracturePattern, :2072-2096) at an
    impact point on the surface -- once general, once partial (pieces outside the impact sphere stay whole)."""
    import hostapi
    d0 = np.load(os.path.join(HERE, "config1_full_bunny32.npz"))
    from test_oracle_port import load_polyset
    convex, mesh = load_polyset(d0, "convex_"), load_polyset(d0, "mesh_")
    v = d0["verts"][:, :3]
    max_axis = float((v.max(0) - v.min(0)).max())
    cloud, _ = load_obj(os.path.join(REF_MODELS, "sphere.obj"), 0.5)        # m_spherePointCloud (Surtr.cpp:1508)
    impact = v[int(np.argmax(v[:, 2]))].copy()                               # a surface point of the bunny
    d = {"cloud": cloud, "impact": impact, "max_axis_scale": np.float32(max_axis)}
    for name, partial, mean, radius in (("general", False, 1.0, 1.0), ("partial", True, 0.05, 3.0)):
        s = R.seeds_radial(46354, 32, mean)
        off, idx = hostapi.dt3d_neighbors(s)
        c, m, ncomp = R.do_fracture(convex, mesh, s, off, idx, cloud, impact, radius, max_axis, partial)
        d[name + "_seeds"] = s
        d[name + "_radius"] = np.float32(radius)
        d[name + "_ncomp"] = np.int32(ncomp)
        save_polyset(d, name + "_convex_", c)
        save_polyset(d, name + "_mesh_", m)
        print("do_fracture", name, "pieces", c.n, "compounds", ncomp, "untouched", int(c.piece.sum()),
              "pieces per compound", np.bincount(c.cell, minlength=ncomp)[:12], "empty convex", int((c.nverts == 0).sum()))
    for k in list(d):
        if k.endswith("face_off") or k.endswith("face_idx") or k.endswith("planes") or k.endswith("plane_off"):
            del d[k]
    np.savez_compressed(os.path.join(HERE, "do_fracture_bunny.npz"), **d)


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    scalar_kats()
    small_events()
    config1_kdop()
    refit_fixture()
    config1_fixture()
    mesh_fixture()
    config1_full_fixture()
    degenerate_fixture()
    transform_fixture()
    do_fracture_fixture()
    summaries()
