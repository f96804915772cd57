"""Randomised parity sweep: many differently shaped convex pieces against many random plane lists -- rotated pieces,
generic planes, planes through exact vertex coordinates, duplicated planes, long and short lists -- every pair reaching
K3 (cells without bounds), GPU vs the oracle port bit for bit.  The fixtures pin the oracle to the reference build;
this sweep hunts for rare topology cases the fixtures do not contain."""
import numpy as np
import pytest

import common
from common import bits
from oracle import portapi as P

pytestmark = pytest.mark.gpu


def rotation(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q.astype(np.float32)


def random_pieces(rng, seed):
    """Voronoi cells of a few seed sets (4-50 vertices each), some rotated / scaled / moved in float32."""
    sets = []
    for k, n in enumerate((24, 60, 150)):
        cells = common.voronoi(seed * 10 + k, n)
        ps = cells.subset(range(cells.n))
        ps.verts = cells.verts.copy()
        if k:
            r = rotation(rng)
            s = np.float32(rng.uniform(0.5, 3.0))
            t = rng.uniform(-2, 2, 3).astype(np.float32)
            ps.verts[:, :3] = ((cells.verts[:, :3] @ r.T).astype(np.float32) * s).astype(np.float32) + t
        ps.planes = None
        sets.append(ps)
    pieces, _ = common.concat(sets)
    return pieces


def random_cells(rng, pieces, n_cells):
    lo, hi = pieces.verts[:, :3].min(0), pieces.verts[:, :3].max(0)
    planes, off = [], [0]
    for _ in range(n_cells):
        kind = rng.randint(4)
        k = int(rng.choice([1, 2, 3, 5, 8, 13, 21, 40]))
        pl = []
        for _ in range(k):
            n = rng.normal(size=3).astype(np.float32)
            if kind == 0:        # generic plane through a random point of the bounding box, normal of any length
                p = rng.uniform(lo, hi).astype(np.float32)
            elif kind == 1:      # plane through a vertex of some piece (that vertex is not in-plane in general)
                p = pieces.verts[rng.randint(len(pieces.verts)), :3]
            elif kind == 2:      # axis plane through an exact vertex coordinate: that vertex IS in-plane
                v = pieces.verts[rng.randint(len(pieces.verts)), :3]
                ax = rng.randint(3)
                n = np.zeros(3, np.float32)
                n[ax] = np.float32(1.0 if rng.rand() < 0.5 else -1.0)
                pl.append([n[0], n[1], n[2], -n[ax] * v[ax]])
                continue
            else:                # far planes that keep everything, mixed with repeats of earlier ones
                if pl and rng.rand() < 0.4:
                    pl.append(list(pl[rng.randint(len(pl))]))
                    continue
                p = (lo - np.float32(5.0) * (hi - lo)).astype(np.float32) if rng.rand() < 0.5 else rng.uniform(lo, hi).astype(np.float32)
            d = -np.float32(np.float32(n[0] * p[0]) + np.float32(n[1] * p[1]) + np.float32(n[2] * p[2]))
            pl.append([n[0], n[1], n[2], d])
        planes.extend(pl)
        off.append(len(planes))
    return np.asarray(planes, np.float32), np.asarray(off, np.uint32)


N_CELLS = 800


@pytest.mark.parametrize("seed", list(range(1, 17)))
def test_random_pairs_match_oracle(ctx, seed):
    """234 pieces x 800 plane lists = 187 K pairs per seed (3 M in all), in the small tier (a few overflow into the large one)."""
    rng = np.random.RandomState(1000 + seed)
    pieces = random_pieces(rng, seed)
    planes, off = random_cells(rng, pieces, N_CELLS)
    want = P.apply_fracture(pieces, planes, off, cap_frags=pieces.n * N_CELLS + 16)
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(planes, off)
    ctx.fracture_event()
    got = ctx.download()
    c = ctx.counts()
    assert c.n_candidates == pieces.n * N_CELLS
    common.assert_fragments_equal(got, want)
    assert want.n > 10000 and c.n_seq_cuts > 0


@pytest.mark.parametrize("seed", list(range(20, 26)))
def test_random_pairs_large_tiers_match_oracle(ctx, seed):
    """The same random plane lists against the 107-vertex ACH (large tier, shared-memory workspace) and the 2503-vertex
    non-convex bunny mesh (global-memory tier), both rotated and moved."""
    import os
    rng = np.random.RandomState(2000 + seed)
    a = np.load(os.path.join(common.GOLDEN, "config1_bunny32.npz"))
    m = np.load(os.path.join(common.GOLDEN, "bunny_mesh_x32.npz"))
    ach = common.PolySet(a["ach_verts"].copy(), a["ach_vert_off"], a["ach_ring_off"], a["ach_ring"])
    mesh = common.PolySet(m["mesh_verts"].copy(), m["mesh_vert_off"], m["mesh_ring_off"], m["mesh_ring"])
    for ps in (ach, mesh):
        r = rotation(rng)
        t = rng.uniform(-3, 3, 3).astype(np.float32)
        ps.verts[:, :3] = (ps.verts[:, :3] @ r.T).astype(np.float32) + t
    pieces, _ = common.concat([ach, mesh])
    planes, off = random_cells(rng, pieces, 96)
    want = P.apply_fracture(pieces, planes, off, cap_frags=4096, cap_verts=600000)
    ctx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
    ctx.upload_cells(planes, off)
    ctx.fracture_event()
    got = ctx.download()
    c = ctx.counts()
    common.assert_fragments_equal(got, want)
    # ACH pairs: 128-slot warp tier, the few whose result outgrows the small blob go on to the large tier; mesh pairs: global tier
    assert c.n_tier1b == 96 and c.n_tier2 < 96 and c.n_tier3 >= 96 and want.n > 20
    if seed == 20:
        # the same event with the 128-slot tier switched off (test hook): every ACH pair is cut by the large tier's own code
        import os as _os
        from surtr_b200 import FractureContext
        _os.environ["SURTR_DEBUG_NO_TIER1B"] = "1"
        try:
            cx = FractureContext(0)
        finally:
            del _os.environ["SURTR_DEBUG_NO_TIER1B"]
        try:
            cx.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring)
            cx.upload_cells(planes, off)
            cx.fracture_event()
            common.assert_fragments_equal(cx.download(), want)
            assert cx.counts().n_tier2 == 96 and cx.counts().n_tier1b == 0
        finally:
            cx.close()
