"""Host-side helpers of the synthetic workloads (CPU)."""
import numpy as np

import common
from surtr_b200 import synth


def _cell(cs, i):
    v0, v1 = int(cs.vert_off[i]), int(cs.vert_off[i + 1])
    p0, p1 = int(cs.plane_off[i]), int(cs.plane_off[i + 1])
    rings = [cs.ring[cs.ring_off[v]:cs.ring_off[v + 1]].tolist() for v in range(v0, v1)]
    return cs.verts[v0:v1].tobytes(), cs.planes[p0:p1].tobytes(), rings


def test_roll_cells_renumbers_the_cells_and_nothing_else():
    """bench.py lays out its resident input sets as the same pattern with the cells renumbered: cell i of the rolled
    set is cell (i + r) mod n of the original, vertex for vertex, plane for plane, ring for ring."""
    ps = common.voronoi(46354, 64)
    cs = synth.CellSet(ps.verts, ps.vert_off, ps.ring_off, ps.ring, ps.planes, ps.plane_off)
    n = cs.n
    for r in (0, 1, 17, 63, 64, 65):
        rolled = synth.roll_cells(cs, r)
        assert rolled.n == n
        assert rolled.vert_off[-1] == cs.vert_off[-1] and rolled.plane_off[-1] == cs.plane_off[-1]
        assert rolled.ring_off[-1] == cs.ring_off[-1] and len(rolled.ring) == len(cs.ring)
        for i in (0, 1, n // 2, n - 1):
            assert _cell(rolled, i) == _cell(cs, (i + r) % n)
    assert synth.roll_cells(cs, 0) is cs
    # distinct byte streams: that is what makes the bench's input sets different data for the L2
    assert synth.roll_cells(cs, 1).planes.tobytes() != cs.planes.tobytes()


def test_algorithmic_bytes_formula():
    """16*V_in + 4*E2_in + 16*P + 16*V_out + 4*E2_out + 64 per surviving pair."""
    rec = np.zeros(2, dtype=[("piece", np.uint32), ("cell", np.uint32), ("n_verts", np.uint32), ("n_ring", np.uint32)])
    rec["cell"] = [0, 1]
    rec["n_verts"] = [10, 12]
    rec["n_ring"] = [30, 36]
    vo = np.array([0, 8], np.uint32)
    ro = np.arange(0, 25, 3, dtype=np.uint32)
    po = np.array([0, 5, 12], np.uint32)
    want = (16 * 8 + 4 * 24 + 16 * 5 + 16 * 10 + 4 * 30 + 64) + (16 * 8 + 4 * 24 + 16 * 7 + 16 * 12 + 4 * 36 + 64)
    assert synth.algorithmic_bytes(vo, ro, po, rec) == want


def test_bench_clock_sampler_parses_nvidia_smi_lines():
    """bench.py's clocks object: median SM clock under load, the maximum clock and the throttle reasons seen."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.lines = ["0, 1965, 1965, 412.3, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
               "0, 1950, 1965, 690.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
               "0, 1965, 1965, 500.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
               "garbage line", "0, [N/A], 1965, 1, 0, Not Active, Not Active, Not Active, Not Active"]
    c = s.stop()
    assert c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["samples"] == 3 and c["reasons"] == ["sw_power_cap"]
    s2 = bench.ClockSampler(0)
    assert s2.stop()["reasons"] == ["nvidia-smi unavailable"]
