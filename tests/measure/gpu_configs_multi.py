"""Measurement: BASELINE configs 4 and 5 at FULL size sharded over N GPUs of one box.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tests/measure/gpu_configs_multi.py [n_events=4096] [n_objects=4096]

Config 4: 4096 independent fracture events (1000 pieces x 64 cells each); config 5: 4096 unit cubes re-fractured to
depth 3 (64 seeds per level, fragments stay on the device).  Event / object e belongs to rank e mod N
(surtr_b200.sharding.event_shard): every rank cuts its share as ONE batch per level with no
collective on the hot path; the only collective is the NCCL gather of the fragment records to rank 0 at the end,
timed separately.  Times are CUDA-event times of the batch on each rank, maximum over the ranks.  As in
gpu_configs.py the host builds 8 distinct events and tiles them (building 4096 Voronoi sets on the host would take
minutes and is not what is measured).  Rank 0 writes gpurun_out/configs_multi_r1.json."""
import json, os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')   # run from the repo root
import torch
import torch.distributed as dist
from surtr_b200 import FractureContext, sharding
import common

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n_ev = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_obj = int(sys.argv[2]) if len(sys.argv) > 2 else 4096


def reduce(x, op):
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=op)
    return float(t.item())


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


out = {"n_gpus": world}

# ---- config 4 ----
mine = sharding.event_shard(n_ev, world, rank)
base_p = [common.voronoi(1234 + e, 1000) for e in range(8)]
base_c = [common.voronoi(46354 + e, 64) for e in range(8)]
pieces, ev_p = common.concat([base_p[e % 8] for e in mine])
cells, ev_c = common.concat([base_c[e % 8] for e in mine])
ctx = FractureContext(local)
fr = common.run_gpu(ctx, pieces, cells, ev_p, ev_c)      # warm-up: buffers grow, tiers enable
ctx.fracture_event(); ctx.counts()
barrier()
ts = []
for _ in range(7):
    ctx.fracture_event(); ctx.counts(); ts.append(ctx.last_event_ms()[0])
ms = reduce(float(np.median(ts)), dist.ReduceOp.MAX)
frags = reduce(fr.n, dist.ReduceOp.SUM)
g_ms = None
if world > 1:
    rec = torch.from_numpy(fr.rec.view(np.uint8).reshape(-1).copy()).to(dev)
    sharding.gather_variable(rec, 0)                    # NCCL warm-up
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(); parts = sharding.gather_variable(rec, 0); g1.record()
    torch.cuda.synchronize()
    g_ms = g0.elapsed_time(g1)
    if rank == 0:
        assert sum(p.numel() for p in parts) // fr.rec.dtype.itemsize == int(frags)
out["config4"] = {"events": n_ev, "events_per_gpu": len(mine), "fragments": int(frags), "batch_ms_max_over_ranks": ms,
                  "fragments_per_s": frags / (ms * 1e-3), "events_per_s": n_ev / (ms * 1e-3),
                  "record_gather_ms": g_ms, "record_gather_bytes": int(frags) * fr.rec.dtype.itemsize}
if rank == 0:
    print(out["config4"], flush=True)
ctx.close()
del fr, pieces, cells

# ---- config 5 ----
mine = sharding.event_shard(n_obj, world, rank)
cube = common.unit_cube()
levels = common.recursion_levels()
pieces, ev_p0 = common.concat([cube] * len(mine))
lvl_cells = [common.concat([c] * len(mine)) for c in levels]
ctx2 = FractureContext(local)


def recurse():
    ctx2.upload_pieces(pieces.verts, pieces.vert_off, pieces.ring_off, pieces.ring, ev_p0)
    tot, counts = 0.0, []
    for cl, ev_c in lvl_cells:
        ctx2.upload_cells(cl.planes, cl.plane_off, cl.verts, cl.vert_off, ev_c)
        ctx2.fracture_event()
        cc = ctx2.counts()
        tot += ctx2.last_event_ms()[0]
        counts.append(int(cc.n_fragments))
        rec = ctx2.download(geometry=False).rec
        ev_of_frag = np.searchsorted(ev_c, rec["cell"], side="right") - 1
        new_ev = np.concatenate([[0], np.cumsum(np.bincount(ev_of_frag, minlength=len(mine)))]).astype(np.uint32)
        ctx2.fragments_to_pieces(new_ev)
    return tot, counts


recurse()                                               # warm-up
barrier()
runs = [recurse() for _ in range(3)]
tot = float(np.median([r[0] for r in runs]))
counts = runs[0][1]
ms5 = reduce(tot, dist.ReduceOp.MAX)
final = reduce(counts[-1], dist.ReduceOp.SUM)
allf = reduce(sum(counts), dist.ReduceOp.SUM)
out["config5"] = {"objects": n_obj, "objects_per_gpu": len(mine), "per_object": [c // len(mine) for c in counts],
                  "final_fragments": int(final), "fragments_all_levels": int(allf), "sum_event_ms_max_over_ranks": ms5,
                  "final_fragments_per_s": final / (ms5 * 1e-3), "fragments_all_levels_per_s": allf / (ms5 * 1e-3),
                  "objects_per_s": n_obj / (ms5 * 1e-3)}
if rank == 0:
    print(out["config5"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/configs_multi_r1_n{world}.json", "w"), indent=1)
barrier()
if world > 1:
    dist.destroy_process_group()
