"""Config 5: parameter ensemble (1024 parameter sets x 100k points x 365 days, n_ksub=3, alb_scheme "none") with ONE
NCCL all-reduce of the per-particle misfit sums.  Points are sharded over the ranks.
   python tools/bench_ensemble.py [--particles 1024] [--points 100000] [--days 365]
   torchrun --nproc-per-node N tools/bench_ensemble.py ..."""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=1024)
ap.add_argument("--points", type=int, default=100000, help="total grid points (sharded over ranks)")
ap.add_argument("--days", type=int, default=365)
ap.add_argument("--nloop", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
from semic_b200 import driver, engine, surface_physics as sp, synth
from semic_b200._capi import FIELD_ID, check, lib
engine.set_device(local)
L = lib()
lo, hi = driver.shard_bounds(args.points, world, rank)
n = hi - lo
# the library's own communicator: rank 0 makes the id, torch.distributed carries the 128 bytes
comm = None
if world > 1:
    uid = C.create_string_buffer(128)
    if rank == 0:
        check(L.semic_comm_unique_id(uid))
    box = [bytes(uid.raw)]
    dist.broadcast_object_list(box, src=0)
    comm = C.c_void_p()
    check(L.semic_comm_init(C.byref(comm), box[0], world, rank))

st = sp.Surface_State_Class(); sp.surface_alloc(st, n)
check(L.semic_state_synth_mask(st.handle, 7, lo, 0.85, 0.10))
for name, v in (("tsurf", 260.0), ("hsnow", 1.0), ("alb", 0.8), ("alb_snow", 0.8)):
    check(L.semic_state_fill(st.handle, FIELD_ID[name], v))
forcing = engine.Series(n, 9, args.days).synth_forcing(20260101, point0=lo)
# synthetic "validation": a reference trajectory of the default parameters + noise would do; the cost value is not the point here
r = np.random.default_rng(lo)
vali_h = np.ascontiguousarray(r.normal(size=(args.days, 8, n)) * np.array([5, .1, 30, 1e-8, 1e-8, 1e-8, 5, 5])[None, :, None]
                              + np.array([260, .7, 60, 0, 1e-8, 1e-8, 0, 0])[None, :, None])
vali = engine.Series(n, 8, args.days).upload(vali_h)
# latin-hypercube sample of the six tuned parameters (optimize/namelist.ranges), seeded
rng = np.random.default_rng(20260101)
names = ["hcrit", "alb_smax", "amp", "rcrit", "albi", "albl"]
lo_r = np.array([0.0, 0.80, 0.0, 0.0, 0.35, 0.05]); hi_r = np.array([0.5, 0.9, 5.0, 1.0, 0.55, 0.40])
P = args.particles
u = (np.stack([rng.permutation(P) for _ in range(6)], axis=1) + rng.uniform(size=(P, 6))) / P
pop = [{"id": i, "pos": list(lo_r + (hi_r - lo_r) * u[i])} for i in range(P)]
pars = driver.particle_params(names, pop)
bnd = sp.Boundary_Opt_Class()
best = None
for rep in range(args.reps):
    t0 = time.perf_counter()
    sumsq, ms = engine.ensemble_cost(st, pars, bnd, forcing, vali, args.days, args.nloop, comm=comm)
    wall = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
    best = ms if best is None else min(best, ms)
cost = np.sqrt(sumsq)
if rank == 0:
    td = float(P) * args.points * args.days * args.nloop
    print(json.dumps({"metric": "trajectory-days/s (ensemble, config 5)", "value": td / (best * 1e-3), "n_gpus": world,
                      "particles": P, "points": args.points, "days": args.days, "kernel_ms": best, "wall_s_last": wall,
                      "cost_min": float(np.nanmin(cost)), "cost_max": float(np.nanmax(cost)),
                      "collective": "ncclAllReduce(sum, %d doubles)" % P if world > 1 else "none (1 rank)"}))
if dist is not None:
    dist.destroy_process_group()
