"""Run under torchrun by tests/test_multi_gpu.py (one rank per GPU): the ensemble misfit with points sharded over the ranks and
the per-particle sums all-reduced by the LIBRARY's NCCL communicator must equal the single-GPU evaluation of the whole grid."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
from semic_b200 import driver, engine, surface_physics as sp, synth
from semic_b200._capi import check, lib

engine.set_device(local)
L = lib()
NX, NT, NLOOP, P = 5000, 60, 2, 19
forc = synth.forcing(NX, NT, daystride=3)
mask = synth.mask(NX)
mask[mask == 0] = 1                                  # a constant validation series (ocean melt = 0) has no CRMSD
r = np.random.default_rng(0)
vali = np.ascontiguousarray(r.normal(size=(NT, 8, NX)) * np.array([5, .1, 30, 1e-8, 1e-8, 1e-8, 5, 5])[None, :, None]
                            + np.array([260, .7, 60, 0, 1e-8, 1e-8, 0, 0])[None, :, None])
names = ["hcrit", "alb_smax", "amp", "rcrit", "albi", "albl"]
lo_r, hi_r = np.array([0.01, 0.80, 0.5, 0.0, 0.35, 0.05]), np.array([0.5, 0.9, 5.0, 1.0, 0.55, 0.40])
pop = [{"id": i, "pos": list(lo_r + (hi_r - lo_r) * r.uniform(size=6))} for i in range(P)]
init = dict(hsnow=1.0, hice=0.0, alb=0.8, tsurf=260.0, alb_snow=0.8)

uid = C.create_string_buffer(128)
if rank == 0:
    check(L.semic_comm_unique_id(uid))
t = torch.tensor(list(uid.raw), dtype=torch.uint8, device="cuda")
dist.broadcast(t, 0)
uid = C.create_string_buffer(bytes(t.cpu().tolist()), 128)
comm = C.c_void_p()
check(L.semic_comm_init(C.byref(comm), uid, world, rank))
fused = bool(L.semic_comm_p2p(comm, None, 0))
lo, hi = driver.shard_bounds(NX, world, rank)
cost = driver.run_particles(pop, names, np.ascontiguousarray(forc[:, :, lo:hi]), np.ascontiguousarray(vali[:, :, lo:hi]),
                            mask[lo:hi], init, NT, NLOOP, comm=comm)
ms = C.c_float(-1.0)
check(L.semic_comm_last_allreduce_ms(comm, C.byref(ms)))
check(L.semic_comm_free(comm))
# every rank holds the same reduced costs
mine = torch.tensor([cost[p["id"]] for p in pop], dtype=torch.float64, device="cuda")
ref = mine.clone()
dist.broadcast(ref, 0)
same_everywhere = bool(torch.equal(mine, ref))
flags = torch.tensor([1 if same_everywhere else 0], device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    full = driver.run_particles(pop, names, forc, vali, mask, init, NT, NLOOP)
    err = max(abs(cost[k] - full[k]) / abs(full[k]) for k in full)
    print(json.dumps({"world": world, "max_rel_err": err, "same_on_all_ranks": int(flags[0]) == 1, "allreduce_ms": ms.value, "peer_memory": fused,
                      "finite": all(np.isfinite(v) for v in full.values())}))
dist.destroy_process_group()
