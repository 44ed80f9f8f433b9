"""world_size-2 `gloo` tests (CPU) of the multi-rank host logic: contiguous point shards, no data-path
communication for the grid step, and ONE all-reduce of per-particle sums for the ensemble cost."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_partials(pars, forcing, vali, mask, init, ntime, nloop):
    """per-particle sum of squared CRMSDs of the local shard, computed by the CPU oracle"""
    from oracle import binding as ob
    nx = forcing.shape[2]
    res = []
    for par in pars:
        d = {k: getattr(par, k) for k in ob.PAR_DOUBLES if k != "tsticsub"}
        d["n_ksub"] = par.n_ksub
        d["alb_scheme"] = par.alb_scheme
        S, m = ob.new_state(nx)
        m[:] = mask
        for k, v in init.items():
            S[ob.FIDX[k]] = v
        out, _ = ob.run(S, m, ob.make_par(d), ob.make_bnd(), forcing, ntime * nloop, out_days=ntime)
        res.append(ob.particle_cost(out, vali, sumsq=True) if nx else 0.0)
    return np.array(res)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from semic_b200 import driver, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ntime, nloop = 37, 60, 2
    forc = synth.forcing(nx, ntime)
    r = np.random.default_rng(0)
    vali = np.ascontiguousarray(r.normal(size=(ntime, 8, nx)) * np.array([5, .1, 30, 1e-8, 1e-8, 1e-8, 5, 5])[None, :, None]
                                + np.array([260, .7, 60, 0, 1e-8, 1e-8, 0, 0])[None, :, None])
    mask = synth.mask(nx)
    pop = [{"id": i, "pos": [0.05 + 0.1 * i, 0.8 + 0.02 * i, 1.0 + i, 0.1 * i, 0.35 + 0.04 * i, 0.05 + 0.07 * i]} for i in range(5)]
    names = ["hcrit", "alb_smax", "amp", "rcrit", "albi", "albl"]                 # optimize/namelist.ranges
    init = dict(hsnow=1.0, hice=0.0, alb=0.8, tsurf=260.0, alb_snow=0.8)
    lo, hi = driver.shard_bounds(nx, world, rank)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    cost = driver.run_particles(pop, names, np.ascontiguousarray(forc[:, :, lo:hi]), np.ascontiguousarray(vali[:, :, lo:hi]),
                                mask[lo:hi], init, ntime, nloop, allreduce=allreduce, partial_fn=_oracle_partials)
    if rank == 0:
        full = driver.run_particles(pop, names, forc, vali, mask, init, ntime, nloop, partial_fn=_oracle_partials)
        q.put((cost, full))
    dist.barrier()
    dist.destroy_process_group()


def test_ensemble_cost_sharded_over_two_ranks_equals_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    cost, full = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert set(cost) == set(full) == set(range(5))
    for k in cost:
        assert np.isfinite(cost[k]) and abs(cost[k] - full[k]) <= 1e-12 * abs(full[k])


def test_finish_costs_nan_maps_to_max():
    from semic_b200 import driver
    pop = [{"id": 3, "pos": []}, {"id": 9, "pos": []}]
    c = driver.finish_costs(pop, np.array([4.0, np.nan]))
    assert c[3] == 2.0 and c[9] == np.finfo("d").max            # optimize/tools.py:62


def test_reference_arm_under_torchrun_prints_one_json_line():
    """`bench.py --impl reference` launched like our own arm (torchrun, 2 ranks): rank 0 alone runs the CPU reference and
    prints exactly one JSON line; the other rank exits 0 without work"""
    import json
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--cpu-points-per-core", "256", "--cpu-days", "5"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grid-point-days/s" and d["value"] > 0
    from oracle import ref_binding as rb
    assert d["cpu_baseline"]["kind"] == ("reference" if rb.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["world_size"] == 2            # same config as our arm at N = 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
