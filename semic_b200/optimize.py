"""Parameter search over the ensemble misfit -- the workflow of optimize/optimization.py, in process.

The reference drives `./run_particles.x` through files and `mpiexec` once per generation (optimize/tools.py:46-63);
here a generation is ONE call of `cost_fn(pop) -> {id: cost}` (normally `driver.run_particles`, i.e. one
`semic_ensemble_cost` launch over all candidates), so the search logic is all that is left on the host.  The six
searches keep the reference's command names (optimization.py:266-318): rs, lhs, pso, es, ca, hs.  They are written
from the textbook definitions of the methods, for Python 3; candidates are dicts {"id", "pos", "cost"} like the
reference's, the search space is a list of (lo, hi) per parameter (optimize/namelist.ranges), and the best-so-far
line per generation goes to `<alg>_cost.txt` in the reference's layout (tools.py:82-89: "# gen fitness <names>").

    python -m semic_b200.optimize pso --ranges namelist.ranges --namelist optimization.namelist --gens 20 --pop 64
"""
import argparse
import os

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# plumbing shared by the searches
# ---------------------------------------------------------------------------------------------------------------
def read_ranges(path):
    """namelist.ranges: one `name,low,high` per line (tools.py:92-108).  Returns (names, [(lo, hi), ...])."""
    names, space = [], []
    for line in open(path):
        parts = [p.strip() for p in line.split(",")]
        if len(parts) >= 3 and parts[0]:
            names.append(parts[0])
            space.append((float(parts[1]), float(parts[2])))
    return names, space


class CostLog(object):
    """best-so-far per generation, `<prefix>cost.txt` (tools.py:82-89; read by optimize/plot_costs.py)"""

    def __init__(self, path, names):
        self.f = open(path, "w") if path else None
        if self.f:
            self.f.write("# gen fitness " + " ".join(names) + " \n")

    def write(self, gen, best):
        if self.f:
            self.f.write("%i %.6g " % (gen, best["cost"]) + " ".join("%.6g" % p for p in best["pos"]) + " \n")
            self.f.flush()

    def close(self):
        if self.f:
            self.f.close()


def _bounds(space):
    a = np.asarray(space, dtype=np.float64)
    return a[:, 0], a[:, 1]


def _evaluate(cost_fn, cands):
    """one generation = one batch evaluation; ids are made unique within the batch"""
    for i, c in enumerate(cands):
        c["id"] = i
    cost = cost_fn(cands)
    for c in cands:
        c["cost"] = float(cost[c["id"]])
    return cands


def _best(cands, best=None):
    for c in cands:
        if best is None or c["cost"] < best["cost"]:
            best = {"pos": list(c["pos"]), "cost": c["cost"]}
    return best


# ---------------------------------------------------------------------------------------------------------------
# the six searches
# ---------------------------------------------------------------------------------------------------------------
def search_rs(cost_fn, space, max_iter, rng=None, log=None):
    """random search: max_iter independent uniform samples, one batch (optimization.py:13-33)"""
    rng = rng or np.random.default_rng()
    lo, hi = _bounds(space)
    pop = _evaluate(cost_fn, [{"pos": list(lo + (hi - lo) * rng.uniform(size=len(lo)))} for _ in range(max_iter)])
    best = None
    for i, c in enumerate(pop):
        best = _best([c], best)
        if log:
            log.write(i, best)
    return best


def latin_hypercube(space, n, rng):
    """n points, exactly one per stratum of width (hi-lo)/n in every dimension"""
    lo, hi = _bounds(space)
    u = (np.stack([rng.permutation(n) for _ in range(len(lo))], axis=1) + rng.uniform(size=(n, len(lo)))) / n
    return lo + (hi - lo) * u


def search_lhs(cost_fn, space, max_iter, rng=None, log=None):
    """latin hypercube sample evaluated in one batch (optimization.py:35-56)"""
    rng = rng or np.random.default_rng()
    pop = _evaluate(cost_fn, [{"pos": list(x)} for x in latin_hypercube(space, max_iter, rng)])
    best = None
    for i, c in enumerate(pop):
        best = _best([c], best)
        if log:
            log.write(i, best)
    return best


def search_pso(cost_fn, space, max_gens, pop_size, c1=2.0, c2=2.0, rng=None, log=None, inertia=0.72, vel_frac=1.0):
    """global-best particle swarm: v <- w v + c1 r1 (pbest - x) + c2 r2 (gbest - x), velocities and positions clipped
    to the box (optimization.py:58-90 drives the same loop through files)"""
    rng = rng or np.random.default_rng()
    lo, hi = _bounds(space)
    vmax = vel_frac * (hi - lo)
    x = lo + (hi - lo) * rng.uniform(size=(pop_size, len(lo)))
    v = -vmax + 2 * vmax * rng.uniform(size=x.shape)
    pbest, pcost = x.copy(), np.full(pop_size, np.inf)
    best = None
    for gen in range(max_gens):
        pop = _evaluate(cost_fn, [{"pos": list(xi)} for xi in x])
        cost = np.array([c["cost"] for c in pop])
        better = cost < pcost
        pbest[better], pcost[better] = x[better], cost[better]
        best = _best(pop, best)
        if log:
            log.write(gen, best)
        g = np.asarray(best["pos"])
        r1, r2 = rng.uniform(size=x.shape), rng.uniform(size=x.shape)
        v = np.clip(inertia * v + c1 * r1 * (pbest - x) + c2 * r2 * (g - x), -vmax, vmax)
        x = np.clip(x + v, lo, hi)
    return best


def search_es(cost_fn, space, max_gens, pop_size, num_children, rng=None, log=None):
    """(mu + lambda) evolution strategy with self-adapted per-coordinate step sizes (log-normal rule)"""
    rng = rng or np.random.default_rng()
    lo, hi = _bounds(space)
    n = len(lo)
    tau, tau_p = 1.0 / np.sqrt(2.0 * np.sqrt(n)), 1.0 / np.sqrt(2.0 * n)
    pop = [{"pos": list(lo + (hi - lo) * rng.uniform(size=n)), "sigma": list(0.05 * (hi - lo))} for _ in range(pop_size)]
    pop = _evaluate(cost_fn, pop)
    best = _best(pop)
    for gen in range(max_gens):
        kids = []
        for k in range(num_children):
            p = pop[k % pop_size]
            s = np.asarray(p["sigma"]) * np.exp(tau_p * rng.normal() + tau * rng.normal(size=n))
            s = np.minimum(s, hi - lo)
            kids.append({"pos": list(np.clip(np.asarray(p["pos"]) + s * rng.normal(size=n), lo, hi)), "sigma": list(s)})
        kids = _evaluate(cost_fn, kids)
        pop = sorted(pop + kids, key=lambda c: c["cost"])[:pop_size]
        best = _best(pop, best)
        if log:
            log.write(gen, best)
    return best


def search_hs(cost_fn, space, max_gens, mem_size, consid_rate=0.95, adjust_rate=0.7, bandwidth=0.05, batch=None,
              rng=None, log=None):
    """harmony search: every new harmony takes each coordinate from memory (rate consid_rate, pitch-adjusted with rate
    adjust_rate by +-bandwidth of the range) or uniformly at random; it replaces the worst harmony if better.
    `batch` harmonies are improvised and evaluated together per generation (one ensemble launch)."""
    rng = rng or np.random.default_rng()
    lo, hi = _bounds(space)
    n = len(lo)
    batch = batch or mem_size
    mem = _evaluate(cost_fn, [{"pos": list(lo + (hi - lo) * rng.uniform(size=n))} for _ in range(mem_size)])
    best = _best(mem)
    for gen in range(max_gens):
        new = []
        for _ in range(batch):
            pos = np.empty(n)
            for j in range(n):
                if rng.uniform() < consid_rate:
                    pos[j] = mem[rng.integers(len(mem))]["pos"][j]
                    if rng.uniform() < adjust_rate:
                        pos[j] += bandwidth * (hi[j] - lo[j]) * rng.uniform(-1.0, 1.0)
                else:
                    pos[j] = lo[j] + (hi[j] - lo[j]) * rng.uniform()
            new.append({"pos": list(np.clip(pos, lo, hi))})
        new = _evaluate(cost_fn, new)
        mem = sorted(mem + new, key=lambda c: c["cost"])[:mem_size]
        best = _best(mem, best)
        if log:
            log.write(gen, best)
    return best


def search_ca(cost_fn, space, max_gens, pop_size, num_accepted, rng=None, log=None):
    """cultural algorithm: a belief space (situational = best individual, normative = bounding box of the accepted
    individuals) steers the variation of the population; binary tournaments select the next population"""
    rng = rng or np.random.default_rng()
    lo, hi = _bounds(space)
    n = len(lo)
    pop = _evaluate(cost_fn, [{"pos": list(lo + (hi - lo) * rng.uniform(size=n))} for _ in range(pop_size)])
    best = _best(pop)
    nlo, nhi = lo.copy(), hi.copy()                                      # normative knowledge
    for gen in range(max_gens):
        kids = []
        g = np.asarray(best["pos"])
        for k, p in enumerate(pop):
            x = np.asarray(p["pos"])
            step = (nhi - nlo) * rng.normal(scale=0.3, size=n)
            if k % 2:                                                    # situational influence: step towards the best
                step = np.abs(step) * np.sign(g - x)
            kids.append({"pos": list(np.clip(x + step, lo, hi))})
        kids = _evaluate(cost_fn, kids)
        both = pop + kids
        pop = []
        for _ in range(pop_size):                                        # binary tournaments
            a, b = both[rng.integers(len(both))], both[rng.integers(len(both))]
            pop.append(dict(a if a["cost"] <= b["cost"] else b))
        best = _best(pop, best)                                          # situational knowledge
        acc = np.array([c["pos"] for c in sorted(pop, key=lambda c: c["cost"])[:num_accepted]])
        nlo, nhi = acc.min(axis=0), np.maximum(acc.max(axis=0), acc.min(axis=0) + 1e-12 * (hi - lo))
        if log:
            log.write(gen, best)
    return best


SEARCHES = {"rs": search_rs, "lhs": search_lhs, "pso": search_pso, "es": search_es, "ca": search_ca, "hs": search_hs}


# ---------------------------------------------------------------------------------------------------------------
# command line (optimization.py:266-318)
# ---------------------------------------------------------------------------------------------------------------
def semic_cost_fn(nml_file, names, workdir=None):
    """cost function of the SEMIC parameter search: the &driver group of optimization.namelist names the forcing and
    validation tables; every call evaluates the whole population with one ensemble launch on the GPU."""
    from . import driver
    workdir = workdir or os.path.dirname(os.path.abspath(nml_file))
    drv = driver.read_driver_namelist(nml_file)
    nx, ntime, nloop = drv["nx"], drv["ntime"], drv["nloop"]
    rel = lambda p: p if os.path.isabs(p) else os.path.join(workdir, p)
    forc = driver.read_forcing(rel(drv["input_forcing"]), ntime, nx)
    vali = driver.read_validation(rel(drv["validation"]), ntime, nx)
    mask = drv["loi_mask"][:nx]
    init = dict(hsnow=1.0, hice=0.0, alb=vali[0, 1, :], tsurf=vali[0, 0, :], alb_snow=vali[0, 1, :])   # run_particles.f90:77-82
    return lambda pop: driver.run_particles(pop, names, forc, vali, mask, init, ntime, nloop)


def main(argv=None):
    ap = argparse.ArgumentParser(description="SEMIC parameter search on the GPU ensemble misfit")
    ap.add_argument("alg", choices=sorted(SEARCHES))
    ap.add_argument("--ranges", default="namelist.ranges")
    ap.add_argument("--namelist", default="optimization.namelist")
    ap.add_argument("--gens", type=int, default=20)
    ap.add_argument("--pop", type=int, default=64, help="population / sample / memory size")
    ap.add_argument("--children", type=int, default=None, help="es: children per generation (default 2 x pop)")
    ap.add_argument("--accepted", type=int, default=None, help="ca: accepted individuals (default pop / 5)")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--prefix", default=None, help="cost log is <prefix>cost.txt (default <alg>_)")
    a = ap.parse_args(argv)
    names, space = read_ranges(a.ranges)
    cost_fn = semic_cost_fn(a.namelist, names)
    rng = np.random.default_rng(a.seed)
    log = CostLog((a.prefix if a.prefix is not None else a.alg + "_") + "cost.txt", names)
    if a.alg in ("rs", "lhs"):
        best = SEARCHES[a.alg](cost_fn, space, a.pop * a.gens, rng=rng, log=log)
    elif a.alg == "pso":
        best = search_pso(cost_fn, space, a.gens, a.pop, rng=rng, log=log)
    elif a.alg == "es":
        best = search_es(cost_fn, space, a.gens, a.pop, a.children or 2 * a.pop, rng=rng, log=log)
    elif a.alg == "hs":
        best = search_hs(cost_fn, space, a.gens, a.pop, rng=rng, log=log)
    else:
        best = search_ca(cost_fn, space, a.gens, a.pop, a.accepted or max(1, a.pop // 5), rng=rng, log=log)
    log.close()
    print("best cost %.6g" % best["cost"])
    for n, p in zip(names, best["pos"]):
        print("  %s = %.8f" % (n, p))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
