"""CPU tests of the search algorithms (semic_b200/optimize.py) on a toy objective, like the `__main__` self-tests of
the reference's optimization_algorithms/*.py (sphere function), plus the GPU test of the whole workflow."""
import numpy as np
import pytest

from semic_b200 import optimize as opt

SPACE = [(-5.0, 5.0)] * 3


def sphere(pop):
    return {c["id"]: float(np.sum((np.asarray(c["pos"]) - 1.0) ** 2)) for c in pop}


@pytest.mark.parametrize("alg", sorted(opt.SEARCHES))
def test_search_minimises_the_sphere(alg, tmp_path):
    rng = np.random.default_rng(1)
    log = opt.CostLog(str(tmp_path / "cost.txt"), ["a", "b", "c"])
    if alg in ("rs", "lhs"):
        best = opt.SEARCHES[alg](sphere, SPACE, 2000, rng=rng, log=log)
        bound = 0.5
    elif alg == "pso":
        best = opt.search_pso(sphere, SPACE, 60, 30, rng=rng, log=log)
        bound = 1e-3
    elif alg == "es":
        best = opt.search_es(sphere, SPACE, 80, 15, 30, rng=rng, log=log)
        bound = 1e-3
    elif alg == "hs":
        best = opt.search_hs(sphere, SPACE, 150, 20, rng=rng, log=log)
        bound = 5e-2
    else:
        best = opt.search_ca(sphere, SPACE, 80, 40, 8, rng=rng, log=log)
        bound = 5e-2
    log.close()
    assert best["cost"] < bound, (alg, best)
    assert all(lo <= p <= hi for p, (lo, hi) in zip(best["pos"], SPACE))
    rows = [l.split() for l in open(tmp_path / "cost.txt") if not l.startswith("#")]
    costs = [float(r[1]) for r in rows]
    assert len(rows) > 0 and all(b <= a + 1e-15 for a, b in zip(costs, costs[1:]))      # best-so-far never gets worse
    assert abs(costs[-1] - best["cost"]) <= 1e-5 * max(1.0, best["cost"])


def test_latin_hypercube_has_one_point_per_stratum():
    rng = np.random.default_rng(2)
    n = 50
    x = opt.latin_hypercube([(0.0, 1.0), (10.0, 20.0)], n, rng)
    for j, (lo, hi) in enumerate([(0.0, 1.0), (10.0, 20.0)]):
        strata = np.floor((x[:, j] - lo) / (hi - lo) * n).astype(int)
        assert sorted(strata) == list(range(n))


def test_read_ranges(tmp_path):
    p = tmp_path / "namelist.ranges"
    p.write_text("hcrit,0,0.5\nalb_smax,0.80,0.9\namp,0,5\n")                # optimize/namelist.ranges
    names, space = opt.read_ranges(str(p))
    assert names == ["hcrit", "alb_smax", "amp"] and space == [(0.0, 0.5), (0.8, 0.9), (0.0, 5.0)]


@pytest.mark.gpu
def test_pso_on_the_transect_lowers_the_misfit(tmp_path):
    """the reference's workflow (optimization.py pso) end to end on the GPU: files in, cost log out"""
    from tests.test_host_api import _example_dir
    d = _example_dir(tmp_path, nloop=2)
    (d / "optimization.namelist").write_text((d / "example.namelist").read_text())
    (d / "namelist.ranges").write_text("hcrit,0,0.5\nalb_smax,0.80,0.9\namp,0,5\nrcrit,0,1\nalbi,0.35,0.55\nalbl,0.05,0.40\n")
    names, space = opt.read_ranges(str(d / "namelist.ranges"))
    cost_fn = opt.semic_cost_fn(str(d / "optimization.namelist"), names)
    rng = np.random.default_rng(3)
    first = opt.search_lhs(cost_fn, space, 16, rng=rng)
    log = opt.CostLog(str(d / "pso_cost.txt"), names)
    best = opt.search_pso(cost_fn, space, 6, 16, rng=np.random.default_rng(3), log=log)
    log.close()
    assert np.isfinite(best["cost"]) and best["cost"] > 0
    rows = [l.split() for l in open(d / "pso_cost.txt") if not l.startswith("#")]
    assert len(rows) == 6 and float(rows[-1][1]) <= float(rows[0][1])
    assert best["cost"] <= 1.2 * first["cost"]
