"""Multi-GPU tests (skipped on a one-GPU box): the one collective of the path, with the library's own NCCL communicator."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("p2p", [1, 0], ids=["peer_memory", "nccl"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_ensemble_allreduce_over_ranks(nranks, p2p):
    """optimize/run_particles.f90:103-107,165-177 + tools.py:58-62: points sharded over N ranks, ncclAllReduce(sum) of the
    per-particle sums == the single-GPU evaluation of the whole grid (summation order differs: 1e-12), through both exchanges
    of semic_comm: the fused reduction kernel over NVLink peer memory and ncclAllReduce (SEMIC_COMM_P2P=0)"""
    from semic_b200 import engine
    if engine.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_rank_ensemble.py")]
    env = dict(os.environ, SEMIC_COMM_P2P=str(p2p))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print("\n[multi-rank ensemble]", d)
    assert d["world"] == nranks and d["finite"] and d["same_on_all_ranks"]
    assert d["peer_memory"] == bool(p2p)            # the fused peer-memory kernel really ran (or really did not)
    assert d["max_rel_err"] <= 1e-12
    assert 0.0 <= d["allreduce_ms"] < 5000.0          # the first collective of a communicator includes NCCL's lazy set-up
