"""Host <-> device copy bandwidth of the box, per GPU and with 1/2/4/8 GPUs copying at the same time.

  python tools/pcie_probe.py                 # GPU 0: H2D alone, D2H alone, both directions at once
  python tools/pcie_probe.py --all-gpus      # the same with 1, 2, 4, 8 GPUs busy concurrently (one process per GPU,
                                             # pinned buffers first-touched on the GPU's own NUMA node), one JSON line per N

The end-to-end (host buffers) leg of bench.py is bound by these numbers: 72 B in + 64 B out per point-day.  torch is
used for plumbing only (pinned memory, streams).
"""
import argparse
import json
import os
import subprocess
import sys
import time


def child(rank, start_at, seconds, mib):
    import torch
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(rank))
    except Exception:
        pass
    torch.cuda.set_device(rank)
    n = mib << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h1.zero_(); h2.zero_()
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    def both():
        h2d(); d2h()

    res = {}
    for i, (name, f) in enumerate((("h2d", h2d), ("d2h", d2h), ("both_each_dir", both))):
        f(); torch.cuda.synchronize()
        t_phase = start_at + i * (seconds + 1.5)
        while time.time() < t_phase:
            time.sleep(0.001)                      # all ranks enter the phase together (wall-clock rendezvous)
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            f(); torch.cuda.synchronize(); reps += 1
        res[name] = n * reps / (time.perf_counter() - t0) / 1e9
    print(json.dumps({"rank": rank, **res}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--all-gpus", action="store_true")
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--child", type=int, default=-1)
    ap.add_argument("--start-at", type=float, default=0.0)
    a = ap.parse_args()
    if a.child >= 0:
        return child(a.child, a.start_at, a.seconds, a.mib)
    import torch
    ndev = torch.cuda.device_count()
    counts = [n for n in (1, 2, 4, 8) if n <= ndev] if a.all_gpus else [1]
    for n in counts:
        start = time.time() + 25.0                 # time to import torch and pin the buffers
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", str(r), "--start-at", str(start),
                                   "--seconds", str(a.seconds), "--mib", str(a.mib)], stdout=subprocess.PIPE, text=True)
                 for r in range(n)]
        rows = []
        for p in procs:
            out, _ = p.communicate()
            rows += [json.loads(l) for l in out.splitlines() if l.startswith("{")]
        agg = {k: sum(r[k] for r in rows) for k in ("h2d", "d2h", "both_each_dir")}
        print(json.dumps({"gpus_busy": n, "GBps_sum_over_gpus": agg, "GBps_per_gpu": sorted(rows, key=lambda r: r["rank"]),
                          "note": "both_each_dir: H2D and D2H at once, figure per direction; 72 B in + 64 B out per point-day "
                                  "-> e2e ceiling = both_each_dir / 72e-9 point-days/s"}), flush=True)
    if not a.all_gpus:
        os.system("nvidia-smi topo -m | head -12; free -g | head -2; lscpu | grep -E 'Model name|Socket|NUMA node'")


if __name__ == "__main__":
    main()
