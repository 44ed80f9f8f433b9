"""Turns the ncu reports of tools/capture_profiles.sh (gpurun_out/r02_*.ncu-rep) into the tracked summaries under profiles/:
raw metric CSVs of the captured launch, source-page summaries, the launch list with shares, traffic.json, the SASS excerpt.
   python tools/summarise_profiles.py"""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
N, D = 16777216, 365

def ncu_csv(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout

KEEP = re.compile(r"^(gpu__time_duration|dram__bytes_(read|write)\.sum$|dram__throughput\.avg\.pct|sm__inst_executed_pipe_(fp64|alu|fma|xu|lsu|uniform|tma)\.avg\.pct_of_peak_sustained_active"
                  r"|smsp__issue_active\.avg\.pct|smsp__inst_executed\.sum$|launch__(registers_per_thread|block_size|grid_size|shared_mem_per_block_dynamic)"
                  r"|sm__warps_active\.avg\.pct|smsp__cycles_active\.avg$|sm__cycles_elapsed\.avg\.per_second|smsp__average_warp_latency_issue_stalled_.*\.pct|smsp__average_warps_issue_stalled_.*_per_issue_active"
                  r"|smsp__thread_inst_executed_per_inst_executed\.ratio|l1tex__data_bank_conflicts_pipe_lsu\.sum$|sm__throughput\.avg\.pct)")

def raw_summary(name):
    rep = os.path.join(G, name + ".ncu-rep")
    if not os.path.exists(rep):
        return None
    rows = list(csv.reader(io.StringIO(ncu_csv(rep, "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    out = os.path.join(P, name + "_raw.csv")
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name", "Block Size", "Grid Size", "Device", "CC") or KEEP.match(h):
                w.writerow([h, u, v])
    return dict(zip(hdr, vals))

def source_summary(name, n, d):
    rep = os.path.join(G, name + ".ncu-rep")
    src = os.path.join(G, name + "_source.csv")
    open(src, "w").write(ncu_csv(rep, "source"))
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src.py"), src, str(n), str(d)], capture_output=True, text=True).stdout
    open(os.path.join(P, name + "_source_summary.txt"), "w").write(txt)
    return txt

traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of the step kernel from `ncu --set full` captures (profiles/r02_k_run_lean_k*_raw.csv); "
                       "one launch = 16 777 216 points x 365 days, output M1 into a 32-day ring"}
for k in (3, 24):
    m = raw_summary("r02_k_run_lean_k%d" % k)
    if m:
        rd, wr = float(m["dram__bytes_read.sum"]), float(m["dram__bytes_write.sum"])
        scale = 1e9 if rd < 1e6 else 1.0          # ncu prints Gbyte for large values
        traffic["k%d" % k] = {"bytes_per_point_day": (rd + wr) * scale / (N * D), "dram_read_gb": rd * scale / 1e9, "dram_write_gb": wr * scale / 1e9,
                              "point_days": N * D, "source": "profiles/r02_k_run_lean_k%d_raw.csv" % k}
        print(source_summary("r02_k_run_lean_k%d" % k, N, D))
if len(traffic) > 1:
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
if raw_summary("r02_k_ensemble"):
    print(source_summary("r02_k_ensemble", 128 * 100000 // 32 * 32, 365))

# launch list: per-kernel totals and shares
ll = os.path.join(G, "r02_launches.csv")
if os.path.exists(ll):
    lines = [l for l in open(ll) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = re.sub(r"\(.*", "", r[ik])
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    with open(os.path.join(P, "r02_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 3 --warmup 3 --no-cpu --no-extra\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, v in tot.most_common():
            f.write("%s,%d,%.3f,%.4f\n" % (k, cnt[k], v, v / s))
    print(open(os.path.join(P, "r02_launches.csv")).read())

# SASS excerpt: the Blackwell-path mnemonics of the benchmarked kernels
so = os.path.join(ROOT, "semic_b200", "lib", "libsemic_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, out = None, []
want = re.compile(r"UBLKCP|SYNCS\.|ELECT|FENCE\.VIEW\.ASYNC|MUFU\.RCP64H|HMMA|UTC.*MMA")
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    if cur and ("k_run_leanILi1ELi24ELi3ELi3ELb0ELi0E" in cur or "k_ensembleILi0ELi3E" in cur) and want.search(line):
        out.append("%s  %s" % (cur[-48:], re.sub(r"\s+/\*[0-9a-f]{16}\*/", "", line.strip())))
seen, uniq = set(), []
for l in out:
    key = re.sub(r"/\*[0-9a-f]+\*/", "", l)
    key = re.sub(r"R\d+|UR\d+|0x[0-9a-f]+", "", key)
    if key not in seen:
        seen.add(key); uniq.append(l)
with open(os.path.join(P, "r02_sass_k_run_lean.txt"), "w") as f:
    f.write("# cuobjdump -sass semic_b200/lib/libsemic_b200.so: distinct TMA / mbarrier / elect / MUFU lines of k_run_lean<slater,24,3,3> and k_ensemble<none,3>\n"
            "# (1-D TMA bulk copy global->shared = UBLKCP; mbarrier expect-tx / try_wait = SYNCS.*; no HMMA/UTCMMA: nothing on this path is a contraction)\n")
    f.write("\n".join(uniq) + "\n")
print(len(uniq), "SASS lines kept")
