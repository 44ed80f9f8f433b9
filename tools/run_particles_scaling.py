"""run_particles.x on 1 GPU and on every visible GPU: the per-GPU point shards run concurrently (one host thread per device).

Writes a synthetic optimization.namelist problem in the reference's own file formats (ASCII tables, one namelist per
particle) into a scratch directory and runs semic_b200/bin/run_particles.x twice.  python tools/run_particles_scaling.py [nx] [particles] [nloop]
"""
import os, re, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench                                   # parameter values and the synthetic forcing generator of the product side
from semic_b200 import synth


def write_namelist(path, par_d):
    """one particle's &surface_physics group, the file format optimize/tools.py:7-43 writes"""
    with open(path, "w") as f:
        f.write('&surface_physics\n  boundary = "" "" "",\n')
        for k in sorted(par_d):
            v = par_d[k]
            f.write("  %s = %s,\n" % (k, '"%s"' % v if isinstance(v, str) else ("%d" % v if isinstance(v, int) else repr(float(v)))))
        f.write("/\n")

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nloop = int(sys.argv[3]) if len(sys.argv) > 3 else 20
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(root, "semic_b200", "bin", "run_particles.x")
d = tempfile.mkdtemp()
t0 = time.time()
forc = synth.forcing(nx, 365)
r = np.random.default_rng(1)
vali = r.normal(size=(365, 8, nx)) * np.array([5, .1, 30, 1e-8, 1e-8, 1e-8, 5, 5])[None, :, None] + \
    np.array([260, .7, 60, 0, 1e-8, 1e-8, 0, 0])[None, :, None]
np.savetxt(os.path.join(d, "forcing.txt"), forc.reshape(365, -1), fmt="%.7g")
np.savetxt(os.path.join(d, "vali.txt"), vali.reshape(365, -1), fmt="%.7g")
with open(os.path.join(d, "optimization.namelist"), "w") as f:
    f.write('&driver\n nloop = %d,\n ntime = 365,\n nx = %d,\n input_forcing = "forcing.txt",\n validation = "vali.txt",\n/\n' % (nloop, nx))
for i in range(P):
    pd = dict(bench.par_dict("none", 3), hcrit=0.02 + 0.4 * r.uniform(), amp=1.0 + 3 * r.uniform(), rcrit=r.uniform(),
              albi=0.35 + 0.2 * r.uniform(), albl=0.05 + 0.3 * r.uniform(), alb_smax=0.8 + 0.1 * r.uniform())
    write_namelist(os.path.join(d, "p_%06d.nml" % i), pd)
print("inputs written in %.1f s (%d points, %d particles, nloop %d)" % (time.time() - t0, nx, P, nloop), flush=True)
res = {}
for label, env in (("1 GPU", dict(os.environ, CUDA_VISIBLE_DEVICES="0")), ("all GPUs", dict(os.environ))):
    for rep in range(2):
        out = subprocess.run([exe, '"p_"', "0", str(P - 1)], cwd=d, env=env, stdout=subprocess.PIPE, text=True).stdout
        m = re.search(r"on (\d+) GPU\(s\): ([0-9.]+) s", out)
        print(label, "run", rep, m.group(0) if m else out[-300:], flush=True)
        if m:
            res[label] = (int(m.group(1)), float(m.group(2)))
    res[label + " costs"] = np.array([float(open(os.path.join(d, "p_%06d.out" % i)).read()) for i in range(0, P, 97)])
if "1 GPU" in res and "all GPUs" in res:
    a, b = res["1 GPU costs"], res["all GPUs costs"]
    print("speed-up %.2f on %d GPUs; max rel cost difference %.2e" % (res["1 GPU"][1] / res["all GPUs"][1], res["all GPUs"][0],
                                                                     float(np.nanmax(np.abs(a - b) / np.abs(a)))))
