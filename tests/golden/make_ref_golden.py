"""Golden vectors made by THE REFERENCE ITSELF, run here.

The reference's own programs (example/example.f90, optimize/run_particles.f90) are translated to C by
oracle/f90_to_c.py from the sources under /root/reference, compiled (`make -C oracle ref`) and run on the
reference's own shipped data files; what they write is stored here as small .npy fixtures so that the GPU box
(which has no /root/reference) can check the CUDA path against reference output:

  ref_example_out.npy        example.out of `make run_example` as shipped (example.namelist, nloop = 100): [365][8][7]
                             columns tsurf alb swnet smb melt acc shf lhf (example.f90:117-122)
  ref_example_nloop1.npy     the same with nloop = 1 (cold start, single-precision 0.8 initial albedo)
  ref_example_<scheme>.npy   example.f90 on the transect with the parameters of tests/util.SCHEME_PARS[scheme],
                             nloop = 2, for slater / isba / denby / alex / none and n_ksub = 3; slater24 = n_ksub 24
  ref_particles_cost.npy     optimize/run_particles.f90 on optimize/optimization.namelist (nloop = 20) with three
                             particles (parameters in ref_particles_pars.json): the three costs it writes;
                             ref_particles_cost_ice.npy: the same with loi_mask = 2 (ice) at every point

Run from the repo root in the build container:  python tests/golden/make_ref_golden.py
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_binding as rb      # noqa: E402
from tests import util                    # noqa: E402

REF = rb.REFERENCE


def example_tree(tmp, namelist_text):
    d = os.path.join(tmp, "example")
    os.makedirs(d)
    os.symlink(os.path.join(REF, "example", "data"), os.path.join(d, "data"))
    with open(os.path.join(d, "example.namelist"), "w") as f:
        f.write(namelist_text)
    return d


def run_example(namelist_text):
    tmp = tempfile.mkdtemp()
    try:
        d = example_tree(tmp, namelist_text)
        rb.run_program("example", ["example.namelist"], cwd=d)
        return np.loadtxt(os.path.join(d, "example.out")).reshape(365, 8, 7)
    finally:
        shutil.rmtree(tmp)


def main():
    assert rb.have_reference_tree(), "needs /root/reference"
    rb.build(force=True)
    shipped = open(os.path.join(REF, "example", "example.namelist")).read()
    np.save(os.path.join(HERE, "ref_example_out.npy"), run_example(shipped))
    np.save(os.path.join(HERE, "ref_example_nloop1.npy"), run_example(shipped.replace("nloop = 100", "nloop = 1")))
    driver = shipped[:shipped.index("&surface_physics")].replace("nloop = 100", "nloop = 2")
    for scheme, k in [("slater", 3), ("isba", 3), ("denby", 3), ("alex", 3), ("none", 3), ("slater", 24)]:
        pd = dict(util.SCHEME_PARS[scheme], n_ksub=k)
        tmpf = tempfile.mktemp()
        rb.write_namelist(tmpf, pd)
        text = driver + open(tmpf).read()
        os.unlink(tmpf)
        np.save(os.path.join(HERE, "ref_example_%s%s.npy" % (scheme, "24" if k == 24 else "")), run_example(text))
    # run_particles on the shipped optimization.namelist, and on the same with an all-ice mask (land points near complete
    # snow melt are chaotic -- a one-ulp ghost snow layer keeps the :201 gate on -- so only the ice variant can be matched
    # to 1e-9 by an implementation that is not bit-identical)
    pars = [dict(util.PAR_SLATER, alb_scheme="none", hcrit=0.05 + 0.02 * i, amp=2.0 + 0.7 * i, rcrit=0.3 + 0.2 * i,
                 albi=0.4 + 0.05 * i, albl=0.15 + 0.05 * i, alb_smax=0.8 + 0.02 * i) for i in range(3)]
    with open(os.path.join(HERE, "ref_particles_pars.json"), "w") as f:
        json.dump(pars, f, indent=1, sort_keys=True)
    nml = open(os.path.join(REF, "optimize", "optimization.namelist")).read()
    for label, text in (("", nml), ("_ice", nml.replace("loi_mask(1:3) = 1,1,1,", "loi_mask(1:3) = 2,2,2,"))):
        assert label == "" or text != nml
        tmp = tempfile.mkdtemp()
        try:
            example_tree(tmp, shipped)
            opt = os.path.join(tmp, "optimize")
            os.makedirs(opt)
            with open(os.path.join(opt, "optimization.namelist"), "w") as f:
                f.write(text)
            for i, pd in enumerate(pars):
                rb.write_namelist(os.path.join(opt, "p_%06d.nml" % i), pd)
            rb.run_program("run_particles", ['"p_"', "0", "2"], cwd=opt)
            costs = np.array([float(np.loadtxt(os.path.join(opt, "p_%06d.out" % i))) for i in range(3)])
            np.save(os.path.join(HERE, "ref_particles_cost%s.npy" % label), costs)
            print("costs%s" % label, costs)
        finally:
            shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
