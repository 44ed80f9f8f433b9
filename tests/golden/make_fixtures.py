"""Regenerates the binary fixtures under tests/golden/ from the reference's shipped ASCII data.

Run HERE (the build container, where /root/reference exists); the GPU box only sees the .npy files.

  <name>_input.npy   float64 [ntime][9][nx]  columns sf rf swd lwd wind sp rhoa qq t2m
                     (utils.f90:63-67: each variable is nx consecutive columns of a row)
  <name>_vali.npy    float64 [ntime][8][nx]  columns stt alb swnet smb melt acc shf lhf (utils.f90:89-93).
                     These are MAR regional-climate-model data (example/Readme.md:51-52), i.e. tuning
                     targets and boundary overrides -- NOT SEMIC output, hence not golden results.
"""
import os
import sys
import numpy as np

REF = os.environ.get("SEMIC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def convert(name):
    forc = np.loadtxt(os.path.join(REF, "example", "data", name + "_input.txt"))
    vali = np.loadtxt(os.path.join(REF, "example", "data", name + "_output.txt"))
    ntime = forc.shape[0]
    nx = forc.shape[1] // 9
    assert forc.shape[1] == 9 * nx and vali.shape == (ntime, 8 * nx)
    np.save(os.path.join(HERE, name + "_input.npy"), forc.reshape(ntime, 9, nx).astype(np.float64))
    np.save(os.path.join(HERE, name + "_vali.npy"), vali.reshape(ntime, 8, nx).astype(np.float64))
    print(name, "ntime", ntime, "nx", nx)


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["transect", "c01", "c20"]):
        convert(n)
