"""Binary columnar forcing / output files and the host-streamed run over them.

The reference reads whitespace-separated ASCII with list-directed reads (utils.f90:44-95: one record of 9*nx or 8*nx
reals per day) -- fine for 7 points, not for millions.  The binary twin keeps the same logical layout
[day][var][point] as raw little-endian float64 behind a 64-byte header, so a day row of the ASCII file and a day row of
the binary file hold the same numbers in the same order.  `run_files` streams such a file through `semic_run_host`
window by window from pinned staging buffers: disk -> pinned host -> (H2D || kernel || D2H) -> pinned host -> disk.
"""
import os
import struct

import numpy as np

from . import engine
from ._capi import NFORCING, NOUT

MAGIC = b"SEMICB01"
HEADER = 64


def write_series(path, array):
    """array [ndays][nvar][npts] float64 -> binary series file"""
    a = np.ascontiguousarray(array, dtype="<f8")
    nd, nv, npts = a.shape
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<qqq", nd, nv, npts) + b"\0" * (HEADER - 8 - 24))
        a.tofile(f)


def open_series(path, mode="r"):
    """memory map of a binary series file: array [ndays][nvar][npts]"""
    with open(path, "rb") as f:
        h = f.read(HEADER)
    if len(h) < HEADER or h[:8] != MAGIC:
        raise ValueError("%s is not a SEMIC binary series" % path)
    nd, nv, npts = struct.unpack("<qqq", h[8:32])
    return np.memmap(path, dtype="<f8", mode=mode, offset=HEADER, shape=(nd, nv, npts))


def create_series(path, ndays, nvar, npts):
    """preallocates a binary series file and returns its writable memory map"""
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<qqq", ndays, nvar, npts) + b"\0" * (HEADER - 8 - 24))
        f.truncate(HEADER + 8 * ndays * nvar * npts)
    return open_series(path, mode="r+")


def ascii_to_binary(ascii_path, bin_path, ntime, nx, nvar=NFORCING):
    """one of the reference's tables (utils.f90:63-67 / :89-93) -> binary series"""
    a = np.loadtxt(ascii_path, ndmin=2)
    if a.shape[0] < ntime or a.shape[1] != nvar * nx:
        raise ValueError("%s: expected >= %d lines of %d columns, found %s" % (ascii_path, ntime, nvar * nx, a.shape))
    write_series(bin_path, a[:ntime].reshape(ntime, nvar, nx))


def run_files(state, par, bnd, forcing_path, ndays, out_path=None, out_days=0, window_days=16, chunk_days=1):
    """Runs `ndays` days on `state` with forcing streamed from a binary series file (day d uses file row d % ndays(file)),
    writing the last `out_days` days of the 8 daily fields to a binary series at `out_path`.  Host memory in use: two
    pinned windows of `window_days` days (forcing, output), whatever the length of the run.  Returns total GPU ms."""
    forc = open_series(forcing_path)
    nfile, nv, npts = forc.shape
    if nv != NFORCING:
        raise ValueError("forcing series must have %d variables" % NFORCING)
    out = create_series(out_path, out_days, NOUT, npts) if (out_path and out_days > 0) else None
    w = max(1, min(window_days, ndays))
    hf = engine.PinnedArray((w, NFORCING, npts))
    ho = engine.PinnedArray((w, NOUT, npts)) if out is not None else None
    first_out = ndays - out_days
    total, d0 = 0.0, 0
    try:
        while d0 < ndays:
            n = min(w, ndays - d0)
            for k in range(n):                                         # disk -> pinned window
                hf.array[k] = forc[(d0 + k) % nfile]
            o_lo = max(first_out, d0)                                   # output days of this window: [o_lo, d0 + n)
            o_n = max(0, d0 + n - o_lo) if out is not None else 0
            total += engine.run_host(state, par, bnd, hf.array[:n], n, h_out=ho.array[:o_n] if o_n else None,
                                     out_days=o_n, chunk_days=chunk_days)
            if o_n:
                out[o_lo - first_out:o_lo - first_out + o_n] = ho.array[:o_n]
            d0 += n
        if out is not None:
            out.flush()
    finally:
        hf.free()
        if ho is not None:
            ho.free()
    return total
