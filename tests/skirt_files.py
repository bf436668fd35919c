"""Readers for the output files SKIRT writes (used by the fixture generator and by the tests of the C++ shim)."""
import numpy as np


def read_fits_cube(path):
    """Primary HDU of a SKIRT frame file: BITPIX=-32 big-endian float32 (FITSInOut.cpp:160,194) -> [nz][ny][nx]."""
    raw = open(path, "rb").read()
    cards = {}
    pos = 0
    while True:
        block = raw[pos:pos + 2880].decode("ascii")
        pos += 2880
        done = False
        for i in range(0, 2880, 80):
            c = block[i:i + 80]
            if c.startswith("END"):
                done = True
                break
            if "=" in c[:10]:
                cards[c[:8].strip()] = c[10:].split(" / ")[0].strip().strip("'").strip()
        if done:
            break
    nx, ny = int(cards["NAXIS1"]), int(cards["NAXIS2"])
    nz = int(cards.get("NAXIS3", 1))
    data = np.frombuffer(raw, dtype=">f4", count=nx * ny * nz, offset=pos).astype(np.float32)
    return data.reshape(nz, ny, nx), cards


def read_columns(path):
    """A SKIRT column text file (TextOutFile.cpp:81-98): '#' header lines, then %.9e columns."""
    return np.loadtxt(path, comments="#", ndmin=2)
