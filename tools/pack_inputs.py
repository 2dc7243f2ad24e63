#!/usr/bin/env python
"""Pack the reference's ASCII input data for the hot path into configs/inputs.npz.

Run HERE (the container that mounts /root/reference); the GPU box has no
/root/reference, so tests, smoke() and bench.py read the packed arrays and
`cgenie_b200.jobdir.materialise()` writes them back out in the reference's own
file formats (ints verbatim, reals with repr() so the parsed doubles are
bit-identical).

Sources (all under /root/reference/data):
  goldstein/<world>.k1 .psiles .paths   (goldstein.f90:1109-1123, 1502-1577)
  embm/taux_u.interp ... tauy_v.interp  (embm.f90:845-860)
  embm/uncep.silo vncep.silo            (embm.f90:989-995)
  biogem/worjh2_preindustrial/windspeed.dat (biogem_data.f90:1171-1175)
"""
import os
import sys
import numpy as np

REF = os.environ.get("CGENIE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs", "inputs.npz")


def read_numbers(path):
    with open(path) as f:
        return [float(t) for t in f.read().split()]


def read_paths(path):
    with open(path) as f:
        lines = f.read().split("\n")
    npi = [int(t) for t in lines[0].split()]
    trip, ln = [], 1
    for n in npi:
        ln += 1  # skipped line per island (goldstein.f90:1549)
        for _ in range(n):
            trip.append([int(t) for t in lines[ln].split()[:3]])
            ln += 1
    return np.array(npi, dtype=np.int32), np.array(trip, dtype=np.int32).reshape(-1, 3)


def collect():
    """The packed arrays, read from the reference's data files."""
    out = {}
    # worbe2 / worjh2: the target configurations (one island each); p0055c (2 islands), p0251a (3 islands): 36 x 36 x 16 worlds for
    # the multi-island barotropic closure (matmult, goldstein.f90:203-216, 3470-3492)
    for world in ("worbe2", "worjh2", "p0055c", "p0251a"):
        g = os.path.join(REF, "data", "goldstein", world)
        k1 = np.array([int(t) for t in open(g + ".k1").read().split()], dtype=np.int32)
        out[world + "/k1"] = k1.reshape(38, 38)          # file order: first row j=maxj+1
        out[world + "/psiles"] = np.array(read_numbers(g + ".psiles")).reshape(37, 36)
        npi, trip = read_paths(g + ".paths")
        out[world + "/npi"] = npi
        out[world + "/paths"] = trip
    e = os.path.join(REF, "data", "embm")
    for nm in ("taux_u", "tauy_u", "taux_v", "tauy_v"):
        out["winds/" + nm] = np.array(read_numbers(os.path.join(e, nm + ".interp")))
    for nm in ("uncep", "vncep"):
        out["winds/" + nm] = np.array(read_numbers(os.path.join(e, nm + ".silo")))
    # BIOGEM prescribed wind speed, file rows j = maxj..1, i = 1..maxi per row (gem_util.f90:511-536)
    b = os.path.join(REF, "data", "biogem", "worjh2_preindustrial", "windspeed.dat")
    out["biogem/worjh2_windspeed"] = np.array(read_numbers(b)).reshape(36, 36)   # [row = maxj - j][i - 1]
    return out


def main():
    out = collect()
    for k, v in out.items():
        print(k, v.shape, v.dtype)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
