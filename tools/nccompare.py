#!/usr/bin/env python
"""The reference's regression comparer, restated: netCDF files as src/tools/nccompare.f90 compares them (called by
tools/tests.py:149-151 with `-a 6e-15 -r 35`), ASCII files as tools/tests.py:169-209 does.

Semantics (nccompare.f90:200-264): for every variable present in both files, take the largest absolute difference and the
largest "relative" difference = distance, in units of the last place of a 32-bit float, between the two values after BOTH are
cast to single precision (float_compare :256-264: difference of the IEEE bit patterns read as integers).  A variable fails if
max_absdiff > abs_tol AND NOT (max_ulps > 0 and max_reldiff < max_ulps).  Variables of different total size fail.

Used by tools/dump_for_gfortran.py's recipe: anyone with gfortran + netCDF runs the real cGENIE on the recipe's configuration
and compares its restart / series files with the dumps of this repo's oracle and device using the reference's own pass
criterion.  No netCDF library is needed: scipy.io.netcdf_file reads the classic format both sides write."""
import re
import sys

import numpy as np

ABSTOL = 6.0e-15      # tools/tests.py:131
RELTOL = 35           # tools/tests.py:132


def float_ulps(x, y):
    """nccompare.f90:256-264 / tools/tests.py:156-160, element-wise: |bits(float32(x)) - bits(float32(y))| as signed 32-bit
    integers (for two values of the same sign: their distance in single-precision units of the last place)."""
    with np.errstate(over="ignore", invalid="ignore"):
        x4 = np.asarray(x, dtype=np.float64).astype(np.float32)
        y4 = np.asarray(y, dtype=np.float64).astype(np.float32)
    ix = x4.view(np.int32).astype(np.int64)
    iy = y4.view(np.int32).astype(np.int64)
    return np.abs(ix - iy)


def values_differ(a, b, abs_tol=ABSTOL, max_ulps=RELTOL):
    """do_comparison's verdict on two equally sized arrays: (fails, max_absdiff, max_reldiff)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    if a.size == 0:
        return False, 0.0, 0
    max_abs = float(np.max(np.abs(a - b)))
    max_rel = int(np.max(float_ulps(a, b)))
    tiny = float(np.finfo(np.float64).tiny)
    fails = abs_tol > tiny and max_abs > abs_tol and not (max_ulps > 0 and max_rel < max_ulps)
    return bool(fails), max_abs, max_rel


def compare_nc(f1, f2, abs_tol=ABSTOL, max_ulps=RELTOL, verbose=False, out=sys.stdout):
    """True if the files differ (nccompare's exit status 1)."""
    from scipy.io import netcdf_file
    err = False
    with netcdf_file(f1, "r", mmap=False) as a, netcdf_file(f2, "r", mmap=False) as b:
        common = [n for n in a.variables if n in b.variables]
        for n in common:
            va, vb = a.variables[n], b.variables[n]
            if va.data.dtype.kind in "SUc" or vb.data.dtype.kind in "SUc":
                continue                                  # nccompare reads everything as REAL; text variables are not compared
            da, db = np.asarray(va.data, dtype=np.float64), np.asarray(vb.data, dtype=np.float64)
            if da.size != db.size:
                err = True
                if verbose:
                    print("**ERROR: Differing sizes in variable %s" % n, file=out)
                continue
            fails, mabs, mrel = values_differ(da, db, abs_tol, max_ulps)
            if fails:
                err = True
                if verbose:
                    print("**ERROR: Differing values in variable %s (max abs %.6g, %d float32 ulps)" % (n, mabs, mrel), file=out)
            elif verbose and mabs > abs_tol:
                print("Max. abs. diff. = %.6g but max. rel. diff. = %d < %d   (%s)" % (mabs, mrel, max_ulps, n), file=out)
    if err:
        print("Files %s and %s differ" % (f1, f2), file=out)
    return err


_FP = r"[+-]?(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?"
_FPLINE = re.compile(r"^(" + _FP + r")((\s*,\s*|\s+)" + _FP + r")*$")


def compare_ascii(f1, f2, abs_tol=ABSTOL, max_ulps=RELTOL, out=sys.stdout):
    """tools/tests.py:169-209: line by line; lines that are lists of numbers are compared with the two tolerances."""
    with open(f1) as a, open(f2) as b:
        la, lb = a.read().split("\n"), b.read().split("\n")
    la = [x.strip() for x in la]
    lb = [x.strip() for x in lb]
    while la and la[-1] == "":
        la.pop()
    while lb and lb[-1] == "":
        lb.pop()
    if len(la) != len(lb):
        print("Files %s and %s differ in length" % (f1, f2), file=out)
        return True
    for x, y in zip(la, lb):
        if x == y:
            continue
        if not (_FPLINE.match(x) and _FPLINE.match(y)):
            print("Files %s and %s are different" % (f1, f2), file=out)
            return True
        xs = [float(t) for t in x.replace(",", " ").split()]
        ys = [float(t) for t in y.replace(",", " ").split()]
        if len(xs) != len(ys) or values_differ(xs, ys, abs_tol, max_ulps)[0]:
            print("Files %s and %s are different" % (f1, f2), file=out)
            return True
    return False


def file_compare(f1, f2, **kw):
    """netCDF by magic number, ASCII otherwise (tools/tests.py:212-230)."""
    with open(f1, "rb") as fh:
        magic = fh.read(3)
    return compare_nc(f1, f2, **kw) if magic == b"CDF" else compare_ascii(f1, f2, **{k: v for k, v in kw.items() if k != "verbose"})


def main(argv):
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("-a", type=float, default=ABSTOL, help="absolute tolerance (default: the reference test suite's 6e-15)")
    ap.add_argument("-r", type=int, default=RELTOL, help="relative tolerance in float32 ulps (default 35)")
    ap.add_argument("-v", action="store_true")
    ap.add_argument("file_a")
    ap.add_argument("file_b")
    a = ap.parse_args(argv)
    return 1 if file_compare(a.file_a, a.file_b, abs_tol=a.a, max_ulps=a.r, verbose=a.v) else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
