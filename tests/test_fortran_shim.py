"""Static checks of the ISO_C_BINDING layer (fortran/*.f90) against include/cgenie_b200.h.  The image has no Fortran compiler,
so what can be checked is checked on the text: every BIND(C) interface names a function the header declares, with the same
number of arguments, passing scalars by VALUE where C takes them by value and references / C_PTR where C takes pointers;
every cg_* function the shim modules call has an interface; and the ctypes table of cgenie_b200/_lib.py has the header's arity."""
import glob
import os
import re

from cgenie_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_prototypes():
    hdr = open(os.path.join(ROOT, "include", "cgenie_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char \*|int64_t|long long)\s*\*?\s*(cg_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        if args == ["void"] or args == [""]:
            args = []
        out[m.group(1)] = args
    return out


def fortran_interfaces():
    src = open(os.path.join(ROOT, "fortran", "cgenie_b200_c.f90")).read()
    src = re.sub(r"&\s*\n\s*&?", " ", src)
    out = []
    for m in re.finditer(r"FUNCTION\s+(\w+)\s*\(([^)]*)\)\s*BIND\(C,\s*NAME='(\w+)'\)(.*?)END FUNCTION", src, flags=re.S | re.I):
        dummies = [d.strip() for d in m.group(2).split(",") if d.strip()]
        decl = {}
        for line in m.group(4).split("\n"):
            line = line.split("!")[0]
            if "::" not in line or "IMPORT" in line.upper():
                continue
            spec, names = line.split("::")
            for n in names.split(","):
                decl[n.strip().split("(")[0]] = spec.upper()
        out.append((m.group(1), m.group(3), dummies, decl))
    return out


def test_fortran_interfaces_match_header():
    protos = c_prototypes()
    assert len(protos) >= 60
    itf = fortran_interfaces()
    assert len(itf) >= 20
    for fname, cname, dummies, decl in itf:
        assert cname in protos, "%s binds %s, which the header does not declare" % (fname, cname)
        cargs = protos[cname]
        assert len(dummies) == len(cargs), (fname, dummies, cargs)
        for d, c in zip(dummies, cargs):
            spec = decl.get(d)
            assert spec is not None, "%s: dummy %s has no declaration" % (fname, d)
            by_value = "VALUE" in spec
            c_is_pointer = "*" in c or "[" in c
            if c_is_pointer:
                # a C pointer is either a C_PTR passed by value or a Fortran entity passed by reference
                assert ("C_PTR" in spec and by_value) or not by_value, (fname, d, spec, c)
            else:
                assert by_value, "%s: %s must be passed by VALUE (C takes '%s')" % (fname, d, c)
                if "double" in c:
                    assert "C_DOUBLE" in spec, (fname, d, spec)
                elif "int64_t" in c:
                    assert "C_INT64_T" in spec, (fname, d, spec)
                elif re.match(r"(const )?int\b", c):
                    assert "C_INT)" in spec or "C_INT," in spec or spec.strip().endswith("C_INT)"), (fname, d, spec)


def test_shim_modules_call_declared_functions():
    declared = {f for f, _, _, _ in fortran_interfaces()}
    used = set()
    for path in glob.glob(os.path.join(ROOT, "fortran", "*.f90")):
        src = open(path).read()
        src = "\n".join(line.split("!")[0] for line in src.split("\n"))
        used |= set(re.findall(r"\b(cg_[a-z_0-9]+)\s*\(", src))
    helpers = {"cg_check", "cg_ensure_handle"}
    types = {n for n in used if n.endswith("_io")}
    missing = sorted(n for n in used - declared - helpers - types if not re.search(r"SUBROUTINE\s+%s|FUNCTION\s+%s" % (n, n),
                     open(os.path.join(ROOT, "fortran", "cgenie_b200_c.f90")).read(), flags=re.I))
    assert not missing, "shim modules call undeclared C functions: %s" % missing
    # the shim modules and the procedures the coupler's wrappers call through them
    bio = open(os.path.join(ROOT, "fortran", "biogem_b200.f90")).read()
    for mod, procs in (("biogem_b200", ["step_biogem", "biogem_tracercoupling", "biogem_climate", "biogem_climate_sol", "biogem_forcing"]),
                       ("atchem_b200", ["step_atchem", "cpl_flux_ocnatm", "cpl_comp_atmocn", "cpl_comp_EMBM"]),
                       ("sedgem_b200", ["cpl_flux_ocnsed", "cpl_comp_ocnsed"]), ("rokgem_b200", ["reinit_flux_rokocn"])):
        m = re.search(r"^MODULE %s\b(.*?)^END MODULE %s\b" % (mod, mod), bio, flags=re.S | re.M)
        assert m, mod
        for p in procs:
            assert re.search(r"SUBROUTINE %s\b" % p, m.group(1)), (mod, p)


def test_shims_can_link_next_to_the_reference_modules():
    """shim modules named like the reference's (MODULE goldstein ...) could neither replace nor coexist with them.
    Now every shim module has its own name (*_b200), every C binding label is declared exactly once, and
    fortran/use_b200.py rewrites only the USE line of the hot-path wrappers."""
    names, labels = [], []
    for path in glob.glob(os.path.join(ROOT, "fortran", "*.f90")):
        src = "\n".join(line.split("!")[0] for line in open(path).read().split("\n"))
        names += re.findall(r"^\s*MODULE\s+(\w+)\s*$", src, flags=re.M | re.I)
        labels += re.findall(r"BIND\(C,\s*NAME='(\w+)'\)", src, flags=re.I)
    assert len(names) == 8 and all(n.lower().endswith("_b200") or n == "cgenie_b200_c" for n in names), names
    assert len(set(n.lower() for n in names)) == len(names)
    assert len(labels) == len(set(labels)), sorted(x for x in labels if labels.count(x) > 1)
    import importlib.util
    spec = importlib.util.spec_from_file_location("use_b200", os.path.join(ROOT, "fortran", "use_b200.py"))
    ub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ub)
    # a file shaped like src/wrappers/genie_loop_wrappers.f90: one USE line per wrapper, other wrappers left alone
    text = "MODULE genie_loop_wrappers\n  USE genie_global\nCONTAINS\n"
    for w, (ref, shim, procs) in ub.WRAPPERS.items():
        text += "  SUBROUTINE %s\n    USE %s\n    IMPLICIT NONE\n    CALL %s(x)\n  END SUBROUTINE %s\n" % (w, ref.upper() if w == "embm_wrapper" else ref, procs[0], w)
    text += "  SUBROUTINE goldstein_save_restart_wrapper\n    USE goldstein_data\n  END SUBROUTINE goldstein_save_restart_wrapper\n"
    text += "  SUBROUTINE diag_biogem_timeslice_wrapper\n    USE biogem\n  END SUBROUTINE diag_biogem_timeslice_wrapper\nEND MODULE\n"
    new, changes, missing = ub.rewrite(text)
    assert not missing and len(changes) == len(ub.WRAPPERS)
    assert "USE goldstein_b200, ONLY: step_goldstein" in new and "USE goldstein_data" in new
    assert new.count("USE biogem\n") == 1                       # the diagnostic wrapper keeps the reference module
    assert ub.rewrite(new)[1] == []                              # idempotent
    # every procedure the script routes to a shim module is defined there with that name
    allsrc = "".join(open(pth).read() for pth in glob.glob(os.path.join(ROOT, "fortran", "*.f90")))
    for w, (ref, shim, procs) in ub.WRAPPERS.items():
        m = re.search(r"^MODULE %s\b(.*?)^END MODULE %s\b" % (shim, shim), allsrc, flags=re.S | re.M)
        assert m, shim
        for p in procs:
            assert re.search(r"SUBROUTINE %s\b" % p, m.group(1), flags=re.I), (shim, p)
            assert re.search(r"PUBLIC\s*::.*\b%s\b" % p, m.group(1), flags=re.I), (shim, p)


def test_ctypes_table_has_header_arity():
    protos = c_prototypes()
    for name, (_, argtypes) in _lib.SYMBOLS.items():
        assert name in protos, name
        assert len(argtypes) == len(protos[name]), (name, len(argtypes), protos[name])


def test_shim_procedures_take_the_reference_argument_lists():
    """The drop-in must take exactly what genie_loop_wrappers.f90 passes.  tests/golden/ref_signatures.json holds the dummy-argument
    lists of the reference's hot-path procedures (name, type, rank, INTENT; parsed from the reference's own Fortran with
    numpy.f2py.crackfortran by tools/make_golden.py); the shim modules, parsed the same way, must agree argument for argument:
    same names in the same order, same type and rank, and no INTENT that contradicts the reference's."""
    import contextlib
    import io
    import json
    from numpy.f2py import crackfortran
    crackfortran.verbose = 0
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_signatures.json")))["procedures"]

    def walk(b, out):
        if b.get("block") in ("subroutine", "function"):
            out[b["name"].lower()] = b
        for c in b.get("body", []):
            walk(c, out)

    procs = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "fortran", "*_b200.f90"))):
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            blocks = crackfortran.crackfortran([path])
        for b in blocks:
            assert b["block"] == "module" and b["name"].endswith("_b200"), (path, b["name"])
            walk(b, procs)
    assert len(ref) == 16
    for name, sig in ref.items():
        assert name in procs, "no shim for %s (%s)" % (name, sig["file"])
        p = procs[name]
        assert [a.lower() for a in p["args"]] == [a["name"] for a in sig["args"]], name
        for a in sig["args"]:
            v = p["vars"][[x for x in p["args"] if x.lower() == a["name"]][0]]
            assert v.get("typespec") == a["type"], (name, a["name"], v.get("typespec"), a["type"])
            assert len(v.get("dimension", [])) == a["rank"], (name, a["name"], v.get("dimension"), a["rank"])
            mine = sorted(v.get("intent") or [])
            mine = [x for x in mine if x in ("in", "out", "inout")]
            if a["intent"] and mine:
                # the shim may be stricter about what it leaves alone, never write what the reference only reads
                assert not (a["intent"] == ["in"] and mine != ["in"]), (name, a["name"], mine, a["intent"])
