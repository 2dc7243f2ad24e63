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
    # the module / procedure names the unchanged genie_loop_wrappers.f90 USEs for the coupler calls of the BIOGEM block
    bio = open(os.path.join(ROOT, "fortran", "biogem_b200.f90")).read()
    for mod, procs in (("biogem", ["step_biogem", "biogem_tracercoupling", "biogem_climate"]),
                       ("atchem", ["step_atchem", "cpl_flux_ocnatm", "cpl_comp_atmocn"]),
                       ("sedgem", ["cpl_flux_ocnsed", "cpl_comp_ocnsed"]), ("rokgem", ["reinit_flux_rokocn"])):
        m = re.search(r"^MODULE %s\b(.*?)^END MODULE %s\b" % (mod, mod), bio, flags=re.S | re.M)
        assert m, mod
        for p in procs:
            assert re.search(r"SUBROUTINE %s\b" % p, m.group(1)), (mod, p)


def test_ctypes_table_has_header_arity():
    protos = c_prototypes()
    for name, (_, argtypes) in _lib.SYMBOLS.items():
        assert name in protos, name
        assert len(argtypes) == len(protos[name]), (name, len(argtypes), protos[name])
