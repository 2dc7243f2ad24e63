#!/usr/bin/env python
"""Point the coupler's hot-path wrappers at the B200 shim modules.

    python fortran/use_b200.py <cgenie>/src/wrappers/genie_loop_wrappers.f90 [--check] [--revert]

The reference's wrappers each carry their own USE line (src/wrappers/genie_loop_wrappers.f90:7-468: `SUBROUTINE
goldstein_wrapper / USE goldstein / ... CALL step_goldstein(...)`).  This script rewrites exactly that line in the wrappers
listed below to `USE <module>_b200, ONLY: <procedure>`; nothing else in the reference changes -- the modules themselves stay in
the build for their initialise / end / restart / diagnostic routines, the call statements and genie_global's arrays are
untouched, and the SConscript only gains fortran/*.f90 and -lcgenie_b200.  A backup (.orig) is kept; --revert restores it;
--check reports what would change.  The file is edited in place because the reference's sources are never copied into this
repository."""
import re
import sys

# wrapper subroutine -> (module the reference USEs, shim module, procedures the wrapper calls)
WRAPPERS = {
    "surflux_wrapper": ("embm", "embm_b200", ["surflux"]),
    "embm_wrapper": ("embm", "embm_b200", ["step_embm"]),
    "gold_seaice_wrapper": ("gold_seaice", "gold_seaice_b200", ["step_seaice"]),
    "goldstein_wrapper": ("goldstein", "goldstein_b200", ["step_goldstein"]),
    "cpl_flux_ocnatm_wrapper": ("atchem", "atchem_b200", ["cpl_flux_ocnatm"]),
    "cpl_flux_ocnsed_wrapper": ("sedgem", "sedgem_b200", ["cpl_flux_ocnsed"]),
    "cpl_comp_ocnsed_wrapper": ("sedgem", "sedgem_b200", ["cpl_comp_ocnsed"]),
    "reinit_flux_rokocn_wrapper": ("rokgem", "rokgem_b200", ["reinit_flux_rokocn"]),
    "biogem_wrapper": ("biogem", "biogem_b200", ["step_biogem"]),
    "biogem_tracercoupling_wrapper": ("biogem", "biogem_b200", ["biogem_tracercoupling"]),
    "biogem_forcing_wrapper": ("biogem", "biogem_b200", ["biogem_forcing"]),
    "biogem_climate_wrapper": ("biogem", "biogem_b200", ["biogem_climate"]),
    "biogem_climate_sol_wrapper": ("biogem", "biogem_b200", ["biogem_climate_sol"]),
    "atchem_wrapper": ("atchem", "atchem_b200", ["step_atchem"]),
    "cpl_comp_atmocn_wrapper": ("atchem", "atchem_b200", ["cpl_comp_atmocn"]),
    "cpl_comp_EMBM_wrapper": ("atchem", "atchem_b200", ["cpl_comp_EMBM"]),
}


def rewrite(text):
    """Returns (new text, list of (wrapper, old USE line, new USE line)); raises if a listed wrapper has no such USE line."""
    lines = text.split("\n")
    changes, current = [], None
    for n, ln in enumerate(lines):
        m = re.match(r"\s*SUBROUTINE\s+(\w+)", ln, flags=re.I)
        if m and not re.match(r"\s*END\s+SUBROUTINE", ln, flags=re.I):
            current = m.group(1)
            continue
        if re.match(r"\s*END\s+SUBROUTINE", ln, flags=re.I):
            current = None
            continue
        key = next((w for w in WRAPPERS if current and w.lower() == current.lower()), None)
        if key is None:
            continue
        ref_mod, shim_mod, procs = WRAPPERS[key]
        m = re.match(r"(\s*)USE\s+(\w+)\s*$", ln, flags=re.I)
        if m and m.group(2).lower() == ref_mod.lower():
            new = "%sUSE %s, ONLY: %s" % (m.group(1), shim_mod, ", ".join(procs))
            changes.append((key, ln.strip(), new.strip()))
            lines[n] = new
    missing = sorted(set(WRAPPERS) - {c[0] for c in changes})
    return "\n".join(lines), changes, missing


def main(argv):
    if not argv or argv[0].startswith("-"):
        print(__doc__)
        return 2
    path, flags = argv[0], set(argv[1:])
    if "--revert" in flags:
        open(path, "w").write(open(path + ".orig").read())
        print("restored", path)
        return 0
    text = open(path).read()
    new, changes, missing = rewrite(text)
    for w, old, nw in changes:
        print("%-32s %-22s -> %s" % (w, old, nw))
    if missing:
        already = [w for w in missing if re.search(r"USE\s+%s\b" % WRAPPERS[w][1], text, flags=re.I)]
        rest = [w for w in missing if w not in already]
        if rest:
            print("not found (wrapper absent or its USE line differs):", ", ".join(rest))
            return 1
    if "--check" not in flags and changes:
        open(path + ".orig", "w").write(text)
        open(path, "w").write(new)
        print("rewrote %s (%d USE lines; backup %s.orig)" % (path, len(changes), path))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
