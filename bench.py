#!/usr/bin/env python
"""bench.py -- ensemble model-years per wall-hour of the cGENIE hot path on B200.

One "step" = one model year of every ensemble member resident on this rank's GPU: nyear ocean
steps (tstepo_flux + co + momentum), 5*nyear EMBM steps, nyear surflux and sea-ice steps, nyear/2
BIOGEM steps (step_biogem + tracer coupling + climate) and nyear/2 ATCHEM steps.  Members are sharded across ranks with no collective
on the timestep path (weak scaling: members per GPU fixed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--members M] [--spinup-years Y] [--impl reference]

Prints ONE JSON line (see the task contract): value = whole-job model-years/hour with state
resident in HBM; e2e = the same through the per-module C-ABI entry points the Fortran host calls,
with the year's state uploaded from / downloaded to pinned host memory inside the timed region;
roofline = tracer-step kernel (tstepo_flux) algorithmic bytes / CUDA-event time vs the measured
HBM peak; cpu_baseline = the CPU oracle on the host cores (bounded sample).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIG = "eb_go_gs_ac_bg_36x36x16"
WORKLOAD = ("eb_go_gs_ac_bg 36x36x16 worjh2: EMBM + GOLDSTEIN + sea ice + BIOGEM (16 ocean tracers, 1N1T_PO4MM, "
            "13C/14C, CFCs) + ATCHEM, parameter-perturbed ensemble sharded by member "
            "(BASELINE config #4: the 1024-member ensemble's shard at 2 GPUs, 512 members/GPU, held per GPU at every N -- one "
            "library handle, the kernels run over 128-member tiles; --members 128 is round 1's shard)")
MEMBERS_PER_GPU = 512
# --config N: the other BASELINE.json configurations (the default line, N = 4, is the one the metric is quoted on)
#   job configuration, members per GPU, BIOGEM, default untimed spin-up years, workload text
CONFIGS = {
    1: ("eb_go_gs_36x36x8", 1, False, 10, "eb_go_gs physics only (EMBM + GOLDSTEIN + sea ice) 36x36x8 worbe2, 100 ocean steps / year, "
        "SINGLE member (BASELINE config #1, the reference's own CPU test job)"),
    2: (CONFIG, 1, True, 100, "eb_go_gs_ac_bg 36x36x16 worjh2 with BIOGEM (16 ocean tracers) + ATCHEM, SINGLE member "
        "(BASELINE config #2: 100-year spin-up on one B200)"),
    3: (CONFIG, 64, True, 100, "eb_go_gs_ac_bg 36x36x16 worjh2 with BIOGEM + ATCHEM, 64-member parameter-perturbation ensemble on one "
        "B200 (BASELINE config #3)"),
    4: (CONFIG, MEMBERS_PER_GPU, True, 100, WORKLOAD),
}
from cgenie_b200.sharding import PERTURBED, PERTURBED_BIOGEM, SEED, perturbation_table, shard  # noqa: E402  (pure numpy)


def config_dict(workload, M, I, J, K, L, nyear, variant, biogem, spinup_years, member_stride=None, adrag_group=16):
    """The `config` object of the JSON line; the reference arm prints the same object for the same workload."""
    ms = member_stride or ((M + 31) // 32) * 32
    working_set = (2 * L + 6) * I * J * K * 8 * ms
    in_l2 = working_set < 100e6
    return {"workload": workload, "members_per_gpu": M, "member_stride": ms, "grid": [I, J, K], "tracers": L, "nyear": nyear,
            "tracer_variant": variant, "perturbed": PERTURBED + (PERTURBED_BIOGEM if biogem else []), "seed": SEED,
            "adrag_groups": ("adrag is perturbed per group of %d members (members of a group share one barotropic factorisation)" % adrag_group)
                            if adrag_group > 1 else "adrag is perturbed per member (one barotropic factorisation each)",
            "l2": ("working set %.0f MB per GPU (two ts buffers + u + rho) exceeds the 126 MB L2" if not in_l2 else
                   "working set %.1f MB per GPU (two ts buffers + u + rho) fits the 126 MB L2: L2-resident run") % (working_set / 1e6),
            "step": "one model year of every member: %d koverall iterations" % (5 * nyear),
            "state": "%d model years of untimed spin-up from the uniform initial state, then the warm-up years" % spinup_years}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(members, variant, note=False):
    """dram__bytes_read.sum + dram__bytes_write.sum per tstepo launch (flux + convection kernels) from the committed
    `ncu --set full` capture of this configuration (profiles/ncu_traffic.json), or None when none was taken.  note=True: which
    capture that is (round, kernels, the entry's own remark)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        for row in reversed(json.load(open(p))["captures"]):   # latest capture of this configuration
            if row["members"] == members and row["variant"] == variant and row["config"] == CONFIG:
                if note:
                    return "capture %s (%s)%s" % (row.get("round"), ", ".join(sorted(row.get("kernels", {}))),
                                                  "; " + row["note"] if row.get("note") else "")
                return row["dram_bytes_per_tstepo_launch"]
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.rows, self.stop_flag = dev, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for q, n in enumerate(names) if any(len(r) > 2 + q and r[2 + q].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------- CPU (oracle) arms
def _oracle_worker(args):
    years, params = args
    from oracle_lib import Oracle
    o = Oracle("worjh2", maxk=16, maxl=16, nyear=96, **{k: v for k, v in params.items() if not k.startswith("par_bio")})
    o.biogem_setup(**{k: v for k, v in params.items() if k.startswith("par_bio")})
    t0 = time.perf_counter()
    o.run(int(round(years * 96 * 5)))
    dt = time.perf_counter() - t0
    o.close()
    return dt


def cpu_oracle_rate(years_per_core, cores):
    """Aggregate model-years/hour of `cores` single-threaded oracle processes, one member each."""
    import oracle_lib
    oracle_lib.lib()  # build once before forking
    tab = perturbation_table(max(cores, 16), biogem=True)
    jobs = [(years_per_core, {k: float(v[c]) for k, v in tab.items()}) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    return cores * years_per_core / (wall / 3600.0), wall


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ypc = 4.0  # bounded sample: each step = 4 model years of one member on every host core (~5 s)
    for _ in range(args.warmup):
        cpu_oracle_rate(0.5, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_rate(ypc, cores)
    wall = time.perf_counter() - t0
    value = cores * ypc * args.steps / (wall / 3600.0)
    sample = "%d oracle processes (one per host core) x %.2f model-year per step" % (cores, ypc)
    print(json.dumps({
        "impl": "reference", "metric": "ensemble model-years/wall-hour", "value": value, "unit": "model-years/hour",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(WORKLOAD, MEMBERS_PER_GPU, 36, 36, 16, 16, 96, "col", True, 100),
        "note": "reference arm = the C restatement of the reference (oracle/, -O3 -funroll-loops, no FMA) on the host cores, not a "
                "gfortran build; it starts from the initial state (no spin-up: the CPU cost of a model year does not depend on it)",
        "cpu_baseline": {"value": value, "unit": "model-years/hour", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "model-years/hour", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------- BASELINE config #5
def config5_fields(t, k1, I, J, K, L):
    """tracers c_l = 1 + 0.1 sin(2 pi i / I) cos(pi j / J) (k / K) (1 + l / L), T and S from an analytic
    stratified profile, horizontal velocities from a prescribed stream function (non-divergent) with w from continuity as velc
    computes it (goldstein.f90:3668-3678).  One member, Fortran-shaped: ts (K+2, J+2, I+2, L), u (K, J+1, I+1, 3)."""
    c, cv, ds, dz = t.const("c"), t.const("cv"), t.const("ds"), t.const("dz")
    rdphi = t.const("scalars")[1]
    kk, jj, ii = np.meshgrid(np.arange(K + 2), np.arange(J + 2), np.arange(I + 2), indexing="ij")
    ts = np.zeros((K + 2, J + 2, I + 2, L))
    ts[..., 0] = 2.0 + 18.0 * (kk / (K + 1.0)) ** 2 + 1.5 * np.cos(2 * np.pi * ii / I) * np.sin(np.pi * jj / J)
    ts[..., 1] = 0.3 * np.sin(2 * np.pi * ii / I) * (kk / (K + 1.0)) - 0.1 * np.cos(np.pi * jj / J)
    base = 0.1 * np.sin(2 * np.pi * ii / I) * np.cos(np.pi * jj / J) * (kk / K)
    for l in range(2, L):
        ts[..., l] = 1.0 + base * (1 + l / L)
    ts[K + 1] = 0.0
    ts[:, :, 0, :] = ts[:, :, I, :]
    ts[:, :, I + 1, :] = ts[:, :, 1, :]
    jv, iv = np.arange(J + 1), np.arange(I + 1)
    psi2 = np.sin(np.pi * np.minimum(jv, J - 2) / (J - 2))[:, None] ** 2 * (1 + 0.5 * np.sin(2 * np.pi * iv / I))[None, :]
    psi = 0.02 * (np.arange(K + 1) / K)[:, None, None] * psi2[None]                                  # (K+1, J+1, I+1)
    u = np.zeros((K, J + 1, I + 1, 3))
    k1e = np.maximum(k1[1:J + 1, 1:I + 1], k1[1:J + 1, 2:I + 2])                                       # east face open from this level
    k1n = np.maximum(k1[1:J + 1, 1:I + 1], k1[2:J + 2, 1:I + 1])
    lev = np.arange(1, K + 1)[:, None, None]
    ue = -c[1:J + 1][None, :, None] * (psi[1:, 1:, 1:] - psi[1:, :-1, 1:]) / ds[1:J + 1][None, :, None]
    vn = (psi[1:, 1:, 1:] - psi[1:, 1:, :-1]) * rdphi / np.where(cv[1:J + 1] != 0, cv[1:J + 1], 1.0)[None, :, None]
    u[:, 1:, 1:, 0] = np.where(lev >= k1e[None], ue, 0.0)
    u[:, 1:, 1:, 1] = np.where((lev >= k1n[None]) & (np.arange(1, J + 1) < J)[None, :, None], vn, 0.0)
    u[:, :, 0, 0] = u[:, :, I, 0]
    tv1 = (u[:, 1:, 1:, 0] - u[:, 1:, :-1, 0]) * rdphi / c[1:J + 1][None, :, None]
    tv2 = (u[:, 1:, 1:, 1] * cv[1:J + 1][None, :, None] - u[:, :-1, 1:, 1] * cv[0:J][None, :, None]) / ds[1:J + 1][None, :, None]
    wet = lev >= k1[1:J + 1, 1:I + 1][None]
    w = -np.cumsum(np.where(wet, dz[1:K + 1][:, None, None] * (tv1 + tv2), 0.0), axis=0)
    w[K - 1] = 0.0
    u[:, 1:, 1:, 2] = np.where(wet, w, 0.0)
    return ts, u


def run_config5(args, rank, world, local):
    """Stand-alone tracer step (tstepo_flux + co) on the synthetic 128 x 128 x 32 grid with 40 tracers, M members per GPU, through
    cg_tracer_create / set / step (the generic-shape kernels: the column kernel is compiled for 36 x 36 x 16, L = 16 only)."""
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cgenie_b200 import TracerStep
    I, J, K, L = 128, 128, 32, 40
    M = args.members or 64
    k1 = np.ones((J + 2, I + 2), dtype=np.int32)
    k1[0, :] = 94
    k1[J - 1:J + 2, :] = 92                                    # 2-cell polar land cap
    t = TracerStep(I, J, K, L, k1, n_members=M, device=local, diff1=2000.0, diff2=1e-5, nyear=96)
    t.set_tracer_variant(args.variant)      # col: the tracer-window column kernels (member strides 32 / 64 / 128), else the generic kernels
    ts1, u1 = config5_fields(t, k1, I, J, K, L)
    fac = (1.0 + 0.01 * np.arange(M))[:, None, None, None, None]          # members differ by a scale of the passive tracers
    ts = np.repeat(ts1[None], M, axis=0)
    ts[..., 2:] *= fac
    u = np.repeat(u1[None], M, axis=0)
    flux = np.zeros((M, J, I, 2))
    t.set(ts=ts, u=u, tsflux=flux)
    for _ in range(max(args.warmup, 3)):
        t.step(1)
    t.synchronize()
    if world > 1:
        dist.barrier()
    t.launch_count(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    t.timer_start()
    t.step(args.steps)
    ms = t.timer_stop_ms()
    launches = t.launch_count()
    clocks = sampler.summary()
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    n_wet = int(np.sum(np.clip(K - k1[1:J + 1, 1:I + 1] + 1, 0, None)[k1[1:J + 1, 1:I + 1] <= K]))
    bytes_per_launch = n_wet * (16 * L + 32) * M
    avg_ms = ms / args.steps
    peak, peak_src = peaks()
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
    # end to end through the same entry points with host arrays: upload the state, one step, download ts / rho / cost
    t0 = time.perf_counter()
    t.set(ts=ts, u=u, tsflux=flux)
    t.step(1)
    got = t.fetch()
    e2e_s = time.perf_counter() - t0
    assert np.isfinite(got[0]).all()
    out = {"metric": "tracer-step HBM GB/s vs peak", "value": world * achieved, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": avg_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": "synthetic 128x128x32 grid, 40 tracers, stand-alone tracer step (tstepo_flux + co), all-wet flat bottom "
                                  "with a 2-cell polar land cap (BASELINE config #5)", "members_per_gpu": M, "grid": [I, J, K], "tracers": L,
                      "tracer_variant": t.tracer_variant_active(),
                      "l2": "working set %.1f GB per GPU (two ts buffers + u + rho) exceeds the 126 MB L2" % ((2 * L + 6) * I * J * K * 8 * t.member_stride / 1e9),
                      "step": "one tstepo of every member"},
           "clocks": clocks, "gpu_launches": launches,
           "roofline": {"bound": "hbm", "kernel": ("tstepo = 3 x k_tstep_colx (tracer windows 16 + 12 + 12) + k_co_col" if t.tracer_variant_active() == "col"
                                                    else "tstepo = k_tstepo_flux_coop + k_co_fast2 (generic-shape kernels)"), "achieved": achieved,
                        "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms},
           "e2e": {"value": world * bytes_per_launch / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(ts.nbytes + u.nbytes + flux.nbytes),
                   "d2h_bytes_per_step": int(sum(a.nbytes for a in got)),
                   "path": "cg_tracer_set / cg_tracer_step / cg_tracer_get with pageable host arrays (one step)"}}
    t.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    os.close(json_fd)


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--members", type=int, default=None, help="ensemble members per GPU (default: the configuration's)")
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configuration (4 = default line; 1, 2 = single member: L2-resident; 5 = synthetic "
                         "128x128x32 x 40 tracers, stand-alone tracer step)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--variant", default="col", choices=["col", "fast", "strict"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-groups", type=int, default=0, help="end-to-end leg: member groups per GPU (default 1)")
    ap.add_argument("--adrag-group", type=int, default=16, help="members per distinct adrag value (1: every member its own barotropic factors)")
    ap.add_argument("--e2e-serial", action="store_true", help="end-to-end leg: only the serial exchange (upload, compute, download one after the other)")
    ap.add_argument("--e2e-dense", action="store_true", help="end-to-end leg: ship ts dense (all cells) instead of the wet cells only")
    ap.add_argument("--spinup-years", type=int, default=None,
                    help="untimed model years from the uniform initial state before the warm-up (config #2's 100-year spin-up: "
                         "the convective adjustment is data dependent, 70 %% of all cells mix in a 4-year-old ocean; ~7 s)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.config == 5:
        return run_config5(args, rank, world, local)
    cfgname, cfg_members, biogem, cfg_spin, workload = CONFIGS[args.config]
    if args.members is None:
        args.members = cfg_members
    if args.spinup_years is None:
        args.spinup_years = cfg_spin

    # stdout carries exactly one JSON line: NCCL prints its version banner with printf on some boxes, so file descriptor 1
    # points at stderr while the job runs and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from cgenie_b200 import Ensemble, materialise
    M = args.members
    pert = shard(perturbation_table(M * world, biogem=biogem, adrag_group=args.adrag_group), rank, world, M)
    tmp = tempfile.mkdtemp(prefix="cgenie_job_")
    materialise(tmp, cfgname)
    e = Ensemble(tmp, n_members=M, device=local, perturb=pert)
    e.set_tracer_variant(args.variant)
    if os.environ.get("CG_BENCH_FUSE"):      # A/B knob: tracer coupling fused into the BIOGEM step kernel (cg_set_biogem_fusion)
        e.set_biogem_fusion(True)
    kyear = e.nyear * e.ndta
    L, I, J, K = e.maxl, e.maxi, e.maxj, e.maxk

    # ---- device-resident throughput
    if args.spinup_years > 0:
        e.run(kyear * args.spinup_years)
        e.synchronize()
    for _ in range(args.warmup):
        e.run(kyear)
    e.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e.launch_count(reset=True)
    barrier()
    e.timer_start()
    for _ in range(args.steps):
        e.run(kyear)
    ms = e.timer_stop_ms()
    barrier()
    launches = e.launch_count()
    clocks = sampler.summary()
    ms = max_over_ranks(ms)
    value = world * M * args.steps / (ms / 3.6e6)
    bad = int(e.health().sum())

    # ---- roofline of the dominant kernel: instrumented pass over the same step (CUDA events per launch family)
    e.profile(True)
    e.run(kyear)
    e.profile(False)
    fam = {f: e.profile_get(f) for f in ("tstepo_flux", "co", "momentum", "embm", "surflux", "seaice", "biogem")}
    k1 = e.iconst("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    n_wet = int(np.sum(np.clip(K - k1 + 1, 0, None)[k1 <= K]))
    bytes_per_launch = n_wet * (16 * L + 32) * M          # B_tr x members of one launch
    t_ms, t_n = fam["tstepo_flux"]
    c_ms, c_n = fam["co"]
    nstep = e.nyear                                         # tracer steps in the instrumented year
    avg_ms = (t_ms + c_ms) / max(nstep, 1)                  # tstepo = flux + convection kernels of one step (B_tr covers both)
    peak, peak_src = peaks()
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
    working_set = (2 * L + 6) * I * J * K * 8 * e.member_stride
    in_l2 = working_set < 100e6
    kern = {"col": "tstepo = k_tstep_col + k_co_col", "fast": "tstepo = k_tstepo_flux_coop + k_co_fast2 (+ k_sst)",
            "strict": "tstepo = k_tstepo_flux_strict + k_co_strict (+ k_sst)"}[e.tracer_variant_active()]
    roofline = {"bound": "hbm", "kernel": kern + " (%s variant), SURVEY 8d B_tr" % e.tracer_variant_active(),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(M, e.tracer_variant_active()) if args.config == 4 else None,
                "traffic_source": ncu_traffic(M, e.tracer_variant_active(), note=True) if args.config == 4 else None, "peak_source": peak_src,
                "launches_per_step": (t_n + c_n) / max(nstep, 1),
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
                "family_ms_per_year": {k: v[0] for k, v in fam.items()}}
    # the other kernel families of the step against the same roofline (SURVEY 8d algorithmic bytes per unit; serialised,
    # instrumented pass: CUDA events around each family's launches)
    LS = 9
    nbg = e.nyear // 2
    others = {}
    fam_bytes = {"biogem": (n_wet * 8 * (5 * L + 4 * LS) * M, nbg, "BIOGEM / ATCHEM block: k_bg_step (+ k_tc_partial / k_tc_sum / k_tc_factors / "
                            "k_tc_apply, k_bg_climate, k_bg_atchem1/2), N_wet x 8 x (5 L + 4 Ls) B per member-block"),
                 "momentum": ((n_wet * (8 + 24 + 24) + 8 * (I * (J + 1)) * (I + 1 + I + 2)) * M, e.nyear,
                              "velc + jbar + wind + barotropic solve + island: N_wet x 56 B + 8 x 1332 x 75 B of factors per member-step"),
                 "embm": (I * J * 8 * 8 * M, 5 * e.nyear, "k_embm (tstipa + step_embm): 1296 x 8 x 8 B per member and atmosphere step")}
    for f, (nbytes, ncall, what) in fam_bytes.items():
        if not biogem and f == "biogem":
            continue
        t = fam[f][0] / max(ncall, 1)
        if t > 0:
            others[f] = {"bound": "hbm", "kernel": what, "algorithmic_bytes_per_call": nbytes, "avg_call_ms": t,
                         "achieved": nbytes / (t * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": nbytes / (t * 1e-3) / 1e9 / peak}
    roofline["other_families"] = others
    if in_l2:   # single-member runs: the whole state (%.1f MB) lives in the 126 MB L2, so `achieved` is L2, not HBM, bandwidth
        roofline["note"] = ("working set %.1f MB is L2 resident: achieved = L2 GB/s of the tracer step (north_star: 'L2 GB/s for "
                            "single-member runs'); the step is launch / latency bound at this size, not bandwidth bound" % (working_set / 1e6))

    # ---- end to end through the per-module C-ABI entry points with host buffers
    # The members of this GPU run as G independent groups (own library handle, own host thread: cgenie_b200.EnsembleGroups'
    # arrangement).  Every model year each group uploads its year's state from pinned host memory, makes the 480 iterations
    # of module calls the Fortran host makes, and downloads the state again.  With G = 2 one group's copies cross PCIe (both
    # directions) while the other group computes; with G = 1 (--e2e-groups 1, the round-1 form) copies and compute alternate.
    nyear_, variant_active, member_stride = e.nyear, e.tracer_variant_active(), e.member_stride
    # measured (profiles/README_r2.md): two groups are no faster at N = 1 (5.72 vs 5.81 M) nor at N = 8 (35.2 vs 36.0 M), so one
    # group is the default and the exchange moves fewer bytes instead (wet cells only)
    G = args.e2e_groups if args.e2e_groups else 1
    genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / nyear_
    clock_tick = int(round(1000.0 * genie_timestep))
    dts_bg = float(2 * 5) * genie_timestep
    if G == 1:
        parts = [e]
    else:
        e.close()
        parts = []
        for g_ in range(G):
            lo, hi = g_ * (M // G), (g_ + 1) * (M // G)
            parts.append(Ensemble(tmp, n_members=M // G, device=local, perturb={k: np.ascontiguousarray(v[lo:hi]) for k, v in pert.items()}))
            parts[-1].set_tracer_variant(args.variant)

    def in_threads(fn):
        err = []

        def work(q):
            try:
                fn(q)
            except Exception as ex:      # noqa: BLE001
                err.append(ex)
        th = [threading.Thread(target=work, args=(q,)) for q in range(len(parts))]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        if err:
            raise err[0]

    if G > 1:   # the groups reach the bench state (same untimed spin-up as above) side by side
        in_threads(lambda q: (parts[q].run(kyear * (args.spinup_years + 1)), parts[q].synchronize()))
    pins, k0 = [], []
    for p_ in parts:
        size = {n: ((p_.wet_size(n) if (n == "ts" and not args.e2e_dense) else p_.field_size(n)) * p_.member_stride) for n in ("ts", "tq", "varice")}
        try:
            pin = {n: torch.empty(size[n], dtype=torch.float64).pin_memory().numpy() for n in size}
        except Exception:
            pin = {n: np.empty(size[n]) for n in size}
        for n in pin:
            (p_.get_all_wet if (n == "ts" and not args.e2e_dense) else p_.get_all)(n, out=pin[n])
        pins.append(pin)
        k0.append(((args.spinup_years + args.warmup + args.steps + 1) if G == 1 else (args.spinup_years + 1)) * kyear)
    h2d = sum(a.nbytes for pin in pins for a in pin.values())
    d2h = h2d

    def e2e_year(q):
        p_, pin = parts[q], pins[q]
        for n in pin:
            (p_.put_all_wet if (n == "ts" and not args.e2e_dense) else p_.put_all)(n, pin[n])
        p_.put_all("varice1", pin["varice"])
        p_.put_all("tq1", pin["tq"])
        for k in range(1, kyear + 1):
            if k % 5 == 1:
                p_.surflux()
            p_.step_embm()
            if k % 5 == 0:
                p_.step_seaice()
                p_.step_goldstein()
            if biogem and k % 10 == 0:   # conv_kocn_kbiogem = conv_kocn_katchem = 2 (genie.f90:352-447)
                clock = (k0[q] + k) * clock_tick
                p_.biogem_forcing(clock)
                p_.biogem_step(dts_bg, clock)
                p_.biogem_tracercoupling()
                p_.biogem_climate()
                p_.atchem_step(dts_bg)
        k0[q] += kyear
        for n in pin:
            (p_.get_all_wet if (n == "ts" and not args.e2e_dense) else p_.get_all)(n, out=pin[n])

    in_threads(e2e_year)
    barrier()
    t0 = time.perf_counter()
    nrep = max(1, min(args.steps, 3))
    in_threads(lambda q: [e2e_year(q) for _ in range(nrep)])
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * M * nrep / (e2e_s / 3600.0)
    e2e_serial = None
    e2e_path_note = ""
    if G == 1 and not args.e2e_serial:
        # Double-buffered exchange (cg_exchange_*): every model year the GPU takes a NEW batch of members' state from pinned host
        # memory and returns the finished batch's -- an ensemble larger than the GPU's shard, run batch after batch.  The inputs of
        # batch n+1 cross PCIe into a device staging buffer on a copy stream while batch n computes, the results of batch n cross
        # back while batch n+1 computes; committing / packing are device-to-device passes on the compute stream.  Same bytes per
        # year in both directions as the serial form above (kept as e2e.serial_value), same 480 iterations of module calls.
        e2e_serial = e2e_val
        p_ = parts[0]
        wetf = lambda n: (n == "ts" and not args.e2e_dense)
        def pinned_like(a):
            try:
                return torch.empty(a.size, dtype=torch.float64).pin_memory().numpy()
            except Exception:
                return np.empty(a.size)
        bufs_in = [pins[0], {n: pinned_like(a) for n, a in pins[0].items()}]       # two batches' inputs, two batches' results
        bufs_out = [{n: pinned_like(a) for n, a in pins[0].items()} for _ in range(2)]
        for n in bufs_in[1]:
            bufs_in[1][n][:] = pins[0][n]

        def stage_in(year):
            for n, a in bufs_in[year % 2].items():
                p_.exchange_begin_upload(n, a, wet=wetf(n))

        def e2e_year_db(year, last):
            for n in bufs_in[0]:
                p_.exchange_commit_upload(n, also={"varice": "varice1", "tq": "tq1"}.get(n))
            if not last:
                stage_in(year + 1)                 # the next batch starts crossing PCIe now
            for k in range(1, kyear + 1):
                if k % 5 == 1:
                    p_.surflux()
                p_.step_embm()
                if k % 5 == 0:
                    p_.step_seaice()
                    p_.step_goldstein()
                if biogem and k % 10 == 0:
                    clock = (k0[0] + k) * clock_tick
                    p_.biogem_forcing(clock)
                    p_.biogem_step(dts_bg, clock)
                    p_.biogem_tracercoupling()
                    p_.biogem_climate()
                    p_.atchem_step(dts_bg)
            k0[0] += kyear
            for n, a in bufs_out[year % 2].items():
                p_.exchange_begin_download(n, a, wet=wetf(n))

        stage_in(0)
        e2e_year_db(0, False)                      # untimed: fills the pipeline
        p_.exchange_wait()
        p_.synchronize()
        barrier()
        t0 = time.perf_counter()
        for y in range(1, nrep + 1):
            e2e_year_db(y, y == nrep)
        p_.exchange_wait()                         # the last batch's results are on the host
        p_.synchronize()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_val = world * M * nrep / (e2e_s / 3600.0)
        assert all(np.isfinite(a).all() for a in bufs_out[nrep % 2].values())
        e2e_path_note = ("; double buffered (cg_exchange_*): a new batch of members' state every model year, batch n+1 crosses PCIe on a "
                         "copy stream while batch n computes, results of batch n while n+1 computes; serial_value = the same exchange "
                         "without overlap")
    bad += sum(int(p_.health().sum()) for p_ in parts) if G > 1 else 0

    out = {
        "metric": "ensemble model-years/wall-hour", "value": value, "unit": "model-years/hour", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(workload, M, I, J, K, L, nyear_, variant_active, biogem, args.spinup_years, member_stride, args.adrag_group),
        "clocks": clocks, "gpu_launches": launches, "blown_up_members": bad, "roofline": roofline,
        "e2e": {"value": e2e_val, "unit": "model-years/hour", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "groups": G, "years_timed": nrep, "serial_value": e2e_serial,
                "path": "per-module C-ABI calls (surflux/step_embm/step_seaice/step_goldstein/biogem_*/atchem), state in/out of pinned host "
                        "per year; the GPU's members run as %d group(s) of %d (one library handle + host thread each), so one group's copies "
                        "overlap the other's compute" % (G, M // G) if G > 1 else
                        "per-module C-ABI calls (surflux/step_embm/step_seaice/step_goldstein/biogem_*/atchem), state in/out of pinned host per "
                        "year (ts: %s)%s" % ("all cells" if args.e2e_dense else "wet cells only, packed on the device", e2e_path_note)},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == 4:   # the CPU arm is timed next to the N=1 line only
        cores = os.cpu_count() or 1
        rate, wall = cpu_oracle_rate(10.0, cores)
        out["cpu_baseline"] = {"value": rate, "unit": "model-years/hour", "cores": cores, "kind": "port",
                               "sample": "%d oracle processes (one member per host core) x 10 model-years, %.1f s wall" % (cores, wall)}
    for p_ in parts:
        p_.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    os.close(json_fd)


if __name__ == "__main__":
    main()
