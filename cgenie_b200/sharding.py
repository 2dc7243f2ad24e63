"""Ensemble sharding across GPUs: members are independent, so each rank owns a contiguous block of
members and there is NO collective on the timestep path.  The only exchanges are the
barrier / max-reduction around timed regions and an optional final gather of per-member diagnostics."""
import numpy as np

SEED = 20261017
BASE = dict(diff1=2000.0, diff2=1.0e-5, adrag=2.5, scf=2.0, diffamp1=5.0e6, diffamp2=1.0e6, betaz2=0.4, betam2=0.4)
# BIOGEM parameters of SURVEY.md 8d; appended so that the physics columns of the table keep their values
BASE_BIOGEM = dict(par_bio_k0_PO4=2.0e-6, par_bio_remin_POC_eL1=500.0, par_bio_red_POC_CaCO3=0.2)
PERTURBED = list(BASE)
PERTURBED_BIOGEM = list(BASE_BIOGEM)
ADRAG_GROUP = 16  # members per barotropic factorisation (adrag is perturbed per group)


def perturbation_table(n_total, seed=SEED, biogem=False, adrag_group=None):
    """Member m scales each whitelisted parameter by U(0.8, 1.25); member 0 is the unperturbed control
    (SURVEY.md 8d).  Deterministic in (n_total, seed); the first n rows do not depend on n_total.
    adrag_group: members per distinct drag value (default ADRAG_GROUP = 16: members with equal adrag share one barotropic
    factorisation on the device; 1 = every member its own, as SURVEY 8d words it)."""
    ag = ADRAG_GROUP if adrag_group is None else int(adrag_group)
    tab = {}
    base = dict(BASE, **BASE_BIOGEM) if biogem else BASE
    for q, k in enumerate(PERTURBED + (PERTURBED_BIOGEM if biogem else [])):
        rng = np.random.default_rng([seed, q])
        f = rng.uniform(0.8, 1.25, size=n_total)
        if k == "adrag":
            f = np.repeat(f[::ag], ag)[:n_total]
            f[:ag] = 1.0
        f[0] = 1.0
        tab[k] = base[k] * f
    return tab


def shard_bounds(rank, world, members_per_rank):
    """Contiguous block [lo, hi) of global member indices owned by `rank` (weak scaling: fixed members per rank)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * members_per_rank, (rank + 1) * members_per_rank


def shard(table, rank, world, members_per_rank):
    lo, hi = shard_bounds(rank, world, members_per_rank)
    return {k: np.ascontiguousarray(v[lo:hi]) for k, v in table.items()}


def max_over_ranks(x, dist=None, device=None):
    """Max of a host scalar over all ranks (timing: the slowest rank defines the step)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_diagnostics(local, dist=None, device=None):
    """Optional end-of-run gather of per-member diagnostics [members_per_rank, n] -> [world*members_per_rank, n]."""
    import torch
    t = torch.as_tensor(np.ascontiguousarray(local), dtype=torch.float64, device=device or "cpu")
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return t.cpu().numpy()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.cat(out, 0).cpu().numpy()
