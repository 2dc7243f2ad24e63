"""More than 128 members in ONE handle (member stride 256 / 512: the fixed-shape kernels run over 128-member tiles, the flux kernel
stages its rows through 2-D tensor maps) -- run with -m gpu on a B200.

Members are independent and every kernel does the same arithmetic per member whatever the stride, so member m of a 256-member
handle must be BIT-IDENTICAL to member m % 128 of a 128-member handle holding members [128 * (m // 128), ...) of the same
perturbation table -- through the production path (col variant, BIOGEM / ATCHEM on, cg_run's concurrent schedule)."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise
from cgenie_b200.sharding import perturbation_table
from test_gpu_biogem import CFG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [256, 300])
def test_member_tiles_bit_identical_to_128_member_handles(built, tmp_path, M):
    materialise(str(tmp_path), CFG)
    tab = perturbation_table(M, biogem=True)
    nk = 5 * 24     # a quarter of a model year: 24 ocean steps, 12 BIOGEM / ATCHEM blocks
    names = ("ts", "u", "tq", "varice", "ocn", "bio_part", "atm", "carbH")
    lanes = {0: [0, 77, 127], 1: [0, 31, 127 if M >= 256 else 0], 2: [0, M - 257]}
    with Ensemble(str(tmp_path), n_members=M, perturb=tab) as big:
        assert big.member_stride == (256 if M <= 256 else 512)
        big.set_tracer_variant("col")
        assert big.tracer_variant_active() == "col"
        big.run(nk)
        assert int(big.health().sum()) == 0
        got = {}
        for g in range((M + 127) // 128):
            for lane in lanes[g]:
                if 128 * g + lane < M:
                    got[(g, lane)] = {n: big.get(n, 128 * g + lane) for n in names}
    for g in range((M + 127) // 128):
        lo, hi = 128 * g, min(M, 128 * (g + 1))
        with Ensemble(str(tmp_path), n_members=hi - lo, perturb={k: np.ascontiguousarray(v[lo:hi]) for k, v in tab.items()}) as e:
            e.set_tracer_variant("col")
            e.run(nk)
            for lane in lanes[g]:
                if (g, lane) in got:
                    for n in names:
                        assert np.array_equal(e.get(n, lane), got[(g, lane)][n]), (M, g, lane, n)


def test_col_per_cell_on_a_512_member_handle(built, tmp_path):
    """the per-cell proof of tests/test_gpu_col_proof.py on a 512-member handle (tile form of the flux kernel against 'strict')"""
    from col_check import check_step
    materialise(str(tmp_path), CFG)
    M = 512
    tab = perturbation_table(M, biogem=True)
    with Ensemble(str(tmp_path), n_members=M, perturb=tab) as e:
        e.set_tracer_variant("col")
        e.run(5 * 96)
        st = check_step(e, tol=1e-10, what="512 members, ocean step 96:")
        assert st["convecting_columns"] > 0
        assert int(e.health().sum()) == 0
