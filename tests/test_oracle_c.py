"""The C oracle (oracle/c/sem_oracle.c, the timed CPU baseline) against the golden reference outputs
and the NumPy oracle."""

import numpy as np
import pytest

from conftest import MESH_CASES, load_golden, rel_inf
from oracle import c_oracle as co
from oracle import sem_oracle as so


def _geom9(g):
    return g["geom6"][:, [[0, 1, 2], [1, 3, 4], [2, 4, 5]], :]


@pytest.mark.parametrize("case", MESH_CASES)
def test_c_oracle_matches_reference(case):
    g = load_golden(case)
    N, D, geom = int(g["N"]), g["D"], _geom9(g)
    assert rel_inf(co.elliptic_ax(geom, g["p"], N, D), g["ax_raw"]) < 1e-14
    g2l, gs = g["global_to_local"], g["global_start"]
    assert rel_inf(co.gather_scatter(g["p"], g2l, gs), so.gather_scatter(g["p"], g2l, gs)) < 1e-15
    x, it = co.cg(geom, N, D, g["mask"], g2l, gs, g["rmult"], g["b"], 1e-8, 10000)
    assert abs(it - int(g["cg_niter_1e-08"])) <= 1 and rel_inf(x, g["cg_x_1e-08"]) < 1e-8
    dinv = g["mask"] / so.gather_scatter(g["diag_local"], g2l, gs)
    x, it = co.cg(geom, N, D, g["mask"], g2l, gs, g["rmult"], g["b"], 1e-10, 10000, dinv=dinv)
    assert abs(it - int(g["pcg_niter_1e-10"])) <= 1 and rel_inf(x, g["pcg_x_1e-10"]) < 1e-8
    assert co.num_threads() >= 1
