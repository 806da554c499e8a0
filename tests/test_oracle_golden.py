"""Pin the CPU oracle (oracle/sem_oracle.py) to outputs of the unmodified
reference stored in tests/golden/*.npz (made by tests/golden/make_golden.py)."""

import numpy as np
import pytest

from conftest import MESH_CASES, load_golden, rel_inf
from oracle import sem_oracle as so

UNIT_N = (1, 2, 3, 4, 7, 10, 15)


@pytest.mark.parametrize("N", UNIT_N)
def test_one_d_inputs(units, N):
    z, w = so.gauss_lobatto(N)
    # eig-based nodes: same LAPACK on the same image, but allow last-ulp noise across CPUs
    assert np.allclose(z, units[f"z_{N}"], rtol=0, atol=4e-15)
    assert np.allclose(w, units[f"w_{N}"], rtol=1e-13, atol=0)
    D = so.lagrange_derivative_matrix(units[f"z_{N}"])
    assert np.array_equal(D, units[f"D_{N}"])
    assert np.allclose(so.reference_mass_matrix_3d(N), units[f"B_{N}"], rtol=1e-13)
    z1, _ = so.gauss_lobatto(1)
    assert np.array_equal(so.lagrange(units[f"z_{N}"], z1), units[f"J_{N}"])


@pytest.mark.parametrize("N", UNIT_N)
def test_gradient_and_transpose(units, N):
    n = N + 1
    D = units[f"D_{N}"]
    got = np.stack(so.gradient(units[f"grad_U_{N}"], n, D))
    assert np.array_equal(got, units[f"grad_out_{N}"])
    W = units[f"gradT_W_{N}"]
    assert np.array_equal(so.gradient_transpose(W[0], W[1], W[2], n, D), units[f"gradT_out_{N}"])


def test_kron(units):
    got = so.kron(units["kron_Sz"], units["kron_Sy"], units["kron_Sx"], units["kron_U"])
    assert np.array_equal(got, units["kron_out"])


def test_generic_cg_pcg(units):
    A, b = units["it_A"], units["it_b"]
    x, it = so.cg(lambda v: A @ v, b, tol=1e-12, maxit=100)
    assert it == int(units["it_cg_niter"]) and np.array_equal(x, units["it_cg_x"])
    dA = np.diag(A).copy()
    x, it = so.pcg(lambda v: A @ v, lambda r: r / dA, b, tol=1e-10, maxit=100)
    assert it == int(units["it_pcg_niter"]) and np.array_equal(x, units["it_pcg_x"])


def test_cg_known_answers():
    # reference tests/cg_test.py:6-34
    A = np.array([[4.0, 1.0], [1.0, 3.0]])
    x, it = so.cg(lambda v: A @ v, np.array([1.0, 2.0]))
    assert it <= 2 and (np.round(x, 4) == np.array([0.0909, 0.6364])).all()
    A = np.array([1.0, 2.0, -1.0, 2.0, 2.0, 2.0, 1.0, -1.0, 2.0]).reshape(3, 3)
    A = (A + A.T) / 2.0
    b = np.array([2.0, 12.0, 5.0])
    x, it = so.cg(lambda v: A @ v, b)
    assert it <= 3 and np.allclose(x, np.linalg.solve(A, b), atol=1e-8)


@pytest.mark.parametrize("case", MESH_CASES)
def test_mesh_setup(case):
    g = load_golden(case)
    N = int(g["N"])
    vx, vy, vz = so.element_vertices(g["points"], g["hexes"])
    assert np.array_equal(vx, g["vx"]) and np.array_equal(vz, g["vz"])
    xe, ye, ze = so.physical_coordinates(vx, vy, vz, N)
    for got, ref in ((xe, g["xe"]), (ye, g["ye"]), (ze, g["ze"])):
        assert np.allclose(got, ref, rtol=0, atol=1e-14)
    # numbering: bit-exact when fed the reference's own coordinates
    for fn in (so.global_numbering_literal, so.global_numbering):
        gid, g2l, gs = fn(g["xe"], g["ye"], g["ze"])
        assert gid.dtype == np.int32 and np.array_equal(gid, g["global_id"])
        assert np.array_equal(g2l, g["global_to_local"])
        assert np.array_equal(gs, g["global_start"])
    assert np.array_equal(so.multiplicity_inverse(g["global_to_local"], g["global_start"]), g["rmult"])
    assert np.array_equal(so.dirichlet_mask(g["xe"], g["ye"], g["ze"]), g["mask"])
    geom, jaco, mass = so.geometric_factors(g["xe"], g["ye"], g["ze"], N)
    geom6 = np.stack([geom[:, 0, 0], geom[:, 0, 1], geom[:, 0, 2], geom[:, 1, 1], geom[:, 1, 2], geom[:, 2, 2]], axis=1)
    scale = np.abs(g["geom6"]).max()
    assert np.max(np.abs(geom6 - g["geom6"])) <= 1e-13 * scale
    assert np.allclose(jaco, g["jaco"], rtol=1e-12)
    assert np.allclose(mass[0], g["mass0"], rtol=1e-13)


def _geom9(g):
    g6 = g["geom6"]
    E, _, Np = g6.shape
    idx = [[0, 1, 2], [1, 3, 4], [2, 4, 5]]
    return np.stack([np.stack([g6[:, idx[a][b]] for b in range(3)], axis=1) for a in range(3)], axis=1)


@pytest.mark.parametrize("case", MESH_CASES)
def test_ax_and_gather_scatter(case):
    g = load_golden(case)
    N, D, geom = int(g["N"]), g["D"], _geom9(g)
    assert np.array_equal(so.elliptic_ax(geom, g["p"], N, D), g["ax_raw"])
    assert rel_inf(so.elliptic_ax_batched(geom, g["p"], N, D), g["ax_raw"]) < 1e-14
    g2l, gs = g["global_to_local"], g["global_start"]
    pc = so.gather_scatter(g["p"], g2l, gs) * g["rmult"]
    assert rel_inf(pc, g["p_cont"]) < 1e-15
    assert rel_inf(so.gather_scatter_literal(g["p"], g2l, gs), so.gather_scatter(g["p"], g2l, gs)) < 1e-15
    full = so.gather_scatter(so.elliptic_ax(geom, g["p_cont"], N, D) * g["mask"], g2l, gs)
    assert rel_inf(full, g["ax_full"]) < 1e-15
    # closed-form Jacobi diagonal vs the operator itself
    diag = so.jacobi_diagonal(geom, N, D)
    assert np.allclose(diag, g["diag_local"], rtol=1e-13)
    rng = np.random.default_rng(0)
    for q in rng.integers(0, g["p"].size, size=3):
        e = np.zeros(g["p"].size)
        e[q] = 1.0
        assert abs(so.elliptic_ax(geom, e, N, D)[q] - diag[q]) <= 1e-12 * abs(diag[q])


@pytest.mark.parametrize("case", MESH_CASES)
def test_cg_and_pcg(case):
    g = load_golden(case)
    N, D, geom = int(g["N"]), g["D"], _geom9(g)
    g2l, gs, rmult, mask = g["global_to_local"], g["global_start"], g["rmult"], g["mask"]
    ax = lambda p: so.elliptic_ax(geom, p, N, D)
    for tol in ("1e-08", "1e-10"):
        x, it = so.elliptic_cg(ax, mask, g2l, gs, rmult, g["b"], tol=float(tol), maxit=10000)
        assert it == int(g["cg_niter_" + tol])
        assert rel_inf(x, g["cg_x_" + tol]) < 1e-13
    diag = so.gather_scatter(so.jacobi_diagonal(geom, N, D), g2l, gs)
    dinv = mask / diag
    for tol in ("1e-08", "1e-10"):
        x, it = so.elliptic_pcg_jacobi(ax, dinv, mask, g2l, gs, rmult, g["b"], tol=float(tol), maxit=10000)
        assert abs(it - int(g["pcg_niter_" + tol])) <= 1
        if tol == "1e-10":
            assert rel_inf(x, g["pcg_x_" + tol]) < 1e-9


@pytest.mark.parametrize("N", [3, 7])
def test_fdm_restatement_inverts_the_interior_blocks(N):
    """The reference's FDM sketch (example.py:71-114,155-171,186-190) cannot be run (example.py is broken: undefined
    names at :33,36,49), so its restatement is pinned to what it must do: on straight box elements the interior
    block of the operator is inverted exactly (A M^-1 r = r on the interior nodes), and `fast_kron` equals the
    dense Kronecker product it abbreviates."""
    rng = np.random.default_rng(3)
    Sx, Sy, Sz = rng.standard_normal((3, 4)), rng.standard_normal((2, 5)), rng.standard_normal((4, 3))
    U = rng.standard_normal(3 * 5 * 4)
    assert np.allclose(so.fast_kron(Sz, Sy, Sx, U), np.kron(Sz, np.kron(Sy, Sx)) @ U, rtol=1e-13, atol=1e-13)
    pts, hexes = so.box_mesh(3, 2, 2)
    vx, vy, vz = so.element_vertices(pts, hexes)
    xe, ye, ze = so.physical_coordinates(vx, vy, vz, N)
    geom, _, _ = so.geometric_factors(xe, ye, ze, N)
    D = so.reference_derivative_matrix(N)
    S, lam = so.fdm_1d(N)
    esc = so.fdm_element_scales(vx, vy, vz)
    E, n = len(hexes), N + 1
    r = np.zeros((E, n, n, n))
    r[:, 1:-1, 1:-1, 1:-1] = rng.standard_normal((E, n - 2, n - 2, n - 2))
    z = so.fdm_apply(r.reshape(-1), N, S, lam, esc, np.zeros(r.size))
    Az = so.elliptic_ax_batched(geom, z, N, D).reshape(E, n, n, n)
    assert np.abs(Az[:, 1:-1, 1:-1, 1:-1] - r[:, 1:-1, 1:-1, 1:-1]).max() < 1e-11
