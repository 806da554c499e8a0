"""GPU parity tests: the CUDA path (through the C ABI / ctypes) against the
golden outputs of the unmodified reference and against the CPU oracle on
seeded inputs.  Tolerances are the ones BASELINE.json states: numbering
bit-exact, Ax relative inf-norm <= 1e-12, CG iteration count +-1, solution
<= 1e-10 relative."""

import numpy as np
import pytest

from conftest import MESH_CASES, load_golden, rel_inf
from oracle import sem_oracle as so

pytestmark = pytest.mark.gpu

AX_TOL = 1e-12
UNIT_N = (1, 2, 3, 4, 7, 10, 15)


def golden_mesh(g, own_geometry=False):
    """Product Mesh carrying the reference's own arrays (same D, G, numbering
    and mask feed both paths) or, with own_geometry, built by the product."""
    from sempy_b200.mesh import Mesh

    m = Mesh.from_arrays(g["points"], g["hexes"])
    N = int(g["N"])
    m.find_physical_coordinates(N)
    if not own_geometry:
        m.xe, m.ye, m.ze = g["xe"].copy(), g["ye"].copy(), g["ze"].copy()
    m.establish_global_numbering()
    m.calc_geometric_factors()
    if not own_geometry:
        m.geom6 = g["geom6"].copy()
    m.setup_mask()
    return m


@pytest.mark.parametrize("N", UNIT_N)
def test_gradient_kernels(units, N):
    from sempy_b200.gradient import gradient, gradient_transpose

    n = N + 1
    got = np.stack(gradient(units[f"grad_U_{N}"], n, units[f"D_{N}"]))
    assert rel_inf(got, units[f"grad_out_{N}"]) < 1e-13
    W = units[f"gradT_W_{N}"]
    assert rel_inf(gradient_transpose(W[0], W[1], W[2], n, units[f"D_{N}"]), units[f"gradT_out_{N}"]) < 1e-13
    # default D (rebuilt from the order, as the reference does) and batched input
    got = np.stack(gradient(units[f"grad_U_{N}"], n))
    assert rel_inf(got, units[f"grad_out_{N}"]) < 1e-11
    batch = np.stack([units[f"grad_U_{N}"], 2.0 * units[f"grad_U_{N}"]])
    ur, us, ut = gradient(batch, n, units[f"D_{N}"])
    assert ur.shape == batch.shape and rel_inf(ur[1], 2.0 * units[f"grad_out_{N}"][0]) < 1e-13


def test_kron_kernel(units):
    from sempy_b200.kron import kron, kron_2d

    Sz, Sy, Sx, U = units["kron_Sz"], units["kron_Sy"], units["kron_Sx"], units["kron_U"]
    assert rel_inf(kron(Sz, Sy, Sx, U), units["kron_out"]) < 1e-14
    assert rel_inf(kron(Sz, Sy, Sx, U, order="C"), units["kron_out"]) < 1e-14
    ref_f = np.einsum("ai,bj,ijk,ck->abc", Sz, Sy, U.reshape((2, 3, 4), order="F"), Sx).reshape(-1, order="F")
    assert rel_inf(kron(Sz, Sy, Sx, U, order="F"), ref_f) < 1e-14
    U2 = U[:12]
    assert rel_inf(kron_2d(Sy, Sx, U2), (Sy @ U2.reshape(3, 4) @ Sx.T).reshape(-1)) < 1e-14


def test_generic_cg_known_answers(units):
    """reference tests/cg_test.py:6-34 and a fixed SPD system, through the CUDA vector kernels."""
    from sempy_b200.iterative import cg, pcg

    A = np.array([[4.0, 1.0], [1.0, 3.0]])
    x, it = cg(lambda v: A @ v, np.array([1.0, 2.0]))
    assert it <= 2 and (np.round(x, 4) == np.array([0.0909, 0.6364])).all()
    A3 = np.array([1.0, 2.0, -1.0, 2.0, 2.0, 2.0, 1.0, -1.0, 2.0]).reshape(3, 3)
    A3 = (A3 + A3.T) / 2.0
    b3 = np.array([2.0, 12.0, 5.0])
    x, it = cg(lambda v: A3 @ v, b3)
    assert it <= 3 and np.allclose(x, np.linalg.solve(A3, b3), atol=1e-8)
    A, b = units["it_A"], units["it_b"]
    x, it = cg(lambda v: A @ v, b, tol=1e-12, maxit=100)
    assert abs(it - int(units["it_cg_niter"])) <= 1 and rel_inf(x, units["it_cg_x"]) < 1e-10
    dA = np.diag(A).copy()
    x, it = pcg(lambda v: A @ v, lambda r: r / dA, b, tol=1e-10, maxit=100)
    assert abs(it - int(units["it_pcg_niter"])) <= 1 and rel_inf(x, units["it_pcg_x"]) < 1e-9
    x, it = cg(lambda v: A @ v, np.zeros(12))  # early-out, iterative.py:15
    assert it == 0 and not x.any()
    with pytest.raises(AssertionError):
        pcg(lambda v: v, lambda v: v, np.zeros((2, 2)))
    with pytest.raises(ZeroDivisionError):
        cg(lambda v: 1 / 0, np.ones(3))


@pytest.mark.parametrize("case", MESH_CASES)
def test_setup_matches_reference(case):
    g = load_golden(case)
    m = golden_mesh(g, own_geometry=True)
    for got, ref in ((m.xe, g["xe"]), (m.ye, g["ye"]), (m.ze, g["ze"])):
        assert np.max(np.abs(got - ref)) < 1e-14
    assert np.max(np.abs(m.geom6 - g["geom6"])) <= 1e-12 * np.abs(g["geom6"]).max()
    assert np.allclose(m.get_jaco(), g["jaco"].reshape(-1), rtol=1e-11)
    assert m.get_geom().shape == (g["hexes"].shape[0], 3, 3, m.Np)
    assert np.allclose(m.get_mass()[: m.Np], g["mass0"], rtol=1e-13)
    assert m.global_id.max() == g["global_id"].max()
    assert np.array_equal(np.sort(m.get_rmult()), np.sort(g["rmult"]))


@pytest.mark.parametrize("case", MESH_CASES)
def test_ax_gs_mask_vs_reference(case):
    from sempy_b200.elliptic import elliptic_ax

    g = load_golden(case)
    m = golden_mesh(g)
    assert np.array_equal(m.global_id, g["global_id"]) and np.array_equal(m.get_rmult(), g["rmult"])
    assert np.array_equal(m.get_mask_ids(), g["mask_ids"])
    ax = elliptic_ax(m, g["p"])
    assert ax is not g["p"] and rel_inf(ax, g["ax_raw"]) <= AX_TOL
    x = g["p"].copy()
    assert m.dssum(x) is x  # in place, returns its argument (mesh.py:444-447)
    assert rel_inf(x * m.get_rmult(), g["p_cont"]) < 1e-15
    full = m.dssum(m.apply_mask(elliptic_ax(m, g["p_cont"])))
    assert rel_inf(full, g["ax_full"]) <= AX_TOL
    y = g["p"].copy()
    m.apply_mask(y)
    assert np.array_equal(y, g["p"] * g["mask"])


@pytest.mark.parametrize("case", MESH_CASES)
def test_cg_vs_reference(case):
    from sempy_b200.elliptic import elliptic_cg, elliptic_pcg

    g = load_golden(case)
    m = golden_mesh(g)
    for tol in ("1e-08", "1e-10"):
        x, it = elliptic_cg(m, g["b"], tol=float(tol), maxit=10000)
        assert abs(it - int(g["cg_niter_" + tol])) <= 1, (it, int(g["cg_niter_" + tol]))
        # two exact-arithmetic-equivalent CG runs stopped at tol already differ by a fraction of tol
        # in x, more if they stop one iteration apart; the 1e-10 solution
        # bar is checked on the tight solve below
        assert rel_inf(x, g["cg_x_" + tol]) < float(tol)
    x, it = elliptic_pcg(m, g["b"], tol=1e-10, maxit=10000)
    assert abs(it - int(g["pcg_niter_1e-10"])) <= 1
    assert rel_inf(x, g["pcg_x_1e-10"]) < 1e-8
    # tight solve: the solution itself to 1e-10 against the oracle at the same tolerance (tol = 1e-10 is
    # where the +-1 iteration bar is meaningful: at 1e-12 the recursive residual of high-order cases sits
    # at the FP64 floor and equally valid summation orders stop up to 2 iterations apart)
    geom = so_geom9(g)
    axo = lambda p: so.elliptic_ax_batched(geom, p, int(g["N"]), g["D"])
    xo, ito = so.elliptic_cg(axo, g["mask"], g["global_to_local"], g["global_start"], g["rmult"], g["b"], tol=1e-10, maxit=5000)
    x, it = elliptic_cg(m, g["b"], tol=1e-10, maxit=5000)
    assert abs(it - ito) <= 1
    if it != ito:  # stopped one iteration apart: compare the iterates after the same number of steps
        x, it = elliptic_cg(m, g["b"], tol=0.0, maxit=ito)
    assert rel_inf(x, xo) < 1e-10


def so_geom9(g):
    g6 = g["geom6"]
    idx = [[0, 1, 2], [1, 3, 4], [2, 4, 5]]
    return g6[:, idx, :]


def test_cg_edge_cases_and_verbose(capsys):
    from sempy_b200.elliptic import elliptic_cg, elliptic_cg_loopy

    g = load_golden("box008_N7")
    m = golden_mesh(g)
    x, it = elliptic_cg(m, np.zeros_like(g["b"]))  # early-out (elliptic.py:43)
    assert it == 0 and not x.any()
    x, it = elliptic_cg(m, g["b"], tol=1e-8, maxit=0)
    assert it == 0 and not x.any()
    x, it = elliptic_cg(m, g["b"], tol=1e-8, maxit=5)
    assert it == 5
    x, it = elliptic_cg(m, g["b"], tol=1e-8, maxit=3, verbose=1)
    out = capsys.readouterr().out.splitlines()
    assert out[0].startswith("Initial rnorm=") and len(out) == 4 and out[1].startswith("niter=0 r0=") and " pap=" in out[3]
    # the printed scalars (elliptic.py:63-68) are the reference's: r0, r1, alpha, beta, pAp per iteration
    hist = []
    geom = so_geom9(g)
    so.elliptic_cg(lambda v: so.elliptic_ax(geom, v, int(g["N"]), g["D"]), g["mask"], g["global_to_local"],
                   g["global_start"], g["rmult"], g["b"], tol=1e-8, maxit=3, history=hist)
    for line, ref in zip(out[1:], hist):
        vals = [float(tok.split("=")[1]) for tok in line.split()[1:]]
        assert np.allclose(vals, ref, rtol=1e-11, atol=0)
    assert abs(float(out[0].split("=")[1]) - hist[0][0]) <= 1e-12 * hist[0][0]
    xl, itl = elliptic_cg_loopy(m, g["b"], tol=1e-8, maxit=10000)
    assert itl == int(g["cg_niter_1e-08"]) or abs(itl - int(g["cg_niter_1e-08"])) <= 1
    b = g["b"].copy()
    elliptic_cg(m, b, tol=1e-8, maxit=10)
    assert np.array_equal(b, g["b"])  # inputs are not mutated
    with pytest.raises(ValueError):
        elliptic_cg(m, g["b"][:-1])


def test_device_tensors_and_operator():
    torch = pytest.importorskip("torch")
    from sempy_b200.elliptic import elliptic_ax, elliptic_cg
    from sempy_b200.iterative import cg

    g = load_golden("box008_N7")
    m = golden_mesh(g)
    p = torch.from_numpy(g["p"]).cuda()
    ax = elliptic_ax(m, p)
    assert ax.is_cuda and rel_inf(ax.cpu().numpy(), g["ax_raw"]) <= AX_TOL
    x = torch.from_numpy(g["p"]).cuda()
    m.dssum(x)
    assert rel_inf(x.cpu().numpy() * g["rmult"], g["p_cont"]) < 1e-15
    xb, it = elliptic_cg(m, torch.from_numpy(g["b"]).cuda(), tol=1e-8, maxit=10000)
    assert xb.is_cuda and abs(it - int(g["cg_niter_1e-08"])) <= 1
    # iterative.cg with the device operator: assembled, masked A on a continuous vector
    A = m.operator(mask=True, assemble=True)
    xg, itg = cg(A, g["b"], tol=1e-10, maxit=10000)
    # oracle: the same generic CG (sempy/iterative.py:4-42, plain dots) on the assembled, masked operator
    geom = so_geom9(g)
    A_o = lambda v: so.gather_scatter(so.elliptic_ax_batched(geom, v, int(g["N"]), g["D"]) * g["mask"],
                                      g["global_to_local"], g["global_start"])
    xo, ito = so.cg(A_o, g["b"], tol=1e-10, maxit=10000)
    assert abs(itg - ito) <= 1 and rel_inf(xg, xo) < 1e-9


@pytest.mark.parametrize("dims,perturb,N", [((6, 6, 6), 0.0, 7), ((5, 4, 3), 0.3, 7), ((3, 3, 3), 0.3, 15), ((4, 3, 2), 0.3, 10), ((3, 2, 2), 0.25, 1), ((7, 5, 3), 0.2, 2), ((2, 2, 2), 0.3, 5),
                                            ((3, 2, 2), 0.3, 8), ((2, 3, 2), 0.3, 9), ((2, 2, 3), 0.3, 11), ((2, 2, 2), 0.3, 12), ((2, 2, 2), 0.3, 13), ((2, 2, 2), 0.3, 14),
                                            ((5, 4, 3), 0.3, 6), ((7, 6, 5), 0.2, 3), ((6, 5, 4), 0.3, 4), ((9, 7, 5), 0.2, 2)])
def test_generated_meshes_vs_oracle(dims, perturb, N):
    """Seeded inputs on generated boxes (straight and deformed, every kernel
    family: Nq=8 tuned, Nq=16, generic) against the CPU oracle."""
    from sempy_b200.elliptic import elliptic_ax, elliptic_cg, elliptic_pcg
    from sempy_b200.mesh import box_mesh

    m = box_mesh(*dims, perturb=perturb, seed=1)
    m.find_physical_coordinates(N)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    E = m.get_num_elems()
    D = np.ascontiguousarray(__import__("sempy_b200.derivative", fromlist=["x"]).reference_derivative_matrix(N))
    geom = m.get_geom()
    rng = np.random.default_rng(5)
    p = rng.standard_normal(E * m.Np)
    ax_o = so.elliptic_ax_batched(geom, p, N, D)
    assert rel_inf(elliptic_ax(m, p), ax_o) <= AX_TOL
    g2l, gs = m.get_global_to_local_map()
    x = p.copy()
    m.dssum(x)
    assert rel_inf(x, so.gather_scatter(p, g2l, gs)) < 1e-15
    assert np.array_equal(m.get_rmult(), so.multiplicity_inverse(g2l, gs))
    X, Y, Z = m.get_x(), m.get_y(), m.get_z()
    b = 3 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z) * m.get_mass() * m.get_jaco()
    b = m.apply_mask(m.dssum(b))
    axo = lambda v: so.elliptic_ax_batched(geom, v, N, D)
    xo, ito = so.elliptic_cg(axo, m.mask, g2l, gs, m.get_rmult(), b, tol=1e-10, maxit=2000)
    xg, itg = elliptic_cg(m, b, tol=1e-10, maxit=2000)
    assert abs(itg - ito) <= 1
    if itg != ito:
        xg, itg = elliptic_cg(m, b, tol=0.0, maxit=ito)
    assert rel_inf(xg, xo) < 1e-10
    dinv = m.mask / so.gather_scatter(so.jacobi_diagonal(geom, N, D), g2l, gs)
    xo, ito = so.elliptic_pcg_jacobi(axo, dinv, m.mask, g2l, gs, m.get_rmult(), b, tol=1e-10, maxit=2000)
    xg, itg = elliptic_pcg(m, b, tol=1e-10, maxit=2000)
    assert abs(itg - ito) <= 1
    if itg != ito:
        xg, itg = elliptic_pcg(m, b, tol=0.0, maxit=ito)
    assert rel_inf(xg, xo) < 1e-10


@pytest.mark.parametrize("case", MESH_CASES)
def test_device_setup_matches_reference(case):
    """semb_numbering is bit-exact against the reference's numbering on the reference's coordinates;
    semb_geom_factors matches the reference's geometric factors to 1e-12."""
    from sempy_b200 import _capi
    from sempy_b200 import mesh as pm
    from sempy_b200.quadrature import gauss_lobatto
    import ctypes as C

    g = load_golden(case)
    gid, g2l, gs = pm.global_numbering_device(g["xe"], g["ye"], g["ze"])
    assert gid.dtype == np.int32 and np.array_equal(gid, g["global_id"])
    assert np.array_equal(g2l, g["global_to_local"]) and np.array_equal(gs, g["global_start"])
    N = int(g["N"])
    E, Np = g["xe"].shape
    _, w = gauss_lobatto(N)
    g6, J = np.empty((E, 6, Np)), np.empty((E, Np))
    D = np.ascontiguousarray(g["D"])
    _capi.check(_capi.lib().semb_geom_factors(0, N + 1, D.ctypes.data, w.ctypes.data, E, g["xe"].ctypes.data,
                                              g["ye"].ctypes.data, g["ze"].ctypes.data, g6.ctypes.data, J.ctypes.data,
                                              _capi.HOST))
    assert np.max(np.abs(g6 - g["geom6"])) <= 1e-12 * np.abs(g["geom6"]).max()
    assert np.allclose(J, g["jaco"], rtol=1e-11)


def test_device_setup_large_and_signed_zero(monkeypatch):
    """A generated 12^3, N=7 box (884 736 nodes) numbers identically on host and device, including
    coordinates that are -0.0 on one copy and +0.0 on another; the whole set-up also runs with
    SEMB_DEVICE_SETUP=1 and solves the same problem."""
    from sempy_b200 import mesh as pm
    from sempy_b200.elliptic import elliptic_pcg

    m = pm.box_mesh(12, perturb=0.3, seed=3)
    m.find_physical_coordinates(7)
    xe, ye, ze = m.xe.copy(), m.ye.copy(), m.ze.copy()
    xe[xe == 0.0] = np.where(np.arange((xe == 0.0).sum()) % 2 == 0, 0.0, -0.0)
    h = pm.global_numbering_from_coordinates(xe, ye, ze)
    d = pm.global_numbering_device(xe, ye, ze)
    for a, b in zip(h, d):
        assert np.array_equal(a, b)
    sols = []
    for flag in ("0", "1"):
        monkeypatch.setenv("SEMB_DEVICE_SETUP", flag)
        mm = pm.box_mesh(6, perturb=0.3, seed=3)
        mm.find_physical_coordinates(7)
        mm.establish_global_numbering()
        mm.calc_geometric_factors()
        mm.setup_mask()
        X, Y, Z = mm.get_x(), mm.get_y(), mm.get_z()
        b = 3 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z) * mm.get_mass() * mm.get_jaco()
        b = mm.apply_mask(mm.dssum(b))
        x, it = elliptic_pcg(mm, b, tol=1e-10, maxit=2000)
        sols.append((x, it, mm.global_id.copy(), mm.geom6.copy()))
    assert np.array_equal(sols[0][2], sols[1][2])
    assert np.max(np.abs(sols[0][3] - sols[1][3])) <= 1e-12 * np.abs(sols[0][3]).max()
    assert abs(sols[0][1] - sols[1][1]) <= 1 and rel_inf(sols[1][0], sols[0][0]) < 1e-9


@pytest.fixture(scope="module")
def big_mesh():
    """BASELINE config 2 size: 32^3 elements, N=7 (16.8 M local nodes)."""
    from sempy_b200.mesh import box_mesh

    m = box_mesh(32, perturb=0.3, seed=0)
    m.find_physical_coordinates(7)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    return m


def test_full_size_properties(big_mesh):
    """Size-independent properties at 32^3 N=7 on device tensors: unique node
    count, symmetry, null space, gather-scatter idempotence, partition of
    unity of rmult, manufactured solution."""
    torch = pytest.importorskip("torch")
    from sempy_b200.elliptic import elliptic_ax, elliptic_pcg

    m = big_mesh
    n = m.get_num_elems() * m.Np
    assert m.global_start.size - 1 == 225 ** 3
    gen = torch.Generator(device="cuda").manual_seed(3)
    u = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    v = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    au, av = elliptic_ax(m, u), elliptic_ax(m, v)
    s1, s2 = torch.dot(v, au).item(), torch.dot(u, av).item()
    assert abs(s1 - s2) <= 1e-12 * max(abs(s1), abs(s2))  # A = A^T element by element
    assert torch.dot(u, au).item() > 0
    lin = elliptic_ax(m, 2.0 * u - 3.0 * v)
    assert (lin - (2.0 * au - 3.0 * av)).abs().max().item() <= 1e-12 * au.abs().max().item()
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    assert elliptic_ax(m, ones).abs().max().item() <= 1e-11 * au.abs().max().item()  # constants in the null space
    rm = torch.from_numpy(m.get_rmult()).cuda()
    w = u.clone()
    m.dssum(w)
    w2 = (w * rm).clone()
    m.dssum(w2)
    assert (w2 - w).abs().max().item() <= 1e-14 * w.abs().max().item()  # dssum(rmult * dssum(x)) = dssum(x)
    assert abs(rm.sum().item() - 225 ** 3) < 1e-6  # sum of 1/multiplicity = number of unique nodes
    X, Y, Z = (torch.from_numpy(a).cuda() for a in (m.get_x(), m.get_y(), m.get_z()))
    exact = torch.sin(np.pi * X) * torch.sin(np.pi * Y) * torch.sin(np.pi * Z)
    b = 3 * np.pi ** 2 * exact * torch.from_numpy(m.get_mass() * m.get_jaco()).cuda()
    m.dssum(b)
    m.apply_mask(b)
    x, it = elliptic_pcg(m, b, tol=1e-10, maxit=2000)
    assert 0 < it < 2000
    assert (x - exact).abs().max().item() < 1e-6  # spectral accuracy on the deformed mesh
    # result is continuous and zero on the boundary
    xc = x.clone()
    m.dssum(xc)
    assert (xc * rm - x).abs().max().item() < 1e-12
    xm = x.clone()
    m.apply_mask(xm)
    assert torch.equal(xm, x)


def test_bitwise_repeatable_and_independent_of_the_gs_form():
    """Every sum of the solver has a fixed order (no floating-point atomics): two solves give the same
    bits, and so does the solve with the pairwise gather-scatter as a separate in-place pass
    (gs_pull = 0) instead of inside the update kernel (a + b commutes exactly), and the solve whose kernels are
    launched one after the other (pdl = 0) instead of chained by programmatic dependent launch."""
    from sempy_b200 import _capi
    from sempy_b200.elliptic import elliptic_cg, elliptic_pcg
    from sempy_b200.mesh import box_mesh

    m = box_mesh(6, 5, 4, perturb=0.3, seed=7)
    m.find_physical_coordinates(7)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    rng = np.random.default_rng(2)
    b = m.apply_mask(m.dssum(rng.standard_normal(m.get_num_elems() * m.Np)))
    lib = _capi.lib()
    runs = []
    try:
        # (gs_pull, pdl): the default twice, the in-place gather-scatter, plain instead of programmatic launches
        for pull, pdl in ((1, 1), (1, 1), (0, 1), (1, 0)):
            _capi.check(lib.semb_set_tuning(b"gs_pull", pull))
            _capi.check(lib.semb_set_tuning(b"pdl", pdl))
            runs.append((elliptic_pcg(m, b, tol=1e-10, maxit=500), elliptic_cg(m, b, tol=1e-9, maxit=40)))
    finally:
        _capi.check(lib.semb_set_tuning(b"gs_pull", 1))
        _capi.check(lib.semb_set_tuning(b"pdl", 1))
    for (xp, itp), (xc, itc) in runs[1:]:
        assert itp == runs[0][0][1] and itc == runs[0][1][1]
        assert np.array_equal(xp, runs[0][0][0]) and np.array_equal(xc, runs[0][1][0])
    x1 = m.dssum(b.copy())
    x2 = m.dssum(b.copy())
    assert np.array_equal(x1, x2)


def test_gslib_handle():
    """GS(0).setup(ids) / .gs(x, gs_double, gs_add) (sempy/mesh.py:437-445) on arbitrary ids, id 0 = not shared."""
    torch = pytest.importorskip("torch")
    from sempy_b200.gslib_wrapper import GS, gs_add, gs_double

    rng = np.random.default_rng(8)
    n = 50000
    ids = rng.integers(0, 9000, n).astype(np.int32)  # multiplicities from 1 to ~15, some zeros
    x = rng.standard_normal(n)
    ref = x.copy()
    for k in np.unique(ids[ids > 0]):
        sel = ids == k
        ref[sel] = x[sel].sum()
    gs = GS(0)
    gs.setup(ids)
    y = x.copy()
    assert gs.gs(y, gs_double, gs_add) is None
    assert rel_inf(y, ref) < 1e-14 and np.array_equal(y[ids == 0], x[ids == 0])
    t = torch.from_numpy(x).cuda()
    gs.gs(t, gs_double, gs_add)
    assert np.array_equal(t.cpu().numpy(), y)
    with pytest.raises(ValueError):
        gs.gs(x[:-1].copy(), gs_double, gs_add)


def test_unmodified_reference_mesh_runs_on_the_cuda_gather_scatter():
    """The reference's own sempy/mesh.py (verbatim copy under oracle/_ref, see oracle/make_ref.py) with
    sempy_b200.gslib_wrapper in place of gslib: numbering, rmult and dssum as in the golden fixtures."""
    from oracle import make_ref, ref_stubs

    if not make_ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import sempy_b200.gslib_wrapper as gsw

    ref_stubs.install(make_ref.DEST, gslib=gsw)
    try:
        from sempy.mesh import load_mesh as ref_load_mesh

        g = load_golden("box008_N7")
        mesh = ref_load_mesh("box008.msh")
        mesh.find_physical_coordinates(7)
        mesh.establish_global_numbering()
        assert isinstance(mesh.gs, gsw.GS)
        assert np.array_equal(mesh.global_id, g["global_id"]) and np.array_equal(mesh.get_rmult(), g["rmult"])
        x = g["p"].copy()
        assert mesh.dssum(x) is x
        assert rel_inf(x * mesh.get_rmult(), g["p_cont"]) < 1e-15
    finally:
        ref_stubs.install(make_ref.DEST)


@pytest.mark.parametrize("dims,perturb,N", [((4, 3, 3), 0.0, 7), ((3, 3, 3), 0.3, 7), ((2, 2, 3), 0.3, 10), ((3, 2, 2), 0.2, 3), ((2, 2, 2), 0.3, 1)])
def test_fdm_preconditioner_vs_oracle(dims, perturb, N):
    """PCG with the fast-diagonalisation preconditioner against the oracle's restatement of the reference's
    sketch (example.py:71-114,155-171,186-190) applied per element; on straight boxes the interior solves
    are exact, so the iteration count must drop below Jacobi's."""
    from sempy_b200.elliptic import elliptic_pcg, elliptic_pcg_fdm
    from sempy_b200.mesh import box_mesh

    m = box_mesh(*dims, perturb=perturb, seed=4)
    m.find_physical_coordinates(N)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    D = np.ascontiguousarray(__import__("sempy_b200.derivative", fromlist=["x"]).reference_derivative_matrix(N))
    geom = m.get_geom()
    g2l, gs = m.get_global_to_local_map()
    X, Y, Z = m.get_x(), m.get_y(), m.get_z()
    b = 3 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z) * m.get_mass() * m.get_jaco()
    b = m.apply_mask(m.dssum(b))
    axo = lambda v: so.elliptic_ax_batched(geom, v, N, D)
    dinv = m.mask / so.gather_scatter(so.jacobi_diagonal(geom, N, D), g2l, gs)
    S, lam = so.fdm_1d(N) if N > 1 else (np.zeros((0, 0)), np.zeros(0))
    esc = so.fdm_element_scales(m.x, m.y, m.z)
    minv = lambda r: so.fdm_apply(r, N, S, lam, esc, dinv)
    xo, ito = so.elliptic_pcg_fdm(axo, minv, m.mask, g2l, gs, m.get_rmult(), b, tol=1e-10, maxit=2000)
    xg, itg = elliptic_pcg_fdm(m, b, tol=1e-10, maxit=2000)
    assert abs(itg - ito) <= 1
    if itg != ito:
        xg, itg = elliptic_pcg_fdm(m, b, tol=0.0, maxit=ito)
    assert rel_inf(xg, xo) < 1e-10
    xj, itj = elliptic_pcg(m, b, tol=1e-10, maxit=2000)
    assert rel_inf(xg, xj) < 1e-7
    if N >= 3:
        assert itg < itj, (itg, itj)


@pytest.mark.parametrize("N", [7, 4, 10])
def test_irregular_multiplicities_vs_oracle(N):
    """A box with elements removed (L-shaped in every direction): nodes with 3, 5, 6 and 7 copies exercise the
    CSR gather-scatter class in its in-place form (dssum) and its compact form (inside the CG loop)."""
    from sempy_b200 import _capi
    from sempy_b200.elliptic import elliptic_ax, elliptic_cg, elliptic_pcg
    from sempy_b200.mesh import Mesh, box_mesh_arrays

    pts, hexes = box_mesh_arrays(3, 3, 2, perturb=0.2, seed=9)
    keep = np.ones(len(hexes), bool)
    keep[[0, 4, 13, 17]] = False  # corners of the two layers and the centre of the upper one
    m = Mesh.from_arrays(pts, hexes[keep])
    m.find_physical_coordinates(N)
    m.establish_global_numbering()
    m.calc_geometric_factors()
    m.setup_mask()
    g2l, gs = m.get_global_to_local_map()
    counts = np.unique(np.diff(gs))
    assert {3, 5, 6, 7} & set(counts.tolist()), counts  # irregular segment lengths are present
    D = np.ascontiguousarray(__import__("sempy_b200.derivative", fromlist=["x"]).reference_derivative_matrix(N))
    geom = m.get_geom()
    rng = np.random.default_rng(12)
    p = rng.standard_normal(m.get_num_elems() * m.Np)
    x = p.copy()
    m.dssum(x)
    assert rel_inf(x, so.gather_scatter(p, g2l, gs)) < 1e-15
    assert np.array_equal(m.get_rmult(), so.multiplicity_inverse(g2l, gs))
    b = m.apply_mask(m.dssum(rng.standard_normal(p.size)))
    axo = lambda v: so.elliptic_ax_batched(geom, v, N, D)
    xo, ito = so.elliptic_cg(axo, m.mask, g2l, gs, m.get_rmult(), b, tol=1e-10, maxit=3000)
    lib = _capi.lib()
    sols = []
    for pull in (1, 0):
        _capi.check(lib.semb_set_tuning(b"gs_pull", pull))
        xg, itg = elliptic_cg(m, b, tol=1e-10, maxit=3000)
        assert abs(itg - ito) <= 1
        if itg != ito:
            xg, itg = elliptic_cg(m, b, tol=0.0, maxit=ito)
        assert rel_inf(xg, xo) < 1e-10
        sols.append(xg)
    _capi.check(lib.semb_set_tuning(b"gs_pull", 1))
    assert np.array_equal(sols[0], sols[1])
    dinv = m.mask / so.gather_scatter(so.jacobi_diagonal(geom, N, D), g2l, gs)
    xo, ito = so.elliptic_pcg_jacobi(axo, dinv, m.mask, g2l, gs, m.get_rmult(), b, tol=1e-10, maxit=3000)
    xg, itg = elliptic_pcg(m, b, tol=1e-10, maxit=3000)
    assert abs(itg - ito) <= 1
    if itg != ito:
        xg, itg = elliptic_pcg(m, b, tol=0.0, maxit=ito)
    assert rel_inf(xg, xo) < 1e-10


def test_unmodified_reference_driver_runs_on_the_product(capsys):
    """BASELINE config 1: the reference's own poisson.py (verbatim copy, oracle/_ref/poisson.py), with
    `sempy.mesh` / `sempy.elliptic` resolved to the product's modules of the same names.  The script's own
    asserts (solution against the analytic one, host CG against the "loopy" device CG) must hold."""
    import os
    import runpy
    import sys

    from oracle import make_ref

    script = os.path.join(make_ref.DEST, "poisson.py")
    if not os.path.isfile(script):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import sempy_b200
    import sempy_b200.elliptic
    import sempy_b200.mesh

    saved = {k: sys.modules.get(k) for k in ("sempy", "sempy.elliptic", "sempy.mesh")}
    sys.modules.update({"sempy": sempy_b200, "sempy.elliptic": sempy_b200.elliptic, "sempy.mesh": sempy_b200.mesh})
    try:
        ns = runpy.run_path(script)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    out = capsys.readouterr().out
    assert "CG: iters (host/device):" in out and ns["niter"] == ns["niter_loopy"] > 0
    assert type(ns["mesh"]).__module__ == "sempy_b200.mesh"


def test_misaligned_device_vectors_are_refused():
    """The vector kernels use 128-bit accesses: a contiguous CUDA view with an odd element offset must raise
    (ValueError from the Python checks, RuntimeError from the C ABI) instead of faulting inside a kernel."""
    torch = pytest.importorskip("torch")
    from sempy_b200.elliptic import elliptic_ax, elliptic_cg

    g = load_golden("box008_N7")
    m = golden_mesh(g)
    n = g["p"].size
    base = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    view = base[1:]
    view.copy_(torch.from_numpy(g["p"]))
    assert view.is_contiguous() and view.data_ptr() % 16 == 8
    with pytest.raises(ValueError):
        m.dssum(view)
    with pytest.raises(ValueError):
        m.apply_mask(view)
    with pytest.raises(RuntimeError, match="16-byte aligned"):
        elliptic_ax(m, view)
    with pytest.raises(RuntimeError, match="16-byte aligned"):
        elliptic_cg(m, view, tol=1e-8, maxit=5)
    # the context is still usable afterwards (no sticky CUDA error)
    ok = elliptic_ax(m, torch.from_numpy(g["p"]).cuda())
    assert rel_inf(ok.cpu().numpy(), g["ax_raw"]) <= AX_TOL


def test_setup_order_and_missing_pieces_raise():
    """Calls before their set-up step fail loudly, as the reference's attribute errors would."""
    from sempy_b200.elliptic import elliptic_ax, elliptic_cg
    from sempy_b200.mesh import box_mesh

    m = box_mesh(2, 2, 2)
    with pytest.raises(RuntimeError):
        m.dssum(np.zeros(8))  # no order yet
    m.find_physical_coordinates(3)
    n = m.get_num_elems() * m.Np
    with pytest.raises(RuntimeError, match="establish_global_numbering"):
        m.dssum(np.zeros(n))
    with pytest.raises(RuntimeError, match="calc_geometric_factors"):
        elliptic_ax(m, np.zeros(n))
    m.establish_global_numbering()
    m.calc_geometric_factors()
    with pytest.raises(RuntimeError, match="setup_mask"):
        m.apply_mask(np.zeros(n))
    b = m.dssum(np.random.default_rng(1).standard_normal(n))
    x, it = elliptic_cg(m, b, tol=1e-8, maxit=3)  # the mask is optional for the solver
    assert it == 3 and np.isfinite(x).all()
