"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``sempy`` from /root/reference with oracle/ref_stubs.py standing in
for the reference's two un-vendored imports, drives the reference's own public
API (sempy/mesh.py Mesh set-up, sempy/elliptic.py elliptic_ax / elliptic_cg,
sempy/iterative.py cg / pcg) and stores inputs and outputs as small fixtures.
The fixtures travel to the GPU box; the reference does not.
"""

from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_stubs, sem_oracle  # noqa: E402

sempy = ref_stubs.install("/root/reference")

from sempy.derivative import reference_derivative_matrix  # noqa: E402
from sempy.elliptic import elliptic_ax, elliptic_cg  # noqa: E402
from sempy.gradient import gradient, gradient_transpose  # noqa: E402
from sempy.iterative import cg, pcg  # noqa: E402
from sempy.kron import kron  # noqa: E402
from sempy.mass import reference_mass_matrix_3d  # noqa: E402
from sempy.mesh import Mesh  # noqa: E402
from sempy.quadrature import gauss_lobatto  # noqa: E402
from sempy.interpolation import lagrange  # noqa: E402


def run_case(name, msh_path, N, cg_tols=(1e-8, 1e-10), with_pcg=True):
    mesh = Mesh(msh_path)
    mesh.find_physical_coordinates(N)
    xe0, ye0, ze0 = mesh.xe.copy(), mesh.ye.copy(), mesh.ze.copy()
    mesh.establish_global_numbering()
    mesh.calc_geometric_factors()
    mesh.setup_mask()
    E, Np = mesh.get_num_elems(), mesh.Np
    geom = mesh.get_geom()
    for a in range(3):
        for b in range(3):
            assert np.array_equal(geom[:, a, b], geom[:, b, a])
    geom6 = np.stack([geom[:, 0, 0], geom[:, 0, 1], geom[:, 0, 2], geom[:, 1, 1], geom[:, 1, 2], geom[:, 2, 2]], axis=1)
    g2l, gstart = mesh.get_global_to_local_map()
    D = reference_derivative_matrix(N)
    z, w = gauss_lobatto(N)

    rng = np.random.default_rng(1234)
    p = rng.standard_normal(E * Np)
    p_cont = mesh.dssum(p.copy()) * mesh.get_rmult()
    ax_raw = elliptic_ax(mesh, p)
    ax_cont = elliptic_ax(mesh, p_cont)
    ax_full = mesh.dssum(mesh.apply_mask(ax_cont.copy()))

    X, Y, Z = mesh.get_x(), mesh.get_y(), mesh.get_z()
    J, B = mesh.get_jaco(), mesh.get_mass()
    x_exact = mesh.apply_mask(np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z))
    b = 3 * np.pi * np.pi * np.sin(np.pi * X) * np.sin(np.pi * Y) * np.sin(np.pi * Z)
    b = b * B * J
    b = mesh.dssum(b)
    b = mesh.apply_mask(b)

    out = dict(
        N=N, points=None, hexes=None,
        vx=mesh.x, vy=mesh.y, vz=mesh.z,
        xe=xe0, ye=ye0, ze=ze0,
        geom6=geom6, jaco=mesh.jaco, mass0=mesh.mass[0],
        global_id=mesh.global_id, global_to_local=g2l, global_start=gstart,
        rmult=mesh.get_rmult(), mask=mesh.mask, mask_ids=mesh.get_mask_ids(),
        D=D, gll_z=z, gll_w=w,
        p=p, p_cont=p_cont, ax_raw=ax_raw, ax_cont=ax_cont, ax_full=ax_full,
        b=b, x_exact=x_exact,
    )
    pts, hexes, _ = sem_oracle.read_gmsh22(msh_path)
    out["points"], out["hexes"] = pts, hexes

    for tol in cg_tols:
        x_cg, niter = elliptic_cg(mesh, b.copy(), tol=tol, maxit=10000)
        tag = "%g" % tol
        out["cg_x_" + tag] = x_cg
        out["cg_niter_" + tag] = niter
        print(f"  {name}: elliptic_cg tol={tol} niter={niter} err_vs_exact={np.abs(x_cg - x_exact).max():.3e}")

    if with_pcg:
        # Jacobi-PCG pinned through the reference's own pcg() on GLOBAL vectors
        # (SURVEY.md section 8a row 9): A_g = M_g Q^T ax(Q .), Minv = M_g/(Q^T diag).
        gid = mesh.global_id.reshape(-1).astype(np.int64) - 1
        ng = gid.max() + 1
        mask_g = np.ones(ng)
        mask_g[gid[mesh.mask == 0]] = 0.0

        def A_g(ug):
            return mask_g * np.bincount(gid, weights=elliptic_ax(mesh, ug[gid]), minlength=ng)

        diag_l = sem_oracle.jacobi_diagonal(geom, N, D)
        # cross-check the closed form against the reference operator on a few unit vectors
        for q in rng.integers(0, E * Np, size=4):
            eq = np.zeros(E * Np)
            eq[q] = 1.0
            assert abs(elliptic_ax(mesh, eq)[q] - diag_l[q]) <= 1e-12 * abs(diag_l[q])
        diag_g = np.bincount(gid, weights=diag_l, minlength=ng)
        dinv_g = mask_g / diag_g

        b_g = np.zeros(ng)
        b_g[gid] = b  # b is continuous
        for tol in cg_tols:
            xg, niter = pcg(A_g, lambda r: dinv_g * r, b_g, tol=tol, maxit=10000)
            tag = "%g" % tol
            out["pcg_x_" + tag] = xg[gid]
            out["pcg_niter_" + tag] = niter
            xg2, niter2 = cg(A_g, b_g, tol=tol, maxit=10000)
            out["gcg_niter_" + tag] = niter2
            print(f"  {name}: pcg(jacobi, global) tol={tol} niter={niter}; global cg niter={niter2}")
        out["diag_local"] = diag_l

    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"wrote {name}.npz  E={E} N={N}")


def unit_fixtures():
    """Single-element pieces: gradient / gradient_transpose / kron / 1-D inputs."""
    out = {}
    rng = np.random.default_rng(7)
    for N in (1, 2, 3, 4, 7, 10, 15):
        n = N + 1
        U = rng.standard_normal(n ** 3)
        W = rng.standard_normal((3, n ** 3))
        ur, us, ut = gradient(U, n)
        out[f"grad_U_{N}"] = U
        out[f"grad_out_{N}"] = np.stack([ur, us, ut])
        out[f"gradT_W_{N}"] = W
        out[f"gradT_out_{N}"] = gradient_transpose(W[0], W[1], W[2], n)
        out[f"D_{N}"] = reference_derivative_matrix(N)
        z, w = gauss_lobatto(N)
        out[f"z_{N}"], out[f"w_{N}"] = z, w
        out[f"B_{N}"] = reference_mass_matrix_3d(N)
        z1, _ = gauss_lobatto(1)
        out[f"J_{N}"] = lagrange(z, z1)
    Sz, Sy, Sx = rng.standard_normal((5, 2)), rng.standard_normal((4, 3)), rng.standard_normal((6, 4))
    U = rng.standard_normal(2 * 3 * 4)
    out["kron_Sz"], out["kron_Sy"], out["kron_Sx"], out["kron_U"] = Sz, Sy, Sx, U
    out["kron_out"] = kron(Sz, Sy, Sx, U, order="C")
    # iterative.cg / pcg on a fixed SPD system
    A = rng.standard_normal((12, 12))
    A = A @ A.T + 12 * np.eye(12)
    bb = rng.standard_normal(12)
    x, it = cg(lambda v: A @ v, bb, tol=1e-12, maxit=100)
    out["it_A"], out["it_b"], out["it_cg_x"], out["it_cg_niter"] = A, bb, x, it
    dA = np.diag(A).copy()
    x, it = pcg(lambda v: A @ v, lambda r: r / dA, bb, tol=1e-10, maxit=100)
    out["it_pcg_x"], out["it_pcg_niter"] = x, it
    np.savez_compressed(os.path.join(HERE, "units.npz"), **out)
    print("wrote units.npz")


def main():
    unit_fixtures()
    ref_meshes = "/root/reference/sempy/meshes"
    for f, N in (("box001", 7), ("box002", 7), ("box004", 7), ("box008", 7), ("box001", 10), ("box004", 10), ("box002", 3), ("box001", 15)):
        run_case(f"{f}_N{N}", os.path.join(ref_meshes, f + ".msh"), N)
    with tempfile.TemporaryDirectory() as tmp:
        for tag, dims, perturb, N in (("gen222", (2, 2, 2), 0.0, 3), ("pert333", (3, 3, 3), 0.3, 4), ("pert322", (3, 2, 2), 0.3, 7)):
            pts, hexes = sem_oracle.box_mesh(*dims, perturb=perturb, seed=0)
            path = os.path.join(tmp, tag + ".msh")
            sem_oracle.write_gmsh22(path, pts, hexes)
            run_case(f"{tag}_N{N}", path, N)


if __name__ == "__main__":
    main()
