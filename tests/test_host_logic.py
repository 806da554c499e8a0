"""CPU-only tests of the product's host side (no CUDA calls): 1-D inputs,
gmsh I/O, box generator, global numbering (bit-exact vs the reference's output
on the reference's own coordinates), mask rule."""

import numpy as np
import pytest

from conftest import MESH_CASES, load_golden
from sempy_b200 import mesh as pm
from sempy_b200.derivative import lagrange_derivative_matrix, reference_derivative_matrix
from sempy_b200.interpolation import lagrange
from sempy_b200.mass import reference_mass_matrix_1d, reference_mass_matrix_2d, reference_mass_matrix_3d
from sempy_b200.quadrature import gauss_lobatto

UNIT_N = (1, 2, 3, 4, 7, 10, 15)


@pytest.mark.parametrize("N", UNIT_N)
def test_inputs_match_reference(units, N):
    z, w = gauss_lobatto(N)
    assert np.allclose(z, units[f"z_{N}"], rtol=0, atol=4e-15)
    assert np.allclose(w, units[f"w_{N}"], rtol=1e-13)
    assert np.array_equal(lagrange_derivative_matrix(units[f"z_{N}"]), units[f"D_{N}"])
    assert np.allclose(reference_derivative_matrix(N), units[f"D_{N}"], rtol=1e-12, atol=1e-13)
    assert np.allclose(reference_mass_matrix_3d(N), units[f"B_{N}"], rtol=1e-13)
    z1, _ = gauss_lobatto(1)
    assert np.array_equal(lagrange(units[f"z_{N}"], z1), units[f"J_{N}"])
    assert reference_mass_matrix_1d(N).shape == (N + 1, N + 1)
    assert reference_mass_matrix_2d(N).shape == ((N + 1) ** 2,)


@pytest.mark.parametrize("case", MESH_CASES)
def test_numbering_bit_exact(case):
    g = load_golden(case)
    gid, g2l, gs = pm.global_numbering_from_coordinates(g["xe"], g["ye"], g["ze"])
    assert gid.dtype == np.int32 and g2l.dtype == np.int64 and gs.dtype == np.int64
    assert np.array_equal(gid, g["global_id"])
    assert np.array_equal(g2l, g["global_to_local"])
    assert np.array_equal(gs, g["global_start"])
    assert gid.min() == 1 and gid.max() == gs.size - 1


@pytest.mark.parametrize("case", MESH_CASES)
def test_mask_rule(case):
    g = load_golden(case)
    m = pm.bounding_box_mask(g["xe"], g["ye"], g["ze"])
    assert np.array_equal(m, g["mask"])
    assert np.array_equal(np.flatnonzero(m == 0), g["mask_ids"])


@pytest.mark.parametrize("case", MESH_CASES)
def test_gmsh_roundtrip_and_topology(case, tmp_path):
    g = load_golden(case)
    path = str(tmp_path / "m.msh")
    pm.write_gmsh22(path, g["points"], g["hexes"])
    m = pm.Mesh(path)
    assert m.get_num_elems() == g["hexes"].shape[0] and m.get_ndim() == 3 and m.get_num_verts() == 8
    assert np.array_equal(m.x, g["vx"]) and np.array_equal(m.y, g["vy"]) and np.array_equal(m.z, g["vz"])
    assert np.array_equal(m.get_elem_to_vert_map(), g["hexes"])
    m2 = pm.Mesh.from_arrays(g["points"], g["hexes"])
    assert np.array_equal(m2.x, m.x)


def test_mesh_errors(tmp_path):
    with pytest.raises(Exception, match="Only strings"):
        pm.Mesh(3)
    p = tmp_path / "quad.msh"
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n1 0 0 0\n2 1 0 0\n3 1 1 0\n4 0 1 0\n$EndNodes\n"
                 "$Elements\n1\n1 3 2 1 1 1 2 3 4\n$EndElements\n")
    with pytest.raises(Exception, match="2D not supported"):
        pm.Mesh(str(p))


def test_box_generator_counts():
    pts, hexes = pm.box_mesh_arrays(3, 2, 4)
    assert pts.shape == (4 * 3 * 5, 3) and hexes.shape == (24, 8)
    # element 0 spans [0,1/3]x[0,1/2]x[0,1/4] with gmsh vertex order
    v = pts[hexes[0]]
    assert np.allclose(v[0], [0, 0, 0]) and np.allclose(v[1], [1 / 3, 0, 0]) and np.allclose(v[2], [1 / 3, 0.5, 0])
    assert np.allclose(v[4], [0, 0, 0.25]) and np.allclose(v[6], [1 / 3, 0.5, 0.25])
    # i-fastest element order
    assert np.allclose(pts[hexes[1, 0]], [1 / 3, 0, 0]) and np.allclose(pts[hexes[3, 0]], [0, 0.5, 0])
    p2, _ = pm.box_mesh_arrays(3, 3, 3, perturb=0.3, seed=0)
    p0, _ = pm.box_mesh_arrays(3, 3, 3)
    moved = np.any(p2 != p0, axis=1)
    assert moved.sum() == 8  # only the 2x2x2 interior vertices move
    assert np.abs(p2 - p0).max() <= 0.15 / 3 + 1e-15


def test_numbering_generated_boxes():
    # straight and perturbed boxes number to exactly (nN+1)^3 unique nodes; coordinates from the oracle
    from oracle import sem_oracle as so

    for dims, pert, N in (((2, 3, 2), 0.0, 3), ((3, 3, 3), 0.3, 4)):
        pts, hexes = pm.box_mesh_arrays(*dims, perturb=pert)
        xe, ye, ze = so.physical_coordinates(*so.element_vertices(pts, hexes), N)
        gid, g2l, gs = pm.global_numbering_from_coordinates(xe, ye, ze)
        assert gs.size - 1 == np.prod([d * N + 1 for d in dims])
        gid_o, g2l_o, gs_o = so.global_numbering_literal(xe, ye, ze)
        assert np.array_equal(gid, gid_o) and np.array_equal(g2l, g2l_o) and np.array_equal(gs, gs_o)


def test_bundled_meshes_load():
    for name, ne in (("box001.msh", 1), ("box002.msh", 2), ("box004.msh", 4), ("box008.msh", 8)):
        m = pm.load_mesh(name)
        assert m.get_num_elems() == ne and m.get_ndim() == 3
        assert m.x.min() == 0.0 and m.z.max() == 1.0


def test_bundled_meshes_equal_the_reference_fixtures():
    """load_mesh("box00x.msh") must mean the reference's mesh of that name: same vertices (bit for bit) and
    the same, partly rotated, element orientations (sempy/meshes/box008.msh:41-72) as the golden fixtures
    recorded from the reference's reader -- and as the reference's own files where they are present."""
    import os

    from conftest import load_golden
    from sempy_b200 import mesh as pm

    here = os.path.dirname(pm.__file__)
    for name in ("box001", "box002", "box004", "box008"):
        pts, hexes, _ = pm.read_gmsh22(os.path.join(here, "meshes", name + ".msh"))
        g = load_golden(name + "_N7")
        assert np.array_equal(pts, g["points"]) and np.array_equal(hexes, g["hexes"])
        ref = os.path.join("/root/reference/sempy/meshes", name + ".msh")
        if os.path.exists(ref):
            rp, rh, _ = pm.read_gmsh22(ref)
            assert np.array_equal(pts, rp) and np.array_equal(hexes, rh)


@pytest.mark.parametrize("N", [2, 3, 7, 10, 15])
def test_fdm_host_setup_matches_the_oracle(N):
    """sempy_b200.fdm (symmetric eigen-solve on the B-scaled interior matrix) against the oracle's restatement
    of the reference sketch (scipy.linalg.eigh(A, B), example.py:92-100): same eigenvalues, the same
    preconditioner matrix S diag(1/lambda) S^T (eigenvectors are only defined up to sign), S^T B S = I."""
    from oracle import sem_oracle as so
    from sempy_b200 import fdm
    from sempy_b200.mass import reference_mass_matrix_1d

    S, lam = fdm.fdm_1d(N)
    So, lamo = so.fdm_1d(N)
    assert np.allclose(lam, lamo, rtol=1e-11) and (lam > 0).all()
    P, Po = S @ np.diag(1.0 / lam) @ S.T, So @ np.diag(1.0 / lamo) @ So.T
    assert np.abs(P - Po).max() <= 1e-11 * np.abs(Po).max()
    B = reference_mass_matrix_1d(N)[1:-1, 1:-1]
    assert np.abs(S.T @ B @ S - np.eye(N - 1)).max() < 1e-12


def test_fdm_element_scales():
    from oracle import sem_oracle as so
    from sempy_b200 import fdm, mesh as pm

    pts, hexes = pm.box_mesh_arrays(3, 2, 4, perturb=0.3, seed=5)
    vx, vy, vz = pts[hexes, 0], pts[hexes, 1], pts[hexes, 2]
    got = fdm.element_scales(vx, vy, vz)
    assert np.allclose(got, so.fdm_element_scales(vx, vy, vz), rtol=1e-13)
    # a straight 3 x 2 x 4 box: a = (2 / h)^2 per direction, scale = 8 / volume
    pts, hexes = pm.box_mesh_arrays(3, 2, 4)
    sc = fdm.element_scales(pts[hexes, 0], pts[hexes, 1], pts[hexes, 2])
    assert np.allclose(sc, [[36.0, 16.0, 64.0, 8.0 * 24.0]] * 24)
