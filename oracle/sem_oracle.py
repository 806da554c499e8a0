"""CPU oracle for the SemPy Poisson hot path -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference algorithm (SealUofI/SemPy,
pure Python).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it.  Nothing under ``sempy_b200/`` imports it, and the product path raises when
its CUDA library is missing instead of falling back to anything in here.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (with stand-ins for its two un-vendored
imports, ``meshio`` and ``gslib_wrapper``) and stores its outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those files.  The one boundary that no reference test pins
directly is ``gslib_wrapper.GS.gs`` (third-party, absent, version not stated by
the reference): only the summed value is defined there, not the summation
order, so ``gather_scatter`` is pinned to 1 ulp-level agreement, not bitwise.
The FDM functions at the end (``fdm_*``, ``fast_kron``, a "next" row of the scope) are the exception --
parity UNPINNED against reference outputs: the reference's ``example.py`` does not run (undefined names
at ``:33,36,49``), so they are pinned to the properties the sketch defines (dense Kronecker product,
exact inverse of the interior blocks on straight elements; ``tests/test_oracle_golden.py``).

Each function cites the reference lines it follows (paths relative to the
reference checkout).
"""

from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# 1-D building blocks (inputs of the hot path)
# --------------------------------------------------------------------------


def gauss_lobatto(p):
    """GLL nodes and weights on [-1, 1]; follows sempy/quadrature.py:8-40.

    Interior nodes are the eigenvalues of the (p-1)x(p-1) symmetric Jacobi
    matrix (np.linalg.eig, then sorted); weights 2/(p(p+1) L_p(z)^2) with L_p
    from the three-term recurrence in the same operation order.
    """
    n = p + 1
    z = np.zeros(n)
    w = np.zeros(n)
    z[0], z[-1] = -1.0, 1.0
    if p == 2:
        z[1] = 0.0
    elif p > 2:
        m = p - 1
        T = np.zeros((m, m))
        for i in range(1, m):
            off = 0.5 * np.sqrt((i * (i + 2.0)) / ((i + 0.5) * (i + 1.5)))
            T[i - 1, i] = off
            T[i, i - 1] = off
        lam, _ = np.linalg.eig(T)
        z[1:p] = np.sort(lam)
    w[0] = w[-1] = 2.0 / (p * n)
    for i in range(1, p):
        x = z[i]
        lo, hi = 1, x
        nxt = hi
        for j in range(1, p):
            nxt = x * hi * (2.0 * j + 1.0) / (j + 1.0) - lo * j / (j + 1.0)
            lo, hi = hi, nxt
        w[i] = 2.0 / (p * n * nxt * nxt)
    return z, w


def lagrange_derivative_matrix(x):
    """Differentiation matrix on nodes x; follows sempy/derivative.py:6-31."""
    n = x.size
    a = np.ones(n)
    for i in range(n):
        for j in range(n):
            if j != i:
                a[i] = a[i] * (x[i] - x[j])
    a = 1.0 / a
    D = np.zeros((n, n))
    for i in range(n):
        D[i, :] = a[i] * (x[i] - x)
        D[i, i] = 1
    for j in range(n):
        D[:, j] = D[:, j] / a[j]
    D = 1.0 / D
    for i in range(n):
        D[i, i] = 0
        D[i, i] = -sum(D[i, :])
    return D


def reference_derivative_matrix(p):
    """sempy/derivative.py:34-37."""
    z, _ = gauss_lobatto(p)
    return lagrange_derivative_matrix(z)


def reference_mass_matrix_3d(p):
    """Diagonal reference mass w_i w_j w_k, i fastest; sempy/mass.py:23-33."""
    _, w = gauss_lobatto(p)
    return (w[:, None, None] * (w[None, :, None] * w[None, None, :])).reshape(-1)


def lagrange(x_out, x_in):
    """Interpolation matrix in extended precision; sempy/interpolation.py:4-29."""
    n_in = x_in.size
    a = np.ones(n_in, dtype=np.longdouble)
    for i in range(n_in):
        for j in range(n_in):
            if j != i:
                a[i] = a[i] * (x_in[i] - x_in[j])
    a = 1.0 / a
    J = np.zeros((x_out.size, n_in))
    s = np.ones(n_in, dtype=np.longdouble)
    t = np.ones(n_in, dtype=np.longdouble)
    for i, x in enumerate(x_out):
        for j in range(1, n_in):
            s[j] = s[j - 1] * (x - x_in[j - 1])
            t[n_in - j - 1] = t[n_in - j] * (x - x_in[n_in - j])
        J[i, :] = a * s * t
    return J


# --------------------------------------------------------------------------
# Tensor-product contractions on one element
# --------------------------------------------------------------------------


def kron(Sz, Sy, Sx, U):
    """(Sz (x) Sy (x) Sx) u with C ordering; sempy/kron.py:17-43 (einsum branch).

    ``order=None`` of the reference crashes on numpy >= 2 (kron.py:24-26); the
    layout every caller assumes is C order, which is what is computed here.
    """
    mz, my, mx = Sz.shape[1], Sy.shape[1], Sx.shape[1]
    V = np.asarray(U).reshape(mz, my, mx)
    V = np.einsum("ai,bj,ijk,ck->abc", Sz, Sy, V, Sx, optimize=True)
    return V.reshape(-1)


def gradient(U, n, D):
    """r/s/t derivatives of one element; sempy/gradient.py:6-23.

    Ur[kj,i] = sum_m U[kj,m] D[i,m];  Us[k,j,i] = sum_m D[j,m] U[k,m,i];
    Ut[k,ji] = sum_m D[k,m] U[m,ji].  D is passed in (the reference rebuilds
    it on every call, gradient.py:7).
    """
    Ur = U.reshape(n * n, n) @ D.T
    V = U.reshape(n, n, n)
    Us = np.empty((n, n, n))
    for k in range(n):
        Us[k] = D @ V[k]
    Ut = D @ U.reshape(n, n * n)
    return Ur.reshape(-1), Us.reshape(-1), Ut.reshape(-1)


def gradient_transpose(Wr, Ws, Wt, n, D):
    """Transposed contractions and their sum; sempy/gradient.py:39-56."""
    A = Wr.reshape(n * n, n) @ D
    V = Ws.reshape(n, n, n)
    B = np.empty((n, n, n))
    for k in range(n):
        B[k] = D.T @ V[k]
    C = D.T @ Wt.reshape(n, n * n)
    return A.reshape(-1) + B.reshape(-1) + C.reshape(-1)


# --------------------------------------------------------------------------
# Mesh set-up (inputs of the hot path)
# --------------------------------------------------------------------------


def read_gmsh22(fname):
    """Minimal gmsh 2.2 ASCII reader: what ``meshio.read`` gives ``Mesh.read``
    (sempy/mesh.py:102-122): points (nv,3) and the type-5 (8-node hex) cells,
    0-based.  Format as in sempy/meshes/box008.msh."""
    with open(fname) as fh:
        tok = fh.read().split("\n")
    it = iter(tok)
    points, hexes, quads = None, [], []
    for line in it:
        line = line.strip()
        if line == "$Nodes":
            nv = int(next(it))
            points = np.zeros((nv, 3))
            for _ in range(nv):
                f = next(it).split()
                points[int(f[0]) - 1] = [float(f[1]), float(f[2]), float(f[3])]
        elif line == "$Elements":
            ne = int(next(it))
            for _ in range(ne):
                f = [int(v) for v in next(it).split()]
                etype, ntags = f[1], f[2]
                nodes = f[3 + ntags:]
                if etype == 5:
                    hexes.append([v - 1 for v in nodes])
                elif etype == 3:
                    quads.append([v - 1 for v in nodes])
    return points, np.array(hexes, dtype=np.int64).reshape(-1, 8), np.array(quads, dtype=np.int64).reshape(-1, 4)


def element_vertices(points, hexes):
    """x,y,z of the 8 vertices per element in gmsh order; mesh.py:141-163."""
    return points[hexes, 0], points[hexes, 1], points[hexes, 2]


def physical_coordinates(vx, vy, vz, N):
    """GLL node coordinates of every element; sempy/mesh.py:247-282.

    gmsh vertex order -> lexicographic (i fastest) via [0,1,3,2,4,5,7,6], then
    trilinear interpolation kron(J,J,J,.) with J = lagrange(z_N, z_1).
    """
    z1, _ = gauss_lobatto(1)
    zN, _ = gauss_lobatto(N)
    J = lagrange(zN, z1)
    perm = [0, 1, 3, 2, 4, 5, 7, 6]
    E = vx.shape[0]
    Np = (N + 1) ** 3
    xe = np.empty((E, Np))
    ye = np.empty((E, Np))
    ze = np.empty((E, Np))
    for e in range(E):
        xe[e] = kron(J, J, J, vx[e, perm])
        ye[e] = kron(J, J, J, vy[e, perm])
        ze[e] = kron(J, J, J, vz[e, perm])
    return xe, ye, ze


def geometric_factors(xe, ye, ze, N):
    """G (E,3,3,Np), jaco (E,Np), mass (E,Np); sempy/mesh.py:304-356."""
    n = N + 1
    D = reference_derivative_matrix(N)
    B = reference_mass_matrix_3d(N)
    E, Np = xe.shape
    geom = np.zeros((E, 3, 3, Np))
    jaco = np.zeros((E, Np))
    mass = np.zeros((E, Np))
    for e in range(E):
        xr, xs, xt = gradient(xe[e], n, D)
        yr, ys, yt = gradient(ye[e], n, D)
        zr, zs, zt = gradient(ze[e], n, D)
        J = xr * (ys * zt - yt * zs) - yr * (xs * zt - xt * zs) + zr * (xs * yt - ys * xt)
        rx = (ys * zt - yt * zs) / J
        sx = (yt * zr - yr * zt) / J
        tx = (yr * zs - ys * zr) / J
        ry = -(zt * xs - zs * xt) / J
        sy = -(zr * xt - zt * xr) / J
        ty = -(zs * xr - zr * xs) / J
        rz = (xs * yt - xt * ys) / J
        sz = -(xr * yt - xt * yr) / J
        tz = (xr * ys - xs * yr) / J
        g = {
            (0, 0): rx * rx + ry * ry + rz * rz,
            (0, 1): rx * sx + ry * sy + rz * sz,
            (0, 2): rx * tx + ry * ty + rz * tz,
            (1, 1): sx * sx + sy * sy + sz * sz,
            (1, 2): sx * tx + sy * ty + sz * tz,
            (2, 2): tx * tx + ty * ty + tz * tz,
        }
        for (a, b), v in g.items():
            geom[e, a, b] = v * B * J
            geom[e, b, a] = geom[e, a, b]
        jaco[e] = J
        mass[e] = B
    return geom, jaco, mass


def global_numbering_literal(xe, ye, ze):
    """Literal restatement of sempy/mesh.py:389-435 with Python's stable sort.

    Sort all local nodes by exact (x, y, z); walk the sorted list and open a new
    global id whenever the SQUARED distance to the next point exceeds 1e-12;
    ids start at 1.  Returns (global_id int32 (E,Np), global_to_local int64,
    global_start int64).  Only for small meshes (Python objects).
    """
    E, Np = xe.shape
    size = E * Np
    xs, ys, zs = xe.reshape(-1), ye.reshape(-1), ze.reshape(-1)
    order = sorted(range(size), key=lambda q: (xs[q], ys[q], zs[q]))
    gid = np.zeros(size, dtype=np.int32)
    starts = [0]
    cur = 1
    for pos in range(size - 1):
        a, b = order[pos], order[pos + 1]
        gid[a] = cur
        d2 = (xs[a] - xs[b]) ** 2 + (ys[a] - ys[b]) ** 2 + (zs[a] - zs[b]) ** 2
        if d2 > 1e-12:
            cur += 1
            starts.append(pos + 1)
    gid[order[-1]] = cur
    starts.append(size)
    return gid.reshape(E, Np), np.array(order, dtype=np.int64), np.array(starts, dtype=np.int64)


def global_numbering(xe, ye, ze):
    """Vectorised form of the same rule (stable lexsort + adjacent d^2 > 1e-12);
    used for meshes too large for the literal version.  Bit-identical to
    ``global_numbering_literal`` (tests/test_oracle_golden.py)."""
    E, Np = xe.shape
    xs, ys, zs = xe.reshape(-1), ye.reshape(-1), ze.reshape(-1)
    order = np.lexsort((zs, ys, xs)).astype(np.int64)
    sx, sy, sz = xs[order], ys[order], zs[order]
    d2 = (sx[:-1] - sx[1:]) ** 2 + (sy[:-1] - sy[1:]) ** 2 + (sz[:-1] - sz[1:]) ** 2
    brk = d2 > 1e-12
    ids_sorted = np.concatenate(([1], 1 + np.cumsum(brk)))
    gid = np.empty(E * Np, dtype=np.int32)
    gid[order] = ids_sorted.astype(np.int32)
    starts = np.concatenate(([0], np.nonzero(brk)[0] + 1, [E * Np])).astype(np.int64)
    return gid.reshape(E, Np), order, starts


def gather_scatter(x, global_to_local, global_start):
    """Direct-stiffness sum: every entry becomes the sum over all entries with
    the same global id.  Semantics of gslib's gs(x, gs_double, gs_add) as SemPy
    uses it (sempy/mesh.py:437-447); stated in-tree by python_gather_scatter,
    sempy/loopy/loopy_kernels.py:26-39 (sum in segment order).  Returns a new
    array."""
    seg_sum = np.add.reduceat(x[global_to_local], global_start[:-1])
    out = np.empty_like(x)
    out[global_to_local] = np.repeat(seg_sum, np.diff(global_start))
    return out


def gather_scatter_literal(x, global_to_local, global_start):
    """Loop form, same summation order as python_gather_scatter
    (loopy_kernels.py:26-39): left-to-right inside each segment."""
    out = np.zeros_like(x)
    for s in range(global_start.size - 1):
        lo, hi = global_start[s], global_start[s + 1]
        acc = 0.0
        for q in range(lo, hi):
            acc += x[global_to_local[q]]
        for q in range(lo, hi):
            out[global_to_local[q]] = acc
    return out


def multiplicity_inverse(global_to_local, global_start):
    """rmult = 1 / (number of local copies); sempy/mesh.py:440-442."""
    ones = np.ones(global_to_local.size)
    return 1.0 / gather_scatter(ones, global_to_local, global_start)


def dirichlet_mask(xe, ye, ze):
    """0/1 mask, 0 where a coordinate is within 1e-8 of the bounding box;
    sempy/mesh.py:449-477."""
    tol = 1e-8
    m = np.ones(xe.shape)
    for c in (xe, ye, ze):
        m[np.abs(c - np.min(c)) < tol] = 0
        m[np.abs(c - np.max(c)) < tol] = 0
    return m.reshape(-1)


# --------------------------------------------------------------------------
# The hot path: Ax, CG, PCG
# --------------------------------------------------------------------------


def elliptic_ax(geom, p, N, D):
    """Per-element Poisson operator, unassembled, unmasked;
    sempy/elliptic.py:9-27 (loop form, same operation order per element)."""
    n = N + 1
    E = geom.shape[0]
    Np = n ** 3
    P = p.reshape(E, Np)
    out = np.zeros_like(P)
    for e in range(E):
        pr, ps, pt = gradient(P[e], n, D)
        g = geom[e]
        wr = g[0, 0] * pr + g[0, 1] * ps + g[0, 2] * pt
        ws = g[1, 0] * pr + g[1, 1] * ps + g[1, 2] * pt
        wt = g[2, 0] * pr + g[2, 1] * ps + g[2, 2] * pt
        out[e] = gradient_transpose(wr, ws, wt, n, D)
    return out.reshape(-1)


def elliptic_ax_batched(geom, p, N, D):
    """Same operator for all elements at once (batched matmuls).  Differs from
    ``elliptic_ax`` only by floating-point association (<= 1e-15 relative);
    used where the loop form is too slow and as the timed CPU baseline."""
    n = N + 1
    E = geom.shape[0]
    U = p.reshape(E, n, n, n)
    pr = U @ D.T
    ps = D @ U
    pt = np.einsum("km,emji->ekji", D, U, optimize=True)
    g = geom.reshape(E, 3, 3, n, n, n)
    wr = g[:, 0, 0] * pr + g[:, 0, 1] * ps + g[:, 0, 2] * pt
    ws = g[:, 1, 0] * pr + g[:, 1, 1] * ps + g[:, 1, 2] * pt
    wt = g[:, 2, 0] * pr + g[:, 2, 1] * ps + g[:, 2, 2] * pt
    out = wr @ D + D.T @ ws + np.einsum("mk,emji->ekji", D, wt, optimize=True)
    return out.reshape(-1)


def elliptic_cg(ax, mask, g2l, gstart, rmult, b, tol=1e-12, maxit=100, history=None):
    """CG on local (element-major) vectors; sempy/elliptic.py:30-73.

    ``ax`` is a callable p -> unassembled A p.  Weighted dots sum rmult*r*r;
    pAp is taken between the masked local Ap and p BEFORE the gather-scatter.
    """
    norm_b = np.dot(rmult * b, b)
    TOL = max(tol * tol * norm_b, tol * tol)
    r = b
    rdotr = np.dot(rmult * r, r)
    x = 0 * b
    niter = 0
    if rdotr < 1.0e-20:
        return x, niter
    p = r
    while niter < maxit and rdotr > TOL:
        Ap = ax(p) * mask
        pAp = np.dot(Ap, p)
        alpha = rdotr / pAp
        Ap = gather_scatter(Ap, g2l, gstart)
        x = x + alpha * p
        r = r - alpha * Ap
        rdotr0 = rdotr
        rdotr = np.dot(rmult * r, r)
        beta = rdotr / rdotr0
        if history is not None:
            history.append((rdotr0, rdotr, alpha, beta, pAp))
        p = r + beta * p
        niter += 1
    return x, niter


def cg(A, b, tol=1e-12, maxit=100):
    """Generic CG with plain dots; sempy/iterative.py:4-42."""
    norm_b = np.dot(b, b)
    TOL = max(tol * tol * norm_b, tol * tol)
    r = b
    rdotr = np.dot(r, r)
    x = 0 * b
    niter = 0
    if rdotr < 1.0e-20:
        return x, niter
    p = r
    while niter < maxit and rdotr > TOL:
        Ap = A(p)
        pAp = np.dot(p, Ap)
        alpha = rdotr / pAp
        x = x + alpha * p
        r = r - alpha * Ap
        rdotr0 = rdotr
        rdotr = np.dot(r, r)
        beta = rdotr / rdotr0
        p = r + beta * p
        niter += 1
    return x, niter


def pcg(A, Minv, b, tol=1e-8, maxit=100):
    """Generic PCG, stops on r.z; sempy/iterative.py:45-85."""
    assert b.ndim == 1
    x = np.zeros(b.size)
    norm_b = np.dot(b, b)
    TOL = max(tol * tol * norm_b, tol * tol)
    r = b
    niter = 0
    z = Minv(r)
    rdotz = np.dot(r, z)
    p = z
    while niter < maxit and rdotz > TOL:
        Ap = A(p)
        pAp = np.dot(p, Ap)
        alpha = rdotz / pAp
        x = x + alpha * p
        r = r - alpha * Ap
        z = Minv(r)
        rdotz0 = rdotz
        rdotz = np.dot(r, z)
        beta = rdotz / rdotz0
        p = z + beta * p
        niter += 1
    return x, niter


def jacobi_diagonal(geom, N, D):
    """Unassembled diagonal of the element operator:
    diag[k,j,i] = sum_m D[m,i]^2 Grr[k,j,m] + sum_m D[m,j]^2 Gss[k,m,i]
                + sum_m D[m,k]^2 Gtt[m,j,i]
                + 2 (D_ii D_jj Grs + D_ii D_kk Grt + D_jj D_kk Gst)[k,j,i].
    There is no reference function for it; it is checked against the operator
    itself (tests/test_oracle_golden.py: e_q . A e_q)."""
    n = N + 1
    E = geom.shape[0]
    g = geom.reshape(E, 3, 3, n, n, n)
    D2 = D * D
    dd = np.diag(D)
    diag = np.einsum("mi,ekjm->ekji", D2, g[:, 0, 0])
    diag += np.einsum("mj,ekmi->ekji", D2, g[:, 1, 1])
    diag += np.einsum("mk,emji->ekji", D2, g[:, 2, 2])
    di = dd[None, None, None, :]
    dj = dd[None, None, :, None]
    dk = dd[None, :, None, None]
    diag += 2.0 * (di * dj * g[:, 0, 1] + di * dk * g[:, 0, 2] + dj * dk * g[:, 1, 2])
    return diag.reshape(-1)


def elliptic_pcg_jacobi(ax, dinv, mask, g2l, gstart, rmult, b, tol=1e-8, maxit=100):
    """Jacobi-PCG on local vectors.  It is sempy/iterative.py:45-85 (``pcg``)
    applied to the assembled operator, written on element-major vectors the way
    sempy/elliptic.py:30-73 writes CG: z = dinv*r, dots weighted by rmult, pAp
    before the gather-scatter, stop on r.z <= max(tol^2 b.b, tol^2) with the
    b.b of pcg taken in the same weighted sense.  ``dinv`` = mask/dssum(diag).
    Equivalence with ``pcg`` on global vectors is tested."""
    norm_b = np.dot(rmult * b, b)
    TOL = max(tol * tol * norm_b, tol * tol)
    x = np.zeros(b.size)
    r = b
    niter = 0
    z = dinv * r
    rdotz = np.dot(rmult * r, z)
    p = z
    while niter < maxit and rdotz > TOL:
        Ap = ax(p) * mask
        pAp = np.dot(Ap, p)
        alpha = rdotz / pAp
        Ap = gather_scatter(Ap, g2l, gstart)
        x = x + alpha * p
        r = r - alpha * Ap
        z = dinv * r
        rdotz0 = rdotz
        rdotz = np.dot(rmult * r, z)
        beta = rdotz / rdotz0
        p = z + beta * p
        niter += 1
    return x, niter


# --------------------------------------------------------------------------
# Fast-diagonalisation preconditioner: the reference's single-element sketch
# (example.py:71-114 fdm_d_inv, :155-171 fast_kron, :186-190 precon_fdm) applied
# element by element to the INTERIOR nodes, point Jacobi on the element surfaces.
# --------------------------------------------------------------------------


def fdm_1d(N):
    """example.py:71-100 with the restriction R dropping both end nodes: Ah = Dh^T Bh Dh symmetrised,
    generalised eigenproblem of the restricted pair (scipy.linalg.eigh(A, B) as the reference calls it)."""
    import scipy.linalg as sla

    _, w = gauss_lobatto(N)
    Bh = np.diag(w)
    Dh = reference_derivative_matrix(N)
    Ah = Dh.T @ Bh @ Dh
    Ah = 0.5 * (Ah + Ah.T)
    I = np.identity(N + 1)
    R = I[1:-1, :]
    A, B = R @ Ah @ R.T, R @ Bh @ R.T
    L, S = sla.eigh(A, B)
    return S.real, L.real


def fast_kron(Sz, Sy, Sx, U):
    """example.py:155-171."""
    nx, mx = Sx.shape
    ny, my = Sy.shape
    nz, mz = Sz.shape
    U = U.reshape((my * mz, mx))
    U = np.dot(U, Sx.T)
    U = U.reshape((mz, my, nx))
    V = np.zeros((mz, ny, nx))
    for i in range(mz):
        V[i, :, :] = np.dot(Sy, U[i, :, :])
    V = V.reshape((mz, nx * ny))
    U = np.dot(Sz, V)
    return U.reshape((nx * ny * nz,))


def fdm_element_scales(vx, vy, vz):
    """a_r, a_s, a_t = (2/L)^2 and 8/(L_r L_s L_t) per element, L = mean edge length per local direction
    (vertices in gmsh order, sempy/mesh.py:272-274 gives the lexicographic permutation)."""
    lex = [0, 1, 3, 2, 4, 5, 7, 6]
    out = np.zeros((vx.shape[0], 4))
    for e in range(vx.shape[0]):
        v = np.stack([vx[e, lex], vy[e, lex], vz[e, lex]], axis=1)  # [k*4 + j*2 + i]
        L = []
        for stride in (1, 2, 4):
            starts = [q for q in range(8) if not (q // stride) % 2]
            L.append(np.mean([np.linalg.norm(v[q + stride] - v[q]) for q in starts]))
        out[e] = [(2 / L[0]) ** 2, (2 / L[1]) ** 2, (2 / L[2]) ** 2, 8.0 / (L[0] * L[1] * L[2])]
    return out


def fdm_apply(r, N, S, lam, esc, dinv):
    """z = M^-1 r: example.py:186-190 (R, S^T(x)S^T(x)S^T, dinv, S(x)S(x)S, R^T) on each element's interior,
    z = dinv * r on its surface."""
    n = N + 1
    E = esc.shape[0]
    r3 = r.reshape(E, n, n, n)
    z = (dinv * r).reshape(E, n, n, n).copy()
    if n > 2:
        m = n - 2
        for e in range(E):
            ax, ay, az, sc = esc[e]
            d = sc / (ax * lam[None, None, :] + ay * lam[None, :, None] + az * lam[:, None, None])
            ri = np.ascontiguousarray(r3[e, 1:-1, 1:-1, 1:-1]).reshape(-1)
            bb = fast_kron(S.T, S.T, S.T, ri) * d.reshape(-1)
            z[e, 1:-1, 1:-1, 1:-1] = fast_kron(S, S, S, bb).reshape(m, m, m)
    return z.reshape(-1)


def elliptic_pcg_fdm(ax, minv, mask, g2l, gstart, rmult, b, tol=1e-8, maxit=100):
    """sempy/iterative.py:45-85 on local vectors (as elliptic_pcg_jacobi) with z = minv(r)."""
    norm_b = np.dot(rmult * b, b)
    TOL = max(tol * tol * norm_b, tol * tol)
    x = np.zeros(b.size)
    r = b
    niter = 0
    z = minv(r)
    rdotz = np.dot(rmult * r, z)
    p = z
    while niter < maxit and rdotz > TOL:
        Ap = ax(p) * mask
        pAp = np.dot(Ap, p)
        alpha = rdotz / pAp
        Ap = gather_scatter(Ap, g2l, gstart)
        x = x + alpha * p
        r = r - alpha * Ap
        z = minv(r)
        rdotz0 = rdotz
        rdotz = np.dot(rmult * r, z)
        beta = rdotz / rdotz0
        p = z + beta * p
        niter += 1
    return x, niter


# --------------------------------------------------------------------------
# Synthetic meshes shared by tests and bench (not in the reference)
# --------------------------------------------------------------------------


def box_mesh(nx, ny, nz, perturb=0.0, seed=0):
    """Vertices and hexes of an nx*ny*nz box on [0,1]^3:
    gmsh hex node order, elements i-fastest / k-slowest; ``perturb`` moves
    interior vertices by perturb*h*(U(0,1)-0.5) per axis."""
    gx = np.linspace(0.0, 1.0, nx + 1)
    gy = np.linspace(0.0, 1.0, ny + 1)
    gz = np.linspace(0.0, 1.0, nz + 1)
    Z, Y, X = np.meshgrid(gz, gy, gx, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if perturb:
        rng = np.random.default_rng(seed)
        h = np.array([1.0 / nx, 1.0 / ny, 1.0 / nz])
        shift = perturb * h * (rng.random(pts.shape) - 0.5)
        kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
        interior = ((ii > 0) & (ii < nx) & (jj > 0) & (jj < ny) & (kk > 0) & (kk < nz)).ravel()
        pts[interior] += shift[interior]

    def vid(i, j, k):
        return (k * (ny + 1) + j) * (nx + 1) + i

    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    hexes = np.stack(
        [vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k),
         vid(i, j, k + 1), vid(i + 1, j, k + 1), vid(i + 1, j + 1, k + 1), vid(i, j + 1, k + 1)],
        axis=1,
    ).astype(np.int64)
    return pts, hexes


def write_gmsh22(fname, points, hexes):
    """Write vertices + 8-node hexes as gmsh 2.2 ASCII (the format of
    sempy/meshes/*.msh) so the same file can feed the reference and the
    product."""
    with open(fname, "w") as fh:
        fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(points))
        for q, (x, y, z) in enumerate(points):
            fh.write("%d %s %s %s\n" % (q + 1, repr(float(x)), repr(float(y)), repr(float(z))))
        fh.write("$EndNodes\n$Elements\n%d\n" % len(hexes))
        for q, h in enumerate(hexes):
            fh.write("%d 5 2 2 1 %s\n" % (q + 1, " ".join(str(int(v) + 1) for v in h)))
        fh.write("$EndElements\n")
