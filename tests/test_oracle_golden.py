"""Pins the CPU oracle (oracle/heat_oracle.c) — CPU only.

The reference ships no golden vectors and cannot be built here ("parity unpinned" against the
binary), so the pins are: the hand-checkable 3x3 system, golden numbers produced by the
independent numpy/scipy restatement + a sparse direct solve (tests/golden/make_golden.py),
the analytic P1 solution, structural properties, and METIS known answers."""
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spl

from conftest import mesh_path

MESH_NAMES = ["rectangle-tris-boundary", "bolted_bracket", "tet-cube-heat", "mitchell_tri"]


def test_hand_checked_3x3(oracle):
    # A=[[5,0,-1],[0,4,-1],[-1,-1,5]], B=[500,450,300]
    m = oracle.read_exodus(mesh_path("rectangle-tris-boundary"))
    s = oracle.assemble(m, oracle.GRAPH_LAPLACIAN)
    assert s.n == 3 and s.nnz == 7
    np.testing.assert_array_equal(s.csr().toarray(), [[5, 0, -1], [0, 4, -1], [-1, -1, 5]])
    np.testing.assert_array_equal(s.b, [500, 450, 300])
    x, it, ach, _ = oracle.pcg(s, tol=1e-12)
    np.testing.assert_allclose(x, [122.527472527472, 140.659340659340, 112.637362637362], rtol=1e-12)
    assert it <= 3


@pytest.mark.parametrize("name", MESH_NAMES)
@pytest.mark.parametrize("mode_name", ["graph", "p1"])
def test_oracle_matches_golden(oracle, golden, name, mode_name):
    mode = oracle.GRAPH_LAPLACIAN if mode_name == "graph" else oracle.P1_FEM
    g = golden[name][mode_name]
    m = oracle.read_exodus(mesh_path(name))
    s = oracle.assemble(m, mode)
    A = s.csr()
    assert (s.n, s.nnz) == (g["n"], g["nnz"])
    assert A.diagonal().sum() == pytest.approx(g["trace"], rel=1e-13)
    assert s.b.sum() == pytest.approx(g["sum_b"], rel=1e-12)
    assert abs(A - A.T).sum() <= 1e-10 * abs(A).sum()
    x = spl.spsolve(A.tocsc(), s.b)
    for k, v in (("x_min", x.min()), ("x_max", x.max()), ("x_mean", x.mean()), ("x_norm2", np.linalg.norm(x))):
        assert v == pytest.approx(g[k], rel=1e-9), k
    np.testing.assert_allclose(x[:3], g["x_head"], rtol=1e-9)


@pytest.mark.parametrize("name", MESH_NAMES)
@pytest.mark.parametrize("mode", [0, 1])
def test_c_oracle_equals_numpy_restatement(oracle, name, mode):
    m = oracle.read_exodus(mesh_path(name))
    s = oracle.assemble(m, mode)
    A, b, r2o = oracle.assemble_np(m, mode)
    Ac = s.csr()
    np.testing.assert_array_equal(Ac.indptr, A.indptr)        # pattern bit-exact
    np.testing.assert_array_equal(Ac.indices, A.indices)
    np.testing.assert_array_equal(s.red2orig, r2o)
    scale = np.abs(A.data).max()
    if mode == 0:
        np.testing.assert_array_equal(Ac.data, A.data)         # integer-valued
        np.testing.assert_array_equal(s.b, b)
    else:
        assert np.abs(Ac.data - A.data).max() <= 1e-12 * scale
        assert np.abs(s.b - b).max() <= 1e-12 * max(1.0, np.abs(b).max())


def test_structural_properties_graph(oracle):
    # symmetric; A.1 = number of Dirichlet neighbours; diag = full degree
    m = oracle.read_exodus(mesh_path("tet-cube-heat"))
    s = oracle.assemble(m, oracle.GRAPH_LAPLACIAN)
    A = s.csr()
    assert abs(A - A.T).sum() == 0
    rowsum = np.asarray(A.sum(1)).ravel()
    bc = s.node_bc
    ndir = np.zeros(s.n)
    orig2red = np.full(m.num_nodes, -1)
    orig2red[s.red2orig] = np.arange(s.n)
    edges = set()
    for a in range(4):
        for b_ in range(4):
            if a != b_:
                edges.update(zip(m.conn[:, a].tolist(), m.conn[:, b_].tolist()))
    for u, v in edges:
        if orig2red[u] >= 0 and not np.isnan(bc[v]):
            ndir[orig2red[u]] += 1
    np.testing.assert_array_equal(rowsum, ndir)
    assert (A.diagonal() >= 3).all()


def test_p1_reproduces_linear_field(oracle):
    # exact solution T = 550 - 90 x on tet-cube-heat.exo
    m = oracle.read_exodus(mesh_path("tet-cube-heat"))
    s = oracle.assemble(m, oracle.P1_FEM)
    x, it, ach, hist = oracle.pcg(s, tol=1e-12, max_iters=2000)
    f = oracle.scatter_field(s, x)
    assert np.abs(f - (550.0 - 90.0 * m.x)).max() < 1e-7
    assert ach <= 1e-12


def test_pcg_iteration_counts(oracle):
    # BASELINE.md §2 (survey-derived): Jacobi-PCG, x0 = 0
    expect = {("tet-cube-heat", 0): (105, 136), ("tet-cube-heat", 1): (140, 180), ("bolted_bracket", 0): (123, 137)}
    for (name, mode), (i8, i10) in expect.items():
        s = oracle.assemble(oracle.read_exodus(mesh_path(name)), mode)
        assert abs(oracle.pcg(s, tol=1e-8)[1] - i8) <= 1
        assert abs(oracle.pcg(s, tol=1e-10)[1] - i10) <= 1


def test_bug_compat_d1(oracle, golden):
    # the reference's off-by-one (ExodusIO.hpp:220) for 1 rank
    m = oracle.read_exodus(mesh_path("tet-cube-heat"))
    A, b, _ = oracle.assemble_np(m, oracle.GRAPH_LAPLACIAN, bug_compat_d1=True)
    g = golden["tet-cube-heat"]["graph_bug_compat_d1"]
    assert (A.shape[0], A.nnz, A.diagonal().sum()) == (19248, 274791, 260686)
    assert abs(A - A.T).sum() == 22
    assert g["n"] == 19248


def test_cube_mesh_properties(oracle):
    # SURVEY.md Appendix E
    nx = ny = nz = 6
    m = oracle.cube_mesh(nx, ny, nz)
    assert m.conn.shape == (6 * 5 ** 3, 4)
    P = np.stack([m.x, m.y, m.z], 1)[m.conn]
    vol = np.abs(np.linalg.det(P[:, 1:] - P[:, :1])) / 6
    assert vol.sum() == pytest.approx(1000.0, rel=1e-12)
    s = oracle.assemble(m, oracle.GRAPH_LAPLACIAN)
    a, b_, c = nx - 2, ny, nz
    edges = ((a - 1) * b_ * c + a * (b_ - 1) * c + a * b_ * (c - 1) + (a - 1) * (b_ - 1) * c
             + a * (b_ - 1) * (c - 1) + (a - 1) * b_ * (c - 1) + (a - 1) * (b_ - 1) * (c - 1))
    assert s.n == a * b_ * c and s.nnz == s.n + 2 * edges
    lens = np.diff(s.row_ptr)
    assert lens.max() == 15
    # P1 on the Kuhn cube: exactly linear T = 1000 - 900 i/(nx-1)
    sp1 = oracle.assemble(m, oracle.P1_FEM)
    x, *_ = oracle.pcg(sp1, tol=1e-13, max_iters=500)
    f = oracle.scatter_field(sp1, x)
    i = np.arange(nx * ny * nz) % nx
    assert np.abs(f - (1000.0 - 900.0 * i / (nx - 1))).max() < 1e-8


def test_chebyshev_pcg_converges(oracle):
    s = oracle.assemble(oracle.read_exodus(mesh_path("bolted_bracket")), 0)
    xj, itj, *_ = oracle.pcg(s, tol=1e-10)
    xc, itc, *_ = oracle.pcg(s, tol=1e-10, prec=oracle.PREC_CHEBYSHEV, cheb_degree=3, cheb_lambda_max=2.0)
    assert itc < itj
    assert np.abs(xc - xj).max() <= 1e-7 * np.abs(xj).max()


def test_spmv_matches_scipy(oracle):
    s = oracle.assemble(oracle.read_exodus(mesh_path("bolted_bracket")), 1)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, s.n)
    np.testing.assert_allclose(oracle.spmv(s, x), s.csr() @ x, rtol=0, atol=1e-12)


def test_metis_known_answers(oracle, golden):
    # SURVEY.md §8c(v) with the bundled METIS and the reference's arguments (ExodusIO.hpp:1615)
    m = oracle.read_exodus(mesh_path("rectangle-tris"))
    obj, epart, npart = oracle.metis_part_mesh_dual(m.conn, m.num_nodes, 2, 2)
    assert obj == 2
    assert epart.tolist() == [0, 1, 1, 1, 0, 0, 0, 1]
    assert npart.tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 0]
    for key, rec in golden["metis_part_mesh_dual"].items():
        name, nparts = key.split(":")
        mm = oracle.read_exodus(mesh_path(name))
        ncommon = 2 if mm.conn.shape[1] == 3 else 3
        obj, epart, npart = oracle.metis_part_mesh_dual(mm.conn, mm.num_nodes, ncommon, int(nparts))
        assert obj == rec["objval"]
        assert np.bincount(epart, minlength=int(nparts)).tolist() == rec["epart_hist"]


def test_get_matrix_is_the_whole_mesh_laplacian(oracle):
    """IO::getMatrix (ExodusIO.hpp:1379-1425): -1 off-diagonals, diagonal = row entries - 1, nodesets
    not applied => every row sums to zero (singular, as the reference says at :729-731)."""
    mesh = oracle.read_exodus(mesh_path("bolted_bracket"))
    s = oracle.get_matrix(mesh)
    A = s.csr()
    assert s.n == mesh.num_nodes and np.array_equal(s.red2orig, np.arange(s.n))
    assert abs(A - A.T).sum() == 0 and np.abs(np.asarray(A.sum(1))).max() == 0
    np.testing.assert_array_equal(A.diagonal(), np.diff(s.row_ptr) - 1)
    assert np.all(s.b == 0)


def test_power_method_oracle(oracle):
    """PowerMethod::run (ExodusMatrixTest.cpp:56-129, 500 iterations / 1e-2 at :163) against scipy's
    largest eigenvalue; the stop index is a multiple of the report frequency 50."""
    import scipy.sparse.linalg as spl
    s = oracle.get_matrix(oracle.read_exodus(mesh_path("bolted_bracket")))
    z0 = oracle.hash_vector(np.arange(s.n), 12345)
    lam, res, it, conv = oracle.power_method(s, z0, 500, 1e-2)
    lmax = spl.eigsh(s.csr(), k=1, which="LA", return_eigenvectors=False)[0]
    assert conv and it % 50 == 0 and res < 1e-2
    assert 0 < lmax - lam <= 1e-3 * lmax                      # Rayleigh quotient from below
    lam2, res2, it2, conv2 = oracle.power_method(s, z0, 7, 0.0)
    assert (it2, conv2) == (7, False) and lam2 < lam


def test_hash_vector_twin(oracle):
    v = oracle.hash_vector(np.arange(1000), 12345)
    assert v.min() >= -1 and v.max() < 1 and abs(v.mean()) < 0.1 and len(np.unique(v)) == 1000


@pytest.mark.parametrize("name,mode", [("bolted_bracket", 0), ("tet-cube-heat", 1), ("mitchell_tri", 1)])
def test_ilu0_oracle_reproduces_A_on_its_pattern(oracle, name, mode):
    """Defining property of ILU(0): (L U)_ij = A_ij wherever A has an entry."""
    import scipy.sparse as sp
    s = oracle.assemble(oracle.read_exodus(mesh_path(name)), mode)
    lu = oracle.ilu0(s)
    rows = np.repeat(np.arange(s.n), np.diff(s.row_ptr))
    L = sp.csr_matrix((np.where(s.col < rows, lu, 0.0), s.col, s.row_ptr), shape=(s.n, s.n)) + sp.identity(s.n)
    U = sp.csr_matrix((np.where(s.col >= rows, lu, 0.0), s.col, s.row_ptr), shape=(s.n, s.n))
    A = s.csr()
    patt = A.copy()
    patt.data[:] = 1.0
    assert abs((L @ U).multiply(patt) - A).max() <= 1e-12 * abs(A).max()
    # the two triangular solves invert L U
    v = oracle.hash_vector(np.arange(s.n), 3)
    z = oracle.ilu0_apply(s, lu, v)
    assert np.abs(L @ (U @ z) - v).max() <= 1e-10


@pytest.mark.parametrize("prec,restart", [(3, 300), (3, 1), (1, 30), (0, 50), (2, 20)])
def test_gmres_oracle_matches_direct_solve(oracle, golden, prec, restart):
    """Right-preconditioned restarted GMRES (the reference's literal Krylov method, BelosMueLuSolver.cpp:102-109)
    against the golden direct-solve statistics; restart = 1 is the reference's one-iteration-per-solve loop."""
    s = oracle.assemble(oracle.read_exodus(mesh_path("bolted_bracket")), 0)
    x, it, ach, conv = oracle.gmres(s, prec=prec, restart=restart, max_iters=3000, tol=1e-10, cheb_degree=3)
    g = golden["bolted_bracket"]["graph"]
    assert conv and ach <= 1e-10
    assert abs(x.mean() - g["x_mean"]) <= 1e-7 * abs(g["x_mean"]) and abs(np.linalg.norm(x) - g["x_norm2"]) <= 1e-7 * g["x_norm2"]
    if restart == 1:          # GMRES(1) = minimal-residual iteration: the residual never grows
        res = []
        xk = np.zeros(s.n)
        for _ in range(5):
            xk = oracle.gmres(s, x0=xk, prec=prec, restart=1, max_iters=1, tol=0.0)[0]
            res.append(np.linalg.norm(s.b - s.csr() @ xk))
        assert all(b <= a for a, b in zip(res, res[1:]))


def test_pcg_with_ilu0_oracle(oracle):
    s = oracle.assemble(oracle.read_exodus(mesh_path("tet-cube-heat")), oracle.P1_FEM)
    x, it, *_ = oracle.pcg(s, prec=oracle.PREC_ILU0, tol=1e-10, max_iters=1000)
    xj, itj, *_ = oracle.pcg(s, tol=1e-10, max_iters=1000)
    assert it < itj / 2 and np.abs(x - xj).max() <= 1e-8 * np.abs(xj).max()


@pytest.mark.parametrize("name", ["bolted_bracket", "mitchell_tri", "rectangle-tris"])
def test_get_matrix_matches_golden(oracle, golden, name):
    """Whole-mesh Laplacian of IO::getMatrix against the independent scipy assembly in the golden file, and
    the power method of ExodusMatrixTest against scipy's largest eigenvalue."""
    g = golden["get_matrix"][name]
    s = oracle.get_matrix(oracle.read_exodus(mesh_path(name)))
    assert (s.n, s.nnz) == (g["n"], g["nnz"]) and s.csr().diagonal().sum() == g["trace"]
    lam, res, it, conv = oracle.power_method(s, oracle.hash_vector(np.arange(s.n), 12345), 2000, 1e-6)
    assert -1e-12 * lam <= g["lambda_max"] - lam <= 1e-5 * g["lambda_max"], (lam, g["lambda_max"], it)


@pytest.mark.parametrize("dims", [(3, 2, 2), (5, 4, 3), (9, 7, 5), (12, 9, 11), (17, 17, 17), (33, 17, 16)])
@pytest.mark.parametrize("mode", [0, 1])
def test_cube_analytic_equals_explicit(oracle, dims, mode):
    """oracle_cube_assemble (closed-form Kuhn connectivity: the CPU baseline's assembler at sizes whose explicit
    mesh does not fit the host) is the SAME system as oracle_assemble on the explicit cube mesh, to the last bit
    (zero signs included) — so it inherits every pin of oracle_assemble."""
    a = oracle.assemble(oracle.cube_mesh(*dims), mode)
    b = oracle.cube_assemble(*dims, mode)
    assert (a.n, a.nnz) == (b.n, b.nnz)
    for f in ("row_ptr", "col", "val", "b", "red2orig"):
        assert getattr(a, f).tobytes() == getattr(b, f).tobytes(), f
    c = oracle.cube_assemble(*dims, mode, copy=False)        # the zero-copy wrapper bench.py uses
    assert np.array_equal(c.val, a.val)
    c.free()


@pytest.mark.parametrize("dims", [(3, 2, 2), (7, 5, 6), (12, 9, 11), (5, 4, 3)])
def test_generated_kuhn_row_arithmetic_is_the_oracles(oracle, dims):
    """The partially evaluated element arithmetic of the cube assembly kernel (tools/gen_kuhn_rows.py ->
    csrc/kuhn_rows_generated.cuh: exact-zero operands eliminated, common subexpressions shared) evaluated in IEEE
    doubles on the CPU gives every row of the oracle's system bit for bit — zero signs included — and the header in
    the tree is what the generator emits today."""
    import io
    import struct
    import sys as _sys
    from contextlib import redirect_stderr, redirect_stdout
    from conftest import ROOT
    _sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_kuhn_rows as G
    nx, ny, nz = dims
    ref = oracle.cube_assemble(nx, ny, nz, oracle.P1_FEM)
    w = nx - 2
    for r in range(ref.n):
        i, j, k = r % w + 1, (r // w) % ny, r // (w * ny)
        X = [-5.0 + 10.0 * (i - 1 + t) / (nx - 1) for t in range(3)]
        Y = [-5.0 + 10.0 * (j - 1 + t) / (ny - 1) for t in range(3)]
        Z = [-5.0 + 10.0 * (k - 1 + t) / (nz - 1) for t in range(3)]
        v, b = G.evaluate(X, Y, Z, j >= 1, j <= ny - 2, k >= 1, k <= nz - 2, i == 1, i == nx - 2)
        row = []
        for s in range(15):
            code, sg = (s - 7, 1) if s > 7 else (7 - s, -1)
            ii, jj, kk = i + sg * (code & 1), j + sg * ((code >> 1) & 1), k + sg * ((code >> 2) & 1)
            if 1 <= ii <= nx - 2 and 0 <= jj < ny and 0 <= kk < nz:
                row.append(v[s])
        assert np.array(row).tobytes() == ref.val[ref.row_ptr[r]:ref.row_ptr[r + 1]].tobytes(), (dims, r)
        assert struct.pack("d", b) == struct.pack("d", ref.b[r]), (dims, r)
    if dims == (7, 5, 6):
        out = io.StringIO()
        with redirect_stdout(out), redirect_stderr(io.StringIO()):
            G.main()
        with open(os.path.join(ROOT, "domain-decomposed-pde-solver_b200", "csrc", "kuhn_rows_generated.cuh")) as f:
            assert f.read() == out.getvalue(), "kuhn_rows_generated.cuh is stale: re-run tools/gen_kuhn_rows.py"
