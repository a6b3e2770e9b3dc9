"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(include/heat_b200.h via ctypes), against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): sparsity pattern / maps bit-exact; matrix entries within 1e-12
relative (we get bit-exact: the element arithmetic is contraction-free on both sides); solution
within 1e-8 relative at relative residual <= 1e-10; iteration counts within +-2.
"""
import os

import numpy as np
import pytest

from conftest import mesh_path

pytestmark = pytest.mark.gpu

EXO = ["rectangle-tris-boundary", "bolted_bracket", "tet-cube-heat", "mitchell_tri"]
SOL_RTOL = 1e-8      # north_star: solution within 1e-8 relative
RES_TOL = 1e-10      # ... at relative residual <= 1e-10
ITER_SLACK = 2       # ... iteration counts within +-2


@pytest.fixture(scope="module")
def hb():
    import heat_b200
    if heat_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    return heat_b200


@pytest.fixture()
def io(hb):
    h = hb.IO(0)
    yield h
    h.close()


def _assemble_exo(hb, io, oracle, name, mode):
    io.open(mesh_path(name), True)
    A, X, B = io.assemble(mode, hb.PART_METIS_KWAY)
    ref = oracle.assemble(oracle.read_exodus(mesh_path(name)), mode)
    return A, X, B, ref


@pytest.mark.parametrize("name", EXO)
@pytest.mark.parametrize("mode", [0, 1])
def test_assemble_matches_oracle_bit_exact(hb, io, oracle, name, mode):
    A, X, B, ref = _assemble_exo(hb, io, oracle, name, mode)
    rp, col, val = A.csr()
    mi = A.info
    assert (mi.n_global, mi.nnz_global, mi.n_owned, mi.n_ghost) == (ref.n, ref.nnz, ref.n, 0)
    np.testing.assert_array_equal(rp, ref.row_ptr)          # sparsity pattern bit-exact
    np.testing.assert_array_equal(col, ref.col)
    np.testing.assert_array_equal(A.red2orig(), ref.red2orig)
    np.testing.assert_array_equal(val, ref.val)              # entries: bar 1e-12 relative, measured bit-exact
    np.testing.assert_array_equal(B.numpy(), ref.b)
    assert np.all(X.numpy() == 0.0)


@pytest.mark.parametrize("name", EXO)
def test_assemble_matches_golden(hb, io, oracle, golden, name):
    for mode, key in ((0, "graph"), (1, "p1")):
        A, X, B, _ = _assemble_exo(hb, io, oracle, name, mode)
        g = golden[name][key]
        rp, col, val = A.csr()
        diag = val[col == np.repeat(np.arange(len(rp) - 1), np.diff(rp))]
        assert (A.info.n_global, A.info.nnz_global) == (g["n"], g["nnz"])
        assert diag.sum() == pytest.approx(g["trace"], rel=1e-13)
        assert B.numpy().sum() == pytest.approx(g["sum_b"], rel=1e-12)


@pytest.mark.parametrize("name", ["bolted_bracket", "tet-cube-heat"])
@pytest.mark.parametrize("mode", [0, 1])
def test_spmv_bit_exact(hb, io, oracle, name, mode):
    A, X, B, ref = _assemble_exo(hb, io, oracle, name, mode)
    x = A.hash_vector(12345)
    y = A.new_vector()
    io.spmv(A, x, y)
    xh = x.numpy()
    assert np.abs(xh).max() <= 1.0 and abs(xh.mean()) < 0.05
    np.testing.assert_array_equal(y.numpy(), oracle.spmv(ref, xh))


@pytest.mark.parametrize("name", ["rectangle-tris-boundary", "bolted_bracket", "tet-cube-heat", "mitchell_tri"])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("solver", [0, 1])
def test_solve_matches_oracle(hb, io, oracle, golden, name, mode, solver):
    A, X, B, ref = _assemble_exo(hb, io, oracle, name, mode)
    res = io.solve(A, X, B, solver=solver, prec=hb.PREC_JACOBI, max_iters=2000, tol=RES_TOL, check_every=16)
    x_ref, it_ref, ach_ref, _ = oracle.pcg(ref, tol=RES_TOL, max_iters=2000)
    x = X.numpy()
    assert res.converged and res.achieved_tol <= RES_TOL
    assert abs(res.iters - it_ref) <= ITER_SLACK, (res.iters, it_ref)
    assert np.abs(x - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()
    # true residual, not the recurrence
    r = ref.b - ref.csr() @ x
    assert np.linalg.norm(r) <= 10 * RES_TOL * np.linalg.norm(ref.b)
    g = golden[name]["graph" if mode == 0 else "p1"]
    assert x.mean() == pytest.approx(g["x_mean"], rel=1e-7)
    assert x.min() == pytest.approx(g["x_min"], rel=1e-6) and x.max() == pytest.approx(g["x_max"], rel=1e-6)


def test_p1_analytic_linear_field(hb, io, oracle):
    # T = 550 - 90 x on tet-cube-heat.exo
    io.open(mesh_path("tet-cube-heat"), True)
    A, X, B = io.assemble(hb.OP_P1_FEM)
    res = io.solve(A, X, B, max_iters=3000, tol=1e-12)
    m = oracle.read_exodus(mesh_path("tet-cube-heat"))
    f = io.nodal_field(X, m.num_nodes)
    assert res.converged
    assert np.abs(f - (550.0 - 90.0 * m.x)).max() < 1e-7
    assert set(np.unique(f[m.nodesets[100]])) == {100.0} and set(np.unique(f[m.nodesets[1000]])) == {1000.0}


def test_iteration_counts_baseline_table(hb, io, oracle):
    # BASELINE.md §2
    expect = {("tet-cube-heat", 0): (105, 136), ("tet-cube-heat", 1): (140, 180), ("bolted_bracket", 0): (123, 137)}
    for (name, mode), (i8, i10) in expect.items():
        A, X, B, _ = _assemble_exo(hb, io, oracle, name, mode)
        assert abs(io.solve(A, X, B, max_iters=1000, tol=1e-8).iters - i8) <= ITER_SLACK
        X.fill(0.0)
        assert abs(io.solve(A, X, B, max_iters=1000, tol=1e-10).iters - i10) <= ITER_SLACK


def test_max_iters_and_status(hb, io, oracle):
    A, X, B, ref = _assemble_exo(hb, io, oracle, "bolted_bracket", 0)
    res = io.solve(A, X, B, max_iters=7, tol=1e-14, check_every=3)
    x_ref, it_ref, ach_ref, hist = oracle.pcg(ref, tol=1e-14, max_iters=7)
    assert (res.iters, res.converged) == (7, False) and it_ref == 7
    assert res.achieved_tol == pytest.approx(ach_ref, rel=1e-9)
    assert np.abs(X.numpy() - x_ref).max() <= 1e-12 * np.abs(x_ref).max()
    # zero iterations: x untouched, achieved_tol = 1
    X.fill(0.0)
    res0 = io.solve(A, X, B, max_iters=0, tol=1e-10)
    assert res0.iters == 0 and not res0.converged and res0.achieved_tol == pytest.approx(1.0)


def test_solve_host_end_to_end(hb, io, oracle):
    A, X, B, ref = _assemble_exo(hb, io, oracle, "tet-cube-heat", 1)
    xh = np.zeros(ref.n)
    res = io.solve_host(A, ref.b.copy(), xh, max_iters=1000, tol=RES_TOL)
    x_ref = oracle.pcg(ref, tol=RES_TOL, max_iters=1000)[0]
    assert res.converged and np.abs(xh - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()


def test_solve_host_batch_equals_single_calls(hb, io, oracle):
    """heat_solve_host_batch (copies of neighbouring systems overlapped with each solve) returns, for every system, the
    bits of a heat_solve_host call on it; zero start where no x0 is given."""
    A, X, B, ref = _assemble_exo(hb, io, oracle, "tet-cube-heat", 1)
    rng = np.random.default_rng(4)
    bs = [ref.b * (1.0 + 0.1 * k) + rng.standard_normal(ref.n) * (k % 2) for k in range(5)]
    x0s = [None, rng.standard_normal(ref.n), None, np.zeros(ref.n), rng.standard_normal(ref.n)]
    outs = [np.empty(ref.n) for _ in bs]
    rs = io.solve_host_batch(A, bs, x0s, outs, max_iters=400, tol=RES_TOL)
    for k, b in enumerate(bs):
        x = np.zeros(ref.n) if x0s[k] is None else x0s[k].copy()
        r1 = io.solve_host(A, b, x, max_iters=400, tol=RES_TOL)
        assert (rs[k].iters, rs[k].converged) == (r1.iters, r1.converged)
        np.testing.assert_array_equal(outs[k], x)
        x_ref = oracle.pcg(type(ref)(ref.n, ref.row_ptr, ref.col, ref.val, b, ref.red2orig, ref.node_bc),
                           x0=np.zeros(ref.n) if x0s[k] is None else x0s[k], tol=RES_TOL, max_iters=400)[0]
        assert np.abs(outs[k] - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()
    assert io.solve_host_batch(A, [], None, []) == []


def test_chebyshev_matches_oracle(hb, io, oracle):
    A, X, B, ref = _assemble_exo(hb, io, oracle, "bolted_bracket", 0)
    res = io.solve(A, X, B, prec=hb.PREC_CHEBYSHEV, cheb_degree=3, cheb_lambda_max=2.0, max_iters=500, tol=RES_TOL)
    x_ref, it_ref, *_ = oracle.pcg(ref, tol=RES_TOL, prec=oracle.PREC_CHEBYSHEV, cheb_degree=3, cheb_lambda_max=2.0)
    assert res.converged and abs(res.iters - it_ref) <= ITER_SLACK
    assert np.abs(X.numpy() - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()
    # estimated lambda_max (power iterations) still converges to the same solution
    X.fill(0.0)
    res2 = io.solve(A, X, B, prec=hb.PREC_CHEBYSHEV, cheb_degree=2, cheb_lambda_max=0.0, max_iters=500, tol=RES_TOL)
    assert res2.converged and np.abs(X.numpy() - x_ref).max() <= 1e-7 * np.abs(x_ref).max()


@pytest.mark.parametrize("degree", [1, 2, 3, 4])
@pytest.mark.parametrize("name,mode", [("tet-cube-heat", 1), ("bolted_bracket", 0)])
def test_chebyshev_fused_path_matches_oracle_and_plain_path(hb, oracle, name, mode, degree):
    """Chebyshev-PCG with the polynomial folded into the CG kernels (cheb_xr_first / cheb_step, the path the multi-GPU
    runs take over peer memory; here its one-GPU form) against the oracle's Ifpack2-style recurrences and against the
    plain kernel sequence (HEAT_CHEB_FUSED=0)."""
    ref = oracle.assemble(oracle.read_exodus(mesh_path(name)), mode)
    x_ref, it_ref, *_ = oracle.pcg(ref, tol=RES_TOL, prec=oracle.PREC_CHEBYSHEV, cheb_degree=degree, cheb_lambda_max=2.2, max_iters=3000)
    got = {}
    for fused in ("1", "0"):
        old = os.environ.get("HEAT_CHEB_FUSED")
        os.environ["HEAT_CHEB_FUSED"] = fused
        try:
            io = hb.IO(0)
            io.open(mesh_path(name), True)
            A, X, B = io.assemble(mode)
            res = io.solve(A, X, B, prec=hb.PREC_CHEBYSHEV, cheb_degree=degree, cheb_lambda_max=2.2, max_iters=3000, tol=RES_TOL, check_every=7)
        finally:
            os.environ.pop("HEAT_CHEB_FUSED") if old is None else os.environ.__setitem__("HEAT_CHEB_FUSED", old)
        assert res.converged and abs(res.iters - it_ref) <= ITER_SLACK, (fused, res, it_ref)
        assert np.abs(X.numpy() - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max(), fused
        got[fused] = res.iters
        io.close()
    assert abs(got["1"] - got["0"]) <= 1


def test_no_preconditioner(hb, io, oracle):
    A, X, B, ref = _assemble_exo(hb, io, oracle, "bolted_bracket", 1)
    res = io.solve(A, X, B, prec=hb.PREC_NONE, max_iters=3000, tol=RES_TOL)
    x_ref, it_ref, *_ = oracle.pcg(ref, tol=RES_TOL, prec=oracle.PREC_NONE, max_iters=3000)
    assert res.converged and abs(res.iters - it_ref) <= max(ITER_SLACK, it_ref // 50)
    assert np.abs(X.numpy() - x_ref).max() <= 1e-7 * np.abs(x_ref).max()


# ---- synthetic Kuhn cubes (BASELINE.json configs[2..4]) -------------------------------------------
@pytest.mark.parametrize("dims", [(5, 5, 5), (9, 9, 9), (17, 17, 17), (7, 5, 6), (3, 2, 2), (33, 9, 5)])
@pytest.mark.parametrize("mode", [0, 1])
def test_cube_analytic_and_explicit_match_oracle(hb, oracle, dims, mode):
    nx, ny, nz = dims
    ref = oracle.assemble(oracle.cube_mesh(nx, ny, nz), mode)
    for explicit in (False, True):
        io = hb.IO(0)
        io.mesh_cube(nx, ny, nz, explicit)
        A, X, B = io.assemble(mode)
        rp, col, val = A.csr()
        np.testing.assert_array_equal(rp, ref.row_ptr)
        np.testing.assert_array_equal(col, ref.col)
        np.testing.assert_array_equal(val, ref.val)
        np.testing.assert_array_equal(B.numpy(), ref.b)
        np.testing.assert_array_equal(A.red2orig(), ref.red2orig)
        assert A.info.nnz_global == ref.nnz
        io.close()


@pytest.mark.parametrize("dims", [(9, 9, 9), (40, 33, 29), (66, 3, 2), (130, 5, 4), (12, 9, 11)])
@pytest.mark.parametrize("mode", [0, 1])
def test_cube_direct_sell_assembly(hb, oracle, dims, mode):
    """Analytic cubes are assembled STRAIGHT into the SpMV format (cube_sell_kernel: SELL values + byte indices +
    tables + diagonal + row lengths, no CSR).  The CSR a parity export asks for is rebuilt from those arrays
    (sell_to_csr) and must be the oracle's, bit for bit; so must the SpMV that streams them; and the CSR-first path
    (HEAT_CUBE_ASSEMBLY=csr) and the int32 column stream (HEAT_SPMV_CIDX=0) must give the same bits."""
    ref = oracle.assemble(oracle.cube_mesh(*dims), mode)
    xg = oracle.hash_vector(np.arange(ref.n), 31)
    y_ref = oracle.spmv(ref, xg)
    seen = {}
    for tag, env in (("direct", {}), ("direct-int32", {"HEAT_SPMV_CIDX": "0"}), ("csr-first", {"HEAT_CUBE_ASSEMBLY": "csr"})):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            io = hb.IO(0)
            io.mesh_cube(*dims)
            A, X, B = io.assemble(mode)
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        mi = A.info
        assert bool(mi.csr_resident) == (tag == "csr-first"), tag               # direct: no CSR until asked
        assert mi.col_index_bytes == (4 if tag == "direct-int32" else 1), tag
        x, y = A.hash_vector(31), A.new_vector()
        np.testing.assert_array_equal(x.numpy(), xg)
        io.spmv(A, x, y)                                                       # before any export: streams what the kernel wrote
        np.testing.assert_array_equal(y.numpy(), y_ref, err_msg=tag)
        np.testing.assert_array_equal(B.numpy(), ref.b, err_msg=tag)
        bytes_before = mi.matrix_bytes
        rp, col, val = A.csr()
        np.testing.assert_array_equal(rp, ref.row_ptr, err_msg=tag)
        np.testing.assert_array_equal(col, ref.col, err_msg=tag)
        assert val.tobytes() == ref.val.tobytes(), tag                         # zero signs included
        assert A.info.csr_resident == 1 and A.info.nnz_local == ref.nnz
        if tag == "direct":
            assert A.info.matrix_bytes > bytes_before                          # the export built the CSR
        res = io.solve(A, X, B, max_iters=2000, tol=RES_TOL)
        seen[tag] = (res.iters, X.numpy().tobytes())
        io.close()
    # same system, same SpMV bits; the int32 kernel runs another grid shape, so the p.Ap partial sums are added in
    # another order: iteration counts equal, solutions equal to rounding
    its = {v[0] for v in seen.values()}
    xs = [np.frombuffer(v[1]) for v in seen.values()]
    assert len(its) == 1 and all(np.abs(x - xs[0]).max() <= 1e-12 * np.abs(xs[0]).max() for x in xs)
    assert seen["direct"] == seen["csr-first"]              # same kernels downstream of assembly: same bits


@pytest.mark.parametrize("mode,solver", [(0, 0), (1, 0), (1, 1)])
def test_cube_solve_matches_oracle(hb, oracle, mode, solver):
    n = 17
    ref = oracle.assemble(oracle.cube_mesh(n, n, n), mode)
    io = hb.IO(0)
    io.mesh_cube(n, n, n)
    A, X, B = io.assemble(mode)
    res = io.solve(A, X, B, solver=solver, max_iters=1000, tol=RES_TOL)
    x_ref, it_ref, *_ = oracle.pcg(ref, tol=RES_TOL)
    assert res.converged and abs(res.iters - it_ref) <= ITER_SLACK
    assert np.abs(X.numpy() - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()
    if mode == 1:   # discrete P1 solution is exactly linear
        f = io.nodal_field(X, n ** 3)
        i = np.arange(n ** 3) % n
        assert np.abs(f - (1000.0 - 900.0 * i / (n - 1))).max() < 1e-6
    io.close()


def test_full_size_256_properties(hb):
    """BASELINE.json configs[2] at FULL size (16.6 M DOF) through size-independent properties."""
    n = 256
    io = hb.IO(0)
    io.mesh_cube(n, n, n)
    A, X, B = io.assemble(hb.OP_P1_FEM)
    mi = A.info
    assert (mi.n_global, mi.nnz_global) == (16_646_144, 248_130_550)
    # (1) the exact discrete solution is linear in i: A x* = b  =>  residual ~ rounding
    gid = np.arange(mi.n_owned)
    xstar = 1000.0 - 900.0 * ((gid % (n - 2)) + 1) / (n - 1)
    xs = A.new_vector().set(xstar)
    y = A.new_vector()
    io.spmv(A, xs, y)
    b = B.numpy()
    assert np.abs(y.numpy() - b).max() <= 1e-9 * max(1.0, np.abs(b).max())
    # (2) linearity of the SpMV: A(2u + 3v) = 2Au + 3Av
    u, v = A.hash_vector(1), A.hash_vector(2)
    yu, yv, yw = A.new_vector(), A.new_vector(), A.new_vector()
    io.spmv(A, u, yu); io.spmv(A, v, yv)
    w = A.new_vector().set(2.0 * u.numpy() + 3.0 * v.numpy())
    io.spmv(A, w, yw)
    lhs, rhs = yw.numpy(), 2.0 * yu.numpy() + 3.0 * yv.numpy()
    assert np.abs(lhs - rhs).max() <= 1e-12 * np.abs(rhs).max()
    # (3) PCG drives the recurrence AND the true residual down; both solvers agree
    r1 = io.solve(A, X, B, solver=hb.SOLVER_CG, max_iters=3000, tol=1e-10)
    assert r1.converged
    io.spmv(A, X, y)
    assert np.linalg.norm(y.numpy() - b) <= 1e-9 * np.linalg.norm(b)
    assert np.abs(X.numpy() - xstar).max() <= 1e-5
    X2 = A.new_vector()
    r2 = io.solve(A, X2, B, solver=hb.SOLVER_CG_SINGLE_REDUCE, max_iters=3000, tol=1e-10)
    assert r2.converged and abs(r2.iters - r1.iters) <= ITER_SLACK
    io.close()


def test_full_size_512_headline_config_properties(hb):
    """BASELINE.json configs[3] on ONE GPU at FULL size — the workload bench.py times (512^3 nodes, 133.7 M DOF, 2.0 G
    nonzeros, byte-indexed SELL stream) — through size-independent properties: the documented counts, the exact
    discrete solution (A x* = b to rounding: one SpMV checks all 2 G matrix entries and the RHS against the analytic
    field), linearity of the SpMV, symmetry (u.Av = v.Au), and 100 PCG iterations whose recurrence residual equals the
    true one and which reduce the energy norm of the error."""
    n = 512
    io = hb.IO(0)
    io.mesh_cube(n, n, n)
    A, X, B = io.assemble(hb.OP_P1_FEM)
    mi = A.info
    assert (mi.n_global, mi.nnz_global, mi.n_owned) == (133_693_440, 1_999_132_662, 133_693_440)
    assert mi.col_index_bytes == 1                                         # the bench's column stream
    gid = np.arange(mi.n_owned)
    xstar = 1000.0 - 900.0 * ((gid % (n - 2)) + 1) / (n - 1)
    del gid
    xs, y = A.new_vector().set(xstar), A.new_vector()
    io.spmv(A, xs, y)
    b = B.numpy()
    bmax = np.abs(b).max()
    assert np.abs(y.numpy() - b).max() <= 1e-9 * bmax
    del xstar
    u, v = A.hash_vector(1), A.hash_vector(2)
    yu, yv = A.new_vector(), A.new_vector()
    io.spmv(A, u, yu); io.spmv(A, v, yv)
    uh, vh, yuh, yvh = u.numpy(), v.numpy(), yu.numpy(), yv.numpy()
    assert abs(np.dot(uh, yvh) - np.dot(vh, yuh)) <= 1e-10 * (np.linalg.norm(uh) * np.linalg.norm(yvh))   # symmetry
    xs.set(2.0 * uh + 3.0 * vh)
    io.spmv(A, xs, y)
    rhs = 2.0 * yuh + 3.0 * yvh
    assert np.abs(y.numpy() - rhs).max() <= 1e-12 * np.abs(rhs).max()                                       # linearity
    del uh, vh, yuh, yvh, rhs
    r = io.cg_iterations(A, X, B, 100, solver=hb.SOLVER_CG, prec=hb.PREC_JACOBI)
    assert r.iters == 100
    io.spmv(A, X, y)
    resid = b - y.numpy()
    res = np.linalg.norm(resid) / np.linalg.norm(b)
    assert abs(res - r.achieved_tol) <= 1e-6 * res                         # the recurrence residual IS the true residual
    # CG minimises the energy norm of the error over the Krylov space: ||x* - x_100||_A^2 = (x* - x).(b - A x) must have
    # dropped below ||x*||_A^2 = x*.b (the residual 2-norm itself need not be monotone)
    gid = np.arange(mi.n_owned)
    err = 1000.0 - 900.0 * ((gid % (n - 2)) + 1) / (n - 1)
    e0 = float(np.dot(err, b))
    err -= X.numpy()
    e100 = float(np.dot(err, resid))
    assert 0.0 < e100 < e0, (e100, e0)
    io.close()


def test_full_size_256_graph_rowsums(hb):
    n = 256
    io = hb.IO(0)
    io.mesh_cube(n, n, n)
    A, X, B = io.assemble(hb.OP_GRAPH_LAPLACIAN)
    ones = A.new_vector().fill(1.0)
    y = A.new_vector()
    io.spmv(A, ones, y)
    # A.1 = number of Dirichlet neighbours; B = sum of their values => checksum of checksums
    rows = y.numpy()
    assert rows.min() >= 0 and rows.max() <= 5 and np.all(rows == np.round(rows))
    gid = np.arange(A.info.n_owned)
    i = (gid % (n - 2)) + 1
    assert np.all(rows[(i > 1) & (i < n - 2)] == 0)
    b = B.numpy()
    assert np.all(b[i == 1] == 1000.0 * rows[i == 1]) and np.all(b[i == n - 2] == 100.0 * rows[i == n - 2])
    io.close()


def test_write_solution_round_trip(hb, io, oracle, tmp_path):
    from scipy.io import netcdf_file
    io.open(mesh_path("bolted_bracket"), True)
    A, X, B = io.assemble(hb.OP_GRAPH_LAPLACIAN)
    out = str(tmp_path / "solution.exo")
    io.create(out)
    io.decompose(4)
    io.solve(A, X, B, max_iters=5, tol=1e-14)
    io.writeSolution(X, 0)
    io.solve(A, X, B, max_iters=500, tol=1e-10)
    io.writeSolution(X, 1)
    nc = netcdf_file(out, "r", mmap=False)
    name = b"".join(nc.variables["name_nod_var"].data[0]).split(b"\x00")[0].decode()
    assert name == "Steady-State Heat Solution"                       # ExodusIO.hpp:2032
    vals = np.array(nc.variables["vals_nod_var1"].data)
    assert vals.shape == (2, 4098)
    np.testing.assert_array_equal(np.array(nc.variables["time_whole"].data), [0.0, 1.0])
    ref = oracle.assemble(oracle.read_exodus(mesh_path("bolted_bracket")), 0)
    f = oracle.scatter_field(ref, X.numpy())
    np.testing.assert_array_equal(vals[1], f)
    assert nc.dimensions["num_el_blk"] == 4
    assert np.array(nc.variables["eb_prop1"].data).tolist() == [0, 1, 2, 3]     # block ids start at 0 (D12)
    nc.close()


def test_solve_trajectory_writes_every_k_iterations(hb, io, oracle, tmp_path):
    """belosSolver's per-iteration writeSolution (BelosMueLuSolver.cpp:113-133) in ONE Krylov run: frame k
    is the oracle's iterate after 25(k+1) iterations, the last frame is the converged solution, and the
    run is bit-identical to a plain solve (no restart)."""
    from scipy.io import netcdf_file
    io.open(mesh_path("tet-cube-heat"), True)
    A, X, B = io.assemble(hb.OP_P1_FEM)
    ref = oracle.assemble(oracle.read_exodus(mesh_path("tet-cube-heat")), oracle.P1_FEM)
    out = str(tmp_path / "trajectory.exo")
    io.create(out)
    io.decompose(2)
    res, frames = io.solve_trajectory(A, X, B, write_every=25, max_iters=1000, tol=RES_TOL)
    x_ref, it_ref, *_ = oracle.pcg(ref, tol=RES_TOL, max_iters=1000)
    assert res.converged and abs(res.iters - it_ref) <= ITER_SLACK
    assert frames == -(-res.iters // 25)
    nc = netcdf_file(out, "r", mmap=False)
    vals = np.array(nc.variables["vals_nod_var1"].data)
    np.testing.assert_array_equal(np.array(nc.variables["time_whole"].data), np.arange(frames, dtype=np.float64))
    nc.close()
    assert vals.shape == (frames, ref.node_bc.shape[0])
    np.testing.assert_array_equal(vals[-1], oracle.scatter_field(ref, X.numpy()))
    scale = np.abs(x_ref).max()
    for k in (0, 2, frames - 2):
        xk = oracle.pcg(ref, tol=0.0, max_iters=25 * (k + 1))[0]
        assert np.abs(vals[k] - oracle.scatter_field(ref, xk)).max() <= SOL_RTOL * scale, k
    err = [np.abs(vals[k] - vals[-1]).max() for k in range(frames)]
    assert err[0] > err[frames // 2] > err[-2] > 0
    x_traj = X.numpy().copy()
    X.fill(0.0)
    io.solve(A, X, B, max_iters=1000, tol=RES_TOL)
    np.testing.assert_array_equal(X.numpy(), x_traj)                 # same Krylov run, no restart
    # exactly on a poll boundary: no duplicate frame; max_iters == 0: one frame (x0)
    X.fill(0.0)
    _, fr = io.solve_trajectory(A, X, B, write_every=10, first_timestep=frames, max_iters=30, tol=0.0)
    assert fr == 3
    X.fill(0.0)
    _, fr0 = io.solve_trajectory(A, X, B, write_every=5, first_timestep=frames + 3, max_iters=0, tol=0.0)
    assert fr0 == 1
    nc = netcdf_file(out, "r", mmap=False)
    assert nc.variables["vals_nod_var1"].data.shape[0] == frames + 4
    nc.close()


# ---- edge cases ------------------------------------------------------------------------------------
def test_mesh_without_nodesets_is_the_full_laplacian(hb, io, oracle):
    """2blocks.exo: two TETRA4 blocks, no nodesets -> every node is a DOF, B = 0, A singular (A.1 = 0);
    the status test fires before the first iteration (r0 = 0)."""
    io.open(mesh_path("2blocks"), True)
    A, X, B = io.assemble(hb.OP_GRAPH_LAPLACIAN)
    ref = oracle.assemble(oracle.read_exodus(mesh_path("2blocks")), 0)
    rp, col, val = A.csr()
    np.testing.assert_array_equal(rp, ref.row_ptr); np.testing.assert_array_equal(col, ref.col)
    np.testing.assert_array_equal(val, ref.val)
    assert A.info.n_global == 34 and np.all(B.numpy() == 0)
    ones, y = A.new_vector().fill(1.0), A.new_vector()
    io.spmv(A, ones, y)
    assert np.all(y.numpy() == 0)
    res = io.solve(A, X, B, max_iters=10, tol=1e-10)
    assert res.iters == 0 and res.converged and np.all(X.numpy() == 0)


def test_mesh_from_memory_and_overlapping_nodesets(hb, io, oracle):
    """heat_mesh_set + a node in two nodesets: lowest id wins for RHS and output (D2, FIXED)."""
    m = oracle.cube_mesh(5, 4, 3)
    ns = {7: np.flatnonzero(m.x == -5.0), 3: np.flatnonzero((m.x == -5.0) & (m.y == -5.0)), 50: np.flatnonzero(m.x == 5.0)}
    m.nodesets = ns
    ref = oracle.assemble(m, 1)
    io.mesh_set(m.x, m.y, m.z, m.conn, ns)
    A, X, B = io.assemble(hb.OP_P1_FEM)
    rp, col, val = A.csr()
    np.testing.assert_array_equal(col, ref.col); np.testing.assert_array_equal(val, ref.val)
    np.testing.assert_array_equal(B.numpy(), ref.b)
    io.solve(A, X, B, max_iters=200, tol=1e-12)
    f = io.nodal_field(X, m.num_nodes)
    assert np.all(f[ns[3]] == 3.0) and np.all(f[np.setdiff1d(ns[7], ns[3])] == 7.0) and np.all(f[ns[50]] == 50.0)


def test_p1_rejects_unsupported_elements(hb, io):
    x = np.array([0.0, 1.0, 2.0, 3.0])
    conn = np.array([[0, 1], [1, 2], [2, 3]], dtype=np.int32)           # 2-node "bars"
    io.mesh_set(x, x * 0, None, conn, {1: [0], 2: [3]}, num_dim=1)
    A, X, B = io.assemble(hb.OP_GRAPH_LAPLACIAN)                        # cliques of any size are fine
    np.testing.assert_array_equal(A.csr()[2], [2.0, -1.0, -1.0, 2.0])
    np.testing.assert_array_equal(B.numpy(), [1.0, 2.0])
    with pytest.raises(hb.HeatError, match="P1_FEM needs"):
        io.assemble(hb.OP_P1_FEM)


def test_hex_like_cliques_graph_mode(hb, io, oracle):
    """8-node elements (HEX8 connectivity treated as cliques, as assemble does for any element type)."""
    nx = 4
    idx = lambda i, j, k: i + nx * (j + nx * k)  # noqa: E731
    conn = np.array([[idx(i, j, k), idx(i + 1, j, k), idx(i + 1, j + 1, k), idx(i, j + 1, k),
                      idx(i, j, k + 1), idx(i + 1, j, k + 1), idx(i + 1, j + 1, k + 1), idx(i, j + 1, k + 1)]
                     for k in range(nx - 1) for j in range(nx - 1) for i in range(nx - 1)], dtype=np.int32)
    g = np.arange(nx ** 3)
    x, y, z = (g % nx).astype(float), ((g // nx) % nx).astype(float), (g // nx ** 2).astype(float)
    ns = {10: np.flatnonzero(x == 0), 20: np.flatnonzero(x == nx - 1)}
    m = oracle.Mesh(x, y, z, conn, ns, "HEX8", 3, [len(conn)])
    ref = oracle.assemble(m, 0)
    io.mesh_set(x, y, z, conn, ns)
    A, X, B = io.assemble(hb.OP_GRAPH_LAPLACIAN)
    rp, col, val = A.csr()
    np.testing.assert_array_equal(rp, ref.row_ptr); np.testing.assert_array_equal(col, ref.col)
    np.testing.assert_array_equal(val, ref.val); np.testing.assert_array_equal(B.numpy(), ref.b)
    assert A.info.max_row_len == 27 - 9        # interior node: 26 neighbours + itself, minus none... checked vs oracle
    res = io.solve(A, X, B, max_iters=500, tol=1e-10)
    xr = oracle.pcg(ref, tol=1e-10)[0]
    assert res.converged and np.abs(X.numpy() - xr).max() <= 1e-8 * np.abs(xr).max()


def test_spmv_variants_bit_identical(hb, oracle):
    """Every SpMV kernel variant (direct loads, TMA-staged rings) gives the same bits."""
    import subprocess, sys, json as _json
    code = ("import os,sys,hashlib;sys.path.insert(0,os.path.join(%r,'domain-decomposed-pde-solver_b200'));"
            "import heat_b200 as hb;io=hb.IO(0);io.mesh_cube(40,33,29);A,X,B=io.assemble(1);"
            "x=A.hash_vector(7);y=A.new_vector();io.spmv(A,x,y);r=io.cg_iterations(A,X,B,25);"
            "print(hashlib.sha1(y.numpy().tobytes()).hexdigest(), hashlib.sha1(X.numpy().tobytes()).hexdigest())")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = set()
    for v, cidx in [(v, "1") for v in range(6)] + [(5, "0"), (5, "2")]:      # (5, "1"): byte-indexed default; (5, "2"): int16 deltas
        env = dict(os.environ, HEAT_SPMV_VARIANT=str(v), HEAT_SPMV_CIDX=cidx)
        p = subprocess.run([sys.executable, "-c", code % root], capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.add(p.stdout.strip().split()[0])
    assert len(outs) == 1, outs


def test_byte_indexed_column_stream(hb, oracle):
    """The SpMV column stream shrinks to one byte per entry when every 64-row slice has at most 64
    distinct col-row offsets (structured meshes); otherwise the int32 stream is kept.  Either way the
    SpMV is bit-exact against the oracle."""
    for dims in ((40, 33, 29), (5, 5, 5), (66, 3, 2), (3, 2, 2)):
        io = hb.IO(0)
        io.mesh_cube(*dims)
        for mode in (0, 1):
            A, X, B = io.assemble(mode)
            # (3,2,2): 4 rows in one 64-row slice, the 60 tail rows all point at column 0: 60 distinct offsets plus the
            # stencil's do not fit a 64-entry table -> the slice stores int16 deltas
            assert A.info.col_index_bytes == (2 if dims == (3, 2, 2) else 1), dims
            ref = oracle.assemble(oracle.cube_mesh(*dims), mode)
            x, y = A.hash_vector(11), A.new_vector()
            io.spmv(A, x, y)
            np.testing.assert_array_equal(y.numpy(), oracle.spmv(ref, x.numpy()))
            pr = io.power_method(A, 3, 0.0, 5)                       # fused x.y / y.y variant of the same kernel
            lam = oracle.power_method(ref, oracle.hash_vector(np.arange(ref.n), 5), 3, 0.0)[0]
            assert abs(pr.lambda_ - lam) <= 1e-12 * abs(lam)
            for v in (x, y, X, B):
                v.free()
            A.free()
        io.close()
    io = hb.IO(0)
    io.open(mesh_path("tet-cube-heat"), True)                         # unstructured numbering: hundreds of offsets per slice
    A, X, B = io.assemble(0)
    assert A.info.col_index_bytes == 2                                # ... but every offset fits 16 bits: int16 deltas
    io.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_int16_delta_and_mixed_column_streams(hb, oracle, mode):
    """Per-slice choice of the column stream: unstructured numberings (the reference's meshes) stream int16 col-row
    deltas, structured slices 1-byte table indices, and one matrix may mix both; a bandwidth beyond 16 bits keeps int32.
    Whatever the stream, SpMV and assembled system are the oracle's, bit for bit."""
    def check(io, ref, expect):
        A, X, B = io.assemble(mode)
        assert A.info.col_index_bytes == expect, (A.info.col_index_bytes, expect)
        x, y = A.hash_vector(3), A.new_vector()
        io.spmv(A, x, y)
        np.testing.assert_array_equal(y.numpy(), oracle.spmv(ref, x.numpy()))
        pr = io.power_method(A, 3, 0.0, 5)                               # the x.y / y.y variant of the same kernel
        lam = oracle.power_method(ref, oracle.hash_vector(np.arange(ref.n), 5), 3, 0.0)[0]
        assert abs(pr.lambda_ - lam) <= 1e-12 * abs(lam)
        res = io.solve(A, X, B, max_iters=3000, tol=RES_TOL)
        x_ref, it_ref, *_ = oracle.pcg(ref, tol=RES_TOL, max_iters=3000)
        assert res.converged and abs(res.iters - it_ref) <= ITER_SLACK
        assert np.abs(X.numpy() - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()
        rp, col, val = A.csr()
        np.testing.assert_array_equal(col, ref.col); np.testing.assert_array_equal(val, ref.val)

    for name in ("tet-cube-heat", "bolted_bracket"):                     # all slices int16
        io = hb.IO(0)
        io.open(mesh_path(name), True)
        check(io, oracle.assemble(oracle.read_exodus(mesh_path(name)), mode), 2)
        io.close()
    # a structured cube whose upper half is renumbered at random: table-indexed slices below, int16 slices above
    m = oracle.cube_mesh(24, 20, 18)
    N = m.num_nodes
    perm = np.arange(N)
    rng = np.random.default_rng(5)
    hi = np.arange(N // 2, N)
    perm[hi] = rng.permutation(hi)                                       # new id of old node g
    inv = np.empty(N, dtype=np.int64); inv[perm] = np.arange(N)
    m2 = oracle.Mesh(m.x[inv], m.y[inv], m.z[inv], perm[m.conn].astype(np.int32), {k: np.sort(perm[v]) for k, v in m.nodesets.items()})
    io = hb.IO(0)
    io.mesh_set(m2.x, m2.y, m2.z, m2.conn, m2.nodesets)
    check(io, oracle.assemble(m2, mode), 2)
    io.close()
    # a numbering whose bandwidth exceeds 16 bits: int32 ids are kept
    m = oracle.cube_mesh(48, 40, 40)
    N = m.num_nodes
    perm = np.random.default_rng(9).permutation(N)
    inv = np.empty(N, dtype=np.int64); inv[perm] = np.arange(N)
    m3 = oracle.Mesh(m.x[inv], m.y[inv], m.z[inv], perm[m.conn].astype(np.int32), {k: np.sort(perm[v]) for k, v in m.nodesets.items()})
    io = hb.IO(0)
    io.mesh_set(m3.x, m3.y, m3.z, m3.conn, m3.nodesets)
    check(io, oracle.assemble(m3, mode), 4)
    io.close()
    # forced int16 on a structured cube (HEAT_SPMV_CIDX=2): same bits
    old = os.environ.get("HEAT_SPMV_CIDX")
    os.environ["HEAT_SPMV_CIDX"] = "2"
    try:
        io = hb.IO(0)
        io.mesh_cube(40, 33, 29)
        check(io, oracle.assemble(oracle.cube_mesh(40, 33, 29), mode), 2)
        io.close()
    finally:
        os.environ.pop("HEAT_SPMV_CIDX") if old is None else os.environ.__setitem__("HEAT_SPMV_CIDX", old)


def test_cli_driver_end_to_end(hb, oracle, tmp_path):
    """bin/heat_solver: the reference's BelosMueLuSolver call order from the command line."""
    import subprocess
    from scipy.io import netcdf_file
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "domain-decomposed-pde-solver_b200", "bin", "heat_solver")
    out = str(tmp_path / "solution.exo")
    p = subprocess.run([exe, f"--input={mesh_path('tet-cube-heat')}", f"--solution={out}", "--iterations=1000",
                        "--tolerance=1e-10", "--operator=p1", "--partitions=4"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "The Belos solve took 180 iteration(s) to reach a relative residual tolerance" in p.stdout or \
           "The Belos solve took 181" in p.stdout or "The Belos solve took 179" in p.stdout, p.stdout
    nc = netcdf_file(out, "r", mmap=False)
    f = np.array(nc.variables["vals_nod_var1"].data)[0]
    xs = np.array(nc.variables["coordx"].data)
    assert np.abs(f - (550.0 - 90.0 * xs)).max() < 1e-5
    assert nc.dimensions["num_el_blk"] == 4
    nc.close()


@pytest.mark.parametrize("name", ["bolted_bracket", "mitchell_tri", "2blocks", "rectangle-tris"])
def test_get_matrix_and_power_method(hb, io, oracle, golden, name):
    """IO::getMatrix + PowerMethod::run (ExodusIO.hpp:733, ExodusMatrixTest.cpp:56-129) on one GPU:
    the whole-mesh Laplacian bit-exact, lambda within 1e-10 relative, same stop iteration."""
    mesh = oracle.read_exodus(mesh_path(name))
    ref = oracle.get_matrix(mesh)
    io.open(mesh_path(name), True)
    A = io.getMatrix()
    rp, col, val = A.csr()
    assert np.array_equal(rp, ref.row_ptr) and np.array_equal(col, ref.col) and np.array_equal(val, ref.val)
    np.testing.assert_array_equal(A.red2orig(), np.arange(mesh.num_nodes))
    for sid, nodes in mesh.nodesets.items():
        np.testing.assert_array_equal(A.owned_nodeset(sid), np.unique(nodes))
    np.testing.assert_array_equal(io.nodeset_ids(), sorted(mesh.nodesets))
    z0 = oracle.hash_vector(np.arange(ref.n), 12345)
    for niters, tol in ((500, 1e-2), (7, 0.0), (120, 1e-9)):
        lam, res, it, conv = oracle.power_method(ref, z0, niters, tol)
        pr = io.power_method(A, niters, tol, 12345)
        assert (pr.iters, pr.converged) == (it, conv), (pr, it, conv)
        assert abs(pr.lambda_ - lam) <= 1e-10 * abs(lam), (pr, lam)
        assert abs(pr.residual - res) <= 1e-6 * max(res, 1e-12) + 1e-9, (pr, res)
    if name in golden["get_matrix"]:                                # independent pin: scipy's largest eigenvalue
        g = golden["get_matrix"][name]
        assert (A.info.n_global, A.info.nnz_global) == (g["n"], g["nnz"])
        pr = io.power_method(A, 2000, 1e-6, 12345)
        assert -1e-12 * pr.lambda_ <= g["lambda_max"] - pr.lambda_ <= 1e-5 * g["lambda_max"], (pr, g["lambda_max"])
    # P1 stiffness of the whole mesh (Neumann problem): rows sum to ~0
    if mesh.conn.shape[1] in (3, 4):
        K = io.getMatrix(hb.OP_P1_FEM)
        refK = oracle.get_matrix(mesh, oracle.P1_FEM)
        assert np.array_equal(K.csr()[2], refK.val)
        K.free()
    A.free()


def test_power_method_full_size_256(hb):
    """size-independent property at BASELINE configs[2] size: for the P1 operator of the Kuhn cube
    (h x 7-point stencil) Gershgorin bounds lambda_max by 12 h, and the Rayleigh quotient grows
    monotonically with the iteration count."""
    io = hb.IO(0)
    io.mesh_cube(256, 256, 256)
    A, X, B = io.assemble(hb.OP_P1_FEM, hb.PART_SLAB)
    h = 10.0 / 255.0
    lam = [io.power_method(A, k, 0.0, 3).lambda_ for k in (5, 20, 60)]
    assert lam[0] < lam[1] < lam[2] <= 12.0 * h * (1 + 1e-12)
    assert lam[2] > 0.8 * 12.0 * h
    for v in (X, B):
        v.free()
    A.free()
    io.close()


def test_matrix_test_cli(hb, oracle):
    """bin/heat_matrix_test: the reference's ExodusMatrixTest call order and progress lines."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "domain-decomposed-pde-solver_b200", "bin", "heat_matrix_test")
    p = subprocess.run([exe, f"--input={mesh_path('bolted_bracket')}", "--verbose"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    ref = oracle.get_matrix(oracle.read_exodus(mesh_path("bolted_bracket")))
    lam, res, it, conv = oracle.power_method(ref, oracle.hash_vector(np.arange(ref.n), 12345), 500, 1e-2)
    assert f"Converged after {it} iterations" in p.stdout, p.stdout
    assert "Iteration 0:" in p.stdout and "- ||A*q - lambda*q||_2 = " in p.stdout
    got = float(re.search(r"Lambda = ([-0-9.e+]+)", p.stdout).group(1))
    assert abs(got - lam) <= 1e-5 * lam          # printed with 6 significant digits


@pytest.mark.parametrize("name,mode", [("bolted_bracket", 0), ("tet-cube-heat", 1), ("rectangle-tris-boundary", 0)])
def test_ilu0_factors_bit_exact(hb, io, oracle, name, mode):
    """Device ILU(0) (level-scheduled, contraction-free) == the oracle's sequential IKJ factorisation, bit for bit."""
    A, X, B, ref = _assemble_exo(hb, io, oracle, name, mode)
    res = io.solve(A, X, B, solver=hb.SOLVER_GMRES, prec=hb.PREC_ILU0, max_iters=400, tol=RES_TOL)
    lu, nl, nu = A.ilu0()
    np.testing.assert_array_equal(lu, oracle.ilu0(ref))
    assert nl >= 1 and nu >= 1 and nl <= ref.n and nu <= ref.n
    x_ref, it_ref, ach, conv = oracle.gmres(ref, prec=oracle.PREC_ILU0, restart=300, max_iters=400, tol=RES_TOL)
    assert res.converged and conv and abs(res.iters - it_ref) <= ITER_SLACK, (res, it_ref)
    assert np.abs(X.numpy() - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()


@pytest.mark.parametrize("prec,restart", [(3, 1), (1, 30), (0, 50), (2, 20), (3, 10)])
def test_gmres_matches_oracle(hb, io, oracle, prec, restart):
    """Restarted right-preconditioned GMRES against the oracle on tet-cube-heat (P1): iterations +-2,
    solution within 1e-8 relative.  (3, 1) is GMRES(1) + ILU — the iteration the reference's loop performs."""
    A, X, B, ref = _assemble_exo(hb, io, oracle, "tet-cube-heat", 1)
    kw = dict(cheb_degree=3, cheb_lambda_max=2.0)
    res = io.solve(A, X, B, solver=hb.SOLVER_GMRES, prec=prec, gmres_restart=restart, max_iters=2000, tol=RES_TOL, **kw)
    x_ref, it_ref, ach, conv = oracle.gmres(ref, prec=prec, restart=restart, max_iters=2000, tol=RES_TOL, **kw)
    assert res.converged and conv, (res, it_ref, ach)
    # GMRES(1) is a minimal-residual iteration whose step lengths depend non-linearly on the residual: its
    # iteration count moves by ~10 % under perturbations of 1e-14 (measured on the oracle alone: 194, 196,
    # 212, 200 iterations for relative perturbations 0, 1e-16, 1e-14, 1e-12 of b), so only the solution and
    # the order of magnitude of the count can be pinned for it; longer restarts are held to +-2
    slack = it_ref // 5 if restart == 1 else max(ITER_SLACK, it_ref // 100)
    assert abs(res.iters - it_ref) <= slack, (res, it_ref)
    assert np.abs(X.numpy() - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()
    assert res.achieved_tol <= RES_TOL
    # hitting the iteration limit is reported, not an error
    X.fill(0.0)
    r2 = io.solve(A, X, B, solver=hb.SOLVER_GMRES, prec=prec, gmres_restart=restart, max_iters=7, tol=1e-14, **kw)
    xo, ito, acho, convo = oracle.gmres(ref, prec=prec, restart=restart, max_iters=7, tol=1e-14, **kw)
    assert (r2.iters, r2.converged) == (ito, convo) == (7, False)
    assert abs(r2.achieved_tol - acho) <= 1e-6 * acho
    assert np.abs(X.numpy() - xo).max() <= 1e-10 * np.abs(xo).max()


def test_cg_with_ilu0_matches_oracle(hb, io, oracle):
    A, X, B, ref = _assemble_exo(hb, io, oracle, "tet-cube-heat", 1)
    res = io.solve(A, X, B, prec=hb.PREC_ILU0, max_iters=1000, tol=RES_TOL)
    x_ref, it_ref, *_ = oracle.pcg(ref, prec=oracle.PREC_ILU0, tol=RES_TOL, max_iters=1000)
    assert res.converged and abs(res.iters - it_ref) <= ITER_SLACK, (res, it_ref)
    assert np.abs(X.numpy() - x_ref).max() <= SOL_RTOL * np.abs(x_ref).max()


def test_reference_loop_cli(hb, oracle, tmp_path):
    """heat_solver --reference-loop: the loop of BelosMueLuSolver.cpp:113-133 as written — one GMRES
    iteration (right-preconditioned by the incomplete factorisation) per solve(), the field written after
    every pass, and the reference's closing line."""
    import subprocess
    from scipy.io import netcdf_file
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "domain-decomposed-pde-solver_b200", "bin", "heat_solver")
    out = str(tmp_path / "loop.exo")
    n_it = 40
    p = subprocess.run([exe, f"--input={mesh_path('tet-cube-heat')}", f"--solution={out}", f"--iterations={n_it}",
                        "--tolerance=1e-14", "--reference-loop"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert f"The Belos solve took {n_it} iteration(s), but did not converge. Achieved tolerance = " in p.stdout, p.stdout
    nc = netcdf_file(out, "r", mmap=False)
    vals = np.array(nc.variables["vals_nod_var1"].data)
    nc.close()
    assert vals.shape[0] == n_it
    ref = oracle.assemble(oracle.read_exodus(mesh_path("tet-cube-heat")), 0)
    xk = np.zeros(ref.n)
    for _ in range(n_it):
        xk = oracle.gmres(ref, x0=xk, prec=oracle.PREC_ILU0, restart=1, max_iters=1, tol=1e-14)[0]
    fk = oracle.scatter_field(ref, xk)
    assert np.abs(vals[-1] - fk).max() <= 1e-8 * np.abs(fk).max()
    A = ref.csr()
    res = [np.linalg.norm(ref.b - A @ v[ref.red2orig]) for v in vals[::10]]
    assert all(b < a for a, b in zip(res, res[1:]))


def test_cli_dump_matches_reference_format(hb, oracle, tmp_path):
    """heat_solver --dump writes "<prefix><rank>.out" in the format the reference's mpi_output_combiner.py
    consumes (BelosMueLuSolver.cpp:37-84): "[Section]" headers, then "row: [(col,val),...] ~usec~" /
    "row: [val] ~usec~" lines with global ids and ascending columns."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "domain-decomposed-pde-solver_b200", "bin", "heat_solver")
    prefix = str(tmp_path / "mpi-proc-")
    p = subprocess.run([exe, f"--input={mesh_path('rectangle-tris-boundary')}", f"--solution={tmp_path / 's.exo'}",
                        f"--outputPrefix={prefix}", "--dump", "--tolerance=1e-12"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    lines = open(prefix + "0.out").read().splitlines(keepends=True)
    sections, cur, stamps = {}, None, []
    for l in lines:
        if re.match(r"\[.*\]", l):                         # the combiner's section test
            cur = l.strip()[1:-1]
            sections[cur] = []
            continue
        m = re.search(r"~([0-9]*)~\n", l)                   # the combiner's timestamp extraction
        assert m, l
        stamps.append(int(m.group(1)))
        sections[cur].append(re.sub(r"~[0-9]*~", "", l).strip())
    assert list(sections) == ["Laplacian: A", "RHS: B", "Solution: X"]
    assert stamps == sorted(stamps)
    assert sections["Laplacian: A"] == ["0: [(0,5),(2,-1)]", "1: [(1,4),(2,-1)]", "2: [(0,-1),(1,-1),(2,5)]"]
    assert sections["RHS: B"] == ["0: [500]", "1: [450]", "2: [300]"]
    x = [float(re.match(r"\d+: \[(.*)\]", l).group(1)) for l in sections["Solution: X"]]
    np.testing.assert_allclose(x, [122.527472527, 140.659340659, 112.637362637], rtol=1e-5)
