"""CPU ORACLE — Python face (TEST INFRASTRUCTURE ONLY; never imported by the product).

Wraps oracle/_build/libheat_oracle.so (plain-C restatement of ExodusIO.hpp:128-723 and the
CG form of BelosMueLuSolver.cpp:87-139) and adds an INDEPENDENT numpy/scipy restatement
(`assemble_np`) plus a scipy Exodus reader, used to pin the C oracle.

PARITY STATUS.  Assembly / getMatrix / decompose / writeSolution (ExodusIO.hpp): PINNED against the
reference's own code — oracle/_ref/ref_driver runs /root/reference/ExodusIO.hpp, unmodified, on one
rank (oracle/ref_shim/README.md); its outputs on all 17 meshes of the reference's data/ are committed as
tests/golden/ref_pins.json and reproduced bit for bit (tests/test_reference_pins.py).  The Krylov solve
(Belos/Ifpack2, third party, absent) stays "parity unpinned" against a reference binary — pinned against
hand-checked systems, scipy direct solves and analytic P1 solutions instead (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libheat_oracle.so")
_METIS = os.path.join(_HERE, "_build", "libmetis_oracle.so")

GRAPH_LAPLACIAN, P1_FEM = 0, 1
PREC_NONE, PREC_JACOBI, PREC_CHEBYSHEV, PREC_ILU0 = 0, 1, 2, 3


def build(force: bool = False) -> None:
    """Compile the C oracle (gcc).  Building the checker is not using it."""
    if force or not (os.path.exists(_LIB) and os.path.exists(_METIS)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in ("heat_oracle.c", "heat_oracle.h", "Makefile")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE])


class _System(C.Structure):
    _fields_ = [
        ("num_nodes", C.c_int64), ("n", C.c_int64), ("nnz", C.c_int64),
        ("row_ptr", C.POINTER(C.c_int64)), ("col", C.POINTER(C.c_int32)),
        ("val", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double)),
        ("red2orig", C.POINTER(C.c_int64)), ("max_row", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        _lib.oracle_assemble.argtypes = [C.c_int64, dp, dp, dp, C.c_int64, C.c_int, ip, dp, C.c_int, C.POINTER(_System)]
        _lib.oracle_assemble.restype = C.c_int
        _lib.oracle_system_free.argtypes = [C.POINTER(_System)]
        _lib.oracle_cube_mesh.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, ip, dp]
        _lib.oracle_spmv.argtypes = [C.c_int64, lp, ip, dp, dp, dp]
        _lib.oracle_pcg.argtypes = [C.c_int64, lp, ip, dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, dp, dp]
        _lib.oracle_pcg.restype = C.c_int
        _lib.oracle_scatter_field.argtypes = [C.c_int64, dp, C.c_int64, lp, dp, dp]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_num_threads.argtypes = [C.c_int]
        _lib.oracle_pcg_loop_seconds.restype = C.c_double
        _lib.oracle_cube_assemble.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_System)]
        _lib.oracle_cube_assemble.restype = C.c_int
        _lib.oracle_ilu0.argtypes = [C.c_int64, lp, ip, dp, dp]
        _lib.oracle_ilu0.restype = C.c_int
        _lib.oracle_ilu0_apply.argtypes = [C.c_int64, lp, ip, dp, dp, dp]
        _lib.oracle_gmres.argtypes = [C.c_int64, lp, ip, dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                      C.c_int, C.c_double, dp, C.POINTER(C.c_int)]
        _lib.oracle_gmres.restype = C.c_int
        _lib.oracle_power_method.argtypes = [C.c_int64, lp, ip, dp, dp, C.c_int, C.c_double, dp, dp, C.POINTER(C.c_int)]
        _lib.oracle_power_method.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ------------------------------------------------------------------------------------------
# Mesh containers
# ------------------------------------------------------------------------------------------
@dataclass
class Mesh:
    """What IO::assemble reads (ExodusIO.hpp:143-192, :342-359): coords, one connectivity
    table (0-based), nodesets id -> 0-based node arrays."""
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray | None
    conn: np.ndarray                 # [ne, npe] int32, 0-based
    nodesets: dict = field(default_factory=dict)
    elem_type: str = "TETRA"
    num_dim: int = 3
    block_sizes: list = field(default_factory=list)

    @property
    def num_nodes(self):
        return int(self.x.shape[0])

    def node_bc(self) -> np.ndarray:
        """Prescribed temperature per node, NaN for DOF nodes.  A node in several nodesets takes
        the LOWEST id (the reference's RHS rule, ExodusIO.hpp:676-681: ascending std::map + break)."""
        bc = np.full(self.num_nodes, np.nan)
        for sid in sorted(self.nodesets, reverse=True):     # lowest id written last => wins
            bc[np.asarray(self.nodesets[sid], dtype=np.int64)] = float(sid)
        return bc


def _cstr(arr) -> str:
    raw = b"".join(np.asarray(arr).ravel().tolist())
    return raw.split(b"\x00")[0].decode("ascii", "replace").strip()


def read_exodus(path: str) -> Mesh:
    """Independent (scipy.io.netcdf_file) reader of the Exodus-II subset `assemble` consumes."""
    from scipy.io import netcdf_file

    nc = netcdf_file(path, "r", mmap=False)
    v = nc.variables
    N = nc.dimensions["num_nodes"]
    ndim = nc.dimensions["num_dim"]
    if "coordx" in v:
        x = np.array(v["coordx"].data, dtype=np.float64)
        y = np.array(v["coordy"].data, dtype=np.float64) if "coordy" in v else np.zeros(N)
        z = np.array(v["coordz"].data, dtype=np.float64) if "coordz" in v else None
    else:
        c = np.array(v["coord"].data, dtype=np.float64)
        x, y = c[0], (c[1] if ndim > 1 else np.zeros(N))
        z = c[2] if ndim > 2 else None
    nblk = nc.dimensions.get("num_el_blk", 0)
    conns, etype, sizes = [], "", []
    for b in range(1, nblk + 1):
        cv = v[f"connect{b}"]
        conns.append(np.array(cv.data, dtype=np.int32) - 1)
        et = getattr(cv, "elem_type", b"")
        etype = et.decode() if isinstance(et, bytes) else str(et)
        sizes.append(conns[-1].shape[0])
    npe = {c.shape[1] for c in conns}
    if len(npe) != 1:
        raise ValueError(f"mixed nodes-per-element {npe} not supported by the oracle")
    conn = np.ascontiguousarray(np.vstack(conns))
    nodesets = {}
    nns = nc.dimensions.get("num_node_sets", 0)
    if nns:
        ids = np.array(v["ns_prop1"].data, dtype=np.int64)
        for s in range(1, nns + 1):
            if f"node_ns{s}" in v:
                nodesets[int(ids[s - 1])] = np.array(v[f"node_ns{s}"].data, dtype=np.int64) - 1
            else:
                nodesets[int(ids[s - 1])] = np.zeros(0, dtype=np.int64)
    nc.close()
    return Mesh(x, y, z, conn, nodesets, etype.strip(), int(ndim), sizes)


def cube_mesh(nx: int, ny: int, nz: int) -> Mesh:
    """Explicit Kuhn tet cube from the C oracle."""
    N = nx * ny * nz
    ne = 6 * (nx - 1) * (ny - 1) * (nz - 1)
    x, y, z, bc = (np.empty(N) for _ in range(4))
    conn = np.empty((ne, 4), dtype=np.int32)
    lib().oracle_cube_mesh(nx, ny, nz, _p(x, C.c_double), _p(y, C.c_double), _p(z, C.c_double),
                           _p(conn, C.c_int32), _p(bc, C.c_double))
    ns = {1000: np.flatnonzero(bc == 1000.0), 100: np.flatnonzero(bc == 100.0)}
    return Mesh(x, y, z, conn, ns, "TETRA", 3, [ne])


# ------------------------------------------------------------------------------------------
# Assembly
# ------------------------------------------------------------------------------------------
@dataclass
class System:
    n: int
    row_ptr: np.ndarray   # int64 [n+1]
    col: np.ndarray       # int32 [nnz]
    val: np.ndarray       # float64 [nnz]
    b: np.ndarray         # float64 [n]
    red2orig: np.ndarray  # int64 [n]
    node_bc: np.ndarray   # float64 [N]

    @property
    def nnz(self):
        return int(self.row_ptr[-1])

    def csr(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val, self.col, self.row_ptr), shape=(self.n, self.n))

    def free(self):
        """release the C arrays of a cube_assemble(copy=False) system (the numpy views die with them)"""
        c = getattr(self, "_c", None)
        if c is not None:
            self.row_ptr = self.col = self.val = self.b = self.red2orig = None
            lib().oracle_system_free(C.byref(c))
            self._c = None


def assemble(mesh: Mesh, mode: int = GRAPH_LAPLACIAN, node_bc: np.ndarray | None = None) -> System:
    """C oracle of IO::assemble (FIXED semantics)."""
    bc = mesh.node_bc() if node_bc is None else np.ascontiguousarray(node_bc, dtype=np.float64)
    x = np.ascontiguousarray(mesh.x, dtype=np.float64)
    y = np.ascontiguousarray(mesh.y, dtype=np.float64)
    z = None if mesh.z is None else np.ascontiguousarray(mesh.z, dtype=np.float64)
    conn = np.ascontiguousarray(mesh.conn, dtype=np.int32)
    s = _System()
    rc = lib().oracle_assemble(mesh.num_nodes, _p(x, C.c_double), _p(y, C.c_double),
                               _p(z, C.c_double) if z is not None else None, conn.shape[0],
                               conn.shape[1], _p(conn, C.c_int32), _p(bc, C.c_double), mode, C.byref(s))
    if rc:
        raise RuntimeError(f"oracle_assemble failed rc={rc}")
    n, nnz = s.n, s.nnz
    out = System(
        n=n,
        row_ptr=np.ctypeslib.as_array(s.row_ptr, (n + 1,)).copy(),
        col=np.ctypeslib.as_array(s.col, (max(nnz, 1),))[:nnz].copy(),
        val=np.ctypeslib.as_array(s.val, (max(nnz, 1),))[:nnz].copy(),
        b=np.ctypeslib.as_array(s.b, (max(n, 1),))[:n].copy(),
        red2orig=np.ctypeslib.as_array(s.red2orig, (max(n, 1),))[:n].copy(),
        node_bc=bc,
    )
    lib().oracle_system_free(C.byref(s))
    return out


def cube_assemble(nx: int, ny: int, nz: int, mode: int = P1_FEM, copy: bool = True) -> System:
    """The system assemble(cube_mesh(nx, ny, nz), mode) builds, bit for bit, from the closed form of the Kuhn
    cube's connectivity (oracle_cube_assemble) — no explicit mesh, so the 512^3 headline size fits the host.
    copy=False wraps the C arrays without duplicating them (24.5 GB at 512^3); call .free() when done."""
    s = _System()
    rc = lib().oracle_cube_assemble(nx, ny, nz, mode, C.byref(s))
    if rc:
        raise RuntimeError(f"oracle_cube_assemble failed rc={rc}")
    n, nnz = s.n, s.nnz
    wrap = (lambda a: a.copy()) if copy else (lambda a: a)
    out = System(
        n=n,
        row_ptr=wrap(np.ctypeslib.as_array(s.row_ptr, (n + 1,))),
        col=wrap(np.ctypeslib.as_array(s.col, (nnz,))),
        val=wrap(np.ctypeslib.as_array(s.val, (nnz,))),
        b=wrap(np.ctypeslib.as_array(s.b, (n,))),
        red2orig=wrap(np.ctypeslib.as_array(s.red2orig, (n,))),
        node_bc=np.zeros(0),
    )
    if copy:
        lib().oracle_system_free(C.byref(s))
    else:
        out._c = s                    # keeps the C arrays alive until free()
    return out


def assemble_np(mesh: Mesh, mode: int = GRAPH_LAPLACIAN, bug_compat_d1: bool = False):
    """INDEPENDENT numpy/scipy restatement of IO::assemble (different algorithm: global edge
    list + np.unique + COO->CSR), used to pin the C oracle.  Returns (A csr, b, red2orig).

    bug_compat_d1 reproduces defect D1 for ONE rank (ExodusIO.hpp:220 `i < getMaxLocalIndex()`):
    the last node gets no reduced id but still counts as a DOF neighbour, so each of its
    neighbours receives a spurious -1 in column 0 (std::map operator[] default, :598).  Graph
    mode only."""
    import scipy.sparse as sp

    N = mesh.num_nodes
    bc = mesh.node_bc()
    isdof = np.isnan(bc)
    conn = mesh.conn.astype(np.int64)
    ne, npe = conn.shape
    ii, jj = np.meshgrid(np.arange(npe), np.arange(npe), indexing="ij")
    off = ii != jj
    rows = conn[:, ii[off]].ravel()
    cols = conn[:, jj[off]].ravel()
    keep = rows != cols
    pairs = np.unique(np.stack([rows[keep], cols[keep]], 1), axis=0)      # unique directed edges
    rows, cols = pairs[:, 0], pairs[:, 1]

    has_row = isdof.copy()
    if bug_compat_d1:
        if mode != GRAPH_LAPLACIAN:
            raise ValueError("bug-compat is defined for the reference's graph Laplacian only")
        has_row[N - 1] = False
    red = np.full(N, -1, dtype=np.int64)
    red[has_row] = np.arange(int(has_row.sum()))
    n = int(has_row.sum())
    red2orig = np.flatnonzero(has_row)

    rsel = has_row[rows]
    rows, cols = rows[rsel], cols[rsel]
    deg = np.bincount(red[rows], minlength=n).astype(np.float64)           # full degree
    dofn = isdof[cols]
    if mode == GRAPH_LAPLACIAN:
        c_red = red[cols[dofn]]
        c_red = np.where(c_red < 0, 0, c_red)                               # D1: missing -> column 0
        A = sp.coo_matrix((-np.ones(int(dofn.sum())), (red[rows[dofn]], c_red)), shape=(n, n))
        A = (A + sp.diags(deg)).tocsr()
        b = np.bincount(red[rows[~dofn]], weights=bc[cols[~dofn]], minlength=n)
        if not bug_compat_d1:
            # D3 (FIXED): keep the diagonal of rows with no DOF neighbour — sp.diags did that.
            pass
        A.sort_indices()
        return A, b, red2orig

    # ---- P1 FEM: classic vectorised element stiffness, then Dirichlet elimination ----
    X = np.stack([mesh.x, mesh.y, mesh.z if mesh.z is not None else np.zeros(N)], 1)
    P = X[conn]                                                             # [ne, npe, 3]
    if npe == 4:
        J = P[:, 1:, :] - P[:, :1, :]                                       # rows = edges
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)                                             # columns = grad(phi_1..3)
        g = np.concatenate([-Jinv.sum(2, keepdims=True), Jinv], 2).transpose(0, 2, 1)  # [ne,4,3]
        vol = np.abs(det) / 6.0
    elif npe == 3:
        J = (P[:, 1:, :2] - P[:, :1, :2])
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)
        g2 = np.concatenate([-Jinv.sum(2, keepdims=True), Jinv], 2).transpose(0, 2, 1)  # [ne,3,2]
        g = np.concatenate([g2, np.zeros((ne, 3, 1))], 2)
        vol = np.abs(det) / 2.0
    else:
        raise ValueError("P1 needs tets or tris")
    Ke = np.einsum("e,ead,ebd->eab", vol, g, g)
    R = np.repeat(conn[:, :, None], npe, 2).ravel()
    Cc = np.repeat(conn[:, None, :], npe, 1).ravel()
    K = sp.coo_matrix((Ke.ravel(), (R, Cc)), shape=(N, N)).tocsr()
    # pattern = graph pattern (explicit zeros kept), values from K
    patt = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(N, N)).tocsr() + sp.identity(N, format="csr")
    Kdd = K[red2orig][:, red2orig]
    pdd = patt[red2orig][:, red2orig].tocsr()
    pdd.data[:] = 0.0
    A = (pdd + Kdd).tocsr()
    # csr add drops nothing: structural zeros of pdd are kept as explicit entries
    A.sort_indices()
    dir_nodes = np.flatnonzero(~isdof)
    b = -(K[red2orig][:, dir_nodes] @ bc[dir_nodes])
    return A, np.asarray(b).ravel(), red2orig


# ------------------------------------------------------------------------------------------
# Linear algebra
# ------------------------------------------------------------------------------------------
def spmv(sys_: System, x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(sys_.n)
    lib().oracle_spmv(sys_.n, _p(sys_.row_ptr, C.c_int64), _p(sys_.col, C.c_int32),
                      _p(sys_.val, C.c_double), _p(x, C.c_double), _p(y, C.c_double))
    return y


def pcg(sys_: System, x0: np.ndarray | None = None, prec: int = PREC_JACOBI, max_iters: int = 1000,
        tol: float = 1e-8, cheb_degree: int = 1, cheb_lambda_max: float = 2.0, cheb_ratio: float = 30.0):
    """Belos-style classical PCG.  Returns (x, iters, achieved_tol, hist)."""
    x = np.zeros(sys_.n) if x0 is None else np.array(x0, dtype=np.float64)
    hist = np.zeros(max_iters + 1)
    ach = C.c_double(0.0)
    b = np.ascontiguousarray(sys_.b)
    it = lib().oracle_pcg(sys_.n, _p(sys_.row_ptr, C.c_int64), _p(sys_.col, C.c_int32),
                          _p(sys_.val, C.c_double), _p(b, C.c_double), _p(x, C.c_double), prec,
                          cheb_degree, cheb_lambda_max, cheb_ratio, max_iters, tol, C.byref(ach),
                          _p(hist, C.c_double))
    if it < 0:
        raise RuntimeError("oracle_pcg breakdown (p.Ap <= 0)")
    return x, it, ach.value, hist[: it + 1]


def scatter_field(sys_: System, x: np.ndarray) -> np.ndarray:
    """IO::writeSolution's dense nodal array (ExodusIO.hpp:1981-1989, :2045-2055)."""
    N = sys_.node_bc.shape[0]
    f = np.empty(N)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().oracle_scatter_field(N, _p(sys_.node_bc, C.c_double), sys_.n, _p(sys_.red2orig, C.c_int64),
                               _p(x, C.c_double), _p(f, C.c_double))
    return f


def ilu0(sys_: System) -> np.ndarray:
    """ILU(0) factors on the CSR pattern (IKJ order): strictly lower = L (unit diagonal), rest = U."""
    lu = np.empty(max(sys_.nnz, 1))
    rc = lib().oracle_ilu0(sys_.n, _p(sys_.row_ptr, C.c_int64), _p(sys_.col, C.c_int32), _p(sys_.val, C.c_double), _p(lu, C.c_double))
    if rc:
        raise RuntimeError("oracle_ilu0: missing diagonal")
    return lu[: sys_.nnz]


def ilu0_apply(sys_: System, lu: np.ndarray, v: np.ndarray) -> np.ndarray:
    v = np.ascontiguousarray(v, dtype=np.float64)
    lu = np.ascontiguousarray(lu, dtype=np.float64)
    z = np.empty(sys_.n)
    lib().oracle_ilu0_apply(sys_.n, _p(sys_.row_ptr, C.c_int64), _p(sys_.col, C.c_int32), _p(lu, C.c_double), _p(v, C.c_double), _p(z, C.c_double))
    return z


def gmres(sys_: System, x0: np.ndarray | None = None, prec: int = PREC_ILU0, restart: int = 300, max_iters: int = 300,
          tol: float = 1e-8, cheb_degree: int = 1, cheb_lambda_max: float = 2.0, cheb_ratio: float = 30.0):
    """Right-preconditioned restarted GMRES (Belos "GMRES" + setRightPrec, BelosMueLuSolver.cpp:102-109).
    Returns (x, inner iterations, achieved_tol, converged)."""
    x = np.zeros(sys_.n) if x0 is None else np.array(x0, dtype=np.float64)
    b = np.ascontiguousarray(sys_.b)
    ach, conv = C.c_double(0.0), C.c_int(0)
    it = lib().oracle_gmres(sys_.n, _p(sys_.row_ptr, C.c_int64), _p(sys_.col, C.c_int32), _p(sys_.val, C.c_double),
                            _p(b, C.c_double), _p(x, C.c_double), prec, cheb_degree, cheb_lambda_max, cheb_ratio, restart,
                            max_iters, tol, C.byref(ach), C.byref(conv))
    return x, int(it), ach.value, bool(conv.value)


def power_method(sys_: System, z0: np.ndarray, niters: int = 500, tolerance: float = 1.0e-2):
    """PowerMethod::run (ExodusMatrixTest.cpp:56-129).  Returns (lambda, residual, stop_iter, converged)."""
    z = np.array(z0, dtype=np.float64)
    lam, res, conv = C.c_double(0.0), C.c_double(0.0), C.c_int(0)
    it = lib().oracle_power_method(sys_.n, _p(sys_.row_ptr, C.c_int64), _p(sys_.col, C.c_int32), _p(sys_.val, C.c_double),
                                   _p(z, C.c_double), niters, tolerance, C.byref(lam), C.byref(res), C.byref(conv))
    return lam.value, res.value, int(it), bool(conv.value)


def hash_vector(gids, seed: int) -> np.ndarray:
    """numpy twin of the product's counter-based U(-1,1) vector keyed on the global row id
    (splitmix64(g ^ 0x9E3779B97F4A7C15*seed))."""
    M = (1 << 64) - 1
    out = np.empty(len(gids))
    for k, g in enumerate(np.asarray(gids, dtype=np.int64).tolist()):
        z = ((g ^ ((0x9E3779B97F4A7C15 * seed) & M)) + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        out[k] = float(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0
    return out


def get_matrix(mesh: Mesh, mode: int = GRAPH_LAPLACIAN) -> System:
    """The matrix IO::getMatrix ends up with (ExodusIO.hpp:1379-1425): off-diagonals -1 for every
    pair of nodes sharing an element, diagonal = (entries in the row) - 1; nodesets are NOT applied
    (:729-731) — i.e. IO::assemble's operator with no Dirichlet node."""
    return assemble(mesh, mode, node_bc=np.full(mesh.num_nodes, np.nan))


def get_matrix_owners(conn: np.ndarray, epart: np.ndarray, nparts: int, num_nodes: int) -> np.ndarray:
    """Node ownership of IO::getMatrix, restated rank by rank as the reference computes it
    (ExodusIO.hpp:1080-1295, with std::map/std::set semantics; pure Python — small meshes only):
      adjacents (:1089-1097), ghostedNodeMap = pairwise intersections of the ranks' node lists
      (:1113-1135), nodeToFreq (:1163-1168), and the keep/remove decision (:1247-1281): a rank drops a
      node if another sharer has a higher frequency, or an equal one and a lower rank.
    Nodes in no element belong to nobody in the reference; they are given to rank 0 here."""
    conn = np.asarray(conn)
    adj = [dict() for _ in range(nparts)]                    # rank -> node -> set(neighbours)
    for e, nodes in enumerate(conn.tolist()):
        a = adj[int(epart[e])]
        for j in nodes:
            for k in nodes:
                if j != k:
                    a.setdefault(j, set()).add(k)
    node_lists = [sorted({int(v) for e in np.flatnonzero(np.asarray(epart) == r) for v in conn[e]}) for r in range(nparts)]
    freq = []
    for r in range(nparts):                                   # nodeToFreq[id] = rows of `adjacents` containing id
        f = {}
        for row in adj[r].values():
            for v in row:
                f[v] = f.get(v, 0) + 1
        freq.append(f)
    owner = np.zeros(num_nodes, dtype=np.int32)
    claimed = np.zeros(num_nodes, dtype=np.int32)
    sets = [set(l) for l in node_lists]
    for r in range(nparts):
        for v in node_lists[r]:
            mine = freq[r].get(v, 0)
            keep = True
            for i in range(nparts):
                if i == r or v not in sets[i]:
                    continue
                theirs = freq[i].get(v, 0)
                if mine < theirs or (mine == theirs and i < r):
                    keep = False
            if keep:
                owner[v] = r
                claimed[v] += 1
    used = np.zeros(num_nodes, dtype=bool)
    used[np.unique(conn)] = True
    assert np.all(claimed[used] == 1), "ownership rule must give every used node exactly one owner (map->isOneToOne, :1375)"
    return owner


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def pcg_loop_seconds() -> float:
    """wall seconds the iteration loop of the last pcg() call took, set-up excluded"""
    return float(lib().oracle_pcg_loop_seconds())


def set_num_threads(n: int) -> None:
    """OpenMP threads of the C oracle (bench.py: all cores of the affinity mask, whatever the launcher exported)"""
    lib().oracle_set_num_threads(int(n))


# ------------------------------------------------------------------------------------------
# METIS (the bundled static library, called with the reference's arguments)
# ------------------------------------------------------------------------------------------
_metis = None


def metis():
    global _metis
    if _metis is None:
        build()
        _metis = C.CDLL(_METIS)
    return _metis


def metis_part_mesh_dual(conn: np.ndarray, num_nodes: int, ncommon: int, nparts: int):
    """ExodusIO.hpp:1615 — METIS_PartMeshDual(ne, nn, eptr, eind, NULL, NULL, ncommon, nparts,
    NULL, NULL(options), objval, epart, npart) with idx_t=int64 / real_t=float32 (bundled lib)."""
    ne, npe = conn.shape
    eptr = (np.arange(ne + 1, dtype=np.int64) * npe)
    eind = np.ascontiguousarray(conn, dtype=np.int64).ravel()
    epart = np.zeros(ne, dtype=np.int64)
    npart = np.zeros(num_nodes, dtype=np.int64)
    ne_, nn_, nc_, np_, obj = (C.c_int64(ne), C.c_int64(num_nodes), C.c_int64(ncommon), C.c_int64(nparts), C.c_int64(0))
    rc = metis().METIS_PartMeshDual(C.byref(ne_), C.byref(nn_), _p(eptr, C.c_int64), _p(eind, C.c_int64),
                                    None, None, C.byref(nc_), C.byref(np_), None, None, C.byref(obj),
                                    _p(epart, C.c_int64), _p(npart, C.c_int64))
    if rc != 1:
        raise RuntimeError(f"METIS_PartMeshDual rc={rc}")
    return int(obj.value), epart, npart


def metis_part_graph_kway(xadj: np.ndarray, adjncy: np.ndarray, nparts: int):
    """METIS_PartGraphKway with all-default arguments on a CSR graph (no self loops)."""
    n = xadj.shape[0] - 1
    xadj = np.ascontiguousarray(xadj, dtype=np.int64)
    adjncy = np.ascontiguousarray(adjncy, dtype=np.int64)
    part = np.zeros(n, dtype=np.int64)
    if nparts == 1:
        return 0, part
    n_, ncon, np_, obj = C.c_int64(n), C.c_int64(1), C.c_int64(nparts), C.c_int64(0)
    rc = metis().METIS_PartGraphKway(C.byref(n_), C.byref(ncon), _p(xadj, C.c_int64), _p(adjncy, C.c_int64),
                                     None, None, None, C.byref(np_), None, None, None, C.byref(obj),
                                     _p(part, C.c_int64))
    if rc != 1:
        raise RuntimeError(f"METIS_PartGraphKway rc={rc}")
    return int(obj.value), part
