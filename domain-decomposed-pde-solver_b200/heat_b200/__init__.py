"""heat_b200 — ctypes face of libheat_b200.so (include/heat_b200.h).

Host-side mirror of the reference's only first-party interface, `ExodusIO::IO`
(/root/reference/ExodusIO.hpp:83-2225) plus `belosSolver` (BelosMueLuSolver.cpp:87-139):
`IO.open / create / assemble / decompose / writeSolution` keep the reference's names, argument
order and call-order contract; failures raise `HeatError` (the reference returns `false` and
prints to stderr).  All compute runs in the CUDA library — there is NO CPU fallback: importing
works anywhere (so the CPU test suite can check the ABI), but every compute call raises without
a CUDA device or without the built library.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HEAT_B200_LIB: another build of the same library (tools/peer_trace.py uses the -DHEAT_PEER_TRACE build in lib_trace/)
LIB_PATH = os.environ.get("HEAT_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libheat_b200.so")

OP_GRAPH_LAPLACIAN, OP_P1_FEM = 0, 1
SOLVER_CG, SOLVER_CG_SINGLE_REDUCE, SOLVER_GMRES = 0, 1, 2
PREC_NONE, PREC_JACOBI, PREC_CHEBYSHEV, PREC_ILU0 = 0, 1, 2, 3
PART_CONTIGUOUS, PART_METIS_KWAY, PART_SLAB = 0, 1, 2
COMM_ID_BYTES = 128


class HeatError(RuntimeError):
    pass


class SolveOpts(C.Structure):
    _fields_ = [("solver", C.c_int), ("prec", C.c_int), ("max_iters", C.c_int), ("tol", C.c_double),
                ("cheb_degree", C.c_int), ("cheb_lambda_max", C.c_double), ("cheb_ratio", C.c_double),
                ("check_every", C.c_int), ("gmres_restart", C.c_int)]


class SolveInfo(C.Structure):
    _fields_ = [("iters", C.c_int), ("converged", C.c_int), ("achieved_tol", C.c_double),
                ("r0_norm", C.c_double), ("solve_ms", C.c_double)]


class MatrixInfo(C.Structure):
    _fields_ = [("num_nodes", C.c_int64), ("num_elem", C.c_int64), ("n_global", C.c_int64), ("nnz_global", C.c_int64),
                ("n_owned", C.c_int64), ("n_ghost", C.c_int64), ("nnz_local", C.c_int64),
                ("max_row_len", C.c_int32), ("npe", C.c_int32), ("num_dim", C.c_int32), ("num_node_sets", C.c_int32),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("n_neighbors", C.c_int32), ("sell_chunk", C.c_int32),
                ("sell_padded_nnz", C.c_int64), ("n_boundary_slices", C.c_int64), ("n_slices", C.c_int64),
                ("assemble_ms", C.c_double), ("peer_path", C.c_int32), ("col_index_bytes", C.c_int32),
                ("assemble_fill_ms", C.c_double), ("asm_phase_ms", C.c_double * 4), ("csr_resident", C.c_int32),
                ("reserved0", C.c_int32), ("matrix_bytes", C.c_int64)]


class PowerInfo(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("residual", C.c_double), ("iters", C.c_int), ("converged", C.c_int),
                ("solve_ms", C.c_double), ("report_buf", C.POINTER(C.c_double)), ("report_capacity", C.c_int),
                ("report_count", C.c_int)]


class PlanSizes(C.Structure):
    _fields_ = [("n_owned", C.c_int64), ("n_ghost", C.c_int64), ("n_neighbors", C.c_int32), ("n_send", C.c_int64)]


# every symbol include/heat_b200.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "heat_last_error", "heat_version", "heat_device_count", "heat_kernel_launches", "heat_ctx_create", "heat_ctx_set_stream", "heat_ctx_set_output", "heat_open",
    "heat_create", "heat_close", "heat_mesh_set", "heat_mesh_cube", "heat_mesh_nodeset_ids", "heat_comm_unique_id", "heat_comm_init",
    "heat_comm_rank", "heat_assemble", "heat_get_matrix", "heat_node_owners", "heat_matrix_owned_nodeset", "heat_power_method", "heat_solve_opts_default", "heat_solve", "heat_solve_trajectory", "heat_solve_host", "heat_solve_host_batch", "heat_spmv", "heat_spmv_peer",
    "heat_cg_iterations", "heat_decompose", "heat_write_solution", "heat_write_nodal_field", "heat_nodal_field", "heat_scatter_nodal_field", "heat_reference_view_csr", "heat_decompose_partition",
    "heat_matrix_get_info", "heat_matrix_export_csr", "heat_matrix_export_maps", "heat_matrix_export_plan",
    "heat_matrix_export_red2orig", "heat_matrix_export_ilu0", "heat_matrix_free", "heat_vector_create", "heat_vector_size",
    "heat_vector_device_ptr", "heat_vector_set", "heat_vector_get", "heat_vector_fill", "heat_vector_fill_hash",
    "heat_vector_free", "heat_plan_build", "heat_partition_rows",
]

_lib = None


def lib():
    """Load libheat_b200.so.  Raises HeatError (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HeatError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64p, i32p, dp = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.heat_last_error.restype = C.c_char_p
    L.heat_kernel_launches.restype = C.c_ulonglong
    L.heat_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.heat_ctx_set_stream.argtypes = [vp, vp]
    L.heat_ctx_set_output.argtypes = [vp, C.c_int, C.c_int]
    L.heat_open.argtypes = [vp, C.c_char_p, C.c_int]
    L.heat_create.argtypes = [vp, C.c_char_p]
    L.heat_close.argtypes = [vp]
    L.heat_mesh_set.argtypes = [vp, C.c_int64, C.c_int, dp, dp, dp, C.c_int64, C.c_int, i32p, C.c_int, i64p, i64p, i64p]
    L.heat_mesh_cube.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.heat_mesh_nodeset_ids.argtypes = [vp, C.POINTER(C.c_int), i64p]
    L.heat_comm_unique_id.argtypes = [C.c_char_p]
    L.heat_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.heat_comm_rank.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.heat_assemble.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.heat_get_matrix.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.heat_node_owners.argtypes = [C.c_int64, C.c_int64, C.c_int, i32p, i64p, C.c_int, i32p]
    L.heat_matrix_owned_nodeset.argtypes = [vp, C.c_int64, i64p, i64p]
    L.heat_power_method.argtypes = [vp, vp, C.c_int, C.c_double, C.c_uint64, C.POINTER(PowerInfo)]
    L.heat_solve_opts_default.argtypes = [C.POINTER(SolveOpts)]
    L.heat_solve_opts_default.restype = None
    L.heat_solve.argtypes = [vp, vp, vp, vp, C.POINTER(SolveOpts), C.POINTER(SolveInfo)]
    L.heat_solve_trajectory.argtypes = [vp, vp, vp, vp, C.POINTER(SolveOpts), C.c_int, C.c_int, C.POINTER(SolveInfo),
                                        C.POINTER(C.c_int)]
    L.heat_solve_host.argtypes = [vp, vp, vp, vp, C.POINTER(SolveOpts), C.POINTER(SolveInfo)]
    L.heat_solve_host_batch.argtypes = [vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(SolveOpts), C.POINTER(SolveInfo)]
    L.heat_spmv.argtypes = [vp, vp, vp, vp]
    L.heat_spmv_peer.argtypes = [vp, vp, vp, vp, C.c_int, dp, dp]
    L.heat_cg_iterations.argtypes = [vp, vp, vp, vp, C.POINTER(SolveOpts), C.c_int, C.POINTER(SolveInfo)]
    L.heat_decompose.argtypes = [vp, C.c_int]
    L.heat_write_solution.argtypes = [vp, vp, C.c_int]
    L.heat_nodal_field.argtypes = [vp, vp, dp, C.c_int64]
    L.heat_scatter_nodal_field.argtypes = [vp, dp, C.c_int64, dp, C.c_int64]
    L.heat_reference_view_csr.argtypes = [vp, C.c_int64, i64p, i32p, dp, dp, i64p, i64p, i64p, i64p, i32p, dp, dp, i64p, i64p, i64p]
    L.heat_write_nodal_field.argtypes = [vp, dp, C.c_int64, C.c_int]
    L.heat_decompose_partition.argtypes = [vp, C.c_int, i64p, i64p, i64p]
    L.heat_matrix_get_info.argtypes = [vp, C.POINTER(MatrixInfo)]
    L.heat_matrix_export_csr.argtypes = [vp, i64p, i32p, dp]
    L.heat_matrix_export_maps.argtypes = [vp, i64p, i64p, i32p]
    L.heat_matrix_export_plan.argtypes = [vp, i32p, i64p, i32p, i64p]
    L.heat_matrix_export_red2orig.argtypes = [vp, i64p]
    L.heat_matrix_export_ilu0.argtypes = [vp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.heat_matrix_free.argtypes = [vp]
    L.heat_vector_create.argtypes = [vp, vp, C.POINTER(vp)]
    L.heat_vector_size.argtypes = [vp]
    L.heat_vector_size.restype = C.c_int64
    L.heat_vector_device_ptr.argtypes = [vp]
    L.heat_vector_device_ptr.restype = vp
    L.heat_vector_set.argtypes = [vp, vp, vp, C.c_int64]
    L.heat_vector_get.argtypes = [vp, vp, vp, C.c_int64]
    L.heat_vector_fill.argtypes = [vp, vp, C.c_double]
    L.heat_vector_fill_hash.argtypes = [vp, vp, vp, C.c_uint64]
    L.heat_vector_free.argtypes = [vp]
    L.heat_plan_build.argtypes = [C.c_int64, i64p, i32p, i32p, C.c_int, C.c_int, C.POINTER(PlanSizes), i64p, i64p, i32p,
                                  i32p, i64p, i64p, i64p]
    L.heat_partition_rows.argtypes = [C.c_int64, i64p, i32p, C.c_int, C.c_int, i32p]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise HeatError(f"[heat_b200 rc={rc}] {lib().heat_last_error().decode(errors='replace')}")


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def device_count() -> int:
    return int(lib().heat_device_count())


def kernel_launches() -> int:
    """Kernels of libheat_b200 launched by this process so far."""
    return int(lib().heat_kernel_launches())


def _as_pointer(obj):
    """host numpy array, torch tensor (any device) or raw int address -> void*"""
    if obj is None:
        return None
    if isinstance(obj, np.ndarray):
        return C.c_void_p(obj.ctypes.data)
    if hasattr(obj, "data_ptr"):
        return C.c_void_p(obj.data_ptr())
    return C.c_void_p(int(obj))


class Vector:
    """Teuchos::RCP<Tpetra::MultiVector<>> (one column) living on the GPU."""

    def __init__(self, io: "IO", handle):
        self.io, self.h = io, handle

    def __len__(self):
        return int(lib().heat_vector_size(self.h))

    @property
    def device_ptr(self) -> int:
        return int(lib().heat_vector_device_ptr(self.h))

    def numpy(self) -> np.ndarray:
        out = np.empty(len(self), dtype=np.float64)
        _check(lib().heat_vector_get(self.io.h, self.h, _as_pointer(out), len(self)))
        return out

    def set(self, src):
        if isinstance(src, np.ndarray):
            src = np.ascontiguousarray(src, dtype=np.float64)
            n = src.size
        else:
            n = src.numel()
        _check(lib().heat_vector_set(self.io.h, self.h, _as_pointer(src), n))
        return self

    def fill(self, value: float):
        _check(lib().heat_vector_fill(self.io.h, self.h, float(value)))
        return self

    def free(self):
        if self.h:
            lib().heat_vector_free(self.h)
            self.h = None


class Matrix:
    """Teuchos::RCP<Tpetra::CrsMatrix<>>: this rank's rows, device resident (CSR + SELL-C)."""

    def __init__(self, io: "IO", handle):
        self.io, self.h = io, handle

    @property
    def info(self) -> MatrixInfo:
        mi = MatrixInfo()
        _check(lib().heat_matrix_get_info(self.h, C.byref(mi)))
        return mi

    def csr(self):
        """(row_ptr int64, col int32 LOCAL ids, val) host copies."""
        mi = self.info
        rp = np.empty(mi.n_owned + 1, dtype=np.int64)
        col = np.empty(max(mi.nnz_local, 1), dtype=np.int32)
        val = np.empty(max(mi.nnz_local, 1), dtype=np.float64)
        _check(lib().heat_matrix_export_csr(self.h, _ptr(rp, C.c_int64), _ptr(col, C.c_int32), _ptr(val, C.c_double)))
        return rp, col[: mi.nnz_local], val[: mi.nnz_local]

    def maps(self):
        mi = self.info
        owned = np.empty(max(mi.n_owned, 1), dtype=np.int64)
        ghost = np.empty(max(mi.n_ghost, 1), dtype=np.int64)
        owner = np.empty(max(mi.n_ghost, 1), dtype=np.int32)
        _check(lib().heat_matrix_export_maps(self.h, _ptr(owned, C.c_int64), _ptr(ghost, C.c_int64), _ptr(owner, C.c_int32)))
        return owned[: mi.n_owned], ghost[: mi.n_ghost], owner[: mi.n_ghost]

    def plan(self):
        mi = self.info
        k = mi.n_neighbors
        nbr = np.empty(max(k, 1), dtype=np.int32)
        sp = np.zeros(k + 1, dtype=np.int64)
        rp = np.zeros(k + 1, dtype=np.int64)
        _check(lib().heat_matrix_export_plan(self.h, _ptr(nbr, C.c_int32), _ptr(sp, C.c_int64), None, _ptr(rp, C.c_int64)))
        sidx = np.empty(max(int(sp[-1]), 1), dtype=np.int32)
        _check(lib().heat_matrix_export_plan(self.h, None, None, _ptr(sidx, C.c_int32), None))
        return nbr[:k], sp, sidx[: int(sp[-1])], rp

    def red2orig(self) -> np.ndarray:
        out = np.empty(max(self.info.n_owned, 1), dtype=np.int64)
        _check(lib().heat_matrix_export_red2orig(self.h, _ptr(out, C.c_int64)))
        return out[: self.info.n_owned]

    def new_vector(self) -> Vector:
        h = C.c_void_p()
        _check(lib().heat_vector_create(self.io.h, self.h, C.byref(h)))
        return Vector(self.io, h)

    def hash_vector(self, seed: int = 12345) -> Vector:
        v = self.new_vector()
        _check(lib().heat_vector_fill_hash(self.io.h, self.h, v.h, seed))
        return v

    def ilu0(self):
        """(lu values on the CSR pattern, forward levels, backward levels) of HEAT_PREC_ILU0."""
        lu = np.empty(max(self.info.nnz_local, 1), dtype=np.float64)
        nl, nu = C.c_int(0), C.c_int(0)
        _check(lib().heat_matrix_export_ilu0(self.h, _ptr(lu, C.c_double), C.byref(nl), C.byref(nu)))
        return lu[: self.info.nnz_local], nl.value, nu.value

    def owned_nodeset(self, set_id: int) -> np.ndarray:
        """getMatrix's nodeSetMap[set_id] (ExodusIO.hpp:1447-1466): owned nodes of that nodeset, 0-based."""
        cnt = C.c_int64(0)
        _check(lib().heat_matrix_owned_nodeset(self.h, set_id, C.byref(cnt), None))
        out = np.empty(max(cnt.value, 1), dtype=np.int64)
        _check(lib().heat_matrix_owned_nodeset(self.h, set_id, C.byref(cnt), _ptr(out, C.c_int64)))
        return out[: cnt.value]

    def free(self):
        if self.h:
            lib().heat_matrix_free(self.h)
            self.h = None


@dataclass
class PowerResult:
    lambda_: float
    residual: float
    iters: int
    converged: bool
    solve_ms: float
    reports: np.ndarray = None      # rows {iteration, lambda, residual}: what the reference prints every 50 iterations


@dataclass
class SolveResult:
    iters: int
    converged: bool
    achieved_tol: float
    r0_norm: float
    solve_ms: float


class IO:
    """Mirror of `ExodusIO::IO` (ExodusIO.hpp:83).  One IO per process == per GPU == per rank.
    Call order contract of the reference: open -> assemble -> (rank 0) create -> decompose ->
    repeated writeSolution."""

    def __init__(self, device: int = 0, stream=None):
        h = C.c_void_p()
        _check(lib().heat_ctx_create(device, C.byref(h)))
        self.h = h
        if stream is not None:
            self.set_stream(stream)

    # -- plumbing -------------------------------------------------------------------------------
    def set_stream(self, stream):
        """cudaStream_t as an int, or a torch.cuda.Stream."""
        s = getattr(stream, "cuda_stream", stream)
        _check(lib().heat_ctx_set_stream(self.h, C.c_void_p(int(s))))

    def set_output(self, word_size: int = 8, largest_nodeset_id: bool = False):
        """output conventions of a reference build: float32 records (real_t = float, ExodusIO.hpp:104-105) and the
        largest nodeset id in the written field (:1983-1989)"""
        _check(lib().heat_ctx_set_output(self.h, int(word_size), int(bool(largest_nodeset_id))))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(lib().heat_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        _check(lib().heat_comm_init(self.h, rank, nranks, unique_id))

    # -- ExodusIO::IO -----------------------------------------------------------------------------
    def open(self, fname: str, read_only: bool = False) -> bool:          # ExodusIO.hpp:88
        _check(lib().heat_open(self.h, fname.encode(), int(read_only)))
        return True

    def create(self, fname: str) -> bool:                                  # ExodusIO.hpp:103
        _check(lib().heat_create(self.h, fname.encode()))
        return True

    def mesh_set(self, x, y, z, conn, nodesets: dict, num_dim: int = 3):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        zz = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        ids = np.array(sorted(nodesets), dtype=np.int64)
        ptr = np.zeros(len(ids) + 1, dtype=np.int64)
        nodes = []
        for k, sid in enumerate(ids):
            arr = np.asarray(nodesets[int(sid)], dtype=np.int64)
            nodes.append(arr)
            ptr[k + 1] = ptr[k] + arr.size
        flat = np.concatenate(nodes).astype(np.int64) if nodes else np.zeros(0, dtype=np.int64)
        flat = np.ascontiguousarray(flat) if flat.size else np.zeros(1, dtype=np.int64)
        ids_ = ids if ids.size else np.zeros(1, dtype=np.int64)
        _check(lib().heat_mesh_set(self.h, x.size, num_dim, _ptr(x, C.c_double), _ptr(y, C.c_double),
                                   _ptr(zz, C.c_double) if zz is not None else None, conn.shape[0], conn.shape[1],
                                   _ptr(conn, C.c_int32), len(ids), _ptr(ids_, C.c_int64), _ptr(ptr, C.c_int64),
                                   _ptr(flat, C.c_int64)))

    def mesh_cube(self, nx: int, ny: int, nz: int, explicit_mesh: bool = False):
        _check(lib().heat_mesh_cube(self.h, nx, ny, nz, int(explicit_mesh)))

    def nodeset_ids(self) -> np.ndarray:
        cnt = C.c_int(0)
        _check(lib().heat_mesh_nodeset_ids(self.h, C.byref(cnt), None))
        ids = np.zeros(max(cnt.value, 1), dtype=np.int64)
        _check(lib().heat_mesh_nodeset_ids(self.h, C.byref(cnt), _ptr(ids, C.c_int64)))
        return ids[: cnt.value]

    def assemble(self, op_mode: int = OP_GRAPH_LAPLACIAN, partitioner: int = PART_METIS_KWAY):   # ExodusIO.hpp:128
        """-> (A, X, B) like `assemble(&A, &X, &B)`; X is zero (the reference randomises it unseeded)."""
        a, x, b = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().heat_assemble(self.h, op_mode, partitioner, C.byref(a), C.byref(x), C.byref(b)))
        return Matrix(self, a), Vector(self, x), Vector(self, b)

    def getMatrix(self, op_mode: int = OP_GRAPH_LAPLACIAN) -> Matrix:      # ExodusIO.hpp:733
        """-> the whole-mesh Laplacian, rows distributed by element partition + ownership rule;
        the reference's second output (nodeSetMap) is `Matrix.owned_nodeset(id)`."""
        a = C.c_void_p()
        _check(lib().heat_get_matrix(self.h, op_mode, C.byref(a)))
        return Matrix(self, a)

    def power_method(self, A: Matrix, niters: int = 500, tolerance: float = 1.0e-2, seed: int = 12345) -> PowerResult:
        """PowerMethod::run (ExodusMatrixTest.cpp:56-129; defaults 500 / 1e-2 from :163)."""
        info = PowerInfo()
        cap = niters // 50 + 2
        rep = np.zeros(3 * cap)
        info.report_buf, info.report_capacity = _ptr(rep, C.c_double), cap
        _check(lib().heat_power_method(self.h, A.h, niters, tolerance, seed, C.byref(info)))
        return PowerResult(info.lambda_, info.residual, info.iters, bool(info.converged), info.solve_ms,
                           rep[: 3 * info.report_count].reshape(-1, 3))

    def decompose(self, partitions: int) -> bool:                          # ExodusIO.hpp:1496
        _check(lib().heat_decompose(self.h, partitions))
        return True

    def decompose_partition(self, partitions: int, num_elem: int, num_nodes: int):
        epart = np.zeros(num_elem, dtype=np.int64)
        npart = np.zeros(num_nodes, dtype=np.int64)
        obj = C.c_int64(0)
        _check(lib().heat_decompose_partition(self.h, partitions, C.byref(obj), _ptr(epart, C.c_int64), _ptr(npart, C.c_int64)))
        return int(obj.value), epart, npart

    def writeSolution(self, vec: Vector, timestep: int) -> bool:           # ExodusIO.hpp:1972
        _check(lib().heat_write_solution(self.h, vec.h, timestep))
        return True

    def write_nodal_field(self, field, timestep: int) -> bool:
        """the file half of writeSolution (ExodusIO.hpp:2027-2069): host array -> time step `timestep`."""
        f = np.ascontiguousarray(field, dtype=np.float64)
        _check(lib().heat_write_nodal_field(self.h, _ptr(f, C.c_double), f.size, timestep))
        return True

    def scatter_nodal_field(self, x_reduced, num_nodes: int) -> np.ndarray:
        """host solution in reduced-id order -> dense nodal array (nodeset nodes = their id); no GPU involved"""
        x = np.ascontiguousarray(x_reduced, dtype=np.float64)
        out = np.empty(max(num_nodes, 1), dtype=np.float64)
        _check(lib().heat_scatter_nodal_field(self.h, _ptr(x, C.c_double), x.size, _ptr(out, C.c_double), num_nodes))
        return out[:num_nodes]

    def reference_view(self, row_ptr, col, val, b, red2orig):
        """the FIXED system (e.g. Matrix.csr(), B.numpy(), Matrix.red2orig() on one rank) as the reference's own
        IO::assemble returns it for this mesh (defects D1 and D3) -> (row_ptr, col, val, b, idmap_reduced, idmap_original)"""
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        r2o = np.ascontiguousarray(red2orig, dtype=np.int64)
        n = row_ptr.size - 1
        no, nnz, nm = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        head = (self.h, n, _ptr(row_ptr, C.c_int64), _ptr(col, C.c_int32) if col.size else None, _ptr(val, C.c_double) if val.size else None,
                _ptr(b, C.c_double) if b.size else None, _ptr(r2o, C.c_int64) if r2o.size else None, C.byref(no), C.byref(nnz))
        _check(lib().heat_reference_view_csr(*head, None, None, None, None, C.byref(nm), None, None))
        rp = np.zeros(no.value + 1, dtype=np.int64)
        co, va = np.zeros(max(nnz.value, 1), dtype=np.int32), np.zeros(max(nnz.value, 1), dtype=np.float64)
        bo = np.zeros(max(no.value, 1), dtype=np.float64)
        ir, io_ = np.zeros(max(nm.value, 1), dtype=np.int64), np.zeros(max(nm.value, 1), dtype=np.int64)
        _check(lib().heat_reference_view_csr(*head, _ptr(rp, C.c_int64), _ptr(co, C.c_int32), _ptr(va, C.c_double), _ptr(bo, C.c_double),
                                             C.byref(nm), _ptr(ir, C.c_int64), _ptr(io_, C.c_int64)))
        return rp, co[: nnz.value], va[: nnz.value], bo[: no.value], ir[: nm.value], io_[: nm.value]

    def nodal_field(self, vec: Vector, num_nodes: int) -> np.ndarray:
        out = np.empty(num_nodes, dtype=np.float64)
        _check(lib().heat_nodal_field(self.h, vec.h, _ptr(out, C.c_double), num_nodes))
        return out

    # -- belosSolver --------------------------------------------------------------------------------
    @staticmethod
    def solve_opts(**kw) -> SolveOpts:
        o = SolveOpts()
        lib().heat_solve_opts_default(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown solve option {k}")
            setattr(o, k, v)
        return o

    def solve(self, A: Matrix, X: Vector, B: Vector, **kw) -> SolveResult:   # BelosMueLuSolver.cpp:87
        o, info = self.solve_opts(**kw), SolveInfo()
        _check(lib().heat_solve(self.h, A.h, X.h, B.h, C.byref(o), C.byref(info)))
        return SolveResult(info.iters, bool(info.converged), info.achieved_tol, info.r0_norm, info.solve_ms)

    def solve_trajectory(self, A: Matrix, X: Vector, B: Vector, write_every: int = 1, first_timestep: int = 0, **kw):
        """belosSolver with its per-iteration writeSolution (BelosMueLuSolver.cpp:113-133) in one Krylov
        run -> (SolveResult, frames written)."""
        o, info, frames = self.solve_opts(**kw), SolveInfo(), C.c_int(0)
        _check(lib().heat_solve_trajectory(self.h, A.h, X.h, B.h, C.byref(o), write_every, first_timestep,
                                           C.byref(info), C.byref(frames)))
        return SolveResult(info.iters, bool(info.converged), info.achieved_tol, info.r0_norm, info.solve_ms), frames.value

    def solve_host(self, A: Matrix, b_host, x_host, **kw) -> SolveResult:
        """b_host / x_host: numpy arrays or pinned torch CPU tensors (x_host: x0 in, solution out)."""
        o, info = self.solve_opts(**kw), SolveInfo()
        _check(lib().heat_solve_host(self.h, A.h, _as_pointer(b_host), _as_pointer(x_host), C.byref(o), C.byref(info)))
        return SolveResult(info.iters, bool(info.converged), info.achieved_tol, info.r0_norm, info.solve_ms)

    def solve_host_batch(self, A: Matrix, b_hosts, x0_hosts, x_hosts, **kw):
        """A sequence of solves with one matrix through (pinned) host buffers; copies of neighbouring systems overlap
        each solve.  b_hosts / x_hosts: lists of numpy arrays or pinned torch CPU tensors; x0_hosts: list (entries may be
        None = zero start) or None.  Returns one SolveResult per system."""
        n = len(b_hosts)
        o, infos = self.solve_opts(**kw), (SolveInfo * max(n, 1))()
        arr = lambda lst: (C.c_void_p * max(n, 1))(*[(_as_pointer(t).value if t is not None else None) for t in lst])
        bp, xp = arr(b_hosts), arr(x_hosts)
        x0p = arr(x0_hosts) if x0_hosts is not None else None
        _check(lib().heat_solve_host_batch(self.h, A.h, n, bp, x0p, xp, C.byref(o), infos))
        return [SolveResult(i.iters, bool(i.converged), i.achieved_tol, i.r0_norm, i.solve_ms) for i in infos[:n]]

    def cg_iterations(self, A: Matrix, X: Vector, B: Vector, iters: int, **kw) -> SolveResult:
        o, info = self.solve_opts(**kw), SolveInfo()
        _check(lib().heat_cg_iterations(self.h, A.h, X.h, B.h, C.byref(o), iters, C.byref(info)))
        return SolveResult(info.iters, bool(info.converged), info.achieved_tol, info.r0_norm, info.solve_ms)

    def spmv(self, A: Matrix, x: Vector, y: Vector):
        _check(lib().heat_spmv(self.h, A.h, x.h, y.h))

    def spmv_peer(self, A: Matrix, x: Vector, y: Vector, repeat: int = 1):
        """y = A x through the peer-memory halo path (the SpMV launch of the multi-GPU CG loop); collective.
        Returns (global x.y, average kernel ms of launches 2..repeat)."""
        xy, ms = C.c_double(0.0), C.c_double(0.0)
        _check(lib().heat_spmv_peer(self.h, A.h, x.h, y.h, repeat, C.byref(xy), C.byref(ms)))
        return xy.value, ms.value

    def close(self):
        if self.h:
            lib().heat_close(self.h)
            self.h = None


# -- pure-host helpers (no GPU needed) ---------------------------------------------------------------
def partition_rows(row_ptr, col, partitioner: int, nranks: int) -> np.ndarray:
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    n = row_ptr.size - 1
    part = np.zeros(max(n, 1), dtype=np.int32)
    _check(lib().heat_partition_rows(n, _ptr(row_ptr, C.c_int64), _ptr(col, C.c_int32), partitioner, nranks, _ptr(part, C.c_int32)))
    return part[:n]


def node_owners(conn, num_nodes: int, epart, nparts: int) -> np.ndarray:
    """Node-ownership rule of IO::getMatrix (ExodusIO.hpp:1191-1295) from an element partition."""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    epart = np.ascontiguousarray(epart, dtype=np.int64)
    out = np.zeros(max(num_nodes, 1), dtype=np.int32)
    _check(lib().heat_node_owners(num_nodes, conn.shape[0], conn.shape[1], _ptr(conn, C.c_int32), _ptr(epart, C.c_int64),
                                  nparts, _ptr(out, C.c_int32)))
    return out[:num_nodes]


def maps_digest(owned, ghost, ghost_owner, nbr, send_ptr, send_gids, recv_ptr) -> str:
    """sha1 over one rank's owned / ghost / send maps (int64, little endian, in this order) — what bench.py's parity
    block compares with the digests tests/golden/make_parity_golden.py derived on the oracle side."""
    import hashlib
    h = hashlib.sha1()
    for a in (owned, ghost, ghost_owner, nbr, send_ptr, send_gids, recv_ptr):
        h.update(np.ascontiguousarray(np.asarray(a).ravel(), dtype="<i8").tobytes())
        h.update(b"|")
    return h.hexdigest()


def plan_build(row_ptr, col, part, nranks: int, rank: int) -> dict:
    """Owned/ghost/send maps of `rank` (Tpetra conventions) from a global pattern + row partition."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    part = np.ascontiguousarray(part, dtype=np.int32)
    n = row_ptr.size - 1
    sz = PlanSizes()
    args = (n, _ptr(row_ptr, C.c_int64), _ptr(col, C.c_int32), _ptr(part, C.c_int32), nranks, rank)
    _check(lib().heat_plan_build(*args, C.byref(sz), None, None, None, None, None, None, None))
    owned = np.empty(max(sz.n_owned, 1), dtype=np.int64)
    ghost = np.empty(max(sz.n_ghost, 1), dtype=np.int64)
    owner = np.empty(max(sz.n_ghost, 1), dtype=np.int32)
    nbr = np.empty(max(sz.n_neighbors, 1), dtype=np.int32)
    sp = np.zeros(sz.n_neighbors + 1, dtype=np.int64)
    rp = np.zeros(sz.n_neighbors + 1, dtype=np.int64)
    sg = np.empty(max(sz.n_send, 1), dtype=np.int64)
    _check(lib().heat_plan_build(*args, C.byref(sz), _ptr(owned, C.c_int64), _ptr(ghost, C.c_int64), _ptr(owner, C.c_int32),
                                 _ptr(nbr, C.c_int32), _ptr(sp, C.c_int64), _ptr(sg, C.c_int64), _ptr(rp, C.c_int64)))
    return dict(owned=owned[: sz.n_owned], ghost=ghost[: sz.n_ghost], ghost_owner=owner[: sz.n_ghost],
                nbr=nbr[: sz.n_neighbors], send_ptr=sp, send_gids=sg[: sz.n_send], recv_ptr=rp)
