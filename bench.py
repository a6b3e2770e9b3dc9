#!/usr/bin/env python
"""bench.py — headline benchmark of the heat path: Jacobi-PCG iterations/s and SpMV HBM GB/s on the
synthetic 512^3-node Kuhn tet cube (133.7 M DOF, P1 operator), strong-scaled over N B200s.

    python bench.py --gpus 1 --steps K --warmup W                 (N>1: launched by torchrun)
    python bench.py --impl reference ...                           (CPU restatement of the reference, all host cores)
    python bench.py --assemble explicit --nx 256                   (north-star kernel 1: explicit-mesh P1 assembly)
    python bench.py --workload weak --prec chebyshev               (BASELINE.json configs[4])

A "step" = one call of the solver entry point running `--iters-per-step` CG iterations on the
resident system (plus the r0 = b - A x0 set-up SpMV every solve pays).  `value` = CG iterations/s of
the whole job with inputs resident in HBM; `e2e` = the same through the C ABI's host-buffer entry
point (heat_solve_host: H2D of b and x0, D2H of x inside the timed region).  One JSON line on
stdout (rank 0).  Timing: CUDA events on the stream the kernels run on, barrier + synchronize on
both sides, max over ranks.  The working set (>= 38 GB) is >> the 126 MB L2, so no L2 flush.

Before anything is timed every rank runs the PARITY block (`"parity"` in the line; a failure exits non-zero):
A x* = b on the very matrix that is benchmarked (through the NCCL-halo SpMV and, on >1 GPU, through the
peer-memory SpMV of the CG loop), METIS-N / slab-N solves of bolted_bracket.exo and a 33x17x16 cube against the
committed N=1 oracle solutions (tests/golden/parity_multi.*), and the owned / ghost / send-map digests of every
rank against the digests derived on the oracle side.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "domain-decomposed-pde-solver_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "CG iters/s (Jacobi-PCG fp64, 512^3-node P1-tet heat, 133.7M DOF)"
UNIT = "iters/s"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def cube_counts(nx, ny, nz):
    a, b, c = nx - 2, ny, nz
    n = a * b * c
    edges = ((a - 1) * b * c + a * (b - 1) * c + a * b * (c - 1) + (a - 1) * (b - 1) * c + a * (b - 1) * (c - 1)
             + (a - 1) * b * (c - 1) + (a - 1) * (b - 1) * (c - 1))
    return n, n + 2 * edges


def spmv_bytes(n, nnz):          # 12*nnz + 16*n + 4*(n+1)
    return 12 * nnz + 16 * n + 4 * (n + 1)


def cg_iter_bytes(n, nnz):       # SpMV + 88*n of vector traffic
    return spmv_bytes(n, nnz) + 88 * n


def cheb_iter_bytes(n, nnz, k):  # Chebyshev degree k adds (k-1) SpMVs + ~5 vector sweeps per step
    return cg_iter_bytes(n, nnz) + (k - 1) * (spmv_bytes(n, nnz) + 40 * n) + 32 * n


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1048576.0
    except OSError:
        pass
    return 0.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU restatement of the reference path (oracle), ALL host threads,
# on the benchmarked configuration itself (no scaling)
# ---------------------------------------------------------------------------------------------------
def cpu_run(nx, iters, steps, warmup):
    """Times `iters` classical Jacobi-PCG iterations per step of the C/OpenMP oracle on the nx^3 P1 cube — the
    bench's own configuration, un-scaled — with every core of the affinity mask (set explicitly: a launcher such as
    torchrun exports OMP_NUM_THREADS=1).  The system comes from oracle_cube_assemble (closed-form Kuhn connectivity,
    bit-identical to the explicit-mesh oracle, tests/test_oracle_golden.py) because the explicit 512^3 mesh does not
    fit a host; if even the 31 GB system does not fit (MemAvailable < 48 GB) the 256^3 cube (BASELINE.json
    configs[2]) is timed instead and the line says so.  iterations/s = iters / wall time of the iteration loop
    (set-up of a solve excluded, which favours the CPU arm on a short bounded sample)."""
    import oracle as O
    threads = host_threads()
    O.set_num_threads(threads)
    n_cpu = nx
    need_gb = 48.0 * (nx / 512.0) ** 3
    if mem_available_gb() < need_gb:
        n_cpu = 256 if nx > 256 else nx
    t0 = time.time()
    sysm = O.cube_assemble(n_cpu, n_cpu, n_cpu, O.P1_FEM, copy=False)
    t_asm = time.time() - t0
    loop_s = []
    for s in range(warmup + steps):
        O.pcg(sysm, prec=O.PREC_JACOBI, max_iters=iters, tol=0.0)
        if s >= warmup:
            loop_s.append(O.pcg_loop_seconds())
    dt = sum(loop_s) / len(loop_s)
    out = {"its": iters / dt, "ms_per_step": dt * 1e3, "threads": O.num_threads(), "assemble_s": t_asm, "n": sysm.n, "nnz": sysm.nnz,
           "nx_timed": n_cpu, "same_config": n_cpu == nx}
    sysm.free()
    return out


def reference_assembly_sample(n=41):
    """Assembly seconds of THE REFERENCE'S OWN IO::assemble (oracle/_ref/ref_driver = /root/reference/ExodusIO.hpp
    compiled unmodified against single-rank stand-ins, oracle/ref_shim/README.md) on an n^3-node Kuhn cube, one host
    thread — the reported CPU figure beside `assemble_ms`.  None if the binary did not travel."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
        import oracle as O
        import run_ref as R
        m = O.cube_mesh(n, n, n)
        out = R.run_reference("cube.exo", 2, get_matrix=False, timeout=120, arrays=(m.x, m.y, m.z, [("TETRA", m.conn)], m.nodesets))
        t = out["timing"]["assemble_s"]
        return {"kind": "reference", "seconds": t, "nodes": n ** 3, "us_per_node": 1e6 * t / n ** 3, "cores": 1,
                "sample": f"IO::assemble of the reference itself (graph Laplacian, one rank) on a {n}^3-node Kuhn tet cube"}
    except Exception as e:      # noqa: BLE001 — a reported extra, never a reason to lose the bench line
        return {"kind": "reference", "unavailable": f"{type(e).__name__}: {e}"[:200]}


def cpu_baseline_dict(c, nx, iters, steps):
    what = (f"C/OpenMP restatement of the reference path (oracle), {c['nx_timed']}^3-node P1 cube ({c['n']} DOF, {c['nnz']} nnz)"
            + ("" if c["same_config"] else f" — NOT the {nx}^3 configuration: host MemAvailable too small, value is un-scaled")
            + f", {iters} Jacobi-PCG iterations per step x {steps} steps, iteration loop only; {c['threads']} threads; "
            f"CPU assembly (closed-form cube) {c['assemble_s']:.1f} s")
    return {"value": c["its"], "unit": UNIT, "cores": c["threads"], "kind": "port", "sample": what,
            "same_config": c["same_config"], "nx_timed": c["nx_timed"], "assemble_s": c["assemble_s"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nx = args.nx
    n_full, nnz_full = cube_counts(nx, nx, nx)
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    c = cpu_run(nx, args.cpu_iters, steps, warmup)
    cb = cpu_baseline_dict(c, nx, args.cpu_iters, steps)
    cb["assembly_reference"] = reference_assembly_sample()
    metric = METRIC if nx == 512 else f"CG iters/s (jacobi-PCG fp64, {nx}x{nx}x{nx}-node P1-tet heat, {n_full / 1e6:.1f}M DOF)"
    line = {
        "impl": "reference", "metric": metric, "value": c["its"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": c["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {c['nx_timed']}x{c['nx_timed']}x{c['nx_timed']}-node Kuhn tet cube (P1 FEM), jacobi-PCG (cg), CPU restatement",
                   "n_dof": c["n"], "nnz": c["nnz"], "iters_per_step": args.cpu_iters, "same_config_as_gpu_arm": c["same_config"],
                   "note": "CPU restatement of the reference path (C/OpenMP oracle, all host cores); the Trilinos/Belos binary cannot be built here"},
        "cpu_baseline": cb,
        "e2e": {"value": c["its"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
class Job:
    """process group + library context plumbing shared by the modes"""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        if args.gpus != self.world and self.rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={self.world}", file=sys.stderr)
        self.stream = torch.cuda.current_stream()

    def new_io(self, comm=True):
        """a library context on this rank's GPU; comm: joined to a fresh NCCL communicator over all ranks"""
        import heat_b200 as hb
        torch = self.torch
        io = hb.IO(self.local_rank, self.stream)
        if comm and self.world > 1:
            idt = torch.zeros(hb.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                idt.copy_(torch.frombuffer(bytearray(hb.IO.comm_unique_id()), dtype=torch.uint8))
            self.dist.broadcast(idt, 0)
            io.comm_init(self.rank, self.world, bytes(idt.cpu().numpy().tobytes()))
        return io

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.barrier()
        return self.allmax(e0.elapsed_time(e1))


# ---------------------------------------------------------------------------------------------------
# parity block: driver-visible multi-GPU parity (north_star: maps bit-exact, solutions 1e-8, iterations +-2)
# ---------------------------------------------------------------------------------------------------
def parity_block(job, io, A, B, nx, op_is_p1):
    import numpy as np
    import heat_b200 as hb
    world, rank = job.world, job.rank
    out = {"ok": True, "world": world, "checks": {}}

    def record(name, ok, **kw):
        ok_all = job.allmax(0.0 if ok else 1.0) == 0.0          # every rank must agree
        out["checks"][name] = dict(ok=ok_all, **kw)
        out["ok"] = out["ok"] and ok_all

    # (i) A x* = b on the benchmarked slabs against the analytic field (P1 on a Kuhn cube is exact for linear fields:
    #     x* = 1000 - 900 i/(nx-1)), through both SpMV paths
    if op_is_p1:
        i = A.red2orig() % nx
        xs = A.new_vector().set(1000.0 - 900.0 * i / (nx - 1.0))
        ys = A.new_vector()
        b = B.numpy()
        bmax = job.allmax(np.abs(b).max())
        io.spmv(A, xs, ys)
        res = job.allmax(np.abs(ys.numpy() - b).max())
        record("Ax*=b (analytic field, NCCL-halo SpMV)" if world > 1 else "Ax*=b (analytic field)", res <= 1e-9 * bmax,
               max_abs_residual=res, b_inf=bmax, bar="1e-9*|b|inf")
        if world > 1:
            try:
                ys.fill(0.0)
                io.spmv_peer(A, xs, ys)
                resp = job.allmax(np.abs(ys.numpy() - b).max())
                record("Ax*=b (analytic field, PEER-path SpMV of the CG loop)", resp <= 1e-9 * bmax, max_abs_residual=resp,
                       peer_path=bool(A.info.peer_path))
            except RuntimeError as e:
                record("Ax*=b (analytic field, PEER-path SpMV of the CG loop)", os.environ.get("HEAT_COMM") == "nccl",
                       skipped=str(e)[:160])
        for v in (xs, ys):
            v.free()

    # (ii) + (iii) small systems against the committed N=1 oracle solutions and map digests
    with open(os.path.join(GOLDEN, "parity_multi.json")) as f:
        meta = json.load(f)
    gold = np.load(os.path.join(GOLDEN, "parity_multi.npz"))
    cnx, cny, cnz = meta["cube"]
    cases = [("cube_p1", hb.OP_P1_FEM, hb.PART_SLAB), ("cube_graph", hb.OP_GRAPH_LAPLACIAN, hb.PART_SLAB),
             ("bolted_bracket_graph", hb.OP_GRAPH_LAPLACIAN, hb.PART_METIS_KWAY)]
    for name, op, part in cases:
        ent = meta["cases"][name]
        if str(world) not in ent["maps"]:
            record(name, True, skipped=f"no golden digests for {world} ranks")
            continue
        io2 = job.new_io()
        if name.startswith("cube"):
            io2.mesh_cube(cnx, cny, cnz, False)
        else:
            io2.open(os.path.join(GOLDEN, "meshes", "bolted_bracket.exo"), True)
        A2, X2, B2 = io2.assemble(op, part)
        owned, ghost, owner = A2.maps()
        nbr, sp, sidx, rp_ = A2.plan()
        dig = hb.maps_digest(owned, ghost, owner, nbr, sp, owned[sidx], rp_)
        maps_ok = dig == ent["maps"][str(world)][rank]
        res = io2.solve(A2, X2, B2, solver=hb.SOLVER_CG, prec=hb.PREC_JACOBI, max_iters=5000, tol=meta["tol"], check_every=8)
        xg = gold[name + "_x"]
        err = job.allmax(np.abs(X2.numpy() - xg[owned]).max()) / np.abs(xg).max()
        ok = maps_ok and res.converged and abs(res.iters - ent["iters"]) <= 2 and err <= 1e-8
        record(name, ok, maps_bit_exact=bool(job.allmax(0.0 if maps_ok else 1.0) == 0.0), iters=res.iters, iters_n1_oracle=ent["iters"],
               rel_err_vs_n1_oracle=err, peer_path=bool(A2.info.peer_path), n=ent["n"], partition=ent["partition"])
        for v in (X2, B2):
            v.free()
        A2.free()
        io2.close()
    out["note"] = ("solutions vs the committed N=1 oracle solution: 1e-8 relative, iterations +-2 (north_star); not bit-equal across N — "
                   "the dot products are reduced per rank, then in rank order, so the rounding depends on the partition (deterministic "
                   "for a given N)")
    return out


# ---------------------------------------------------------------------------------------------------
def run_assemble_explicit(args):
    """North-star kernel (1) at scale: explicit-connectivity P1 assembly (values_kernel / pattern_* / node->element sort)
    of the nx^3 Kuhn cube materialised on the device, timed by phase with CUDA events, bit-compared with the analytic path."""
    import numpy as np
    import heat_b200 as hb
    job = Job(args)
    nx = args.nx
    n_full, nnz_full = cube_counts(nx, nx, nx)
    N, ne = nx ** 3, 6 * (nx - 1) ** 3
    peak, peak_src = measured_peak()
    runs = []
    io = job.new_io(comm=False)
    io.mesh_cube(nx, nx, nx, True)
    for rep in range(max(args.steps, 1) + 1):
        t0 = time.time()
        A, X, B = io.assemble(hb.OP_P1_FEM, hb.PART_SLAB)
        wall = (time.time() - t0) * 1e3
        mi = A.info
        runs.append({"wall_ms": wall, "assemble_ms": mi.assemble_ms, "phase_ms": list(mi.asm_phase_ms)})
        if rep == 0:
            # parity: the explicit-mesh kernels and the analytic direct-to-SELL kernel build the same system, bit for bit
            ioa = job.new_io(comm=False)
            ioa.mesh_cube(nx, nx, nx, False)
            Aa, Xa, Ba = ioa.assemble(hb.OP_P1_FEM, hb.PART_SLAB)
            same_b = bool(np.array_equal(B.numpy(), Ba.numpy()))
            xs, y1, y2 = A.hash_vector(99), A.new_vector(), Aa.new_vector()
            xa = Aa.new_vector().set(xs.numpy())
            io.spmv(A, xs, y1); ioa.spmv(Aa, xa, y2)
            same_spmv = bool(np.array_equal(y1.numpy(), y2.numpy()))
            same_csr = None
            if nx <= 160:
                same_csr = all(np.array_equal(u, v) for u, v in zip(A.csr(), Aa.csr()))
            parity = {"ok": same_b and same_spmv and same_csr is not False, "rhs_bit_equal": same_b,
                      "spmv_bit_equal_vs_analytic_path": same_spmv, "csr_bit_equal": same_csr}
            for v in (xs, y1, y2, xa, Xa, Ba):
                v.free()
            Aa.free(); ioa.close()
        for v in (X, B):
            v.free()
        A.free()
    io.close()
    best = min(runs[1:], key=lambda r: r["assemble_ms"])
    sort_ms, count_ms, fill_ms, val_ms = best["phase_ms"]
    bytes_alg = 16 * ne + 24 * N + 12 * nnz_full + 8 * n_full        # SURVEY.md §8(d), assembly (P1)
    kernels_ms = sort_ms + count_ms + fill_ms + val_ms
    line = {"metric": f"explicit-mesh P1 assembly seconds ({nx}^3-node Kuhn tet cube, {ne} tets)", "value": kernels_ms * 1e-3, "unit": "s",
            "n_gpus": 1, "steps": args.steps, "warmup": 1, "ms_per_step": best["assemble_ms"], "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"heat_mesh_cube(explicit) -> heat_assemble(P1), {nx}^3 nodes, {ne} tets, n={n_full}, nnz={nnz_full}"},
            "phases_ms": {"n2e_radix_sort": sort_ms, "pattern_count": count_ms, "pattern_fill": fill_ms, "values_kernel": val_ms,
                          "whole_call_events": best["assemble_ms"], "whole_call_wall": best["wall_ms"]},
            "roofline_assembly": {"bound": "hbm", "algorithmic_bytes": bytes_alg, "achieved": bytes_alg / (kernels_ms * 1e-3) / 1e9, "peak": peak,
                                  "unit": "GB/s", "frac": bytes_alg / (kernels_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                                  "values_kernel_gbs": (16 * ne + 24 * N + 8 * nnz_full + 8 * n_full) / (val_ms * 1e-3) / 1e9 if val_ms else None},
            "parity": parity, "runs": runs, "gpu_launches": hb.kernel_launches()}
    print(json.dumps(line), flush=True)
    return 0 if parity["ok"] else 3


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import heat_b200 as hb

    job = Job(args)
    torch, world, rank, dist, stream = job.torch, job.world, job.rank, job.dist, job.stream
    io = job.new_io()

    nx = args.nx
    ny, nz = nx, nx
    if args.workload == "weak":          # BASELINE.json configs[4]: 64 Mi nodes per GPU (512 x 512 x 256 slab each)
        nz = (nx // 2) * world
    n_full, nnz_full = cube_counts(nx, ny, nz)
    op = hb.OP_P1_FEM if args.operator == "p1" else hb.OP_GRAPH_LAPLACIAN
    io.mesh_cube(nx, ny, nz, False)
    t0 = time.time()
    A, X, B = io.assemble(op, hb.PART_SLAB)
    assemble_wall_ms = (time.time() - t0) * 1e3
    mi = A.info
    assert (mi.n_global, mi.nnz_global) == (n_full, nnz_full), (mi.n_global, mi.nnz_global)
    solver = hb.SOLVER_CG_SINGLE_REDUCE if args.solver == "cg1" else hb.SOLVER_CG
    ips = args.iters_per_step
    prec = {"jacobi": hb.PREC_JACOBI, "chebyshev": hb.PREC_CHEBYSHEV, "none": hb.PREC_NONE}[args.prec]
    kw = dict(solver=solver, prec=prec, check_every=ips, cheb_degree=args.cheb_degree, cheb_lambda_max=args.cheb_lambda_max)
    barrier, timed = job.barrier, job.timed

    # ---- (0) parity, before anything is timed -------------------------------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_block(job, io, A, B, nx, op == hb.OP_P1_FEM)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "parity": parity, "error": "parity block failed"}), flush=True)
            barrier()
            return 3

    # ---- (1) device-resident CG steps: the headline `value` -------------------------------------
    def step():
        X.fill(0.0)
        r = io.cg_iterations(A, X, B, ips, **kw)
        assert r.iters == ips, r

    sampler = ClockSampler(job.local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    l0 = hb.kernel_launches()
    ms_total = timed(step, args.steps, 0)
    l1 = hb.kernel_launches()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = ips * args.steps / (ms_total * 1e-3)
    peer_on = bool(A.info.peer_path)               # the peer-memory path is set up by the first multi-GPU solve

    # ---- (2) dominant kernel alone: SpMV launches, CUDA events on the same stream ----------------
    xs, ys = A.hash_vector(12345), A.new_vector()
    n_spmv = max(10, ips)
    if world > 1 and peer_on:
        # the launch the CG loop makes on >1 GPU: one peer-path SpMV over [interior | boundary] slices (halo already delivered)
        io.spmv_peer(A, xs, ys, 4)
        ms_spmv = job.allmax(io.spmv_peer(A, xs, ys, n_spmv + 1)[1])
        spmv_how = "peer-path kernel launched back to back on a delivered halo (heat_spmv_peer), library CUDA events, max over ranks"
    else:
        ms_spmv = timed(lambda: io.spmv(A, xs, ys), n_spmv, 3) / n_spmv
        spmv_how = "heat_spmv launches, CUDA events on the launching stream"
    peak, peak_src = measured_peak()

    def traffic_of(colbytes):
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ent = json.load(f).get(f"sell_spmv_tma_kernel|nx={nx}|gpus={world}|colbytes={colbytes}")
                return ent["traffic_bytes"] if ent else None
        except OSError:
            return None

    def roofline_of(ms, colbytes, how):
        gbs = spmv_bytes(n_full, nnz_full) / (ms * 1e-3) / 1e9 / world
        tr = traffic_of(colbytes)
        return {"bound": "hbm", "kernel": "sell_spmv_tma_kernel (fp64 SELL-64 SpMV, TMA-staged, "
                                          + ("1-byte table-indexed columns)" if colbytes == 1 else "int32 columns)"),
                "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": tr, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": spmv_bytes(n_full, nnz_full) / world, "ms_per_launch": ms, "launches_timed": n_spmv,
                "timed_as": how,
                # the byte-indexed format moves FEWER bytes than the algorithmic CSR figure (12 B/nnz), hence frac > 1 is possible;
                # frac_of_traffic = DRAM bytes actually moved (ncu, per launch) / launch time / peak
                "dram_gbs_from_traffic": (tr / (ms * 1e-3) / 1e9) if tr else None,
                "frac_of_traffic": (tr / (ms * 1e-3) / 1e9 / peak) if tr else None,
                "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)" if tr else None}

    roofline = roofline_of(ms_spmv, mi.col_index_bytes, spmv_how)
    cg_gbs_per_gpu = cg_iter_bytes(n_full, nnz_full) * value / 1e9 / world

    # ---- (3) end to end through the host-buffer entry point --------------------------------------
    n_own = mi.n_owned
    b_host = torch.empty(n_own, dtype=torch.float64).pin_memory()
    x_host = torch.zeros(n_own, dtype=torch.float64).pin_memory()
    b_host.copy_(torch.from_numpy(B.numpy()))

    def step_e2e():
        # x_host carries x0 in and the iterate out; each step continues from the previous step's result (no
        # host-side memset of a 1 GB buffer inside the timed region — that would be harness, not solver, time)
        r = io.solve_host(A, b_host, x_host, max_iters=ips, tol=0.0, **kw)
        assert r.iters == ips

    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps, 1)
    e2e_serial = ips * e2e_steps / (ms_e2e * 1e-3)
    res_check = job.allmax(float(np.abs(x_host.numpy()).max()))
    # the same through the batch entry point (a stream of solves with one matrix, e.g. time stepping): every step still
    # copies its own b and x0 up from pinned memory and its own solution down, but the copies of steps k+1 / k-1 overlap
    # the solve of step k (two staging sets, two copy streams)
    e2e_n = max(4, 2 * e2e_steps)
    x_out = [torch.empty(n_own, dtype=torch.float64).pin_memory() for _ in range(2)]
    x_in = torch.zeros(n_own, dtype=torch.float64).pin_memory()

    def batch_e2e():
        rs = io.solve_host_batch(A, [b_host] * e2e_n, [x_in] * e2e_n, [x_out[k & 1] for k in range(e2e_n)], max_iters=ips, tol=0.0, **kw)
        assert all(r.iters == ips for r in rs)

    ms_batch = timed(batch_e2e, 1, 1)
    e2e_value = ips * e2e_n / (ms_batch * 1e-3)
    del x_out, x_in
    # host<->device copy rate of this rank alone (the same pinned buffers, nothing else running on this rank)
    d_tmp = torch.empty(n_own, dtype=torch.float64, device="cuda")
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record(stream); d_tmp.copy_(b_host, non_blocking=True); ev[1].record(stream); x_host.copy_(d_tmp, non_blocking=True); ev[2].record(stream)
    barrier()
    h2d_gbs, d2h_gbs = 8e-6 * n_own / ev[0].elapsed_time(ev[1]), 8e-6 * n_own / ev[1].elapsed_time(ev[2])
    del d_tmp

    # ---- (4) the general (int32-column) SpMV kernel on the same system, and assembly at steady state ----------
    extra = {}
    if args.workload == "strong" and not args.quick:
        for v in (xs, ys):
            v.free()
        del b_host, x_host
        os.environ["HEAT_SPMV_CIDX"] = "0"
        t0 = time.time()
        A4, X4, B4 = io.assemble(op, hb.PART_SLAB)                    # second assembly: allocations are warm
        wall2 = (time.time() - t0) * 1e3
        os.environ.pop("HEAT_SPMV_CIDX")
        mi4 = A4.info
        x4, y4 = A4.hash_vector(12345), A4.new_vector()
        if world > 1 and peer_on:
            io.spmv_peer(A4, x4, y4, 4)
            ms4 = job.allmax(io.spmv_peer(A4, x4, y4, n_spmv + 1)[1])
        else:
            ms4 = timed(lambda: io.spmv(A4, x4, y4), n_spmv, 3) / n_spmv
        extra["roofline_int32"] = roofline_of(ms4, 4, spmv_how)
        extra["assemble_again"] = {"assemble_ms": mi4.assemble_ms, "assemble_fill_ms": mi4.assemble_fill_ms, "wall_ms": wall2, "col_index_bytes": mi4.col_index_bytes}
        for v in (x4, y4, X4, B4):
            v.free()
        A4.free()
        xs = ys = None

    # assembly roofline: the fill kernel (cube_sell_kernel: straight into SELL + byte indices) against the bytes it must write
    n_loc, nnz_loc = mi.n_owned, mi.nnz_local
    asm_bytes = 12 * nnz_loc + 8 * n_loc                              # matrix entries (value + column id) + right-hand side
    fill_ms = job.allmax(mi.assemble_fill_ms)
    asm_ms = job.allmax(mi.assemble_ms)
    roofline_assembly = {"bound": "hbm", "kernel": "cube_sell_kernel (analytic Kuhn-cube P1 assembly straight into SELL-64 + byte indices, TMA bulk stores)",
                         "algorithmic_bytes_per_gpu": asm_bytes, "ms_fill_kernel": fill_ms, "ms_whole_assemble_events": asm_ms,
                         "ms_whole_assemble_wall": assemble_wall_ms, "achieved": asm_bytes / (fill_ms * 1e-3) / 1e9 if fill_ms else None,
                         "peak": peak, "unit": "GB/s", "frac": asm_bytes / (fill_ms * 1e-3) / 1e9 / peak if fill_ms else None,
                         "bytes_written_actual": mi.sell_padded_nnz * (8 + mi.col_index_bytes) + 25 * n_loc,
                         "matrix_bytes_resident": mi.matrix_bytes, "csr_resident": bool(mi.csr_resident),
                         "note": "write-only stream; the kernel is issue/latency bound at 16 warps per SM (124 registers, 102 KB of staging per CTA), "
                                 "ncu: DRAM 38 % busy, fp64 pipe 39 %.  Every row evaluates its 24 incident tets (partially evaluated: ~125 fp64 "
                                 "operations + 15 divisions) so that the values stay bit-identical to the oracle's"}

    # ---- (5) weak-scaling point (configs[4]) measured after the strong run on >1 GPU ------------------------
    weak = None
    if world > 1 and args.workload == "strong" and not args.quick and not args.no_weak:
        for v in (X, B):
            v.free()
        A.free()
        A = None
        weak = weak_point(job, args, hb)

    # ---- (6) CPU baseline on rank 0 (bounded sample), N == 1 only ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = cpu_run(nx, args.cpu_iters, 2, 1)
        cpu = cpu_baseline_dict(c, nx, args.cpu_iters, 2)
        cpu["assembly_reference"] = reference_assembly_sample()

    if rank == 0:
        metric = METRIC
        if args.workload == "weak" or args.prec != "jacobi" or nx != 512:
            metric = (f"CG iters/s ({args.prec}-PCG fp64, {nx}x{ny}x{nz}-node P1-tet heat, {n_full / 1e6:.1f}M DOF"
                      + (", weak scaling 64Mi nodes/GPU" if args.workload == "weak" else "") + ")")
        config = {"workload": f"synthetic {nx}x{ny}x{nz}-node Kuhn tet cube ({'P1 FEM' if op else 'graph Laplacian'}), "
                              f"slab-partitioned over {world} GPU(s), {args.prec}-PCG ({args.solver})",
                  "comm": "peer-memory" if mi.nranks > 1 and peer_on else ("nccl" if mi.nranks > 1 else "none"),
                  "n_dof": n_full, "nnz": nnz_full, "iters_per_step": ips, "solver": args.solver,
                  "l2_policy": "inputs (>=38 GB per iteration sweep) far exceed the 126 MB L2; no flush",
                  "assemble_ms": mi.assemble_ms, "assemble_fill_ms": mi.assemble_fill_ms, "assemble_wall_ms": assemble_wall_ms,
                  "sell_padding": mi.sell_padded_nnz / max(mi.nnz_local, 1) - 1.0, "spmv_col_index_bytes": mi.col_index_bytes}
        if weak:
            config["weak"] = weak
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.workload == "weak" else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16 * n_own * world, "d2h_bytes_per_step": 8 * n_own * world,
                    "steps": e2e_n, "ms_per_step": ms_batch / e2e_n, "check_max_abs_x": res_check,
                    "api": "heat_solve_host_batch: a stream of solves, each step's H2D (b, x0) and D2H (x) from/to pinned memory inside the "
                           "timed region, overlapped with the neighbouring steps' solves",
                    "serial_value": e2e_serial, "serial_ms_per_step": ms_e2e / e2e_steps,
                    "serial_api": "heat_solve_host: one blocking call per step, copies not overlapped",
                    "rank0_h2d_gbs": h2d_gbs, "rank0_d2h_gbs": d2h_gbs},
            "gpu_launches": l1 - l0,     # counted by the library (heat_kernel_launches), rank 0, timed steps only
            "parity": parity,
            "roofline": roofline,
            "roofline_cg_iteration": ({"bound": "hbm", "achieved": cg_gbs_per_gpu, "peak": peak, "unit": "GB/s",
                                       "frac": cg_gbs_per_gpu / peak,
                                       "algorithmic_bytes_per_iteration": cg_iter_bytes(n_full, nnz_full) / world}
                                      if args.prec == "jacobi" else
                                      {"bound": "hbm", "achieved": cheb_iter_bytes(n_full, nnz_full, args.cheb_degree) * value / 1e9 / world, "peak": peak,
                                       "unit": "GB/s", "frac": cheb_iter_bytes(n_full, nnz_full, args.cheb_degree) * value / 1e9 / world / peak,
                                       "algorithmic_bytes_per_iteration": cheb_iter_bytes(n_full, nnz_full, args.cheb_degree) / world}),
            "roofline_assembly": roofline_assembly,
            "clocks": clocks,
        }
        line.update(extra)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    barrier()
    for v in (xs, ys, X, B):
        if v is not None and v.h:
            v.free()
    if A is not None:
        A.free()
    io.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def weak_point(job, args, hb):
    """BASELINE.json configs[4] next to the strong-scaling line: 64 Mi nodes per GPU (nx x nx x nx/2 slab each),
    Chebyshev(k)-Jacobi PCG.  it/s on all N GPUs, and on ONE GPU running one GPU's share alone (rank 0, the others
    idle) — weak-scaling efficiency = the ratio, both from this process set and build."""
    nx, k = args.nx, args.cheb_degree
    kw = dict(solver=hb.SOLVER_CG, prec=hb.PREC_CHEBYSHEV, cheb_degree=k, cheb_lambda_max=args.cheb_lambda_max or 2.0)
    ips = max(10, args.iters_per_step // 2)

    def run(io, nz, collective):
        io.mesh_cube(nx, nx, nz, False)
        A, X, B = io.assemble(hb.OP_P1_FEM, hb.PART_SLAB)
        peer = bool(A.info.peer_path)

        def step():
            X.fill(0.0)
            r = io.cg_iterations(A, X, B, ips, check_every=ips, **kw)
            assert r.iters == ips, r
        if collective:
            ms = job.timed(step, 2, 1)
        else:
            step()
            job.torch.cuda.synchronize()
            e0, e1 = job.torch.cuda.Event(enable_timing=True), job.torch.cuda.Event(enable_timing=True)
            e0.record(job.stream); step(); step(); e1.record(job.stream)
            job.torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        peer = peer or bool(A.info.peer_path)
        n = A.info.n_global
        for v in (X, B):
            v.free()
        A.free()
        io.close()
        return 2 * ips / (ms * 1e-3), n, peer

    its1 = n1 = None
    if job.rank == 0:
        its1, n1, _ = run(job.new_io(comm=False), nx // 2, False)
    job.barrier()
    itsN, nN, peer = run(job.new_io(), (nx // 2) * job.world, True)
    if job.rank != 0:
        return None
    return {"workload": f"{nx}x{nx}x{(nx // 2) * job.world}-node cube = {nx}x{nx}x{nx // 2} nodes per GPU, chebyshev({k})-PCG",
            "n_dof": nN, "iters_per_s": itsN, "iters_per_s_one_gpu_share_alone": its1, "n_dof_one_gpu": n1,
            "weak_efficiency": itsN / its1, "comm": "peer-memory" if peer else "nccl", "iters_timed": 2 * ips}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=None, help="nodes per cube edge (BASELINE.json configs[3]: 512)")
    ap.add_argument("--iters-per-step", type=int, default=100,
                    help="CG iterations per solver call (a 512^3 solve to 1e-10 needs thousands: SURVEY.md Appendix E)")
    ap.add_argument("--solver", default="cg", choices=["cg", "cg1"])
    ap.add_argument("--operator", default="p1", choices=["p1", "graph"])
    ap.add_argument("--workload", default="strong", choices=["strong", "weak"],
                    help="strong: nx^3 cube over N GPUs (configs[3]); weak: nx*nx*(nx/2) nodes PER GPU (configs[4])")
    ap.add_argument("--assemble", default=None, choices=["explicit"],
                    help="explicit: time the explicit-connectivity P1 assembly of the nx^3 cube (default nx 256) instead of the solve")
    ap.add_argument("--prec", default="jacobi", choices=["jacobi", "chebyshev", "none"])
    ap.add_argument("--cheb-degree", type=int, default=3)
    ap.add_argument("--cheb-lambda-max", type=float, default=0.0)
    ap.add_argument("--cpu-iters", type=int, default=3, help="CPU arm: PCG iterations per step (bounded sample of the same workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the configs[4] weak-scaling point after the strong run")
    ap.add_argument("--quick", action="store_true", help="skip the int32-kernel / re-assembly / weak extras")
    args = ap.parse_args()
    if args.nx is None:
        args.nx = 256 if args.assemble else 512
    if args.impl == "reference":
        return run_reference(args)
    if args.assemble:
        return run_assemble_explicit(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
