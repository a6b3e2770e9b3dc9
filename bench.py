#!/usr/bin/env python
"""bench.py — headline benchmark of the heat path: Jacobi-PCG iterations/s and SpMV HBM GB/s on the
synthetic 512^3-node Kuhn tet cube (133.7 M DOF, P1 operator), strong-scaled over N B200s.

    python bench.py --gpus 1 --steps K --warmup W                 (N>1: launched by torchrun)
    python bench.py --impl reference ...                           (CPU restatement of the reference)

A "step" = one call of the solver entry point running `--iters-per-step` CG iterations on the
resident system (plus the r0 = b - A x0 set-up SpMV every solve pays).  `value` = CG iterations/s of
the whole job with inputs resident in HBM; `e2e` = the same through the C ABI's host-buffer entry
point (heat_solve_host: H2D of b and x0, D2H of x inside the timed region).  One JSON line on
stdout (rank 0).  Timing: CUDA events on the stream the kernels run on, barrier + synchronize on
both sides, max over ranks.  The working set (>= 38 GB) is >> the 126 MB L2, so no L2 flush.
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


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
# reference arm / cpu_baseline: the CPU restatement of the reference path (oracle), all host threads
# ---------------------------------------------------------------------------------------------------
def cpu_run(sample_n, iters, steps, warmup, full_n):
    """Times `iters` classical Jacobi-PCG iterations of the C/OpenMP oracle on a sample_n^3 cube (P1) and
    scales iterations/s to the full workload by the DOF ratio (an iteration is a bandwidth-bound
    O(n) sweep).  Returns (scaled it/s, raw it/s, ms per step, threads, assemble seconds)."""
    import oracle as O
    t0 = time.time()
    mesh = O.cube_mesh(sample_n, sample_n, sample_n)
    sysm = O.assemble(mesh, O.P1_FEM)
    t_asm = time.time() - t0
    times = []
    for s in range(warmup + steps):
        t = time.time()
        O.pcg(sysm, prec=O.PREC_JACOBI, max_iters=iters, tol=0.0)
        if s >= warmup:
            times.append(time.time() - t)
    dt = sum(times) / len(times)
    raw = iters / dt
    return raw * sysm.n / full_n, raw, dt * 1e3, O.num_threads(), t_asm, sysm.n, sysm.nnz


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nx = args.nx
    n_full, nnz_full = cube_counts(nx, nx, nx)
    scaled, raw, ms, threads, t_asm, n_s, nnz_s = cpu_run(args.cpu_sample, args.cpu_iters, max(args.steps, 1), min(args.warmup, 1), n_full)
    sample = (f"{args.cpu_sample}^3-node cube ({n_s} DOF, P1), {args.cpu_iters} Jacobi-PCG iterations per step; "
              f"iterations/s scaled by DOF ratio {n_s}/{n_full}; raw {raw:.2f} it/s; CPU assembly {t_asm:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": scaled, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {nx}x{nx}x{nx}-node Kuhn tet cube (P1 FEM), jacobi-PCG (cg), CPU restatement", "n_dof": n_full, "nnz": nnz_full,
                   "note": "CPU restatement of the reference path (C/OpenMP oracle); the Trilinos/Belos binary cannot be built here"},
        "cpu_baseline": {"value": scaled, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "assembly_reference": reference_assembly_sample()},
        "e2e": {"value": scaled, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import heat_b200 as hb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    stream = torch.cuda.current_stream()
    io = hb.IO(local_rank, stream)
    if world > 1:
        idt = torch.zeros(hb.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(hb.IO.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        io.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))

    nx = args.nx
    ny, nz = nx, nx
    if args.workload == "weak":          # BASELINE.json configs[4]: 64 Mi nodes per GPU (512 x 512 x 256 slab each)
        nz = (nx // 2) * world
    n_full, nnz_full = cube_counts(nx, ny, nz)
    op = hb.OP_P1_FEM if args.operator == "p1" else hb.OP_GRAPH_LAPLACIAN
    io.mesh_cube(nx, ny, nz, False)
    A, X, B = io.assemble(op, hb.PART_SLAB)
    mi = A.info
    assert (mi.n_global, mi.nnz_global) == (n_full, nnz_full), (mi.n_global, mi.nnz_global)
    solver = hb.SOLVER_CG_SINGLE_REDUCE if args.solver == "cg1" else hb.SOLVER_CG
    ips = args.iters_per_step
    prec = {"jacobi": hb.PREC_JACOBI, "chebyshev": hb.PREC_CHEBYSHEV, "none": hb.PREC_NONE}[args.prec]
    kw = dict(solver=solver, prec=prec, check_every=ips, cheb_degree=args.cheb_degree, cheb_lambda_max=args.cheb_lambda_max)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- (1) device-resident CG steps: the headline `value` -------------------------------------
    def step():
        X.fill(0.0)
        r = io.cg_iterations(A, X, B, ips, **kw)
        assert r.iters == ips, r

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launch_mark = {}

    def timed_counted(fn, steps, warmup):        # `timed`, also counting this library's launches inside the timed steps
        for _ in range(warmup):
            fn()
        launch_mark["l0"] = hb.kernel_launches()
        ms = timed(fn, steps, 0)
        launch_mark["l1"] = hb.kernel_launches()
        return ms

    ms_total = timed_counted(step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = ips * args.steps / (ms_total * 1e-3)

    # ---- (2) dominant kernel alone: SpMV launches, CUDA events on the same stream ----------------
    xs, ys = A.hash_vector(12345), A.new_vector()
    n_spmv = max(10, ips)
    ms_spmv = timed(lambda: io.spmv(A, xs, ys), n_spmv, 3) / n_spmv
    peak, peak_src = measured_peak()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f).get(f"sell_spmv_tma_kernel|nx={nx}|gpus={world}|colbytes={mi.col_index_bytes}")
            traffic = ent["traffic_bytes"] if ent else None
    except OSError:
        pass
    spmv_gbs_total = spmv_bytes(n_full, nnz_full) / (ms_spmv * 1e-3) / 1e9        # all ranks together
    per_gpu_gbs = spmv_gbs_total / world
    cg_gbs_per_gpu = cg_iter_bytes(n_full, nnz_full) * value / 1e9 / world

    # ---- (3) end to end through the host-buffer entry point --------------------------------------
    n_own = mi.n_owned
    b_host = torch.empty(n_own, dtype=torch.float64).pin_memory()
    x_host = torch.zeros(n_own, dtype=torch.float64).pin_memory()
    b_host.copy_(torch.from_numpy(B.numpy()))
    o = io.solve_opts(**kw)

    def step_e2e():
        # x_host carries x0 in and the iterate out; each step continues from the previous step's result (no
        # host-side memset of a 1 GB buffer inside the timed region — that would be harness, not solver, time)
        r = io.solve_host(A, b_host, x_host, max_iters=ips, tol=0.0, **kw)
        assert r.iters == ips

    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps, 1)
    e2e_value = ips * e2e_steps / (ms_e2e * 1e-3)
    res_check = float(np.abs(x_host.numpy()).max())

    # ---- (4) CPU baseline on rank 0 (bounded sample), N == 1 only ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        scaled, raw, ms_cpu, threads, t_asm, n_s, nnz_s = cpu_run(args.cpu_sample, args.cpu_iters, 2, 1, n_full)
        cpu = {"value": scaled, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": (f"C/OpenMP restatement of the reference path on a {args.cpu_sample}^3-node cube ({n_s} DOF, P1), "
                          f"{args.cpu_iters} Jacobi-PCG iterations x2; scaled by DOF ratio to {nx}^3; raw {raw:.2f} it/s; "
                          f"CPU assembly {t_asm:.1f} s"),
               "assembly_reference": reference_assembly_sample()}

    if rank == 0:
        metric = METRIC
        if args.workload == "weak" or args.prec != "jacobi" or nx != 512:
            metric = (f"CG iters/s ({args.prec}-PCG fp64, {nx}x{ny}x{nz}-node P1-tet heat, {n_full / 1e6:.1f}M DOF"
                      + (", weak scaling 64Mi nodes/GPU" if args.workload == "weak" else "") + ")")
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.workload == "weak" else "strong",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {nx}x{ny}x{nz}-node Kuhn tet cube ({'P1 FEM' if op else 'graph Laplacian'}), "
                                   f"slab-partitioned over {world} GPU(s), {args.prec}-PCG ({args.solver})",
                       "comm": "peer-memory" if mi.nranks > 1 and A.info.peer_path else ("nccl" if mi.nranks > 1 else "none"),
                       "n_dof": n_full, "nnz": nnz_full, "iters_per_step": ips, "solver": args.solver,
                       "l2_policy": "inputs (>=38 GB per iteration sweep) far exceed the 126 MB L2; no flush",
                       "assemble_ms": mi.assemble_ms, "sell_padding": mi.sell_padded_nnz / max(mi.nnz_local, 1) - 1.0,
                       "spmv_col_index_bytes": mi.col_index_bytes},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16 * n_own * world, "d2h_bytes_per_step": 8 * n_own * world,
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps, "check_max_abs_x": res_check},
            "gpu_launches": launch_mark["l1"] - launch_mark["l0"],     # counted by the library (heat_kernel_launches), rank 0
            "roofline": {"bound": "hbm", "kernel": "sell_spmv_tma_kernel (fp64 SELL-64 SpMV, TMA-staged"
                                                          + (", byte-indexed columns)" if mi.col_index_bytes == 1 else ")"), "achieved": per_gpu_gbs, "peak": peak,
                         "unit": "GB/s", "frac": per_gpu_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": spmv_bytes(n_full, nnz_full) / world, "ms_per_launch": ms_spmv,
                         "launches_timed": n_spmv,
                         # the byte-indexed format moves FEWER bytes than the algorithmic CSR figure, hence frac > 1;
                         # the DRAM rate actually sustained = ncu traffic per launch / the launch time measured here
                         "dram_gbs_from_traffic": (traffic / (ms_spmv * 1e-3) / 1e9) if traffic else None},
            "roofline_cg_iteration": ({"bound": "hbm", "achieved": cg_gbs_per_gpu, "peak": peak, "unit": "GB/s",
                                       "frac": cg_gbs_per_gpu / peak,
                                       "algorithmic_bytes_per_iteration": cg_iter_bytes(n_full, nnz_full) / world}
                                      if args.prec == "jacobi" else None),
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    barrier()
    for v in (xs, ys, X, B):
        v.free()
    A.free()
    io.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=512, help="nodes per cube edge (BASELINE.json configs[3]: 512)")
    ap.add_argument("--iters-per-step", type=int, default=100,
                    help="CG iterations per solver call (a 512^3 solve to 1e-10 needs thousands: SURVEY.md Appendix E)")
    ap.add_argument("--solver", default="cg", choices=["cg", "cg1"])
    ap.add_argument("--operator", default="p1", choices=["p1", "graph"])
    ap.add_argument("--workload", default="strong", choices=["strong", "weak"],
                    help="strong: nx^3 cube over N GPUs (configs[3]); weak: nx*nx*(nx/2) nodes PER GPU (configs[4])")
    ap.add_argument("--prec", default="jacobi", choices=["jacobi", "chebyshev", "none"])
    ap.add_argument("--cheb-degree", type=int, default=3)
    ap.add_argument("--cheb-lambda-max", type=float, default=0.0)
    ap.add_argument("--cpu-sample", type=int, default=160, help="cube edge of the bounded CPU sample")
    ap.add_argument("--cpu-iters", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
