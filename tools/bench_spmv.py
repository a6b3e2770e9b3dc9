#!/usr/bin/env python
"""Times one SpMV variant (HEAT_SPMV_VARIANT) and the CG iteration on a synthetic cube; prints JSON.
Used to pick the default SpMV kernel: run once per variant, compare ms and the y checksum (all
variants must be bit-identical)."""
import argparse, hashlib, json, os, sys

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=512)
ap.add_argument("--ny", type=int, default=0)
ap.add_argument("--nz", type=int, default=0)
ap.add_argument("--variant", type=int, default=5)
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--solver", default="cg")
ap.add_argument("--operator", default="p1")
args = ap.parse_args()
os.environ["HEAT_SPMV_VARIANT"] = str(args.variant)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "domain-decomposed-pde-solver_b200"))
import torch
import heat_b200 as hb

stream = torch.cuda.current_stream()
io = hb.IO(0, stream)
io.mesh_cube(args.nx, args.ny or args.nx, args.nz or args.nx)
A, X, B = io.assemble(hb.OP_P1_FEM if args.operator == "p1" else hb.OP_GRAPH_LAPLACIAN)
mi = A.info
x, y = A.hash_vector(12345), A.new_vector()


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: io.spmv(A, x, y), args.reps)
yh = y.numpy()
bytes_spmv = 12 * mi.nnz_global + 16 * mi.n_global + 4 * (mi.n_global + 1)
solver = hb.SOLVER_CG_SINGLE_REDUCE if args.solver == "cg1" else hb.SOLVER_CG
ips = 30


def step():
    X.fill(0.0)
    io.cg_iterations(A, X, B, ips, solver=solver, check_every=ips)


ms_it = timed(step, 3, 1) / ips
print(json.dumps({"variant": args.variant, "nx": args.nx, "ny": args.ny or args.nx, "nz": args.nz or args.nx, "spmv_ms": ms, "spmv_GBs": bytes_spmv / ms / 1e6,
                  "frac_of_6546.9": bytes_spmv / ms / 1e6 / 6546.9, "cg_ms_per_iter": ms_it, "cg_it_per_s": 1e3 / ms_it,
                  "cg_GBs": (bytes_spmv + 88 * mi.n_global) / ms_it / 1e6, "solver": args.solver,
                  "y_sha1": hashlib.sha1(yh.tobytes()).hexdigest()[:16]}))
