#!/usr/bin/env python
"""Small pass over every kernel family for `compute-sanitizer --tool memcheck` (one GPU, seconds; no torch import).
(On the pool this round ran on, compute-sanitizer is closed; the script also serves as a plain 10-second tour of the API.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "domain-decomposed-pde-solver_b200"))
import heat_b200 as hb

MESH = os.path.join(ROOT, "tests", "golden", "meshes", "bolted_bracket.exo")


def run(io, A, X, B, tag):
    x, y = A.hash_vector(1), A.new_vector()
    io.spmv(A, x, y)
    r1 = io.solve(A, X, B, max_iters=40, tol=1e-10, check_every=7)
    X.fill(0.0)
    r2 = io.solve(A, X, B, prec=hb.PREC_CHEBYSHEV, cheb_degree=3, cheb_lambda_max=2.2, max_iters=30, tol=1e-10)
    X.fill(0.0)
    r3 = io.solve(A, X, B, solver=hb.SOLVER_CG_SINGLE_REDUCE, max_iters=30, tol=1e-10)
    n = A.info.n_owned
    outs = [np.empty(n) for _ in range(3)]
    io.solve_host_batch(A, [B.numpy()] * 3, [None, np.ones(n), None], outs, max_iters=20, tol=1e-10)
    pr = io.power_method(A, 5, 0.0, 3)
    rp, col, val = A.csr()
    print(tag, "col_index_bytes", A.info.col_index_bytes, r1.iters, r2.iters, r3.iters, round(pr.lambda_, 6), len(val), flush=True)


for dims, explicit in (((40, 33, 29), False), ((66, 3, 2), False), ((12, 9, 11), True)):
    for mode in (hb.OP_GRAPH_LAPLACIAN, hb.OP_P1_FEM):
        io = hb.IO(0)
        io.mesh_cube(*dims, explicit)
        A, X, B = io.assemble(mode)
        run(io, A, X, B, f"cube{dims} explicit={explicit} mode={mode}")
        io.close()
for mode in (hb.OP_GRAPH_LAPLACIAN, hb.OP_P1_FEM):
    io = hb.IO(0)
    io.open(MESH, True)
    A, X, B = io.assemble(mode)
    run(io, A, X, B, f"bolted_bracket mode={mode}")
    io.close()
os.environ["HEAT_SPMV_CIDX"] = "0"
io = hb.IO(0)
io.mesh_cube(40, 33, 29)
A, X, B = io.assemble(hb.OP_P1_FEM)
run(io, A, X, B, "cube int32")
io.close()
print("sanitize target done")
