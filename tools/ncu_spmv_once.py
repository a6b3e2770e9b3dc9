#!/usr/bin/env python
"""Minimal target for `ncu --set full -k regex:sell_spmv_tma_kernel -s 3 -c 1`: assembles the synthetic cube and
launches the default SpMV a few times (no torch import, so the process starts in a second)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "domain-decomposed-pde-solver_b200"))
import heat_b200 as hb

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
io = hb.IO(0)
io.mesh_cube(nx, nx, nx)
A, X, B = io.assemble(hb.OP_P1_FEM)
x, y = A.hash_vector(12345), A.new_vector()
for _ in range(6):
    io.spmv(A, x, y)
print("checksum", float(y.numpy()[:1000].sum()), "col_index_bytes", A.info.col_index_bytes)
io.close()
