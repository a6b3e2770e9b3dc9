#!/usr/bin/env python
"""Small single-purpose targets for `ncu --set full -k regex:<kernel> -c 1` (no torch import: the process starts in a
second).    python tools/ncu_targets.py assemble 384      analytic cube assembly (cube_width_kernel, cube_sell_kernel)
            python tools/ncu_targets.py explicit 256      explicit-mesh assembly (radix sort, pattern_*, values_kernel)
            python tools/ncu_targets.py cg 512 [iters]    a few Jacobi-PCG iterations (SpMV, cg_update_xr/p)
            python tools/ncu_targets.py cheb 512 [iters]  a few Chebyshev(3)-PCG iterations (fused kernels)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "domain-decomposed-pde-solver_b200"))
import heat_b200 as hb

what = sys.argv[1]
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
if what == "assemble2":          # steady-state assembly: the same cube four times (byte-indexed x2, int32 x2)
    for cidx in ("1", "1", "0", "0"):
        os.environ["HEAT_SPMV_CIDX"] = cidx
        io = hb.IO(0)
        io.mesh_cube(nx, nx, nx, False)
        A, X, B = io.assemble(hb.OP_P1_FEM)
        mi = A.info
        print("assemble", nx, "col_index_bytes", mi.col_index_bytes, "assemble_ms", round(mi.assemble_ms, 2), "fill_ms", round(mi.assemble_fill_ms, 2), flush=True)
        io.close()
    sys.exit(0)
io = hb.IO(0)
io.mesh_cube(nx, nx, nx, what == "explicit")
A, X, B = io.assemble(hb.OP_P1_FEM)
mi = A.info
print(what, nx, "assemble_ms", mi.assemble_ms, "fill_ms", mi.assemble_fill_ms, "phases", list(mi.asm_phase_ms), "bytes", mi.matrix_bytes)
if what == "cg":
    r = io.cg_iterations(A, X, B, iters)
    print("iters", r.iters, "ms", r.solve_ms)
elif what == "cheb":
    r = io.cg_iterations(A, X, B, iters, prec=hb.PREC_CHEBYSHEV, cheb_degree=3, cheb_lambda_max=2.0)
    print("iters", r.iters, "ms", r.solve_ms)
io.close()
