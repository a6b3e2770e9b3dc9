#!/usr/bin/env python
"""Assembly timing: analytic-cube vs explicit-connectivity kernels on an n^3 cube, and Exodus meshes."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "domain-decomposed-pde-solver_b200"))
import heat_b200 as hb

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, nargs="+", default=[128, 256])
ap.add_argument("--explicit-max", type=int, default=256)
args = ap.parse_args()
for nx in args.nx:
    for explicit in (False, True):
        if explicit and nx > args.explicit_max:
            continue
        for mode in (hb.OP_GRAPH_LAPLACIAN, hb.OP_P1_FEM):
            best = None
            for rep in range(3):
                io = hb.IO(0)
                io.mesh_cube(nx, nx, nx, explicit)
                t = time.time()
                A, X, B = io.assemble(mode)
                wall = (time.time() - t) * 1e3
                mi = A.info
                best = min(best, mi.assemble_ms) if best else mi.assemble_ms
                rec = dict(nx=nx, explicit=explicit, mode=mode, n=mi.n_global, nnz=mi.nnz_global, ne=mi.num_elem)
                io.close()
            ne, N, nnz, n = rec["ne"], nx ** 3, rec["nnz"], rec["n"]
            alg = (16 * ne + 24 * N if explicit else 0) + 12 * nnz + 8 * n
            rec.update(assemble_ms=best, wall_ms_last=wall, alg_GB=alg / 1e9, GBs=alg / best / 1e6)
            print(json.dumps(rec), flush=True)
