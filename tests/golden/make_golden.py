"""Regenerates tests/golden/ from the reference's data files (run in the build container, where
/root/reference exists).  The meshes are the reference's own fixtures; the numbers
are computed with the INDEPENDENT numpy/scipy restatement (oracle.assemble_np + a sparse direct
solve), not with the C oracle they are later used to pin.

    python tests/golden/make_golden.py
"""
import json
import os
import shutil
import sys

import numpy as np
import scipy.sparse.linalg as spl

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import oracle as O  # noqa: E402

REF = "/root/reference/data"
MESHES = ["rectangle-tris-boundary", "rectangle-tris", "2blocks", "bolted_bracket", "tet-cube-heat", "mitchell_tri"]


def stats(A, b, x):
    return {
        "n": int(A.shape[0]), "nnz": int(A.nnz), "trace": float(A.diagonal().sum()),
        "sum_b": float(b.sum()), "asym": float(abs(A - A.T).sum()),
        "x_min": float(x.min()), "x_max": float(x.max()), "x_mean": float(x.mean()),
        "x_norm2": float(np.linalg.norm(x)), "x_head": [float(v) for v in x[:3]],
        "row_sum_abs_max": float(np.abs(np.asarray(A.sum(1)).ravel()).max()),
    }


def main():
    os.makedirs(os.path.join(HERE, "meshes"), exist_ok=True)
    for m in MESHES:
        shutil.copyfile(f"{REF}/{m}.exo", os.path.join(HERE, "meshes", f"{m}.exo"))
        os.chmod(os.path.join(HERE, "meshes", f"{m}.exo"), 0o644)
    gold = {}
    for m in ["rectangle-tris-boundary", "bolted_bracket", "tet-cube-heat", "mitchell_tri"]:
        mesh = O.read_exodus(f"{REF}/{m}.exo")
        entry = {"num_nodes": mesh.num_nodes, "num_elem": int(mesh.conn.shape[0]), "npe": int(mesh.conn.shape[1]),
                 "nodesets": {str(k): int(len(v)) for k, v in mesh.nodesets.items()}}
        for mode, name in ((O.GRAPH_LAPLACIAN, "graph"), (O.P1_FEM, "p1")):
            A, b, r2o = O.assemble_np(mesh, mode)
            x = spl.spsolve(A.tocsc(), b)
            entry[name] = stats(A, b, x)
        A, b, r2o = O.assemble_np(mesh, O.GRAPH_LAPLACIAN, bug_compat_d1=True)
        x = spl.spsolve(A.tocsc(), b)
        entry["graph_bug_compat_d1"] = stats(A, b, x)
        gold[m] = entry
    # small dense golden: the 3x3 hand-checkable system
    mesh = O.read_exodus(f"{REF}/rectangle-tris-boundary.exo")
    A, b, _ = O.assemble_np(mesh, O.GRAPH_LAPLACIAN)
    gold["rectangle-tris-boundary"]["graph"]["A_dense"] = A.toarray().tolist()
    gold["rectangle-tris-boundary"]["graph"]["b"] = b.tolist()
    # METIS known answers with the bundled library and the reference's call (ExodusIO.hpp:1615)
    met = {}
    for m, ncommon in (("rectangle-tris", 2), ("rectangle-tris-boundary", 2), ("2blocks", 3), ("bolted_bracket", 3)):
        mesh = O.read_exodus(f"{REF}/{m}.exo")
        for nparts in (2, 4):
            obj, epart, npart = O.metis_part_mesh_dual(mesh.conn, mesh.num_nodes, ncommon, nparts)
            rec = {"objval": obj, "epart_hist": np.bincount(epart, minlength=nparts).tolist(),
                   "npart_hist": np.bincount(npart, minlength=nparts).tolist()}
            if mesh.conn.shape[0] <= 100:
                rec["epart"] = epart.tolist()
                rec["npart"] = npart.tolist()
            met[f"{m}:{nparts}"] = rec
    gold["metis_part_mesh_dual"] = met
    # IO::getMatrix (ExodusIO.hpp:733-1489): whole-mesh Laplacian statistics from an independent scipy
    # assembly + its largest eigenvalue (scipy eigsh) = what ExodusMatrixTest's power method converges to;
    # node ownership histogram of the rank-by-rank restatement for the METIS element partition
    import scipy.sparse as sp
    gm = {}
    for m, ncommon in (("bolted_bracket", 3), ("mitchell_tri", 2), ("rectangle-tris", 2)):
        mesh = O.read_exodus(f"{REF}/{m}.exo")
        c = mesh.conn.astype(np.int64)
        npe = c.shape[1]
        rows = np.repeat(c, npe, axis=1).ravel()
        cols = np.tile(c, (1, npe)).ravel()
        keep = rows != cols
        Adj = sp.coo_matrix((np.ones(keep.sum()), (rows[keep], cols[keep])), shape=(mesh.num_nodes,) * 2).tocsr()
        Adj.data[:] = 1.0
        Lap = (sp.diags(np.asarray(Adj.sum(1)).ravel()) - Adj).tocsr()
        lmax = float(spl.eigsh(Lap, k=1, which="LA", return_eigenvectors=False, tol=1e-12)[0]) if mesh.num_nodes > 20 else \
            float(np.linalg.eigvalsh(Lap.toarray())[-1])
        rec = {"n": int(Lap.shape[0]), "nnz": int(Lap.nnz), "trace": float(Lap.diagonal().sum()), "lambda_max": lmax}
        for nparts in (2, 4):
            if c.shape[0] < 4 * nparts:
                continue
            _, epart, _ = O.metis_part_mesh_dual(mesh.conn, mesh.num_nodes, ncommon, nparts)
            owners = O.get_matrix_owners(mesh.conn, epart, nparts, mesh.num_nodes)
            rec[f"owner_hist:{nparts}"] = np.bincount(owners, minlength=nparts).tolist()
        gm[m] = rec
    gold["get_matrix"] = gm
    with open(os.path.join(HERE, "golden_values.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "golden_values.json"))


if __name__ == "__main__":
    main()
