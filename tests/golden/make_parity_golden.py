#!/usr/bin/env python
"""Golden fixtures for the multi-GPU parity block of bench.py (and tests/test_gpu_multi.py), made on the ORACLE
side only — nothing of the product is imported here:

  * N=1 reference solutions (Jacobi-PCG of the C oracle to a relative residual of 1e-10, iteration counts) of
      - data/bolted_bracket.exo, graph Laplacian (the reference's own operator; BASELINE.json configs[1]),
      - the 33x17x16-node Kuhn cube, P1 and graph mode (configs[3] in the small: >= 2 k-planes per rank up to 8 ranks);
  * for N = 1, 2, 4, 8 ranks and every rank, the sha1 digest of the owned / ghost / ghost-owner / neighbour / send maps
    (Tpetra conventions, ExodusIO.hpp:252, :656 + the Import plan of fillComplete, :609) derived with an independent
    numpy plan builder from the oracle's pattern and
      - METIS_PartGraphKway called from the oracle side on the row graph (bolted_bracket),
      - the slab rule (cube: k-planes dealt to ranks, the first nz % N ranks get one more).

Run:  python tests/golden/make_parity_golden.py     ->  tests/golden/parity_multi.npz + parity_multi.json
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

CUBE = (33, 17, 16)
RANKS = (1, 2, 4, 8)


def digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(np.asarray(a).ravel(), dtype="<i8").tobytes())
        h.update(b"|")
    return h.hexdigest()


def plan_np(row_ptr, col, part, P, r):
    """owned rows ascending; ghosts grouped by owner ascending, ascending gid inside an owner; send list to q = q's
    ghosts owned by r in q's ghost order (numpy restatement, independent of csrc/partition.cpp)"""
    n = len(row_ptr) - 1
    rows = np.repeat(np.arange(n), np.diff(row_ptr))

    def ghosts_of(q):
        c = np.unique(col[part[rows] == q])
        g = c[part[c] != q]
        order = np.lexsort((g, part[g]))
        return g[order]

    owned = np.flatnonzero(part == r)
    ghost = ghosts_of(r)
    owner = part[ghost]
    nbr_recv = set(np.unique(owner).tolist())
    send = {}
    for q in range(P):
        if q == r:
            continue
        gq = ghosts_of(q)
        mine = gq[part[gq] == r]
        if len(mine):
            send[q] = mine
    nbr = sorted(nbr_recv | set(send))
    send_ptr, recv_ptr, send_gids = [0], [0], []
    for q in nbr:
        s = send.get(q, np.zeros(0, dtype=np.int64))
        send_gids.append(s)
        send_ptr.append(send_ptr[-1] + len(s))
        recv_ptr.append(recv_ptr[-1] + int((owner == q).sum()))
    send_gids = np.concatenate(send_gids) if send_gids else np.zeros(0, dtype=np.int64)
    return owned, ghost, owner, np.array(nbr, dtype=np.int64), np.array(send_ptr), send_gids, np.array(recv_ptr)


def slab_part(red2orig, nx, ny, nz, P):
    k = red2orig // (nx * ny)
    base, rem = divmod(nz, P)
    bounds = np.array([q * base + min(q, rem) for q in range(P + 1)])
    return (np.searchsorted(bounds, k, side="right") - 1).astype(np.int64)


def main():
    arrays, meta = {}, {"cube": list(CUBE), "tol": 1e-10, "cases": {}}
    cases = []
    mesh = O.read_exodus(os.path.join(HERE, "meshes", "bolted_bracket.exo"))
    cases.append(("bolted_bracket_graph", O.assemble(mesh, O.GRAPH_LAPLACIAN), "metis"))
    cmesh = O.cube_mesh(*CUBE)
    cases.append(("cube_p1", O.assemble(cmesh, O.P1_FEM), "slab"))
    cases.append(("cube_graph", O.assemble(cmesh, O.GRAPH_LAPLACIAN), "slab"))
    for name, ref, how in cases:
        x, it, ach, _ = O.pcg(ref, tol=1e-10, max_iters=5000)
        assert ach <= 1e-10
        arrays[name + "_x"] = x
        ent = {"n": int(ref.n), "nnz": int(ref.nnz), "iters": int(it), "x_sha1": hashlib.sha1(x.tobytes()).hexdigest(),
               "partition": how, "maps": {}}
        for P in RANKS:
            if how == "metis":
                rows = np.repeat(np.arange(ref.n), np.diff(ref.row_ptr))
                off = ref.col != rows                                   # METIS graph: no self loops
                xadj = np.concatenate([[0], np.cumsum(np.bincount(rows[off], minlength=ref.n))])
                part = O.metis_part_graph_kway(xadj, ref.col[off], P)[1].astype(np.int64)
            else:
                part = slab_part(ref.red2orig, *CUBE, P)
            ent["maps"][str(P)] = [digest(*plan_np(ref.row_ptr, ref.col.astype(np.int64), part, P, r)) for r in range(P)]
            ent.setdefault("part_sha1", {})[str(P)] = digest(part)
        meta["cases"][name] = ent
    np.savez_compressed(os.path.join(HERE, "parity_multi.npz"), **arrays)
    with open(os.path.join(HERE, "parity_multi.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({k: (v["n"], v["iters"]) for k, v in meta["cases"].items()}))


if __name__ == "__main__":
    main()
