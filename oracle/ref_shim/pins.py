"""pins.py — canonical summaries of what the reference computed (run_ref.run_reference) and of what the oracle /
the product computed, so that the two can be compared bit-exactly through a small committed JSON fixture
(tests/golden/ref_pins.json).  TEST INFRASTRUCTURE (oracle/ref_shim/README.md).

Every array is hashed as little-endian bytes of a fixed dtype (sha256); arrays of at most SMALL items are
stored in full as well, so the tiny meshes stay readable.  Floating-point data that went through the
reference's Exodus calls is float32 (real_t of its METIS build): the product writes float64, so
its values are cast to float32 before hashing — the cast is the only tolerance in these comparisons.
"""
import hashlib

import numpy as np

SMALL = 64


def sha(a, dtype) -> str:
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a).ravel().astype(dtype)).tobytes()).hexdigest()[:32]


def _arr(a, dtype):
    a = np.asarray(a).ravel()
    d = {"n": int(a.size), "sha": sha(a, dtype)}
    if a.size <= SMALL:
        d["data"] = [float(v) if np.dtype(dtype).kind == "f" else int(v) for v in a.astype(dtype)]
    return d


# ---- matrices -----------------------------------------------------------------------------------------------
def summ_csr(n, rows, rowptr, cols, vals) -> dict:
    """rows: ids of the rows that exist (the reference never inserts some rows); columns ascending."""
    rows, rowptr, cols, vals = (np.asarray(v) for v in (rows, rowptr, cols, vals))
    rid = np.repeat(rows, np.diff(rowptr))
    return {"n": int(n), "nrows": int(rows.size), "nnz": int(cols.size), "trace": float(vals[rid == cols].sum()),
            "sum": float(vals.sum()), "rows": _arr(rows, "<i8"), "rowptr": _arr(rowptr, "<i8"), "cols": _arr(cols, "<i8"),
            "vals": _arr(vals, "<f8")}


def summ_scipy(A, row_base: int = 0) -> dict:
    """same summary from a scipy matrix that stores every row; rows with no entry count as absent"""
    A = A.tocsr()
    A.sort_indices()
    keep = np.flatnonzero(np.diff(A.indptr) > 0)
    rowptr = np.r_[0, np.cumsum(np.diff(A.indptr)[keep])]
    return summ_csr(A.shape[0], keep + row_base, rowptr, A.indices + row_base, A.data)


def apply_d1(A, b):
    """What the reference's off-by-one (`i < getMaxLocalIndex()` at ExodusIO.hpp:220/:440 skips the
    last node) does to the FIXED system when the last mesh node is a degree of freedom: its row and column
    disappear, and every entry that referenced it lands in column 0 (`sparseMapping[...]` default-inserts 0 at
    :598) where fillComplete sums it with what is there.  A: scipy FIXED matrix, b: FIXED right-hand side."""
    import scipy.sparse as sp
    A = A.tocsr()
    n = A.shape[0] - 1
    last = A[:n, n].toarray().ravel()
    E = sp.csr_matrix((last, (np.arange(n), np.zeros(n, dtype=np.int64))), shape=(n, n))
    out = (A[:n, :n] + E).tocsr()
    out.sort_indices()
    return out, np.asarray(b)[:n].copy()


def rows_the_reference_inserts(conn, num_nodes: int, nodesets: dict) -> np.ndarray:
    """a DOF row reaches insertGlobalValues only if the node has at least one neighbour that is in no
    nodeset (ExodusIO.hpp:380-386, :591).  -> boolean per DOF node, in reduced (ascending node id) order."""
    conn = np.asarray(conn, dtype=np.int64)
    in_set = np.zeros(num_nodes, dtype=bool)
    for nodes in nodesets.values():
        in_set[np.asarray(nodes, dtype=np.int64)] = True
    has = np.zeros(num_nodes, dtype=bool)
    free_in_elem = (~in_set)[conn]                                   # [ne, npe]
    nfree = free_in_elem.sum(1)
    for a in range(conn.shape[1]):                                   # node a of an element sees the OTHER free nodes of it
        others = nfree - free_in_elem[:, a]
        np.logical_or.at(has, conn[:, a], others > 0)
    return has[~in_set]


def reference_view(A, b, r2o, conn, num_nodes: int, nodesets: dict):
    """The FIXED system (scipy A, RHS b, reduced->original map r2o; every node outside the nodesets is a DOF, rows
    with only Dirichlet neighbours keep their diagonal) seen through the reference's defects, i.e. what the
    reference's own IO::assemble returns for the same mesh:
      D3  rows of DOFs without a DOF neighbour are never inserted and have no id-map entry (B keeps its entry);
      D1  if the last mesh node is a DOF it is dropped: apply_d1.
    -> (A_ref, b_ref, idmap_reduced, idmap_original)"""
    import scipy.sparse as sp
    A = A.tocsr()
    inserted = rows_the_reference_inserts(conn, num_nodes, nodesets)
    n = A.shape[0]
    assert inserted.size == n
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    keep = inserted[rows]
    A = sp.csr_matrix((A.data[keep], (rows[keep], A.indices[keep])), shape=A.shape)
    r2o = np.asarray(r2o)
    last_is_dof = n > 0 and r2o[-1] == num_nodes - 1
    if last_is_dof:
        A, b = apply_d1(A, b)
        inserted, r2o = inserted[:-1], r2o[:-1]
    kept = np.flatnonzero(inserted)
    return A, np.asarray(b), kept, r2o[kept]


def summ_assemble(d: dict) -> dict:
    """d = run_reference(...)["assemble"]"""
    s = {"A": summ_csr(int(d["n"][0]), d["A_rows"], d["A_rowptr"], d["A_cols"], d["A_vals"]),
         "B": _arr(d["B"], "<f8"), "sum_B": float(d["B"].sum()), "B_rows": _arr(d["B_rows"], "<i8"),
         "idmap_reduced": _arr(d["idmap_reduced"], "<i8"), "idmap_original": _arr(d["idmap_original"], "<i8"),
         "nodesets": {str(int(i)): _arr(d[f"ns_{int(i)}"], "<i8") for i in d["ns_ids"]}}
    return s


def summ_dump_text(text: str) -> dict:
    """the "<prefix><rank>.out" file of the reference (printCrsMatrix / printMultiVector, BelosMueLuSolver.cpp:37-84)
    with the ~usec~ time stamps removed, as mpi_output_combiner.py removes them"""
    import re
    lines = [re.sub(r" ~[0-9]*~$", "", l) for l in text.splitlines()]
    s = {"lines": len(lines), "sha": hashlib.sha256("\n".join(lines).encode()).hexdigest()[:32], "head": lines[:3], "tail": lines[-2:]}
    if len(lines) <= SMALL:
        s["text"] = lines
    return s


def parse_power_log(text: str) -> dict:
    """the lines PowerMethod::run prints (ExodusMatrixTest.cpp:107-124)"""
    import re
    reports = [{"iter": int(i), "lambda": float(l), "residual": float(r)} for i, l, r in re.findall(
        r"Iteration (\d+):\s*- lambda = (\S+)\s*- \|\|A\*q - lambda\*q\|\|_2 = (\S+)", text)]
    conv = re.search(r"Converged after (\d+) iterations", text)
    fail = re.search(r"Failed to converge after (\d+) iterations", text)
    return {"reports": reports, "converged": conv is not None, "stop_iter": int(conv.group(1)) if conv else int(fail.group(1))}


def summ_getmatrix(d: dict) -> dict:
    s = {"A": summ_csr(int(d["n"][0]), d["A_rows"], d["A_rowptr"], d["A_cols"], d["A_vals"]),
         "nodesets": {str(int(i)): _arr(d[f"ns_{int(i)}"], "<i8") for i in d["ns_ids"]}}
    if "pm_lambda" in d:
        s["power_method"] = dict(parse_power_log(d["pm_log"].decode()), **{"lambda": float(d["pm_lambda"][0]), "niters": 500,
                                 "tolerance": 1.0e-2, "start": "splitmix64 U(-1,1) keyed on the 0-based row index, seed 12345"})
    return s


# ---- output file ----------------------------------------------------------------------------------------------
def canon_from_shimdump(d: dict) -> dict:
    """what the reference handed to ex_put_* (record names of exodus_shim.cpp's write side)"""
    def i(name, dflt=0):
        return int(d[name][0]) if name in d and len(d[name]) else dflt
    c = {"title": d.get("title", b"").decode("ascii", "replace").rstrip("\x00 "), "num_dim": i("num_dim"), "num_nodes": i("num_nodes"),
         "num_elem": i("num_elem"), "num_el_blk": i("num_el_blk"), "num_node_sets": i("num_node_sets"),
         "num_side_sets": i("num_side_sets"),
         "coords": [d[k] for k in ("coordx", "coordy", "coordz") if k in d],
         "elem_map": d.get("elem_map"), "node_num_map": d.get("node_num_map"), "blocks": [], "nodesets": [], "sidesets": []}
    for k in range(1, i("eb_written") + 1):
        c["blocks"].append({"id": i(f"eb{k}_id"), "type": d[f"eb{k}_type"].decode(), "nelem": i(f"eb{k}_nelem"), "npe": i(f"eb{k}_npe"),
                            "conn": d.get(f"eb{k}_conn", np.zeros(0, np.int32))})
    for k in range(1, i("ns_written") + 1):
        c["nodesets"].append({"id": i(f"ns{k}_id"), "entries": d.get(f"ns{k}_entries", np.zeros(0, np.int32)),
                              "df": d.get(f"ns{k}_df", np.zeros(0, np.float32))})
    for k in range(1, i("ss_written") + 1):
        c["sidesets"].append({"id": i(f"ss{k}_id"), "elems": d.get(f"ss{k}_entries", np.zeros(0, np.int32)),
                              "sides": d.get(f"ss{k}_extra", np.zeros(0, np.int32)), "df": d.get(f"ss{k}_df", np.zeros(0, np.float32))})
    nvar = i("num_nod_var")
    c["var_names"] = [d[f"name_nod_var{v}"].decode() for v in range(1, nvar + 1)]
    nsteps = i("num_time_steps")
    c["times"] = [float(d[f"time{s}"][0]) for s in range(1, nsteps + 1)]
    c["steps"] = [d[f"vals_nod_var1_step{s}"] for s in range(1, nsteps + 1)] if nvar else []
    return c


def canon_from_exodus(path: str) -> dict:
    """the same structure from a real Exodus-II file (the product's output), read with scipy.io.netcdf_file"""
    from scipy.io import netcdf_file
    nc = netcdf_file(path, "r", mmap=False)
    v, dm = nc.variables, nc.dimensions

    def dim(name):
        return int(dm.get(name, 0) or 0)

    def cstr(a):
        return b"".join(np.asarray(a).ravel().tolist()).split(b"\x00")[0].decode("ascii", "replace").rstrip()
    title = nc.title.decode("ascii", "replace") if isinstance(nc.title, bytes) else str(nc.title)
    c = {"title": title.rstrip("\x00 "), "num_dim": dim("num_dim"), "num_nodes": dim("num_nodes"), "num_elem": dim("num_elem"),
         "num_el_blk": dim("num_el_blk"), "num_node_sets": dim("num_node_sets"), "num_side_sets": dim("num_side_sets"),
         "coords": [np.array(v[k].data) for k in ("coordx", "coordy", "coordz") if k in v],
         "elem_map": np.array(v["elem_map"].data) if "elem_map" in v else None,
         "node_num_map": np.array(v["node_num_map"].data) if "node_num_map" in v else None,
         "blocks": [], "nodesets": [], "sidesets": []}
    if "coord" in v and not c["coords"]:
        c["coords"] = [np.array(r) for r in np.array(v["coord"].data)]
    for k in range(1, c["num_el_blk"] + 1):
        cv = v[f"connect{k}"]
        conn = np.array(cv.data)
        c["blocks"].append({"id": int(v["eb_prop1"].data[k - 1]), "type": cv.elem_type.decode().rstrip("\x00 "), "nelem": int(conn.shape[0]),
                            "npe": int(conn.shape[1]), "conn": conn.ravel()})
    for k in range(1, c["num_node_sets"] + 1):
        c["nodesets"].append({"id": int(v["ns_prop1"].data[k - 1]),
                              "entries": np.array(v[f"node_ns{k}"].data) if f"node_ns{k}" in v else np.zeros(0, np.int32),
                              "df": np.array(v[f"dist_fact_ns{k}"].data) if f"dist_fact_ns{k}" in v else np.zeros(0)})
    for k in range(1, c["num_side_sets"] + 1):
        c["sidesets"].append({"id": int(v["ss_prop1"].data[k - 1]),
                              "elems": np.array(v[f"elem_ss{k}"].data) if f"elem_ss{k}" in v else np.zeros(0, np.int32),
                              "sides": np.array(v[f"side_ss{k}"].data) if f"side_ss{k}" in v else np.zeros(0, np.int32),
                              "df": np.array(v[f"dist_fact_ss{k}"].data) if f"dist_fact_ss{k}" in v else np.zeros(0)})
    c["var_names"] = [cstr(r) for r in np.array(v["name_nod_var"].data)] if "name_nod_var" in v else []
    c["times"] = [float(t) for t in np.array(v["time_whole"].data)] if "time_whole" in v else []
    c["steps"] = [np.array(r) for r in np.array(v["vals_nod_var1"].data)] if "vals_nod_var1" in v else []
    nc.close()
    return c


def summ_output(c: dict) -> dict:
    s = {k: c[k] for k in ("title", "num_dim", "num_nodes", "num_elem", "num_el_blk", "num_node_sets", "num_side_sets", "var_names")}
    s["coords"] = [_arr(a, "<f4") for a in c["coords"]]
    s["elem_map"] = _arr(c["elem_map"], "<i8") if c["elem_map"] is not None else None
    s["node_num_map"] = _arr(c["node_num_map"], "<i8") if c["node_num_map"] is not None else None
    s["blocks"] = [{"id": b["id"], "type": b["type"], "nelem": b["nelem"], "npe": b["npe"], "conn": _arr(b["conn"], "<i8")} for b in c["blocks"]]
    s["nodesets"] = [{"id": n["id"], "entries": _arr(n["entries"], "<i8"), "df": _arr(n["df"], "<f4")} for n in c["nodesets"]]
    s["sidesets"] = [{"id": x["id"], "elems": _arr(x["elems"], "<i8"), "sides": _arr(x["sides"], "<i8"), "df": _arr(x["df"], "<f4")}
                     for x in c["sidesets"]]
    s["times"] = [float(np.float32(t)) for t in c["times"]]
    s["steps"] = [_arr(a, "<f4") for a in c["steps"]]
    return s
