"""dump_exo.py — flattens an Exodus-II file (netCDF classic) into the record container that
oracle/ref_shim/exodus_shim.cpp serves to the reference's ExodusIO.hpp, and reads such containers back.
TEST INFRASTRUCTURE (oracle/ref_shim/README.md).

The file is read with scipy.io.netcdf_file, i.e. NOT with the product's own netCDF reader, so the
reference run stays independent of the code it is used to check.

Container: a sequence of   "<name> <dtype> <count>\\n" + count raw little-endian items,
dtype in {i32, f64, f32, str}.

    python oracle/ref_shim/dump_exo.py mesh.exo out_dir      ->  out_dir/mesh.exo.dump
"""
import os
import sys

import numpy as np

_SZ = {"i32": 4, "f64": 8, "f32": 4, "str": 1}
_NP = {"i32": "<i4", "f64": "<f8", "f32": "<f4"}


def _rec(fp, name, dtype, data):
    if dtype == "str":
        raw = data if isinstance(data, bytes) else str(data).encode("ascii", "replace")
        fp.write(f"{name} str {len(raw)}\n".encode())
        fp.write(raw)
        return
    a = np.ascontiguousarray(np.asarray(data).ravel().astype(_NP[dtype]))
    fp.write(f"{name} {dtype} {a.size}\n".encode())
    fp.write(a.tobytes())


def _cstr(arr) -> bytes:
    raw = b"".join(np.asarray(arr).ravel().tolist())
    return raw.split(b"\x00")[0].rstrip()


def dump(exo_path: str, out_dir: str) -> str:
    from scipy.io import netcdf_file

    nc = netcdf_file(exo_path, "r", mmap=False)
    v, d = nc.variables, nc.dimensions
    N, ne = int(d["num_nodes"]), int(d.get("num_elem", 0) or 0)
    ndim = int(d["num_dim"])
    nblk = int(d.get("num_el_blk", 0) or 0)
    nns = int(d.get("num_node_sets", 0) or 0)
    nss = int(d.get("num_side_sets", 0) or 0)
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, os.path.basename(exo_path) + ".dump")
    with open(out, "wb") as fp:
        title = getattr(nc, "title", b"")
        _rec(fp, "title", "str", title if isinstance(title, bytes) else str(title).encode())
        for k, val in (("num_dim", ndim), ("num_nodes", N), ("num_elem", ne), ("num_el_blk", nblk),
                       ("num_node_sets", nns), ("num_side_sets", nss)):
            _rec(fp, k, "i32", [val])
        if "coordx" in v:
            _rec(fp, "coordx", "f64", v["coordx"].data)
            _rec(fp, "coordy", "f64", v["coordy"].data if "coordy" in v else np.zeros(N))
            if "coordz" in v:
                _rec(fp, "coordz", "f64", v["coordz"].data)
        else:
            c = np.array(v["coord"].data, dtype=np.float64)
            for i, nm in enumerate(("coordx", "coordy", "coordz")[:ndim]):
                _rec(fp, nm, "f64", c[i])
        # libexodus hands out the identity when a map is not stored
        _rec(fp, "node_num_map", "i32", v["node_num_map"].data if "node_num_map" in v else np.arange(1, N + 1))
        # ex_get_map reads the element ORDER map "elem_map" only (not the id map "elem_num_map")
        _rec(fp, "elem_map", "i32", v["elem_map"].data if "elem_map" in v else np.arange(1, ne + 1))
        if nblk:
            _rec(fp, "eb_ids", "i32", v["eb_prop1"].data)
        for b in range(1, nblk + 1):
            cv = v[f"connect{b}"]
            et = getattr(cv, "elem_type", b"")
            conn = np.array(cv.data)
            _rec(fp, f"eb{b}_type", "str", (et if isinstance(et, bytes) else str(et).encode()).rstrip(b"\x00 "))
            _rec(fp, f"eb{b}_nelem", "i32", [conn.shape[0]])
            _rec(fp, f"eb{b}_npe", "i32", [conn.shape[1]])
            _rec(fp, f"eb{b}_conn", "i32", conn)
        if nns:
            _rec(fp, "ns_ids", "i32", v["ns_prop1"].data)
        for s in range(1, nns + 1):
            _rec(fp, f"ns{s}_entries", "i32", v[f"node_ns{s}"].data if f"node_ns{s}" in v else [])
            if f"dist_fact_ns{s}" in v:
                _rec(fp, f"ns{s}_df", "f64", v[f"dist_fact_ns{s}"].data)
        if nss:
            _rec(fp, "ss_ids", "i32", v["ss_prop1"].data)
        for s in range(1, nss + 1):
            _rec(fp, f"ss{s}_entries", "i32", v[f"elem_ss{s}"].data if f"elem_ss{s}" in v else [])
            _rec(fp, f"ss{s}_extra", "i32", v[f"side_ss{s}"].data if f"side_ss{s}" in v else [])
            if f"dist_fact_ss{s}" in v:
                _rec(fp, f"ss{s}_df", "f64", v[f"dist_fact_ss{s}"].data)
    nc.close()
    return out


def dump_arrays(name: str, out_dir: str, x, y, z, blocks, nodesets: dict, title: str = "synthetic") -> str:
    """Writes the container for a mesh given as arrays (no Exodus file involved): blocks = [(elem_type, conn
    [ne, npe] 0-based), ...], nodesets = {id: 0-based nodes}.  -> path of "<out_dir>/<name>.dump"; hand the
    reference the file name "<name>"."""
    N = len(x)
    ne = sum(len(c) for _, c in blocks)
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name + ".dump")
    with open(out, "wb") as fp:
        _rec(fp, "title", "str", title.encode())
        for k, val in (("num_dim", 3 if z is not None else 2), ("num_nodes", N), ("num_elem", ne), ("num_el_blk", len(blocks)),
                       ("num_node_sets", len(nodesets)), ("num_side_sets", 0)):
            _rec(fp, k, "i32", [val])
        _rec(fp, "coordx", "f64", x)
        _rec(fp, "coordy", "f64", y)
        if z is not None:
            _rec(fp, "coordz", "f64", z)
        _rec(fp, "node_num_map", "i32", np.arange(1, N + 1))
        _rec(fp, "elem_map", "i32", np.arange(1, ne + 1))
        _rec(fp, "eb_ids", "i32", np.arange(1, len(blocks) + 1))
        for b, (etype, conn) in enumerate(blocks, 1):
            conn = np.asarray(conn)
            _rec(fp, f"eb{b}_type", "str", etype.encode())
            _rec(fp, f"eb{b}_nelem", "i32", [conn.shape[0]])
            _rec(fp, f"eb{b}_npe", "i32", [conn.shape[1]])
            _rec(fp, f"eb{b}_conn", "i32", conn + 1)
        _rec(fp, "ns_ids", "i32", list(nodesets))
        for s, (sid, nodes) in enumerate(nodesets.items(), 1):
            _rec(fp, f"ns{s}_entries", "i32", np.asarray(nodes, dtype=np.int64) + 1)
    return out


def load(path: str) -> dict:
    """Reads a container (a .dump made here or a .shimdump written by the shim's ex_close)."""
    recs = {}
    with open(path, "rb") as fp:
        while True:
            head = fp.readline()
            if not head:
                break
            name, dtype, count = head.decode().split()
            raw = fp.read(int(count) * _SZ[dtype])
            recs[name] = raw if dtype == "str" else np.frombuffer(raw, dtype=_NP[dtype]).copy()
    return recs


if __name__ == "__main__":
    print(dump(sys.argv[1], sys.argv[2]))
