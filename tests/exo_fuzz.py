"""Random small Exodus-II files for the fuzz tests (written with scipy.io.netcdf_file in Exodus' netCDF conventions):
1-3 element blocks of one type (TRI3 / TRI / TETRA / TETRA4 / HEX8) with random connectivity, 0-3 nodesets (random ids,
possibly overlapping, unsorted, with or without distribution factors), 0-2 sidesets, with or without node_num_map /
elem_map, stored as float64 or float32.  Deliberately ugly meshes: isolated nodes, disconnected pieces, DOFs with only Dirichlet neighbours."""
import numpy as np
from scipy.io import netcdf_file


def write_exodus(path, rng, N, blocks, nodesets, sidesets, with_nmap, with_emap, title, ndim=3, real="d"):
    nc = netcdf_file(path, "w", version=2)
    nc.title = title.encode()
    nc.api_version = np.float32(8.03)
    nc.version = np.float32(8.03)
    nc.floating_point_word_size = np.int32(8 if real == "d" else 4)
    nc.file_size = np.int32(1)
    ne = sum(len(c) for _, c in blocks)
    nc.createDimension("time_step", None)
    for name, ln in (("len_string", 33), ("len_line", 81), ("four", 4), ("len_name", 33), ("num_dim", ndim), ("num_nodes", N),
                     ("num_elem", ne), ("num_el_blk", len(blocks))):
        nc.createDimension(name, ln)
    nc.createVariable("time_whole", real, ("time_step",))
    v = nc.createVariable("eb_status", "i", ("num_el_blk",)); v[:] = 1
    v = nc.createVariable("eb_prop1", "i", ("num_el_blk",)); v[:] = np.arange(10, 10 + len(blocks)); v.name = b"ID"
    if nodesets:
        nc.createDimension("num_node_sets", len(nodesets))
        v = nc.createVariable("ns_status", "i", ("num_node_sets",)); v[:] = 1
        v = nc.createVariable("ns_prop1", "i", ("num_node_sets",)); v[:] = [s[0] for s in nodesets]; v.name = b"ID"
    if sidesets:
        nc.createDimension("num_side_sets", len(sidesets))
        v = nc.createVariable("ss_status", "i", ("num_side_sets",)); v[:] = 1
        v = nc.createVariable("ss_prop1", "i", ("num_side_sets",)); v[:] = [s[0] for s in sidesets]; v.name = b"ID"
    for k, (_, nodes, df) in enumerate(nodesets, 1):
        nc.createDimension(f"num_nod_ns{k}", len(nodes))
        v = nc.createVariable(f"node_ns{k}", "i", (f"num_nod_ns{k}",)); v[:] = nodes + 1
        if df is not None:
            v = nc.createVariable(f"dist_fact_ns{k}", real, (f"num_nod_ns{k}",)); v[:] = df
    for k, (_, elems, sides, df) in enumerate(sidesets, 1):
        nc.createDimension(f"num_side_ss{k}", len(elems))
        v = nc.createVariable(f"elem_ss{k}", "i", (f"num_side_ss{k}",)); v[:] = elems + 1
        v = nc.createVariable(f"side_ss{k}", "i", (f"num_side_ss{k}",)); v[:] = sides
        if df is not None:
            nc.createDimension(f"num_df_ss{k}", len(df))
            v = nc.createVariable(f"dist_fact_ss{k}", real, (f"num_df_ss{k}",)); v[:] = df
    for k, (etype, conn) in enumerate(blocks, 1):
        nc.createDimension(f"num_el_in_blk{k}", conn.shape[0])
        nc.createDimension(f"num_nod_per_el{k}", conn.shape[1])
        v = nc.createVariable(f"connect{k}", "i", (f"num_el_in_blk{k}", f"num_nod_per_el{k}")); v[:] = conn + 1
        v.elem_type = etype.encode()
    for nm in ("coordx", "coordy", "coordz")[:ndim]:
        v = nc.createVariable(nm, real, ("num_nodes",)); v[:] = rng.uniform(-5, 5, N)
    v = nc.createVariable("coor_names", "c", ("num_dim", "len_name"))
    for i in range(ndim):
        v[i, 0] = b"xyz"[i:i + 1]
    if with_nmap:
        v = nc.createVariable("node_num_map", "i", ("num_nodes",)); v[:] = rng.permutation(N) + 1
    if with_emap:
        v = nc.createVariable("elem_map", "i", ("num_elem",)); v[:] = rng.permutation(ne) + 1
    nc.close()


def random_case(seed: int):
    """-> (rng, N, blocks, nodesets, sidesets, with_nmap, with_emap, title, partitions)"""
    rng = np.random.default_rng(seed)
    N = int(rng.integers(4, 40))
    etype, npe = [("TRI3", 3), ("TETRA", 4), ("HEX8", 8), ("TRI", 3), ("TETRA4", 4)][int(rng.integers(0, 5))]
    N = max(N, npe + 1)
    blocks = []
    for _ in range(int(rng.integers(1, 4))):
        ne = int(rng.integers(1, 12))
        blocks.append((etype, np.stack([rng.choice(N, size=npe, replace=False) for _ in range(ne)]).astype(np.int32)))
    ne_tot = sum(len(c) for _, c in blocks)
    ids = rng.choice(np.arange(1, 2000), size=6, replace=False)
    nodesets = []
    for k in range(int(rng.integers(0, 4))):
        nodes = np.sort(rng.choice(N, size=int(rng.integers(1, max(2, N // 3))), replace=False))
        if rng.random() < 0.5:
            nodes = rng.permutation(nodes)
        nodesets.append((int(ids[k]), nodes, rng.uniform(0, 2, len(nodes)) if rng.random() < 0.5 else None))
    sidesets = []
    for k in range(int(rng.integers(0, 3))):
        m = int(rng.integers(1, 6))
        sidesets.append((int(ids[3 + k]), rng.integers(0, ne_tot, m), rng.integers(1, 4, m),
                         rng.uniform(0, 1, 2 * m) if rng.random() < 0.5 else None))
    return rng, N, blocks, nodesets, sidesets, bool(rng.random() < 0.5), bool(rng.random() < 0.5), f"fuzz {seed}", int(rng.integers(2, 5))


def write_random(path: str, seed: int) -> int:
    """writes the file of `seed`; -> the number of partitions to decompose it into"""
    rng, N, blocks, nodesets, sidesets, nm, em, title, parts = random_case(seed)
    write_exodus(path, rng, N, blocks, nodesets, sidesets, nm, em, title, real="f" if seed % 3 == 2 else "d")   # every third file is float32
    return parts
