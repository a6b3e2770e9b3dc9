"""CPU-only tests of the host logic of libheat_b200 (no compute calls — there is no CPU fallback):
the C ABI exports, the netCDF/Exodus reader and writer, METIS decomposition, row partitioning and
the owned/ghost/send plan, all against independent numpy / scipy restatements."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import MESHES, ROOT, mesh_path

ALL_MESHES = ["rectangle-tris-boundary", "rectangle-tris", "2blocks", "bolted_bracket", "tet-cube-heat", "mitchell_tri"]


@pytest.fixture(scope="module")
def hb():
    import heat_b200
    if not os.path.exists(heat_b200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return heat_b200


@pytest.fixture()
def host_io(hb):
    """host-only context (device = -1): file I/O and decompose only"""
    h = C.c_void_p()
    assert hb.lib().heat_ctx_create(-1, C.byref(h)) == 0
    io = hb.IO.__new__(hb.IO)
    io.h = h
    yield io
    io.close()


def test_abi_exports_every_declared_symbol(hb):
    header = open(os.path.join(ROOT, "include", "heat_b200.h")).read()
    declared = set(re.findall(r"\b(heat_[a-z0-9_]+)\s*\(", header))
    declared -= {"heat_ctx", "heat_matrix", "heat_vector"}
    L = hb.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(hb.ABI_SYMBOLS) == declared
    assert L.heat_version() == 100


def test_compute_fails_loudly_without_gpu(hb, host_io):
    host_io.open(mesh_path("rectangle-tris-boundary"), True)
    with pytest.raises(hb.HeatError, match="no CPU fallback"):
        host_io.assemble()
    if hb.device_count() == 0:
        with pytest.raises(hb.HeatError, match="no CUDA device"):
            hb.IO(0)


def test_open_errors(hb, host_io, tmp_path):
    with pytest.raises(hb.HeatError, match="cannot open"):
        host_io.open(str(tmp_path / "nope.exo"), True)
    bad = tmp_path / "bad.exo"
    bad.write_bytes(b"not a netcdf file at all")
    with pytest.raises(hb.HeatError, match="not a netCDF classic"):
        host_io.open(str(bad), True)
    hdf = tmp_path / "hdf.exo"
    hdf.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(hb.HeatError, match="HDF5"):
        host_io.open(str(hdf), True)
    trunc = tmp_path / "trunc.exo"
    trunc.write_bytes(open(mesh_path("rectangle-tris-boundary"), "rb").read()[:600])
    with pytest.raises(hb.HeatError):
        host_io.open(str(trunc), True)


@pytest.mark.parametrize("name", ALL_MESHES)
def test_decompose_copies_mesh_and_matches_metis(hb, host_io, oracle, name, tmp_path):
    """IO::decompose (ExodusIO.hpp:1496-1969): METIS_PartMeshDual with the reference's arguments, one
    element block per partition, ids from 0; everything else copied.  Our C++ reader + writer are
    checked against scipy.io.netcdf_file (independent implementation)."""
    from scipy.io import netcdf_file
    src = mesh_path(name)
    host_io.open(src, True)
    m = oracle.read_exodus(src)
    ncommon = 2 if m.conn.shape[1] == 3 else 3
    for parts in (2, 3):
        obj, epart, npart = host_io.decompose_partition(parts, m.conn.shape[0], m.num_nodes)
        o2, e2, n2 = oracle.metis_part_mesh_dual(m.conn, m.num_nodes, ncommon, parts)
        assert obj == o2
        np.testing.assert_array_equal(epart, e2)          # partition bit-exact vs the bundled METIS
        np.testing.assert_array_equal(npart, n2)
        out = str(tmp_path / f"{name}.{parts}.exo")
        host_io.create(out)
        host_io.decompose(parts)
        a, b = netcdf_file(src, "r", mmap=False), netcdf_file(out, "r", mmap=False)
        assert b.dimensions["num_nodes"] == m.num_nodes and b.dimensions["num_elem"] == m.conn.shape[0]
        nonempty = [p for p in range(parts) if (epart == p).any()]
        assert b.dimensions["num_el_blk"] == len(nonempty)
        np.testing.assert_array_equal(b.variables["eb_prop1"].data, np.arange(len(nonempty)))
        for k, p in enumerate(nonempty, 1):
            conn = np.array(b.variables[f"connect{k}"].data)
            np.testing.assert_array_equal(conn, m.conn[epart == p] + 1)
            assert b.variables[f"connect{k}"].elem_type.decode().strip() == m.elem_type
        for v in ("coordx", "coordy", "coordz", "node_num_map", "elem_map", "ns_prop1", "ss_prop1", "qa_records", "coor_names"):
            if v in a.variables:
                np.testing.assert_array_equal(np.array(a.variables[v].data), np.array(b.variables[v].data), err_msg=v)
        for v in a.variables:
            if v.startswith(("node_ns", "dist_fact_ns", "elem_ss", "side_ss", "dist_fact_ss")):
                np.testing.assert_array_equal(np.array(a.variables[v].data), np.array(b.variables[v].data), err_msg=v)
        assert b.title == a.title and b.floating_point_word_size == 8
        a.close(); b.close()
        # the file we wrote can be opened again by our own reader
        h2 = C.c_void_p()
        assert hb.lib().heat_ctx_create(-1, C.byref(h2)) == 0
        io2 = hb.IO.__new__(hb.IO); io2.h = h2
        io2.open(out, True)
        io2.close()


def test_decompose_rejects_unsupported_element_type(hb, host_io):
    x = np.array([0.0, 1.0, 2.0]); conn = np.array([[0, 1], [1, 2]], dtype=np.int32)
    host_io.mesh_set(x, x * 0, None, conn, {1: [0]}, num_dim=2)
    with pytest.raises(hb.HeatError, match="unsupported element type"):
        host_io.decompose_partition(2, 2, 3)


# ---- partition + plan -----------------------------------------------------------------------------
def plan_numpy(row_ptr, col, part, P, rank):
    """Independent restatement of the Tpetra column-map / Import conventions."""
    n = len(row_ptr) - 1
    rows = np.repeat(np.arange(n), np.diff(row_ptr))

    def ghosts(q):
        c = col[(part[rows] == q) & (part[col] != q)]
        c = np.unique(c)
        return c[np.lexsort((c, part[c]))]

    owned = np.flatnonzero(part == rank)
    gh = ghosts(rank)
    nbr, send, recv_ptr, send_ptr = [], [], [0], [0]
    for q in range(P):
        if q == rank:
            continue
        gq = ghosts(q)
        sl = gq[part[gq] == rank]
        nrecv = int((part[gh] == q).sum())
        if len(sl) or nrecv:
            nbr.append(q); send.append(sl)
            send_ptr.append(send_ptr[-1] + len(sl)); recv_ptr.append(recv_ptr[-1] + nrecv)
    return dict(owned=owned, ghost=gh, ghost_owner=part[gh], nbr=np.array(nbr, dtype=np.int32),
                send_ptr=np.array(send_ptr), recv_ptr=np.array(recv_ptr),
                send_gids=np.concatenate(send) if send else np.zeros(0, dtype=np.int64))


@pytest.mark.parametrize("name", ["bolted_bracket", "tet-cube-heat", "mitchell_tri"])
@pytest.mark.parametrize("P", [2, 4, 8])
def test_partition_and_plan_bit_exact(hb, oracle, name, P):
    s = oracle.assemble(oracle.read_exodus(mesh_path(name)), 0)
    # METIS k-way on the row graph: same library, same arguments, from the oracle side
    A = s.csr().copy()
    A.setdiag(0); A.eliminate_zeros()
    obj, part_o = oracle.metis_part_graph_kway(A.indptr, A.indices, P)
    part = hb.partition_rows(s.row_ptr, s.col, hb.PART_METIS_KWAY, P)
    np.testing.assert_array_equal(part, part_o)
    assert np.bincount(part, minlength=P).min() > 0
    # contiguous uniform map (Tpetra::Map(n, 0, comm), ExodusIO.hpp:252)
    pc = hb.partition_rows(s.row_ptr, s.col, hb.PART_CONTIGUOUS, P)
    cnt = np.bincount(pc, minlength=P)
    assert np.all(np.diff(pc) >= 0) and cnt.max() - cnt.min() <= 1 and cnt[0] == -(-s.n // P)
    for prt in (part, pc):
        total_send = 0
        for r in range(P):
            got = hb.plan_build(s.row_ptr, s.col, prt, P, r)
            exp = plan_numpy(s.row_ptr, s.col, prt, P, r)
            for k in exp:
                np.testing.assert_array_equal(got[k], exp[k], err_msg=f"{k} rank {r}")
            total_send += len(got["send_gids"])
            assert np.all(prt[got["send_gids"]] == r)
        assert total_send == sum(len(hb.plan_build(s.row_ptr, s.col, prt, P, r)["ghost"]) for r in range(P))


@pytest.mark.parametrize("name", ["bolted_bracket", "mitchell_tri", "2blocks", "rectangle-tris"])
@pytest.mark.parametrize("P", [2, 4, 8])
def test_get_matrix_node_ownership_rule(hb, oracle, golden, name, P):
    """IO::getMatrix's node ownership (ExodusIO.hpp:1191-1295) from the METIS element partition: the
    host routine against the rank-by-rank restatement in the oracle."""
    mesh = oracle.read_exodus(mesh_path(name))
    if mesh.conn.shape[0] < 2 * P:
        pytest.skip("fewer elements than 2 x parts")
    ncommon = {4: 3, 3: 2, 8: 4}[mesh.conn.shape[1]]
    _, epart, _ = oracle.metis_part_mesh_dual(mesh.conn, mesh.num_nodes, ncommon, P)
    exp = oracle.get_matrix_owners(mesh.conn, epart, P, mesh.num_nodes)
    got = hb.node_owners(mesh.conn, mesh.num_nodes, epart, P)
    np.testing.assert_array_equal(got, exp)
    key = f"owner_hist:{P}"
    if name in golden["get_matrix"] and key in golden["get_matrix"][name]:
        assert np.bincount(got, minlength=P).tolist() == golden["get_matrix"][name][key]
    # a node touched by ONE part only belongs to that part
    for v in range(0, mesh.num_nodes, max(1, mesh.num_nodes // 50)):
        parts = {int(epart[e]) for e in np.flatnonzero((mesh.conn == v).any(1))}
        if len(parts) == 1:
            assert got[v] == parts.pop()
    # the plan built from that ownership covers every row exactly once
    ref = oracle.get_matrix(mesh)
    assert sum(len(hb.plan_build(ref.row_ptr, ref.col, got, P, r)["owned"]) for r in range(P)) == mesh.num_nodes


def test_node_owners_rejects_bad_input(hb):
    conn = np.array([[0, 1, 2], [1, 2, 3]], dtype=np.int32)
    with pytest.raises(hb.HeatError):
        hb.node_owners(conn, 4, np.array([0, 2]), 2)          # part id out of range
    with pytest.raises(hb.HeatError):
        hb.node_owners(conn, 3, np.array([0, 1]), 2)          # node id out of range
    # tie (node 1 and 2 have 2 neighbours in each part) -> lowest part; unused node 4 -> part 0
    np.testing.assert_array_equal(hb.node_owners(conn, 5, np.array([1, 0]), 2), [1, 0, 0, 0, 0])


def test_single_rank_plan_is_trivial(hb, oracle):
    s = oracle.assemble(oracle.read_exodus(mesh_path("rectangle-tris-boundary")), 0)
    p = hb.plan_build(s.row_ptr, s.col, np.zeros(s.n, dtype=np.int32), 1, 0)
    assert len(p["owned"]) == 3 and len(p["ghost"]) == 0 and len(p["nbr"]) == 0


def test_cli_driver_usage(hb):
    import subprocess
    exe = os.path.join(ROOT, "domain-decomposed-pde-solver_b200", "bin", "heat_solver")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode != 0 and "No input file was provided; use the '--input' parameter!" in p.stderr


def test_matrix_test_cli_usage(hb):
    import subprocess
    exe = os.path.join(ROOT, "domain-decomposed-pde-solver_b200", "bin", "heat_matrix_test")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode != 0 and "No input file was provided; use the '--input' parameter!" in p.stderr


def test_write_nodal_field_records_in_place(hb, host_io, tmp_path):
    """The file half of IO::writeSolution (ExodusIO.hpp:2027-2069) without a GPU: the first call defines
    the result variables (file laid out again), later calls append / overwrite single records in place;
    the mesh part of the file must survive every update."""
    from scipy.io import netcdf_file
    src = mesh_path("bolted_bracket")
    out = str(tmp_path / "fields.exo")
    host_io.open(src, True)
    host_io.create(out)
    with pytest.raises(hb.HeatError):
        host_io.write_nodal_field(np.zeros(4098), 0)            # before decompose: no output mesh yet
    host_io.decompose(3)
    N = 4098
    rng = np.random.default_rng(0)
    f = [rng.standard_normal(N) for _ in range(6)]
    host_io.write_nodal_field(f[0], 0)
    size1 = os.path.getsize(out)
    host_io.write_nodal_field(f[1], 1)
    host_io.write_nodal_field(f[2], 2)
    assert os.path.getsize(out) == size1 + 2 * (N * 8 + 8)      # two records appended, nothing else rewritten
    host_io.write_nodal_field(f[3], 1)                          # overwrite step 2 in place (the reference's D7 does this to step 1)
    host_io.write_nodal_field(f[4], 5)                          # skipping steps leaves zero records in between
    with pytest.raises(hb.HeatError):
        host_io.write_nodal_field(f[0][:-1], 6)
    nc = netcdf_file(out, "r", mmap=False)
    vals = np.array(nc.variables["vals_nod_var1"].data)
    tw = np.array(nc.variables["time_whole"].data)
    assert vals.shape == (6, N)
    np.testing.assert_array_equal(vals[0], f[0]); np.testing.assert_array_equal(vals[1], f[3])
    np.testing.assert_array_equal(vals[2], f[2]); np.testing.assert_array_equal(vals[5], f[4])
    assert not vals[3].any() and not vals[4].any()
    np.testing.assert_array_equal(tw, [0, 1, 2, 0, 0, 5])
    ref = netcdf_file(src, "r", mmap=False)
    np.testing.assert_array_equal(np.array(nc.variables["coordx"].data), np.array(ref.variables["coordx"].data))
    assert sum(nc.dimensions[f"num_el_in_blk{b}"] for b in (1, 2, 3)) == ref.dimensions["num_elem"]
    nc.close(); ref.close()


def test_write_nodal_field_records_in_place_float32(hb, host_io, tmp_path):
    """the same record discipline with the reference build's word size 4 (heat_ctx_set_output): records of N*4 + 4 bytes"""
    from scipy.io import netcdf_file
    out = str(tmp_path / "fields32.exo")
    host_io.set_output(4)
    host_io.open(mesh_path("mitchell_tri"), True)
    host_io.create(out)
    host_io.decompose(2)
    N = 828
    rng = np.random.default_rng(1)
    f = [rng.standard_normal(N) * 1e3 for _ in range(4)]
    host_io.write_nodal_field(f[0], 0)
    size1 = os.path.getsize(out)
    host_io.write_nodal_field(f[1], 1)
    host_io.write_nodal_field(f[2], 3)                          # step 3 is skipped: a zero record
    assert os.path.getsize(out) == size1 + 3 * (N * 4 + 4)
    host_io.write_nodal_field(f[3], 0)                          # overwrite step 1 in place
    nc = netcdf_file(out, "r", mmap=False)
    vals, tw = np.array(nc.variables["vals_nod_var1"].data), np.array(nc.variables["time_whole"].data)
    assert vals.dtype.itemsize == 4 and tw.dtype.itemsize == 4 and vals.shape == (4, N) and nc.floating_point_word_size == 4
    np.testing.assert_array_equal(vals[0], f[3].astype(np.float32)); np.testing.assert_array_equal(vals[1], f[1].astype(np.float32))
    np.testing.assert_array_equal(vals[3], f[2].astype(np.float32))
    assert not vals[2].any()
    np.testing.assert_array_equal(tw, [0, 1, 0, 3])
    nc.close()



def test_first_write_at_a_later_timestep_keeps_the_boundary_frame(hb, host_io, tmp_path):
    """IO::writeSolution's first call also writes a boundary-condition-only frame as step 1, time 0 (nodeset nodes hold
    their id, unknowns 0; ExodusIO.hpp:2034-2040).  When the first timestep written is k > 0 that frame survives."""
    from scipy.io import netcdf_file
    import oracle as O
    src = mesh_path("bolted_bracket")
    out = str(tmp_path / "late.exo")
    host_io.open(src, True)
    host_io.create(out)
    host_io.decompose(2)
    N = 4098
    f = np.random.default_rng(3).standard_normal(N)
    host_io.write_nodal_field(f, 2)                             # first call: timestep 2 -> step 3
    nc = netcdf_file(out, "r", mmap=False)
    vals, tw = np.array(nc.variables["vals_nod_var1"].data), np.array(nc.variables["time_whole"].data)
    bc = O.read_exodus(src).node_bc()                           # lowest nodeset id per node (the FIXED output rule), NaN for unknowns
    np.testing.assert_array_equal(vals[0], np.where(np.isnan(bc), 0.0, bc))
    assert not vals[1].any()
    np.testing.assert_array_equal(vals[2], f)
    np.testing.assert_array_equal(tw, [0, 0, 2])
    nc.close()


def _write_nc(path, build):
    from scipy.io import netcdf_file
    nc = netcdf_file(path, "w", version=2)
    build(nc)
    nc.close()


def test_malformed_exodus_files_fail_cleanly(hb, host_io, tmp_path):
    """File-controlled sizes are validated before they are used as indices (advisor finding, round 1): a 1-D connect
    table, a short `coord` array and a short ns_prop1 give an error, not an out-of-bounds read."""
    def base(nc, n_nodes=4):
        nc.createDimension("num_dim", 3); nc.createDimension("num_nodes", n_nodes); nc.createDimension("num_elem", 1)
        nc.createDimension("num_el_blk", 1); nc.createDimension("num_el_in_blk1", 1); nc.createDimension("num_nod_per_el1", 4)

    def flat_connect(nc):
        base(nc)
        for ax in "xyz":
            v = nc.createVariable("coord" + ax, "d", ("num_nodes",)); v[:] = [0.0, 1.0, 0.0, 0.0]
        nc.createDimension("four", 4)
        c = nc.createVariable("connect1", "i", ("four",)); c[:] = [1, 2, 3, 4]            # not [elements][nodes]

    def short_coord(nc):
        base(nc)
        nc.createDimension("short", 7)
        v = nc.createVariable("coord", "d", ("short",)); v[:] = np.arange(7.0)             # 3 x 4 values expected
        c = nc.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1")); c[:] = [[1, 2, 3, 4]]

    def short_ns_prop(nc):
        base(nc)
        for ax in "xyz":
            v = nc.createVariable("coord" + ax, "d", ("num_nodes",)); v[:] = [0.0, 1.0, 0.0, 0.0]
        c = nc.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1")); c[:] = [[1, 2, 3, 4]]
        nc.createDimension("num_node_sets", 3); nc.createDimension("one", 1); nc.createDimension("num_nod_ns1", 1)
        p = nc.createVariable("ns_prop1", "i", ("one",)); p[:] = [7]                       # 3 nodesets, 1 id
        ns = nc.createVariable("node_ns1", "i", ("num_nod_ns1",)); ns[:] = [1]

    for name, build in (("flat", flat_connect), ("coord", short_coord), ("nsprop", short_ns_prop)):
        path = str(tmp_path / f"bad_{name}.exo")
        _write_nc(path, build)
        with pytest.raises(hb.HeatError):
            host_io.open(path, True)
