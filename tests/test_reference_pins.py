"""Pins against THE REFERENCE'S OWN CODE.  tests/golden/ref_pins.json holds digests of what
/root/reference/ExodusIO.hpp (compiled unmodified, one rank, oracle/ref_shim/README.md) computed on every mesh of
its data/ directory: IO::assemble's A, B and id map, IO::getMatrix's matrix, and everything IO::decompose and
IO::writeSolution handed to the Exodus API.  Here

  * the numpy restatement with the reference's off-by-one kept (bug_compat_d1) must reproduce them bit for bit;
  * the C oracle and (on the GPU) the product, which implement the FIXED semantics, must reproduce them after
    the one documented transformation `pins.apply_d1` (identity when the last mesh node is in a nodeset);
  * the product's decompose / writeSolution output FILE must carry the same records, floats compared as
    float32 (the reference writes real_t = float).

The CPU tests run everywhere; with /root/reference present (build container) all 17 meshes are covered and the
fixture is re-derived live, elsewhere the meshes shipped under tests/golden/meshes.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, MESHES, ROOT, mesh_path

sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
import pins as P      # noqa: E402

REF_DATA = "/root/reference/data"
SHIPPED = ["rectangle-tris-boundary", "rectangle-tris", "2blocks", "bolted_bracket", "tet-cube-heat", "mitchell_tri"]


@pytest.fixture(scope="module")
def ref_pins():
    with open(os.path.join(GOLDEN, "ref_pins.json")) as f:
        return json.load(f)


def _all_meshes():
    with open(os.path.join(GOLDEN, "ref_pins.json")) as f:
        names = sorted(k for k in json.load(f) if not k.startswith("_"))
    return [n for n in names if n in SHIPPED or os.path.exists(os.path.join(REF_DATA, n + ".exo"))]


def _path(name):
    p = mesh_path(name)
    return p if os.path.exists(p) else os.path.join(REF_DATA, name + ".exo")


def _same(a: dict, b: dict, keys):
    for k in keys:
        assert a[k] == b[k], (k, a[k] if not isinstance(a[k], dict) else {x: a[k][x] for x in ("n", "sha")},
                              b[k] if not isinstance(b[k], dict) else {x: b[k][x] for x in ("n", "sha")})


CSR_KEYS = ("n", "nrows", "nnz", "trace", "sum", "rows", "rowptr", "cols", "vals")


def _d1_active(mesh) -> bool:
    last = mesh.num_nodes - 1
    return not any(last in set(np.asarray(v).tolist()) for v in mesh.nodesets.values())


# ---- the fixture itself ----------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.exists(os.path.join("/root/reference", "ExodusIO.hpp")), reason="needs /root/reference (build container)")
@pytest.mark.parametrize("name", ["rectangle-tris-boundary", "bolted_bracket", "2blocks"])
def test_fixture_is_what_the_reference_computes_live(ref_pins, name):
    import run_ref as R
    out = R.run_reference(os.path.join(REF_DATA, name + ".exo"), 2)
    assert out["returncode"] == 0, out["stderr"]
    assert P.summ_assemble(out["assemble"]) == ref_pins[name]["assemble"]
    assert P.summ_dump_text(out["dump_text"]) == ref_pins[name]["dump"]
    assert P.summ_getmatrix(out["getmatrix"]) == ref_pins[name]["getmatrix"]
    assert P.summ_output(P.canon_from_shimdump(out["solution"])) == ref_pins[name]["decompose"]["2"]


def test_survey_hand_checked_system_is_the_references(ref_pins):
    """SURVEY.md §8c pin (i), derived there by reading the code — the reference itself agrees"""
    a = ref_pins["rectangle-tris-boundary"]["assemble"]
    A = sp.csr_matrix((a["A"]["vals"]["data"], a["A"]["cols"]["data"], a["A"]["rowptr"]["data"]), shape=(3, 3)).toarray()
    np.testing.assert_array_equal(A, [[5, 0, -1], [0, 4, -1], [-1, -1, 5]])
    assert a["B"]["data"] == [500.0, 450.0, 300.0]
    assert a["idmap_original"]["data"] == [2, 3, 5]
    # Appendix C numbers of the reference as written (one rank, D1 active on tet-cube-heat, inactive on bolted_bracket)
    t, b = ref_pins["tet-cube-heat"]["assemble"], ref_pins["bolted_bracket"]["assemble"]
    assert (t["A"]["n"], t["sum_B"]) == (19248, 2822800.0)
    assert (b["A"]["n"], b["A"]["nnz"], b["A"]["trace"], b["sum_B"]) == (3764, 46222, 43896.0, 8881.0)


# ---- assemble ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", _all_meshes())
def test_numpy_restatement_with_d1_equals_reference(oracle, ref_pins, name):
    mesh = oracle.read_exodus(_path(name))
    A, b, r2o = oracle.assemble_np(mesh, oracle.GRAPH_LAPLACIAN, bug_compat_d1=True)
    want = ref_pins[name]["assemble"]
    _same(P.summ_scipy(A), want["A"], CSR_KEYS)
    assert P._arr(b, "<f8") == want["B"]
    assert P._arr(r2o, "<i8") == want["idmap_original"]
    assert P._arr(np.arange(len(r2o)), "<i8") == want["idmap_reduced"]
    for sid, nodes in mesh.nodesets.items():                       # nodeSetMap cache: 0-based, sorted (std::set)
        assert P._arr(np.unique(nodes), "<i8") == want["nodesets"][str(sid)]


@pytest.mark.parametrize("name", _all_meshes())
def test_c_oracle_fixed_semantics_project_onto_reference(oracle, ref_pins, name):
    mesh = oracle.read_exodus(_path(name))
    s = oracle.assemble(mesh, oracle.GRAPH_LAPLACIAN)
    A, b, r2o = s.csr(), s.b, s.red2orig
    want = ref_pins[name]["assemble"]
    if _d1_active(mesh):
        assert s.n == want["A"]["n"] + 1 and r2o[-1] == mesh.num_nodes - 1      # the node the reference drops
        A, b = P.apply_d1(A, b)
        r2o = r2o[:-1]
    _same(P.summ_scipy(A), want["A"], CSR_KEYS)
    assert P._arr(b, "<f8") == want["B"]
    assert P._arr(r2o, "<i8") == want["idmap_original"]


def _summ_product_view(view):
    """digest of what heat_reference_view_csr returned, in the fixture's terms (rows that exist = rows with entries)"""
    rp, col, val, b, kept, orig = view
    for k in range(len(rp) - 1):
        assert np.all(np.diff(col[rp[k]:rp[k + 1]]) > 0)            # columns ascending, no duplicates
    n = len(rp) - 1
    return P.summ_scipy(sp.csr_matrix((val, col, rp), shape=(n, n))), b, kept, orig


@pytest.mark.parametrize("name", _all_meshes())
def test_product_reference_view_equals_reference(hb, oracle, ref_pins, name):
    """heat_reference_view_csr (csrc/refview.cpp, host code of the product): the FIXED system -> the system the
    reference's own IO::assemble returned, bit for bit, on every mesh (defects D1 / D3 applied inside the product)"""
    mesh = oracle.read_exodus(_path(name))
    s = oracle.assemble(mesh, oracle.GRAPH_LAPLACIAN)
    io = _host_io(hb)
    try:
        io.open(_path(name), True)
        A, b, kept, orig = _summ_product_view(io.reference_view(s.row_ptr, s.col, s.val, s.b, s.red2orig))
        with pytest.raises(hb.HeatError, match="unknowns"):
            io.reference_view(s.row_ptr[:-1], s.col, s.val, s.b[:-1], s.red2orig[:-1])
    finally:
        io.close()
    want = ref_pins[name]["assemble"]
    _same(A, want["A"], CSR_KEYS)
    assert P._arr(b, "<f8") == want["B"]
    assert P._arr(kept, "<i8") == want["idmap_reduced"] and P._arr(orig, "<i8") == want["idmap_original"]


@pytest.mark.parametrize("name", [n for n in _all_meshes() if n != "initialguess"])
def test_c_oracle_get_matrix_equals_reference(oracle, ref_pins, name):
    """IO::getMatrix on one rank: rows and columns are 1-based node ids, nodesets are returned 1-based and are
    NOT applied to the matrix (ExodusIO.hpp:729-731)."""
    mesh = oracle.read_exodus(_path(name))
    s = oracle.get_matrix(mesh, oracle.GRAPH_LAPLACIAN)
    want = ref_pins[name]["getmatrix"]
    _same(P.summ_scipy(s.csr(), row_base=1), want["A"], CSR_KEYS)
    for sid, nodes in mesh.nodesets.items():
        assert P._arr(np.unique(nodes) + 1, "<i8") == want["nodesets"][str(sid)]


@pytest.mark.parametrize("name", [n for n in _all_meshes() if n != "initialguess"])
def test_c_oracle_power_method_equals_reference_loop(oracle, ref_pins, name):
    """The reference's own PowerMethod::run (ExodusMatrixTest.cpp:56-129, compiled from that file) on getMatrix's
    output with the seeded start vector: the oracle must stop at the same iteration with the same verdict, and
    report the same lambda / residual at every 50th iteration up to the summation order of the dots."""
    mesh = oracle.read_exodus(_path(name))
    s = oracle.get_matrix(mesh, oracle.GRAPH_LAPLACIAN)
    want = ref_pins[name]["getmatrix"]["power_method"]
    lam, res, it, conv = oracle.power_method(s, oracle.hash_vector(np.arange(s.n), 12345), want["niters"], want["tolerance"])
    assert (it, conv) == (want["stop_iter"], want["converged"])
    assert lam == pytest.approx(want["lambda"], rel=1e-12)
    assert res == pytest.approx(want["reports"][-1]["residual"], rel=1e-6, abs=1e-12)
    assert [r["iter"] for r in want["reports"]] == [k for k in range(0, want["niters"], 50) if k <= it] + ([want["niters"] - 1] if not conv else [])


def test_reference_rejects_tet4_blocks(ref_pins):
    """`decompose` / `getMatrix` accept only TETRA*, TRI*, HEX* prefixes — initialguess.exo says "TET4"."""
    if "initialguess" not in ref_pins:
        pytest.skip("not pinned")
    e = ref_pins["initialguess"]
    assert "unsupported element type" in e["decompose"]["2"]["failed"][0]
    assert "unsupported element type" in e["getmatrix"]["failed"][0]
    assert e["assemble"]["A"] == ref_pins["bolted_bracket"]["assemble"]["A"]        # assemble takes any clique


# ---- the reference on meshes held in memory: the benchmark's Kuhn cubes and the corners of assemble ------------------
def _syn(ref_pins, prefix):
    return sorted(k for k in ref_pins["_synthetic"] if k.startswith(prefix))


with open(os.path.join(GOLDEN, "ref_pins.json")) as _f:
    _SYN = sorted(json.load(_f)["_synthetic"])
CUBES = [k for k in _SYN if k.startswith("cube:")]
EDGES = [k for k in _SYN if k.startswith("edge:")]


@pytest.mark.parametrize("key", CUBES)
def test_c_oracle_kuhn_cube_equals_reference(oracle, ref_pins, key):
    """The benchmark family: the reference's own assemble on the explicit Kuhn tet cube.  The
    last node lies on the i = nx-1 face (nodeset 100), so D1 is inactive and the reference's system IS the FIXED one."""
    dims = tuple(int(v) for v in key[5:].split("x"))
    s = oracle.assemble(oracle.cube_mesh(*dims), oracle.GRAPH_LAPLACIAN)
    want = ref_pins["_synthetic"][key]["assemble"]
    _same(P.summ_scipy(s.csr()), want["A"], CSR_KEYS)
    assert P._arr(s.b, "<f8") == want["B"]
    assert P._arr(s.red2orig, "<i8") == want["idmap_original"]


def _edge_mesh(oracle, spec):
    x = np.arange(spec["num_nodes"], dtype=np.float64)
    ns = {int(k): np.asarray(v, dtype=np.int64) for k, v in spec["nodesets"].items()}
    conn = np.asarray(spec["conn"], dtype=np.int32)
    return oracle.Mesh(x, 0 * x, 0 * x, conn, ns, "TETRA", 3, [len(conn)])


def _check_edge_against_reference(A, b, r2o, mesh, want):
    """FIXED system (scipy A, b, reduced->original) seen through the reference's defects (pins.reference_view: D3 rows
    never inserted and absent from the id map, D1 drop of the last node) vs the reference's own output"""
    Ar, br, kept, orig = P.reference_view(A, b, r2o, mesh.conn, mesh.num_nodes, mesh.nodesets)
    _same(P.summ_scipy(Ar), want["A"], CSR_KEYS)
    assert P._arr(br, "<f8") == want["B"]                                 # the RHS is complete in the reference too (:671-687)
    assert P._arr(kept, "<i8") == want["idmap_reduced"]
    assert P._arr(orig, "<i8") == want["idmap_original"]


@pytest.mark.parametrize("key", EDGES)
def test_c_oracle_edge_cases_against_reference(oracle, ref_pins, key):
    e = ref_pins["_synthetic"][key]
    m = _edge_mesh(oracle, e["mesh"])
    s = oracle.assemble(m, oracle.GRAPH_LAPLACIAN)
    _check_edge_against_reference(s.csr(), s.b, s.red2orig, m, e["assemble"])


@pytest.mark.parametrize("key", EDGES)
def test_product_reference_view_on_corner_cases(hb, oracle, ref_pins, key):
    e = ref_pins["_synthetic"][key]
    m = _edge_mesh(oracle, e["mesh"])
    s = oracle.assemble(m, oracle.GRAPH_LAPLACIAN)
    io = _host_io(hb)
    try:
        io.mesh_set(m.x, m.y, m.z, m.conn, dict(m.nodesets))
        A, b, kept, orig = _summ_product_view(io.reference_view(s.row_ptr, s.col, s.val, s.b, s.red2orig))
    finally:
        io.close()
    want = e["assemble"]
    _same(A, want["A"], CSR_KEYS)
    assert P._arr(b, "<f8") == want["B"]
    assert P._arr(kept, "<i8") == want["idmap_reduced"] and P._arr(orig, "<i8") == want["idmap_original"]


def test_reference_defects_seen_live(ref_pins):
    """What the reference's writeSolution records show on the corner meshes"""
    syn = ref_pins["_synthetic"]
    # D3: the lone DOF (node 1) is missing from globalIDMap, so its value lands on node 0 — a nodeset node (id 7)
    assert syn["edge:d3_single_dof"]["assemble"]["A"]["nrows"] == 0 and syn["edge:d3_single_dof"]["assemble"]["B"]["data"] == [21.0]
    assert syn["edge:d3_single_dof"]["decompose"]["2"]["steps"][0]["data"] == [0.25, 0.0, 7.0, 7.0, 7.0]
    # a node in no element: row 2 is never inserted, x[2] = 1.25 overwrites node 0
    assert syn["edge:isolated_node"]["decompose"]["2"]["steps"][0]["data"] == [1.25, 0.75, 0.0, 1.75, 2.25, 9.0]
    # D2: the RHS takes the LOWEST id of a node in several nodesets (sum_B = 21), the output field the LARGEST (8 and 5)
    o = syn["edge:overlapping_nodesets"]
    assert o["assemble"]["sum_B"] == 21.0 and o["decompose"]["2"]["steps"][0]["data"] == [8.0, 0.25, 0.75, 1.25, 1.75, 5.0]
    # D7: writeSolution(X, 0) writes step 1 twice, so only two steps exist after writeSolution(X, 0), (X, 1)
    assert o["decompose"]["2"]["times"] == [0.0, 1.0]


# ---- decompose: the product's output file against the reference's Exodus calls --------------------------------------
@pytest.fixture(scope="module")
def hb():
    import heat_b200
    if not os.path.exists(heat_b200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return heat_b200


def _host_io(hb):
    h = C.c_void_p()
    assert hb.lib().heat_ctx_create(-1, C.byref(h)) == 0
    io = hb.IO.__new__(hb.IO)
    io.h = h
    return io


OUT_KEYS = ("title", "num_dim", "num_nodes", "num_elem", "num_el_blk", "num_node_sets", "num_side_sets", "coords", "elem_map",
            "node_num_map", "blocks", "nodesets", "sidesets")


@pytest.mark.parametrize("parts", [2, 4])
@pytest.mark.parametrize("name", [n for n in _all_meshes() if n != "initialguess"])
def test_decompose_file_equals_reference_records(hb, ref_pins, name, parts, tmp_path):
    io = _host_io(hb)
    try:
        io.open(_path(name), True)
        out = str(tmp_path / "solution.exo")
        io.create(out)
        io.decompose(parts)
    finally:
        io.close()
    got = P.summ_output(P.canon_from_exodus(out))
    want = ref_pins[name]["decompose"][str(parts)]
    _same(got, want, OUT_KEYS)


def test_cpp_mirror_decompose_equals_reference_records(hb, ref_pins, tmp_path):
    """include/ExodusIO_b200.hpp — the class a maintainer swaps in for ExodusIO.hpp — compiled with g++ and driven in
    the reference's call order on a host-only context: bool returns, failure messages, the no-CPU-fallback rule,
    and an output file identical (as records) to what the reference's own decompose wrote."""
    import subprocess
    exe = str(tmp_path / "mirror_host_test")
    libdir = os.path.dirname(hb.LIB_PATH)
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_host_test.cpp"),
                    "-o", exe, "-L", libdir, "-lheat_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    for name, parts in (("bolted_bracket", 2), ("rectangle-tris-boundary", 4), ("2blocks", 2)):
        out = str(tmp_path / f"{name}.exo")
        p = subprocess.run([exe, mesh_path(name), out, str(parts)], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0, (p.returncode, p.stderr)
        assert "no CPU fallback" in p.stderr and "ex_open" in p.stderr          # the two expected failures were reported
        _same(P.summ_output(P.canon_from_exodus(out)), ref_pins[name]["decompose"][str(parts)], OUT_KEYS)


@pytest.mark.parametrize("name", [n for n in _all_meshes() if n != "initialguess"])
def test_real4_output_is_the_reference_builds_file(hb, ref_pins, name, tmp_path):
    """heat_ctx_set_output(ctx, 4, ...): the reference creates its output with word size sizeof(real_t) = 4 on the usual
    METIS build (ExodusIO.hpp:104-105).  With that convention the product's file holds float32 records
    that equal the reference's WITHOUT any cast on our side, and it still opens with the product's own reader."""
    from scipy.io import netcdf_file
    io = _host_io(hb)
    out = str(tmp_path / "solution.exo")
    try:
        with pytest.raises(hb.HeatError, match="word size must be 4 or 8"):
            io.set_output(5)
        io.set_output(4)
        io.open(_path(name), True)
        io.create(out)
        io.decompose(2)
        with pytest.raises(hb.HeatError, match="already written with word size 4"):
            io.set_output(8)
    finally:
        io.close()
    nc = netcdf_file(out, "r", mmap=False)
    assert nc.floating_point_word_size == 4
    floats = [v for v in nc.variables.values() if v.data.dtype.kind == "f"]
    assert floats and all(v.data.dtype.itemsize == 4 for v in floats)
    nc.close()
    _same(P.summ_output(P.canon_from_exodus(out)), ref_pins[name]["decompose"]["2"], OUT_KEYS)
    # the float32 file as INPUT of a default (fp64) context: the output word size is the context's, not the input's
    again = _host_io(hb)
    out8 = str(tmp_path / "solution8.exo")
    try:
        again.open(out, True)
        again.create(out8)
        again.decompose(2)
    finally:
        again.close()
    nc = netcdf_file(out8, "r", mmap=False)
    assert nc.floating_point_word_size == 8
    assert all(v.data.dtype.itemsize == 8 for v in nc.variables.values() if v.data.dtype.kind == "f")
    nc.close()
    c4, c8 = P.canon_from_exodus(out), P.canon_from_exodus(out8)
    for a, b in zip(c4["coords"], c8["coords"]):
        np.testing.assert_array_equal(np.asarray(a, dtype=np.float64), b)


def test_cli_heat_decompose_mirrors_the_references_decompose_test(hb, ref_pins, tmp_path):
    """bin/heat_decompose = ExodusIODecomposeTest.cpp: same flags, same messages; its output file against the records
    of the reference's own decompose"""
    import subprocess
    exe = os.path.join(ROOT, "domain-decomposed-pde-solver_b200", "bin", "heat_decompose")
    for args, msg in (([], "No input file was provided; use the '--input' parameter!"),
                      ([f"--input={mesh_path('2blocks')}"], "No output file was provided; use the '--output' parameter!"),
                      ([f"--input={mesh_path('2blocks')}", f"--output={tmp_path / 'x.exo'}"],
                       "Number of partitions to decompose the mesh has not been provided; use the '--partitions' parameter!"),
                      ([f"--input={tmp_path / 'missing.exo'}", f"--output={tmp_path / 'x.exo'}", "--partitions=2"], "Failed to open input Exodus file")):
        p = subprocess.run([exe] + args, capture_output=True, text=True)
        assert p.returncode != 0 and msg in p.stderr, (args, p.stderr)
    for name, parts, extra in (("tet-cube-heat", 4, []), ("mitchell_tri", 2, ["--real4"]), ("2blocks", 4, [])):
        out = str(tmp_path / f"{name}.exo")
        p = subprocess.run([exe, f"--input={mesh_path(name)}", f"--output={out}", f"--partitions={parts}", "--no-verbose"] + extra,
                           capture_output=True, text=True, timeout=120)
        assert p.returncode == 0, p.stderr
        _same(P.summ_output(P.canon_from_exodus(out)), ref_pins[name]["decompose"][str(parts)], OUT_KEYS)


@pytest.mark.parametrize("word_size", [8, 4])
def test_write_nodal_field_equals_reference_records(hb, oracle, ref_pins, tmp_path, word_size):
    """The file half of writeSolution without a GPU: the dense nodal arrays the reference built for the stand-in
    iterates (x[row] = 0.25 + 0.5 row + 4096 step) are rebuilt here from the oracle's id map and written through
    the product; variable name, time values and both records must equal what the reference gave ex_put_*.
    rectangle-tris-boundary has no node in two nodesets and its last node is in a nodeset, so D1/D2 play no part."""
    name = "rectangle-tris-boundary"
    mesh = oracle.read_exodus(mesh_path(name))
    s = oracle.assemble(mesh, oracle.GRAPH_LAPLACIAN)
    io = _host_io(hb)
    out = str(tmp_path / "solution.exo")
    try:
        io.set_output(word_size)
        io.open(mesh_path(name), True)
        io.create(out)
        io.decompose(2)
        for step in (0, 1):
            field = np.where(np.isnan(s.node_bc), 0.0, s.node_bc)
            field[s.red2orig] = 0.25 + 0.5 * np.arange(s.n) + 4096.0 * step
            io.write_nodal_field(field, step)
    finally:
        io.close()
    canon = P.canon_from_exodus(out)
    assert all(np.asarray(a).dtype.itemsize == word_size for a in canon["steps"])
    got = P.summ_output(canon)
    want = ref_pins[name]["decompose"]["2"]
    _same(got, want, ("var_names", "times", "steps"))
    assert want["steps"][1]["data"] == [50.0, 50.0, 4096.25, 4096.75, 50.0, 4097.25, 200.0, 200.0, 200.0]


@pytest.mark.parametrize("name", [n for n in _all_meshes() if n != "initialguess"])
def test_host_scatter_and_write_give_the_references_whole_file(hb, oracle, ref_pins, name, tmp_path):
    """The host-buffer path solution -> nodal field -> file (heat_scatter_nodal_field + heat_write_nodal_field, the code
    heat_write_solution runs after its gather) under the reference's output conventions (float32, largest nodeset id):
    mesh, both result steps, times and variable name equal the reference's records on every mesh.  Where D1 is active
    the reference has no row for the last node and leaves its value 0; that one value is zeroed here, nothing else."""
    mesh = oracle.read_exodus(_path(name))
    n = ref_pins[name]["assemble"]["A"]["n"] + (1 if _d1_active(mesh) else 0)             # FIXED number of unknowns
    io = _host_io(hb)
    out = str(tmp_path / "solution.exo")
    try:
        io.set_output(4, True)
        io.open(_path(name), True)
        io.create(out)
        io.decompose(2)
        with pytest.raises(hb.HeatError, match="unknowns"):
            io.scatter_nodal_field(np.zeros(n + 1), mesh.num_nodes)
        for step in (0, 1):
            field = io.scatter_nodal_field(0.25 + 0.5 * np.arange(n) + 4096.0 * step, mesh.num_nodes)
            if _d1_active(mesh):
                field[-1] = 0.0
            io.write_nodal_field(field, step)
    finally:
        io.close()
    _same(P.summ_output(P.canon_from_exodus(out)), ref_pins[name]["decompose"]["2"], OUT_KEYS + ("var_names", "times", "steps"))


def test_host_scatter_default_convention_uses_the_rhs_id(hb, oracle):
    """default (D2 fixed): a node in several nodesets shows the id its right-hand side used — the lowest"""
    mesh = oracle.read_exodus(mesh_path("bolted_bracket"))
    s = oracle.assemble(mesh, oracle.GRAPH_LAPLACIAN)
    io = _host_io(hb)
    try:
        io.open(mesh_path("bolted_bracket"), True)
        field = io.scatter_nodal_field(np.arange(s.n) + 0.5, mesh.num_nodes)
    finally:
        io.close()
    np.testing.assert_array_equal(field, oracle.scatter_field(s, np.arange(s.n) + 0.5))
    both = np.intersect1d(mesh.nodesets[1], mesh.nodesets[10])
    assert len(both) == 32 and np.all(field[both] == 1.0)


def test_host_scatter_on_cubes_and_memory_meshes(hb, oracle):
    """the other two mesh sources: analytic Kuhn cubes (nodesets 1000 / 100 on the x faces) and heat_mesh_set"""
    io = _host_io(hb)
    try:
        io.mesh_cube(6, 4, 3)
        m = oracle.cube_mesh(6, 4, 3)
        s = oracle.assemble(m, oracle.GRAPH_LAPLACIAN)
        x = np.arange(s.n) + 0.25
        np.testing.assert_array_equal(io.scatter_nodal_field(x, m.num_nodes), oracle.scatter_field(s, x))
        with pytest.raises(hb.HeatError, match="unknowns"):
            io.scatter_nodal_field(x[:-1], m.num_nodes)
        with pytest.raises(hb.HeatError, match="num_nodes"):
            io.scatter_nodal_field(x, m.num_nodes - 1)
        ns = {7: np.array([0, 5]), 3: np.array([5])}
        io.mesh_set(m.x, m.y, m.z, m.conn, ns)
        f = io.scatter_nodal_field(np.arange(m.num_nodes - 2) + 0.5, m.num_nodes)
        assert f[0] == 7.0 and f[5] == 3.0 and f[1] == 0.5 and f[-1] == m.num_nodes - 2.5
        io.set_output(8, True)
        assert io.scatter_nodal_field(np.zeros(m.num_nodes - 2), m.num_nodes)[5] == 7.0        # largest id on request
        io.mesh_set(m.x, m.y, m.z, m.conn, {9: np.array([1])})                                  # same size, other nodesets:
        f = io.scatter_nodal_field(np.zeros(m.num_nodes - 1), m.num_nodes)                     # nothing stale may survive
        assert f[1] == 9.0 and f[0] == 0.0 and f[5] == 0.0
    finally:
        io.close()


# ---- the CUDA path against the reference ---------------------------------------------------------------------------
GPU_MESHES = [n for n in SHIPPED]


@pytest.fixture()
def gpu_io(hb):
    if hb.device_count() < 1:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    h = hb.IO(0)
    yield h
    h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", GPU_MESHES)
def test_gpu_assemble_equals_reference(hb, gpu_io, oracle, ref_pins, name):
    """heat_assemble (device pattern + values + RHS, graph-Laplacian mode) reproduces the reference's A, B and id
    map bit for bit — directly when D1 is inactive, through apply_d1 otherwise."""
    mesh = oracle.read_exodus(mesh_path(name))
    gpu_io.open(mesh_path(name), True)
    A, X, B = gpu_io.assemble(hb.OP_GRAPH_LAPLACIAN, hb.PART_METIS_KWAY)
    rp, col, val = A.csr()
    n = len(rp) - 1
    M, b, r2o = sp.csr_matrix((val, col, rp), shape=(n, n)), B.numpy(), A.red2orig()
    want = ref_pins[name]["assemble"]
    # (1) the product end to end: device assembly -> heat_reference_view_csr (host) = the reference's output
    Av, bv, kept, orig = _summ_product_view(gpu_io.reference_view(rp, col, val, b, r2o))
    _same(Av, want["A"], CSR_KEYS)
    assert P._arr(bv, "<f8") == want["B"]
    assert P._arr(kept, "<i8") == want["idmap_reduced"] and P._arr(orig, "<i8") == want["idmap_original"]
    # (2) the same through the test-side transformation
    if _d1_active(mesh):
        M, b = P.apply_d1(M, b)
        r2o = r2o[:-1]
    _same(P.summ_scipy(M), want["A"], CSR_KEYS)
    assert P._arr(b, "<f8") == want["B"]
    assert P._arr(r2o, "<i8") == want["idmap_original"]


@pytest.mark.gpu
@pytest.mark.parametrize("explicit", [False, True])
@pytest.mark.parametrize("key", CUBES)
def test_gpu_kuhn_cube_equals_reference(hb, gpu_io, ref_pins, key, explicit):
    """The synthetic cubes of the benchmark: `cube_fill_kernel` (analytic connectivity, 24 incident tets recomputed
    per row) and the explicit-mesh kernels, graph-Laplacian mode, against the reference's own assemble."""
    dims = tuple(int(v) for v in key[5:].split("x"))
    gpu_io.mesh_cube(*dims, explicit_mesh=explicit)
    A, X, B = gpu_io.assemble(hb.OP_GRAPH_LAPLACIAN)
    rp, col, val = A.csr()
    n = len(rp) - 1
    want = ref_pins["_synthetic"][key]["assemble"]
    _same(P.summ_scipy(sp.csr_matrix((val, col, rp), shape=(n, n))), want["A"], CSR_KEYS)
    assert P._arr(B.numpy(), "<f8") == want["B"]
    assert P._arr(A.red2orig(), "<i8") == want["idmap_original"]


@pytest.mark.gpu
@pytest.mark.parametrize("key", EDGES)
def test_gpu_edge_cases_against_reference(hb, gpu_io, oracle, ref_pins, key):
    e = ref_pins["_synthetic"][key]
    m = _edge_mesh(oracle, e["mesh"])
    gpu_io.mesh_set(m.x, m.y, m.z, m.conn, {k: v for k, v in m.nodesets.items()}, num_dim=3)
    A, X, B = gpu_io.assemble(hb.OP_GRAPH_LAPLACIAN, hb.PART_METIS_KWAY)
    rp, col, val = A.csr()
    n = len(rp) - 1
    _check_edge_against_reference(sp.csr_matrix((val, col, rp), shape=(n, n)), B.numpy(), A.red2orig(), m, e["assemble"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bolted_bracket", "mitchell_tri", "rectangle-tris", "2blocks"])
def test_gpu_get_matrix_equals_reference(hb, gpu_io, ref_pins, name):
    gpu_io.open(mesh_path(name), True)
    A = gpu_io.getMatrix(hb.OP_GRAPH_LAPLACIAN)
    rp, col, val = A.csr()
    n = len(rp) - 1
    _same(P.summ_scipy(sp.csr_matrix((val, col, rp), shape=(n, n)), row_base=1), ref_pins[name]["getmatrix"]["A"], CSR_KEYS)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bolted_bracket", "mitchell_tri", "tet-cube-heat", "2blocks"])
def test_gpu_power_method_equals_reference_loop(hb, gpu_io, ref_pins, name):
    """heat_power_method (SpMV with fused q.z and z.z, 2 launches per iteration) against the reference's own loop"""
    gpu_io.open(mesh_path(name), True)
    A = gpu_io.getMatrix(hb.OP_GRAPH_LAPLACIAN)
    want = ref_pins[name]["getmatrix"]["power_method"]
    pr = gpu_io.power_method(A, want["niters"], want["tolerance"], 12345)
    assert (pr.iters, pr.converged) == (want["stop_iter"], want["converged"])
    assert pr.lambda_ == pytest.approx(want["lambda"], rel=1e-10)
    assert len(pr.reports) == len(want["reports"])
    for mine, ref in zip(pr.reports, want["reports"]):
        assert int(mine[0]) == ref["iter"]
        assert mine[1] == pytest.approx(ref["lambda"], rel=1e-10)
        assert mine[2] == pytest.approx(ref["residual"], rel=1e-5, abs=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bolted_bracket", "rectangle-tris-boundary"])
def test_gpu_reference_output_conventions_give_the_references_file(hb, gpu_io, ref_pins, name, tmp_path):
    """set_output(4, largest_nodeset_id=True) + the reference's call order: where D1 is inactive (last node in a
    nodeset) the WHOLE output file — mesh, float32 records, both result steps, the 32 nodes of bolted_bracket that sit
    in two nodesets included — equals what the reference handed to the Exodus API, with no exception list."""
    out = str(tmp_path / "solution.exo")
    gpu_io.set_output(4, True)
    gpu_io.open(mesh_path(name), True)
    A, X, B = gpu_io.assemble(hb.OP_GRAPH_LAPLACIAN, hb.PART_METIS_KWAY)
    gpu_io.create(out)
    gpu_io.decompose(2)
    for step in (0, 1):
        X.set(0.25 + 0.5 * np.arange(len(X)) + 4096.0 * step)
        gpu_io.writeSolution(X, step)
    gpu_io.close()
    _same(P.summ_output(P.canon_from_exodus(out)), ref_pins[name]["decompose"]["2"], OUT_KEYS + ("var_names", "times", "steps"))


@pytest.mark.gpu
def test_gpu_cli_heat_assemble_test(hb):
    """bin/heat_assemble_test = ExodusAssembleTest.cpp (open -> assemble, nothing else)"""
    import subprocess
    exe = os.path.join(ROOT, "domain-decomposed-pde-solver_b200", "bin", "heat_assemble_test")
    p = subprocess.run([exe, f"--input={mesh_path('bolted_bracket')}", "--verbose"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "DOF rows: 3764" in p.stdout and "nnz: 46222" in p.stdout, (p.stdout, p.stderr)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode != 0 and "No input file was provided; use the '--input' parameter!" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rectangle-tris-boundary", "bolted_bracket"])
def test_gpu_cli_dump_equals_reference_text(hb, ref_pins, name, tmp_path):
    """`heat_solver --dump` (heat::printCrsMatrix / printMultiVector of include/ExodusIO_b200.hpp) against the text
    the reference's own printCrsMatrix / printMultiVector wrote for its A and B (BelosMueLuSolver.cpp:37-84, compiled
    from that file): identical line for line once the ~usec~ stamps are removed.  Both meshes have their last node
    in a nodeset, so D1 is inactive and the reference's system is the FIXED one."""
    import re
    import subprocess
    exe = os.path.join(ROOT, "domain-decomposed-pde-solver_b200", "bin", "heat_solver")
    prefix = str(tmp_path / "mpi-proc-")
    p = subprocess.run([exe, f"--input={mesh_path(name)}", f"--solution={tmp_path / 's.exo'}", f"--outputPrefix={prefix}", "--dump",
                        "--iterations=5"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    text = open(prefix + "0.out").read()
    mine = text[:text.index("[Solution: X]")]                    # the reference run has no Belos, hence no X section
    assert all(re.search(r" ~[0-9]+~$", l) for l in mine.splitlines() if not l.startswith("["))
    assert P.summ_dump_text(mine) == ref_pins[name]["dump"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rectangle-tris-boundary", "bolted_bracket", "tet-cube-heat", "mitchell_tri"])
def test_gpu_write_solution_equals_reference(hb, gpu_io, oracle, ref_pins, name, tmp_path):
    """open -> assemble -> create -> decompose -> writeSolution(X, 0), writeSolution(X, 1) in the reference's call
    order with the same stand-in iterates; the nodal records in the product's file equal the reference's (as
    float32) except where the documented defects act: the node D1 drops keeps its 0.0 in the reference, and a
    node in several nodesets shows the LARGEST id there (D2) but the id the RHS used (the lowest) here."""
    mesh = oracle.read_exodus(mesh_path(name))
    out = str(tmp_path / "solution.exo")
    gpu_io.open(mesh_path(name), True)
    A, X, B = gpu_io.assemble(hb.OP_GRAPH_LAPLACIAN, hb.PART_METIS_KWAY)
    gpu_io.create(out)
    gpu_io.decompose(2)
    n = len(X)
    for step in (0, 1):
        X.set(0.25 + 0.5 * np.arange(n) + 4096.0 * step)
        gpu_io.writeSolution(X, step)
    r2o = A.red2orig()
    gpu_io.close()
    got = P.canon_from_exodus(out)
    want = ref_pins[name]["decompose"]["2"]
    assert got["var_names"] == want["var_names"] and [float(np.float32(t)) for t in got["times"]] == want["times"]
    member = np.zeros(mesh.num_nodes, dtype=np.int32)
    for nodes in mesh.nodesets.values():
        member[np.unique(nodes)] += 1
    skip = member > 1                                                # D2
    if _d1_active(mesh):
        skip[mesh.num_nodes - 1] = True                              # D1
    if not skip.any():
        _same(P.summ_output(got), want, ("steps",))
    else:
        # the fixture only holds digests for meshes of this size: rebuild the reference's record from its own
        # pinned id map semantics (DOF nodes from x through the id map, nodeset nodes = largest id, dropped node 0)
        for step in (0, 1):
            ref_field = np.zeros(mesh.num_nodes)
            for sid in sorted(mesh.nodesets):                        # ascending std::map: the largest id is written last
                ref_field[np.asarray(mesh.nodesets[sid], dtype=np.int64)] = sid
            nref = ref_pins[name]["assemble"]["A"]["n"]
            ref_field[r2o[:nref]] = 0.25 + 0.5 * np.arange(nref) + 4096.0 * step
            assert P._arr(ref_field, "<f4") == want["steps"][step]   # this reconstruction IS the reference's record
            mine = np.asarray(got["steps"][step])
            np.testing.assert_array_equal(mine[~skip].astype(np.float32), ref_field[~skip].astype(np.float32))
            low = np.full(mesh.num_nodes, np.nan)
            for sid in sorted(mesh.nodesets, reverse=True):
                low[np.asarray(mesh.nodesets[sid], dtype=np.int64)] = sid
            two = member > 1
            np.testing.assert_array_equal(mine[two], low[two])       # FIXED: the lowest id, as in the RHS
