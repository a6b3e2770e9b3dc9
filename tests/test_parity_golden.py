"""The committed multi-rank parity fixtures (tests/golden/parity_multi.*, made on the oracle side by
make_parity_golden.py) against the product's HOST plan logic: heat_partition_rows + heat_plan_build must reproduce
every rank's owned / ghost / send-map digest for 1, 2, 4 and 8 ranks.  (The same digests are what bench.py's parity
block demands of the maps the GPUs actually run with.)"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, mesh_path


@pytest.fixture(scope="module")
def fixtures():
    with open(os.path.join(GOLDEN, "parity_multi.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLDEN, "parity_multi.npz"))


def _slab_part(red2orig, nx, ny, nz, P):
    k = red2orig // (nx * ny)
    base, rem = divmod(nz, P)
    bounds = np.array([q * base + min(q, rem) for q in range(P + 1)])
    return (np.searchsorted(bounds, k, side="right") - 1).astype(np.int32)


@pytest.mark.parametrize("case", ["bolted_bracket_graph", "cube_p1", "cube_graph"])
def test_host_plan_reproduces_golden_map_digests(oracle, fixtures, case):
    import heat_b200 as hb
    meta, arrays = fixtures
    ent = meta["cases"][case]
    if case.startswith("cube"):
        ref = oracle.cube_assemble(*meta["cube"], oracle.P1_FEM if case == "cube_p1" else oracle.GRAPH_LAPLACIAN)
    else:
        ref = oracle.assemble(oracle.read_exodus(mesh_path("bolted_bracket")), oracle.GRAPH_LAPLACIAN)
    assert (ref.n, ref.nnz) == (ent["n"], ent["nnz"])
    for P in (1, 2, 4, 8):
        if ent["partition"] == "metis":
            part = hb.partition_rows(ref.row_ptr, ref.col, hb.PART_METIS_KWAY, P)
        else:
            part = _slab_part(ref.red2orig, *meta["cube"], P)
        assert hb.maps_digest(part, [], [], [], [], [], []).split()[0]      # smoke: digest of the partition itself
        for r in range(P):
            pl = hb.plan_build(ref.row_ptr, ref.col, part, P, r)
            d = hb.maps_digest(pl["owned"], pl["ghost"], pl["ghost_owner"], pl["nbr"], pl["send_ptr"], pl["send_gids"], pl["recv_ptr"])
            assert d == ent["maps"][str(P)][r], (case, P, r)


def test_golden_solutions_are_the_oracles(oracle, fixtures):
    """the stored N=1 solutions are what the oracle computes today (guards a stale fixture)"""
    meta, arrays = fixtures
    ref = oracle.assemble(oracle.read_exodus(mesh_path("bolted_bracket")), oracle.GRAPH_LAPLACIAN)
    x, it, *_ = oracle.pcg(ref, tol=meta["tol"], max_iters=5000)
    assert it == meta["cases"]["bolted_bracket_graph"]["iters"]
    # (OpenMP reductions of the oracle's dot products are not order-stable, so not bit-equal from run to run)
    assert np.abs(x - arrays["bolted_bracket_graph_x"]).max() <= 1e-11 * np.abs(x).max()
    refc = oracle.cube_assemble(*meta["cube"], oracle.P1_FEM)
    xc, itc, *_ = oracle.pcg(refc, tol=meta["tol"], max_iters=5000)
    assert itc == meta["cases"]["cube_p1"]["iters"] and np.abs(xc - arrays["cube_p1_x"]).max() <= 1e-11 * np.abs(xc).max()
    nx = meta["cube"][0]
    i = refc.red2orig % nx
    assert np.abs(xc - (1000.0 - 900.0 * i / (nx - 1))).max() <= 1e-6        # P1 on a Kuhn cube: exactly linear
