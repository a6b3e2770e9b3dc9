"""Property tests (hypothesis) of the host-side multi-GPU logic of libheat_b200 — the owned / ghost / send plan
(Tpetra column-map and Import conventions), the row partitioners and the getMatrix node-ownership
rule (ExodusIO.hpp:1191-1295) — on random inputs, including the shapes the mesh-based tests never produce: ranks
that own nothing, isolated rows, non-symmetric patterns, more ranks than rows.  CPU only."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from test_host_cpu import plan_numpy

SETTINGS = dict(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@pytest.fixture(scope="module")
def hb():
    import heat_b200
    if not os.path.exists(heat_b200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return heat_b200


@st.composite
def pattern_and_partition(draw, symmetric=None):
    n = draw(st.integers(1, 40))
    P = draw(st.integers(1, 6))
    density = draw(st.floats(0.0, 0.4))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    M = rng.random((n, n)) < density
    if symmetric if symmetric is not None else draw(st.booleans()):
        M = M | M.T
    if draw(st.booleans()):
        np.fill_diagonal(M, True)
    rows, cols = np.nonzero(M)
    row_ptr = np.r_[0, np.cumsum(np.bincount(rows, minlength=n))].astype(np.int64)
    # random owner per row; with P > 1 some ranks are often left with no row at all
    part = rng.integers(0, P, size=n).astype(np.int32)
    if draw(st.booleans()) and P > 1:
        part[part == P - 1] = 0                                   # force an empty last rank
    return row_ptr, cols.astype(np.int32), part, P


@settings(**SETTINGS)
@given(pattern_and_partition())
def test_plan_matches_numpy_restatement_on_random_patterns(hb, case):
    row_ptr, col, part, P = case
    n = len(row_ptr) - 1
    owned_all, total_send, total_ghost = [], 0, 0
    plans = [hb.plan_build(row_ptr, col, part, P, r) for r in range(P)]
    for r, got in enumerate(plans):
        exp = plan_numpy(row_ptr, col, part, P, r)
        for k in exp:
            np.testing.assert_array_equal(got[k], exp[k], err_msg=f"{k} rank {r}")
        owned_all.append(got["owned"])
        total_send += len(got["send_gids"]); total_ghost += len(got["ghost"])
        # conventions: owned ascending; ghosts grouped by owner ascending, ascending id within an owner; never own a ghost
        assert np.all(np.diff(got["owned"]) > 0)
        assert np.all(part[got["ghost"]] != r)
        key = got["ghost_owner"].astype(np.int64) * (n + 1) + got["ghost"]
        assert np.all(np.diff(key) > 0)
        assert got["recv_ptr"][-1] == len(got["ghost"]) and got["send_ptr"][-1] == len(got["send_gids"])
    assert np.array_equal(np.sort(np.concatenate(owned_all)), np.arange(n))      # every row has exactly one owner
    assert total_send == total_ghost                                              # what is sent is what is received
    # pairwise: what r sends to q is exactly q's ghosts owned by r, in q's ghost order
    for r, pr in enumerate(plans):
        for j, q in enumerate(pr["nbr"]):
            sent = pr["send_gids"][pr["send_ptr"][j]:pr["send_ptr"][j + 1]]
            pq = plans[int(q)]
            np.testing.assert_array_equal(sent, pq["ghost"][pq["ghost_owner"] == r])
            assert r in pq["nbr"].tolist()                                        # neighbourhood is mutual


@settings(**SETTINGS)
@given(st.integers(0, 50), st.integers(1, 9))
def test_contiguous_partition_is_tpetras_uniform_map(hb, n, P):
    """Tpetra::Map(n, 0, comm) (ExodusIO.hpp:252): ascending blocks, the first n % P ranks hold one row more"""
    row_ptr = np.arange(n + 1, dtype=np.int64)
    col = np.arange(n, dtype=np.int32)
    part = hb.partition_rows(row_ptr, col, hb.PART_CONTIGUOUS, P)
    cnt = np.bincount(part, minlength=P) if n else np.zeros(P, dtype=int)
    assert len(part) == n and (n == 0 or np.all(np.diff(part) >= 0))
    assert cnt.tolist() == [n // P + (1 if r < n % P else 0) for r in range(P)]


@settings(**SETTINGS)
@given(pattern_and_partition(symmetric=True))
def test_metis_partition_is_valid_and_deterministic(hb, case):
    row_ptr, col, _, P = case
    n = len(row_ptr) - 1
    a = hb.partition_rows(row_ptr, col, hb.PART_METIS_KWAY, P)
    b = hb.partition_rows(row_ptr, col, hb.PART_METIS_KWAY, P)
    assert len(a) == n and a.min() >= 0 and a.max() < P
    np.testing.assert_array_equal(a, b)                                           # default seed: run-to-run deterministic
    if P == 1:
        assert not a.any()


@st.composite
def mesh_and_epart(draw):
    N = draw(st.integers(4, 30))
    npe = draw(st.sampled_from([3, 4]))
    ne = draw(st.integers(1, 40))
    P = draw(st.integers(1, 5))
    rng = np.random.default_rng(draw(st.integers(0, 2 ** 31 - 1)))
    conn = np.stack([rng.choice(N, size=npe, replace=False) for _ in range(ne)]).astype(np.int32)
    epart = rng.integers(0, P, size=ne)
    return conn, N, epart, P


@settings(**SETTINGS)
@given(mesh_and_epart())
def test_node_ownership_rule_on_random_meshes(hb, oracle, case):
    """highest number of incident (rank-local) neighbour rows wins, ties go to the lowest rank; every used node gets
    exactly one owner among the ranks that touch it; unused nodes go to rank 0"""
    conn, N, epart, P = case
    got = hb.node_owners(conn, N, epart, P)
    exp = oracle.get_matrix_owners(conn, epart, P, N)
    np.testing.assert_array_equal(got, exp)
    used = np.zeros(N, dtype=bool)
    used[np.unique(conn)] = True
    assert not got[~used].any()
    for v in np.flatnonzero(used):
        touching = {int(epart[e]) for e in np.flatnonzero((conn == v).any(1))}
        assert int(got[v]) in touching
