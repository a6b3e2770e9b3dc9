"""world_size-2 (gloo, CPU) test of the N>1 HOST logic: the owned/ghost/send plan the C ABI builds
(heat_plan_build) drives a halo exchange + distributed Jacobi-PCG written in numpy over
torch.distributed.  No product compute runs here (there is no CPU fallback) — the point is that
the plan is symmetric (every send has its matching recv, in the receiver's ghost order), that the
local numbering convention (owned first, ghosts grouped by owner) reproduces the global operator,
and that one all-reduce of {r.z, w.u, r.r} per iteration (the single-reduce recurrence the CUDA
path uses) converges to the serial oracle's solution in the same number of iterations (+-2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, mesh_path


GETMATRIX = 2
CHEBYSHEV = 3      # METIS partition, the fused Chebyshev(3)-PCG recurrence of cg.cu (numpy twin)


def _worker(rank, world, port, name, partitioner, q):
    try:
        for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "domain-decomposed-pde-solver_b200")):
            sys.path.insert(0, p)
        import heat_b200 as hb
        import oracle as O
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        mesh = O.read_exodus(mesh_path(name))
        if partitioner == GETMATRIX:        # IO::getMatrix: whole-mesh Laplacian, rows by element partition + ownership rule
            ref = O.get_matrix(mesh)
            _, epart, _ = O.metis_part_mesh_dual(mesh.conn, mesh.num_nodes, {4: 3, 3: 2}[mesh.conn.shape[1]], world)
            part = hb.node_owners(mesh.conn, mesh.num_nodes, epart, world)
        else:
            ref = O.assemble(mesh, O.GRAPH_LAPLACIAN)
            part = hb.partition_rows(ref.row_ptr, ref.col, hb.PART_METIS_KWAY if partitioner == CHEBYSHEV else partitioner, world)
        pl = hb.plan_build(ref.row_ptr, ref.col, part, world, rank)
        owned, ghost = pl["owned"], pl["ghost"]
        g2l = {int(g): i for i, g in enumerate(np.concatenate([owned, ghost]))}
        A = ref.csr()[owned]                                    # local rows, global columns
        lcol = np.array([g2l[int(c)] for c in A.indices], dtype=np.int64)
        n_own = len(owned)
        send_local = np.array([g2l[int(g)] for g in pl["send_gids"]], dtype=np.int64)
        assert np.all(send_local < n_own)

        def halo(x):                                            # x: [owned | ghosts]
            reqs, bufs = [], []
            for s, nb in enumerate(pl["nbr"]):
                sb = torch.from_numpy(x[send_local[pl["send_ptr"][s]:pl["send_ptr"][s + 1]]].copy())
                rb = torch.empty(int(pl["recv_ptr"][s + 1] - pl["recv_ptr"][s]), dtype=torch.float64)
                if sb.numel():
                    reqs.append(dist.isend(sb, int(nb)))
                if rb.numel():
                    reqs.append(dist.irecv(rb, int(nb)))
                bufs.append((s, rb))
            for r in reqs:
                r.wait()
            for s, rb in bufs:
                x[n_own + pl["recv_ptr"][s]: n_own + pl["recv_ptr"][s + 1]] = rb.numpy()

        def spmv(x):
            halo(x)
            prod = A.data * x[lcol]
            return np.add.reduceat(prod, A.indptr[:-1])

        def allsum(*v):
            t = torch.tensor(v, dtype=torch.float64)
            dist.all_reduce(t)
            return t.numpy()

        # halo + SpMV reproduce the global operator
        xg = np.random.default_rng(7).uniform(-1, 1, ref.n)
        x = np.zeros(n_own + len(ghost)); x[:n_own] = xg[owned]
        y = spmv(x)
        np.testing.assert_allclose(y, (ref.csr() @ xg)[owned], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(x[n_own:], xg[ghost])     # ghosts landed in ghost order
        if partitioner == GETMATRIX:
            # distributed power method (ExodusMatrixTest.cpp:56-129), 30 iterations, vs the serial oracle
            z0 = O.hash_vector(np.arange(ref.n), 12345)
            z = z0[owned].copy()
            qv = np.zeros(n_own + len(ghost))
            lam = 0.0
            for _ in range(30):
                normz = np.sqrt(allsum(z @ z)[0])
                qv[:n_own] = z / normz
                z = spmv(qv)
                lam = allsum(qv[:n_own] @ z)[0]
            lam_ref, _, it_ref, _ = O.power_method(ref, z0, 30, 0.0)
            q.put((rank, 30, it_ref, abs(lam - lam_ref) / abs(lam_ref), None))
            dist.destroy_process_group()
            return
        if partitioner == CHEBYSHEV:
            # The iteration the CUDA path runs for Chebyshev(k)-PCG (cg.cu: cheb_xr_first / cheb_step / update_p(z)), kernel
            # by kernel, with the halo of every SpMV input delivered before the SpMV and TWO reductions per iteration:
            #   SpMV(p) + p.Ap | xr_first: x += a p, r' = r - a Ap, W = D^-1 r'/theta, Z0 = W | (SpMV(Z_{j-1}); step: W' = c1 W +
            #   c2 D^-1 (r' - A Z_{j-1}), Z_j = Z_{j-1} + W') x (k-1), the last one with r'.Z, r'.r' | p' = Z + beta p
            k, lmax, ratio = 3, 2.2, 30.0
            ca, cb = lmax / ratio, 1.1 * lmax
            delta, theta = 2.0 / (cb - ca), 0.5 * (cb + ca)
            s1 = theta * delta
            dinv = 1.0 / ref.csr().diagonal()[owned]
            b = ref.b[owned]

            def cheb(rv):
                w = dinv * rv / theta
                z = np.zeros(n_own + len(ghost)); z[:n_own] = w
                rho = 1.0 / s1
                for _ in range(1, k):
                    rho_new = 1.0 / (2.0 * s1 - rho)
                    az = spmv(z)                                   # halo of Z, then A Z
                    w = rho_new * rho * w + 2.0 * rho_new * delta * (dinv * (rv - az))
                    z = np.concatenate([z[:n_own] + w, z[n_own:]])  # out of place, like the ping-ponged buffers
                    rho = rho_new
                return z[:n_own]

            X = np.zeros(n_own); r = b.copy()
            z = cheb(r)
            p = np.zeros(n_own + len(ghost)); p[:n_own] = z
            rz, rr = allsum(r @ z, r @ r)
            rr0, it = rr, 0
            while rr > (1e-10 ** 2) * rr0 and it < 2000:
                ap = spmv(p)
                pap = allsum(p[:n_own] @ ap)[0]
                alpha = rz / pap
                X += alpha * p[:n_own]; r = r - alpha * ap          # r out of place
                z = cheb(r)
                rz_new, rr = allsum(r @ z, r @ r)
                p = np.concatenate([z + (rz_new / rz) * p[:n_own], p[n_own:]])
                rz = rz_new
                it += 1
            x_ref, it_ref, _, _ = O.pcg(ref, tol=1e-10, prec=O.PREC_CHEBYSHEV, cheb_degree=k, cheb_lambda_max=lmax, max_iters=2000)
            err = np.abs(X - x_ref[owned]).max() / np.abs(x_ref).max()
            q.put((rank, it, it_ref, float(err), None))
            dist.destroy_process_group()
            return
        # Chronopoulos-Gear PCG, one all-reduce per iteration
        dinv = 1.0 / ref.csr().diagonal()[owned]
        b = ref.b[owned]
        X = np.zeros(n_own); r = b.copy()
        u = np.zeros(n_own + len(ghost)); u[:n_own] = dinv * r
        w = spmv(u)
        gam, dlt, rr = allsum(r @ u[:n_own], w @ u[:n_own], r @ r)
        rr0, p, s = rr, np.zeros(n_own), np.zeros(n_own)
        gam_old = alpha_old = None
        it = 0
        while rr > (1e-10 ** 2) * rr0 and it < 2000:
            beta = 0.0 if it == 0 else gam / gam_old
            alpha = gam / (dlt if it == 0 else dlt - beta * gam / alpha_old)
            p = u[:n_own] + beta * p; s = w + beta * s
            X += alpha * p; r -= alpha * s
            u[:n_own] = dinv * r
            w = spmv(u)
            gam_old, alpha_old = gam, alpha
            gam, dlt, rr = allsum(r @ u[:n_own], w @ u[:n_own], r @ r)
            it += 1
        x_ref, it_ref, _, _ = O.pcg(ref, tol=1e-10, max_iters=2000)
        err = np.abs(X - x_ref[owned]).max() / np.abs(x_ref).max()
        q.put((rank, it, it_ref, float(err), None))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, -1, -1, 1.0, traceback.format_exc()))


@pytest.mark.parametrize("name,partitioner", [("bolted_bracket", 1), ("bolted_bracket", 0), ("mitchell_tri", 1),
                                              ("bolted_bracket", GETMATRIX), ("bolted_bracket", CHEBYSHEV)])
def test_plan_drives_distributed_pcg_gloo(name, partitioner):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29640 + partitioner + (7 if name == "mitchell_tri" else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, partitioner, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, it, it_ref, err, tb in res:
        assert tb is None, tb
        assert abs(it - it_ref) <= 2, (rank, it, it_ref)
        assert err <= 1e-8, (rank, err)
