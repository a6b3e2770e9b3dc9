#!/usr/bin/env python
"""Timeline of the peer-memory CG iteration (debug build: `make -C domain-decomposed-pde-solver_b200
EXTRA=-DHEAT_PEER_TRACE`).  Under torchrun, one rank per GPU:

    HEAT_PEER_TRACE_FILE=gpurun_out/trace_ python -m torch.distributed.run --nproc-per-node N tools/peer_trace.py

Every rank writes <prefix><rank>.txt: one line per iteration with, for SpMV / update_xr / update_p,
{first block in, last block through its peer wait, last block out} in %globaltimer ns.
`python tools/peer_trace.py --analyse gpurun_out/trace_ N` prints where an iteration's time goes.
("last block past its wait" is the LATEST block to get past its peer wait: the vector kernels run their 1184 blocks
in two resident waves, so for them that time is essentially the start of the second wave, not a wait for a peer.)
The NVLink data counters of `nvidia-smi nvlink -gt d` are sampled around the timed solve where the driver exposes them
(on the pods used here every link reads "N/A")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "domain-decomposed-pde-solver_b200"))


def analyse(prefix, n):
    for r in range(n):
        t = np.loadtxt(f"{prefix}{r}.txt", dtype=np.float64).reshape(-1, 3, 3)[20:]      # skip the first iterations
        it = np.diff(t[:, 0, 0]).mean() / 1e3
        out = [f"rank {r}: iteration {it:7.1f} us |"]
        names = ["spmv", "xr", "p"]
        for k in range(3):
            dur = (t[:, k, 2] - t[:, k, 0]).mean() / 1e3
            wait = (t[:, k, 1] - t[:, k, 0]).mean() / 1e3 if t[:, k, 1].any() else float("nan")
            nxt = (t[:, (k + 1) % 3, 0][(1 if k == 2 else 0):] - t[:, k, 2][: (-1 if k == 2 else None)]).mean() / 1e3
            out.append(f" {names[k]}: {dur:6.1f} us (last block past its wait after {wait:5.1f}), gap to next kernel {nxt:5.1f} |")
        print("".join(out))


def nvlink_kib(gpu):
    """(tx KiB, rx KiB) summed over the links of `gpu` from the driver's NVLink data counters (`nvidia-smi nvlink -gt d`);
    None where the counters are not available."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=20).stdout
    except (OSError, subprocess.TimeoutExpired):
        return None
    tx = [int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
    rx = [int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
    return (sum(tx), sum(rx)) if tx and rx else None


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--analyse":
        return analyse(sys.argv[2], int(sys.argv[3]))
    import torch
    import torch.distributed as dist
    import heat_b200 as hb
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    io = hb.IO(local, torch.cuda.current_stream())
    idt = torch.zeros(hb.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(hb.IO.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    io.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
    nx = int(os.environ.get("NX", "512"))
    io.mesh_cube(nx, nx, nx, False)
    A, X, B = io.assemble(hb.OP_P1_FEM, hb.PART_SLAB)
    X.fill(0.0)
    r = io.cg_iterations(A, X, B, 300, check_every=300)           # warm-up + peer set-up
    torch.cuda.synchronize()
    dist.barrier()
    c0 = nvlink_kib(local)
    X.fill(0.0)
    r = io.cg_iterations(A, X, B, 300, check_every=300)
    torch.cuda.synchronize()
    dist.barrier()
    c1 = nvlink_kib(local)
    mi = A.info
    halo = 8 * mi.n_ghost                                          # bytes this rank RECEIVES per SpMV = its ghost entries
    nv = None
    if c0 and c1:
        nv = {"tx_bytes_per_iteration": (c1[0] - c0[0]) * 1024 / 300, "rx_bytes_per_iteration": (c1[1] - c0[1]) * 1024 / 300}
    print(f"rank {rank}: peer path {mi.peer_path}, {r.solve_ms / 300 * 1e3:.1f} us per iteration, ghosts {mi.n_ghost} "
          f"(algorithmic halo in: {halo} B per iteration + {world} x 32 B of reduction stamps x 2), NVLink counters: {nv}", flush=True)
    io.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
