cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2h_topo_${N}gpu.txt 2>&1; nproc; grep MemAvailable /proc/meminfo
(timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2h_bench_${N}gpu.json 2> gpurun_out/r2h_bench_${N}gpu.err; echo "bench rc=$?"); tail -c 300 gpurun_out/r2h_bench_${N}gpu.err
(HEAT_REQUIRE_PEER=1 timeout 1500 $TR --master-port 29511 tests/mgpu_worker.py > gpurun_out/r2h_mgpu_worker_${N}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2h_mgpu_worker_${N}gpu.log); grep -v "^\[W\|^$" gpurun_out/r2h_mgpu_worker_${N}gpu.log | tail -6
(HEAT_B200_LIB=$GRAFT_REPO_ROOT/domain-decomposed-pde-solver_b200/lib_trace/libheat_b200.so HEAT_PEER_TRACE_FILE=gpurun_out/r2h_trace${N}_ timeout 400 $TR --master-port 29513 tools/peer_trace.py > gpurun_out/r2h_peer_trace_${N}gpu.log 2>&1; echo "trace rc=$?"); python tools/peer_trace.py --analyse gpurun_out/r2h_trace${N}_ $N >> gpurun_out/r2h_peer_trace_${N}gpu.log 2>&1; grep -v "^\[W\|^$\|^\*\|OMP" gpurun_out/r2h_peer_trace_${N}gpu.log | tail -20
(timeout 400 $TR --master-port 29514 bench.py --gpus $N --workload weak --prec chebyshev --cheb-lambda-max 2.0 --steps 2 --warmup 1 --iters-per-step 30 --quick --no-parity > gpurun_out/r2h_weak_cheb_${N}gpu.json 2> gpurun_out/r2h_weak_cheb_${N}gpu.err; echo "weak cheb rc=$?")
(timeout 400 $TR --master-port 29515 bench.py --gpus $N --workload weak --steps 2 --warmup 1 --iters-per-step 50 --quick --no-parity > gpurun_out/r2h_weak_jacobi_${N}gpu.json 2> gpurun_out/r2h_weak_jacobi_${N}gpu.err; echo "weak jacobi rc=$?")
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h_*gpu*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), d["e2e"].get("rank0_h2d_gbs"), d["e2e"].get("rank0_d2h_gbs"), "comm", d["config"].get("comm"), "spmv_ms", d["roofline"]["ms_per_launch"], "parity", (d.get("parity") or {}).get("ok"), "weak", d["config"].get("weak"))
    except Exception as e:
        print(f, "ERR", e)
PY
