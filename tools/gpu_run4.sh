cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(HEAT_REQUIRE_PEER=1 timeout 1200 $TR --master-port 29511 tests/mgpu_worker.py > gpurun_out/r2d_mgpu_worker_${N}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2d_mgpu_worker_${N}gpu.log); tail -12 gpurun_out/r2d_mgpu_worker_${N}gpu.log
(timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2d_bench_${N}gpu.json 2> gpurun_out/r2d_bench_${N}gpu.err; echo "bench rc=$?"); tail -c 400 gpurun_out/r2d_bench_${N}gpu.err
(timeout 400 $TR --master-port 29513 bench.py --gpus $N --workload weak --prec chebyshev --cheb-lambda-max 2.0 --steps 2 --warmup 1 --iters-per-step 30 --quick --no-parity > gpurun_out/r2d_weak_cheb_${N}gpu.json 2> gpurun_out/r2d_weak_cheb_${N}gpu.err; echo "weak cheb rc=$?")
(HEAT_COMM=nccl timeout 400 $TR --master-port 29514 bench.py --gpus $N --workload weak --prec chebyshev --cheb-lambda-max 2.0 --steps 2 --warmup 1 --iters-per-step 30 --quick --no-parity > gpurun_out/r2d_weak_cheb_${N}gpu_nccl.json 2> gpurun_out/r2d_weak_cheb_${N}gpu_nccl.err; echo "weak cheb nccl rc=$?")
(timeout 400 $TR --master-port 29515 bench.py --gpus $N --workload weak --steps 2 --warmup 1 --iters-per-step 50 --quick --no-parity > gpurun_out/r2d_weak_jacobi_${N}gpu.json 2> gpurun_out/r2d_weak_jacobi_${N}gpu.err; echo "weak jacobi rc=$?")
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_*gpu*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "comm", d["config"].get("comm"), "spmv_ms", d["roofline"]["ms_per_launch"], "parity", (d.get("parity") or {}).get("ok"), "weak", d["config"].get("weak"))
    except Exception as e:
        print(f, "ERR", e)
PY
