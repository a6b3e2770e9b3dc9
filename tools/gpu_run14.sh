cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_gpu.log); tail -3 gpurun_out/r2n_pytest_gpu.log
for pdl in 1 0; do HEAT_PDL=$pdl CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/bench_spmv.py --nx 512 --ny 512 --nz 64 --reps 50 2>&1 | tail -1; done
(HEAT_REQUIRE_PEER=1 timeout 1200 $TR --master-port 29511 tests/mgpu_worker.py > gpurun_out/r2n_mgpu_worker_${N}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2n_mgpu_worker_${N}gpu.log); grep -v "^\[W\|^$" gpurun_out/r2n_mgpu_worker_${N}gpu.log | tail -3
(timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --quick > gpurun_out/r2n_bench_${N}gpu.json 2> gpurun_out/r2n_bench_${N}gpu.err; echo "bench rc=$?"); python -c "
import json; d=json.loads(open('gpurun_out/r2n_bench_${N}gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['parity']['ok'])"
