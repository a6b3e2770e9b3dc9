cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi nvlink -gt d -i 0 2>&1 | head -8
(HEAT_REQUIRE_PEER=1 timeout 1200 $TR --master-port 29511 tests/mgpu_worker.py > gpurun_out/r2l_mgpu_worker_${N}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2l_mgpu_worker_${N}gpu.log); grep -v "^\[W\|^$" gpurun_out/r2l_mgpu_worker_${N}gpu.log | tail -3
(timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --quick > gpurun_out/r2l_bench_${N}gpu.json 2> gpurun_out/r2l_bench_${N}gpu.err; echo "bench rc=$?"); python -c "
import json; d=json.loads(open('gpurun_out/r2l_bench_${N}gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['parity']['ok'])"
