cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(HEAT_REQUIRE_PEER=1 timeout 1200 $TR --master-port 29511 tests/mgpu_worker.py > gpurun_out/r2e_mgpu_worker_${N}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2e_mgpu_worker_${N}gpu.log); grep -v "^\[W\|^$" gpurun_out/r2e_mgpu_worker_${N}gpu.log | tail -14
for pdl in 1 0; do for nz in 64 512; do HEAT_PDL=$pdl CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/bench_spmv.py --nx 512 --ny 512 --nz $nz --reps 50 >> gpurun_out/r2e_pdl$pdl.log 2>&1; done; done
echo "--- pdl=1"; cat gpurun_out/r2e_pdl1.log; echo "--- pdl=0"; cat gpurun_out/r2e_pdl0.log
for pdl in 1 0; do (HEAT_PDL=$pdl timeout 400 $TR --master-port 2951$pdl bench.py --gpus $N --steps 5 --warmup 3 --quick --no-parity > gpurun_out/r2e_bench_${N}gpu_pdl$pdl.json 2> gpurun_out/r2e_bench_${N}gpu_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"); done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e_*gpu*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "comm", d["config"].get("comm"), "spmv_ms", d["roofline"]["ms_per_launch"])
    except Exception as e:
        print(f, "ERR", e)
PY
