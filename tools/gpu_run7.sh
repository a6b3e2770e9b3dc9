cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0 python tools/ncu_targets.py assemble2 512 > gpurun_out/r2g_assemble2.log 2>&1; cat gpurun_out/r2g_assemble2.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --assemble explicit --nx 256 --steps 2 > gpurun_out/r2g_asm_explicit.json 2> gpurun_out/r2g_asm_explicit.err; echo "asm rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2g_asm_explicit.json').read().strip().splitlines()[-1]); print(d['phases_ms'], d['parity'])"
(CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -q -x -k "cube or byte_indexed or int16 or assemble or smoke" > gpurun_out/r2g_pytest_subset.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_subset.log); tail -3 gpurun_out/r2g_pytest_subset.log
(HEAT_REQUIRE_PEER=1 timeout 1200 $TR --master-port 29511 tests/mgpu_worker.py > gpurun_out/r2g_mgpu_worker_${N}gpu.log 2>&1; echo "worker rc=$?" >> gpurun_out/r2g_mgpu_worker_${N}gpu.log); grep -v "^\[W\|^$" gpurun_out/r2g_mgpu_worker_${N}gpu.log | tail -6
(timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2g_bench_${N}gpu.json 2> gpurun_out/r2g_bench_${N}gpu.err; echo "bench rc=$?"); tail -c 300 gpurun_out/r2g_bench_${N}gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_bench_*gpu*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), d["e2e"].get("rank0_h2d_gbs"), d["e2e"].get("rank0_d2h_gbs"), "comm", d["config"].get("comm"), "spmv_ms", d["roofline"]["ms_per_launch"], "parity", (d.get("parity") or {}).get("ok"), "weak", d["config"].get("weak"))
    except Exception as e:
        print(f, "ERR", e)
PY
