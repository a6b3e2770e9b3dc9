cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2m_bench_${N}gpu.json 2> gpurun_out/r2m_bench_${N}gpu.err; echo "bench rc=$?"); tail -c 300 gpurun_out/r2m_bench_${N}gpu.err
(HEAT_B200_LIB=$GRAFT_REPO_ROOT/domain-decomposed-pde-solver_b200/lib_trace/libheat_b200.so HEAT_PEER_TRACE_FILE=gpurun_out/r2m_trace${N}_ timeout 400 $TR --master-port 29513 tools/peer_trace.py > gpurun_out/r2m_peer_trace_${N}gpu.log 2>&1; echo "trace rc=$?"); python tools/peer_trace.py --analyse gpurun_out/r2m_trace${N}_ $N >> gpurun_out/r2m_peer_trace_${N}gpu.log 2>&1; grep -v "^\[W\|^$\|^\*\|OMP" gpurun_out/r2m_peer_trace_${N}gpu.log | tail -9
(CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 5 --warmup 3 --quick --no-parity --no-cpu-baseline > gpurun_out/r2m_bench_1gpu_samebox.json 2> /dev/null; echo "n1 rc=$?")
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), round(d["e2e"].get("serial_value",0),1), d["e2e"].get("rank0_h2d_gbs"), "comm", d["config"].get("comm"), "spmv_ms", d["roofline"]["ms_per_launch"], "parity", (d.get("parity") or {}).get("ok"), "weak", d["config"].get("weak"))
    except Exception as e:
        print(f, "ERR", e)
PY
