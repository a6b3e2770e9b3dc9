cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_gpu.log); tail -5 gpurun_out/r2i_pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2i_smoke.log); tail -2 gpurun_out/r2i_smoke.log
(timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2i_bench1.json 2> gpurun_out/r2i_bench1.err; echo "bench rc=$?"); tail -c 400 gpurun_out/r2i_bench1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2i_bench1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"], "asm", d["roofline_assembly"]["ms_fill_kernel"], d["roofline_assembly"]["ms_whole_assemble_events"], "cpu", d["cpu_baseline"]["value"], "parity", d["parity"]["ok"])
PY
