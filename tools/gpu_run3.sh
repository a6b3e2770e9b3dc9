cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_gpu.log); tail -15 gpurun_out/r2c_pytest_gpu.log
for alt in 1 0; do for nz in 64 512; do HEAT_CG_ALTERNATE=$alt timeout 200 python tools/bench_spmv.py --nx 512 --ny 512 --nz $nz --reps 50 >> gpurun_out/r2c_alt$alt.log 2>&1; done; done
echo "--- alternate=1"; cat gpurun_out/r2c_alt1.log; echo "--- alternate=0"; cat gpurun_out/r2c_alt0.log
HEAT_SPMV_CIDX=2 timeout 200 python tools/bench_spmv.py --nx 512 --reps 50 > gpurun_out/r2c_i16_512.log 2>&1; cat gpurun_out/r2c_i16_512.log
python tools/ncu_targets.py assemble 512 > gpurun_out/r2c_t_asm.log 2>&1; cat gpurun_out/r2c_t_asm.log
timeout 300 python bench.py --assemble explicit --nx 256 --steps 2 > gpurun_out/r2c_asm_explicit.json 2> gpurun_out/r2c_asm_explicit.err; echo "asm rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2c_asm_explicit.json').read().strip().splitlines()[-1]); print(d['phases_ms'], d['parity'], d['roofline_assembly'])"
