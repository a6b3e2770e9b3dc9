cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.log); tail -5 gpurun_out/r2f_pytest_gpu.log
python tools/ncu_targets.py assemble2 512 > gpurun_out/r2f_assemble2.log 2>&1; cat gpurun_out/r2f_assemble2.log
python tools/bench_spmv.py --nx 180 --ny 180 --nz 2048 --reps 50 > gpurun_out/r2f_i16_u8.log 2>&1; HEAT_SPMV_CIDX=2 python tools/bench_spmv.py --nx 180 --ny 180 --nz 2048 --reps 50 > gpurun_out/r2f_i16_i16.log 2>&1; HEAT_SPMV_CIDX=0 python tools/bench_spmv.py --nx 180 --ny 180 --nz 2048 --reps 50 > gpurun_out/r2f_i16_i32.log 2>&1; cat gpurun_out/r2f_i16_*.log
python tools/ncu_targets.py explicit 256 > gpurun_out/r2f_t_exp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'values_kernel' -c 1 -o gpurun_out/r2f_ncu_values python tools/ncu_targets.py explicit 256 > gpurun_out/r2f_ncu2.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/r2f_t_exp.log
timeout 300 python bench.py --assemble explicit --nx 256 --steps 2 > gpurun_out/r2f_asm_explicit.json 2> gpurun_out/r2f_asm_explicit.err; echo "asm rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2f_asm_explicit.json').read().strip().splitlines()[-1]); print(d['phases_ms'], d['parity'])"
