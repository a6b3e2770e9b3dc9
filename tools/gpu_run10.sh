cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "cube or byte_indexed or int16 or assemble or full_size or reference_pins or smoke" > gpurun_out/r2j_pytest_cube.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_cube.log); tail -4 gpurun_out/r2j_pytest_cube.log
python tools/ncu_targets.py assemble2 512 > gpurun_out/r2j_assemble2.log 2>&1; cat gpurun_out/r2j_assemble2.log
python tools/ncu_targets.py assemble 384 > gpurun_out/r2j_t_asm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cube_sell_kernel -c 1 -o gpurun_out/r2j_ncu_cube_sell python tools/ncu_targets.py assemble 384 > gpurun_out/r2j_ncu1.log 2>&1; echo "ncu rc=$?"
