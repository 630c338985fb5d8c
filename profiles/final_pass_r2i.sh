python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -3 gpurun_out/r2i_pytest.log
python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
ncu --set full --clock-control none -k regex:swept_lane --launch-skip 4 --launch-count 2 -f -o /tmp/r2i_swept python tests/gpu_profile_aux.py > gpurun_out/r2i_ncu_swept.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2i_swept.ncu-rep --page raw --csv > gpurun_out/r2i_aux_raw.csv 2>> gpurun_out/r2i_ncu_swept.log
ls -la gpurun_out/r2i_*
