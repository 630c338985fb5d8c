python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_reference_arm.json 2> gpurun_out/r2f_reference_arm.err; echo "ref rc=$?"
python bench.py --workload comp4 --batch 3700 --steps 2 --warmup 3 > gpurun_out/r2f_bench_comp4.json 2> gpurun_out/r2f_bench_comp4.err; echo "comp4 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"
