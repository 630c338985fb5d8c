#!/bin/bash
# Collects the measured artefacts of a round on one B200 (run through gpurun from the repo root):
#   bash profiles/collect.sh r2       -> gpurun_out/r2_*   (then `python profiles/summarize.py ...` in the build container)
# Order: tests, plain bench, reference arm, and only then the ncu passes of the same commands.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_reference_arm.json 2> $out/${tag}_reference_arm.err
python bench.py --workload eval200 --steps 20 --warmup 5 > $out/${tag}_eval200.json 2> $out/${tag}_eval200.err
# training step and the other BASELINE configs under the same harness
python bench.py --workload dualkuka_train --steps 10 --warmup 3 > $out/${tag}_train_dualkuka.json 2> $out/${tag}_train_dualkuka.err
python bench.py --workload maze2d_train --steps 10 --warmup 3 > $out/${tag}_train_maze2d.json 2> $out/${tag}_train_maze2d.err
python bench.py --workload kuka --batch 5000 --steps 2 --warmup 3 > $out/${tag}_bench_kuka.json 2> $out/${tag}_bench_kuka.err
python bench.py --workload dualkuka --batch 5000 --steps 2 --warmup 3 > $out/${tag}_bench_dualkuka.json 2> $out/${tag}_bench_dualkuka.err
python bench.py --workload comp2 --batch 3700 --steps 2 --warmup 3 > $out/${tag}_bench_comp2.json 2> $out/${tag}_bench_comp2.err
python bench.py --workload comp3 --batch 3700 --steps 2 --warmup 3 > $out/${tag}_bench_comp3.json 2> $out/${tag}_bench_comp3.err
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 2400 --csv --log-file $out/${tag}_train_launches.csv \
    python bench.py --workload dualkuka_train --steps 1 --warmup 3 > $out/${tag}_ncu_train.log 2>&1
# launch list of the bench command (direct launches; one step of the full 10,000-problem workload)
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 800 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
# full capture: one whole eps evaluation (60 conv launches, every shape class) at the bench's chunk size
# (the report of 60 launches is ~200 MB: it is exported to csv on the box and removed; gpurun brings back <= 64 MiB)
ncu --set full --clock-control none -k regex:conv_tc_kernel --launch-skip 120 --launch-count 60 \
    -f -o /tmp/${tag}_conv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > $out/${tag}_ncu_conv.log 2>&1
ncu -i /tmp/${tag}_conv.ncu-rep --page raw --csv > $out/${tag}_conv_raw.csv 2>> $out/${tag}_ncu_conv.log
# full capture of the HBM-bound kernels of the path (sampler update, swept check, Kuka sphere check)
python tests/gpu_profile_aux.py > $out/${tag}_aux_plain.log 2>&1
: > $out/${tag}_aux_raw.csv
for k in step_kernel4 swept_lane kuka_kernel; do
    ncu --set full --clock-control none -k regex:$k --launch-skip 4 --launch-count 2 -f -o /tmp/${tag}_aux_$k \
        python tests/gpu_profile_aux.py > $out/${tag}_ncu_aux_$k.log 2>&1
    ncu -i /tmp/${tag}_aux_$k.ncu-rep --page raw --csv >> $out/${tag}_aux_raw.csv 2>> $out/${tag}_ncu_aux_$k.log
done
du -sh $out
ls -la $out/${tag}_*
