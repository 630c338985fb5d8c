T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 4 --workload kuka --scaling strong --batch 10000 --steps 2 --warmup 3 > gpurun_out/r2_kuka_strong_4gpu.json 2> gpurun_out/r2_kuka_strong_4gpu.err
$T bench.py --gpus 4 --workload dualkuka_train --steps 10 --warmup 3 > gpurun_out/r2_train_dualkuka_4gpu.json 2> gpurun_out/r2_train_dualkuka_4gpu.err
$T bench.py --gpus 4 --steps 1 --warmup 3 > gpurun_out/r2_4gpu_bench.json 2> gpurun_out/r2_4gpu_bench.err
tail -c 600 gpurun_out/r2_kuka_strong_4gpu.json gpurun_out/r2_train_dualkuka_4gpu.json gpurun_out/r2_4gpu_bench.json
