T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
$T bench.py --gpus 8 --workload kuka --scaling strong --batch 10000 --steps 3 --warmup 3 > gpurun_out/r2_kuka_strong_8gpu.json 2> gpurun_out/r2_kuka_strong_8gpu.err
tail -c 400 gpurun_out/r2_kuka_strong_8gpu.json
