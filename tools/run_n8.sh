T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-others --e2e-steps 2 > gpurun_out/n8_cfg2.json 2> gpurun_out/n8_cfg2.err; echo rc=$?
$T bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 > gpurun_out/n8_cfg5.json 2> gpurun_out/n8_cfg5.err; echo rc=$?
tail -c 400 gpurun_out/n8_cfg2.err; head -c 300 gpurun_out/n8_cfg2.json; echo; head -c 300 gpurun_out/n8_cfg5.json
