# Multi-GPU bench lines (one node): bash tools/run_multi.sh N   (under gpurun --gpus N)
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-others --e2e-steps 2 > gpurun_out/n${N}_cfg2.json 2> gpurun_out/n${N}_cfg2.err; echo rc=$?
$T bench.py --gpus $N --workload cfg5 --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 > gpurun_out/n${N}_cfg5.json 2> gpurun_out/n${N}_cfg5.err; echo rc=$?
head -c 200 gpurun_out/n${N}_cfg2.json; echo; head -c 200 gpurun_out/n${N}_cfg5.json
