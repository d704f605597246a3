# Round evidence: GPU tests, the default bench line, launch list + full ncu capture of the headline kernel (cfg2, 1024 users) and cfg5.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
CMD="python bench.py --workload cfg2 --users 1024 --steps 2 --warmup 1 --no-cpu --no-others --e2e-steps 0"
$CMD > gpurun_out/plain_cfg2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv $CMD > gpurun_out/ncu_l_cfg2.log 2>&1
$CMD > gpurun_out/plain_cfg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fd_ws -s 2 -c 1 -f -o gpurun_out/prof_fd_ws_cfg2 $CMD > gpurun_out/ncu_f_cfg2.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
head -c 600 gpurun_out/bench_default.json
