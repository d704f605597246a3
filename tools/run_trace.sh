cp deepmimo_b200/libdmk.so /tmp/libdmk_prod.so
DMK_NVCC_EXTRA=-DDMK_TC_TRACE python -m deepmimo_b200.build --force > gpurun_out/trace_build.log 2>&1
python tools/tc_trace_p.py cfg5 20000 > gpurun_out/tracep_cfg5.txt 2>&1
python tools/tc_trace_p.py cfg2 2048 > gpurun_out/tracep_cfg2.txt 2>&1
cp /tmp/libdmk_prod.so deepmimo_b200/libdmk.so
ncu --set full --clock-control none --import-source on -k regex:fd_tc_persist -s 2 -c 1 -o gpurun_out/prof_tcp_cfg5 python bench.py --workload cfg5 --users 20000 --steps 2 --warmup 1 --no-cpu --no-others --e2e-steps 0 > gpurun_out/ncu_cfg5.log 2>&1
tail -3 gpurun_out/ncu_cfg5.log
cat gpurun_out/tracep_cfg5.txt
