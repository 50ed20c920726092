# one `ncu --set full` capture of the tensor-core sweep (pass 1 and pass 2) on a reduced C3 batch
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tc -s 2 -c 2 -o gpurun_out/prof_tc -f python bench.py --objects 196608 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_tc.log 2>&1
echo ncu exit=$?
