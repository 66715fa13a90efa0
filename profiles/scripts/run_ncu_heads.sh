python -m pytest tests/test_gpu_g_pimg.py -x -q 2>&1 | tail -3
python profiles/heads_only.py > gpurun_out/heads_only.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mlp_chain_kernel|gemm_bf16_tc_kernel" -c 24 -o gpurun_out/r02c_heads python profiles/heads_only.py > gpurun_out/ncu_heads.log 2>&1
ncu -i gpurun_out/r02c_heads.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__grid_size,launch__registers_per_thread,sm__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/r02c_heads_raw.csv 2>&1
tail -3 gpurun_out/ncu_heads.log
