K2='sage_fused|long_partial|long_reduce|collect_long|score_topk|rescore|score_prep|linear_small|score_band|colmean|exact_topk|topk_merge|order_|permute_rows|radix_|scan_kernel|csr_finish'
M='dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread'
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02z_ncu_launches.log 2>&1
timeout 240 ncu --metrics $M --clock-control none -k regex:"$K2" -c 400 --csv --log-file gpurun_out/r02z_c2_step_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02z_ncu_metrics.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'score_topk|sage_fused|long_partial|rescore_kernel' -s 12 -c 12 -o /tmp/c2_top python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02z_ncu_full.log 2>&1
ncu -i /tmp/c2_top.ncu-rep --page raw --csv --print-units base > gpurun_out/r02z_c2_top_full.csv 2>/dev/null
ncu -i /tmp/c2_top.ncu-rep --page details --csv > gpurun_out/r02z_c2_top_details.csv 2>/dev/null
ls -la gpurun_out/r02z*; du -sh gpurun_out
