M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread
for L in lib_base libbpc_b200; do
BPC_LIB=$PWD/bpc_baseline_b200/$L.so ncu --metrics $M --clock-control none -k regex:bpc_crop -s 12 -c 4 --csv --log-file gpurun_out/ncu_ab_$L.csv python tools/crop_sweep.py 60 400 224 16384 > /dev/null 2>&1
done
