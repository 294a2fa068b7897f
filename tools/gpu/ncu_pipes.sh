mkdir -p gpurun_out
M=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,smsp__inst_executed.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
timeout 300 ./scratch/variants17 > gpurun_out/variants17.log 2>&1; cat gpurun_out/variants17.log
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/pipes_variants17.csv ./scratch/variants17 > /dev/null 2>&1
echo "v17 $?"
