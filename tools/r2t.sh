CMD="python bench.py --sampling topk --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$CMD > gpurun_out/r2t_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:par_\|finalize -s 24 -c 24 --csv --log-file gpurun_out/r2t_ncu_par.csv $CMD > gpurun_out/r2t_ncu.log 2>&1
CMD2="python bench.py --sampling nucleus --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$CMD2 > gpurun_out/r2t_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:par_\|finalize -s 22 -c 22 --csv --log-file gpurun_out/r2t_ncu_par_nucleus.csv $CMD2 > gpurun_out/r2t_ncu2.log 2>&1
