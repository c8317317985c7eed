# ncu --set full captures of the kernels added in the last session of round 2 (each after the same command has exited 0 without ncu).
set -x
A="python bench.py --batch 256 --context-min 2048 --context-max 2048 --kv-int8 --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$A > gpurun_out/r2ax_plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:decode_attn_q8 --launch-skip 30 -c 1 -f -o gpurun_out/r2ax_q8_transposed $A > gpurun_out/r2ax_ncu_a.log 2>&1
B="python bench.py --kv-int8 --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$B > gpurun_out/r2ax_plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_persistent --launch-skip 6 -c 1 -f -o gpurun_out/r2ax_persistent_int8 $B > gpurun_out/r2ax_ncu_b.log 2>&1
C="python bench.py --paged 32 --steps 2 --warmup 3 --no-graph --skip-cpu-baseline --no-verify"
$C > gpurun_out/r2ax_plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_persistent --launch-skip 6 -c 1 -f -o gpurun_out/r2ax_persistent_paged32 $C > gpurun_out/r2ax_ncu_c.log 2>&1
ls -la gpurun_out/r2ax_*.ncu-rep
