#!/bin/bash
# Round-end evidence run on ONE B200 (under gpurun): tests, the bench line, ncu launch lists and full captures -> gpurun_out/
# Summaries are made here (no GPU) with tools/launch_summary.py / tools/ncu_report.py / tools/ncu_top.py and committed under profiles/.
set -u
O=gpurun_out
mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/r2_gputests.log 2>&1; tail -3 $O/r2_gputests.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > $O/r2_bench_final.json 2> $O/r2_bench_final.err; tail -c 300 $O/r2_bench_final.err
# launch lists: DEVIAS_BENCH_LAUNCH_LIST=1 = eager, exactly --warmup + --steps steps, no roofline leg (a K400 step takes minutes
# under ncu).  Time only is ONE pass per kernel (~4 min for K400); add ,dram__bytes_read.sum,dram__bytes_write.sum (M=...) for the
# traffic table: that replays every kernel and needs > 10 min for the K400 step.
M=${M:-gpu__time_duration.sum}
export DEVIAS_BENCH_LAUNCH_LIST=1
timeout -s KILL 900 ncu --metrics $M --clock-control none --csv --log-file $O/r2_launches_k400_final.csv python bench.py --no-secondary --no-extras --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --no-graph > $O/r2_ncu_k400.log 2>&1
timeout -s KILL 600 ncu --metrics $M --clock-control none --csv --log-file $O/r2_launches_ucf_final.csv python bench.py --workload ucf --no-extras --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --no-graph > $O/r2_ncu_ucf.log 2>&1
unset DEVIAS_BENCH_LAUNCH_LIST
DEVIAS_ONESHOT=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 12 -f -o $O/r2_gemm_final python tools/bench_gemm.py 32 > /dev/null 2>&1
DEVIAS_ONESHOT=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:flash -c 4 -f -o $O/r2_flash_final python tools/bench_attn.py 32 > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:slot_stream -s 2 -c 2 -f -o $O/r2_slot_final python tools/ncu_slot.py > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:patch_embed_fwd -c 1 -f -o $O/r2_patch_final python bench.py --no-secondary --no-extras --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --no-graph > /dev/null 2>&1
bash tools/evidence_norm.sh
ls -la $O/r2_*final* | awk '{print $5, $9}'
