#!/bin/bash
# ncu --set full captures of the memory-bound helpers (LayerNorm fwd / bwd, bias column sums, flash-backward preparation) and of
# the 64-row slot-side products; summarised with tools/ncu_report.py into profiles/r02_ncu_norm.md
set -u
O=gpurun_out
mkdir -p $O
DEVIAS_ONESHOT=1 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'layernorm|colsum' -s 1 -c 3 -f -o $O/r2_norm python tools/bench_norm.py 32 > $O/r2_norm.log 2>&1
DEVIAS_ONESHOT=1 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'flash_bwd_prep|flash_dq_convert' -c 2 -f -o $O/r2_prep python tools/bench_attn.py 32 > /dev/null 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tile_gemm -s 6 -c 9 -f -o $O/r2_tile python tools/ncu_skinny.py > /dev/null 2>&1
timeout -s KILL 120 python tools/bench_norm.py 32 > $O/r2_norm_bench.log 2>&1
ls -la $O/r2_norm* $O/r2_prep* $O/r2_tile* | awk '{print $5, $9}'
