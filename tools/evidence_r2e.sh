O=gpurun_out; mkdir -p $O
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/r2e_gputests.log 2>&1; tail -2 $O/r2e_gputests.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > $O/r2e_bench.json 2> $O/r2e_bench.err; tail -c 200 $O/r2e_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout -s KILL 600 ncu --metrics $M --clock-control none --csv --log-file $O/r2e_launches_k400.csv python bench.py --no-secondary --no-extras --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --no-graph > $O/r2e_ncu_k400.log 2>&1
timeout -s KILL 400 ncu --metrics $M --clock-control none --csv --log-file $O/r2e_launches_ucf.csv python bench.py --workload ucf --no-extras --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --no-graph > $O/r2e_ncu_ucf.log 2>&1
bash tools/evidence_norm.sh
