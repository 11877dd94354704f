#!/bin/bash
# round-2 final single-GPU batch: tests, benches, profiles
O=gpurun_out/final2; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1
echo "gpu tests rc=$?" | tee $O/status.txt; tail -n 3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?" | tee -a $O/status.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_impl_reference.json 2> $O/bench_impl_reference.err
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_c5_n1.json 2> $O/bench_c5_n1.err
echo "c5 n1 rc=$?" | tee -a $O/status.txt
timeout 900 python tools/bench_configs.py --c4-graphs 500 > $O/bench_configs.jsonl 2> $O/bench_configs.err
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-grad >> $O/bench_configs.jsonl 2>> $O/bench_configs.err
timeout 300 python tools/bench_configs.py --only C4ref --c4-graphs 60 >> $O/bench_configs.jsonl 2>> $O/bench_configs.err
GDB_FORCE_GENERAL=1 timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 >> $O/bench_configs.jsonl 2>> $O/bench_configs.err
for o in rcm random; do timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 --c4-order $o >> $O/bench_configs.jsonl 2>> $O/bench_configs.err; done
timeout 300 python tools/bench_configs.py --only C4 --c4-graphs 100 >> $O/bench_configs.jsonl 2>> $O/bench_configs.err
echo "configs done" | tee -a $O/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --parity-samples 20 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlgk_solve_small -s 3 -c 1 -f -o $O/prof_c3 python tools/profile_c3.py > $O/ncu_c3.log 2>&1
echo "ncu c3 rc=$?" | tee -a $O/status.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlgk_solve -s 1 -c 1 -f -o $O/prof_c4 python tools/profile_c4.py --n-graphs 24 > $O/ncu_c4.log 2>&1
echo "ncu c4 rc=$?" | tee -a $O/status.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py --large > $O/sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" | tee -a $O/status.txt; tail -n 2 $O/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py --large > $O/sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?" | tee -a $O/status.txt; tail -n 2 $O/sanitizer_racecheck.txt
cat $O/status.txt
