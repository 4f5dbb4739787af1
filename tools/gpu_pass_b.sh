#!/bin/bash
# round 2, GPU pass B: tests after the select / threshold / hybrid / pool changes, timings, launch lists
set -o pipefail
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -30 > gpurun_out/r02c_pytest.log; echo pytest $?
timeout 200 python tools/time_pool.py --json gpurun_out/r02c_time_pool.json > gpurun_out/r02c_time_pool.log 2>&1; echo pool $?
timeout 300 python bench.py --workloads c5 --no-cpu-baseline --sub-batches "" --parity 0 --steps 20 --warmup 5 > gpurun_out/r02c_bench_c5.json 2> gpurun_out/r02c_bench_c5.err; echo c5 $?
timeout 300 python bench.py --workloads none --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/r02c_bench_c2.json 2> gpurun_out/r02c_bench_c2.err; echo c2 $?
timeout 300 python bench.py --workload c2s8 --workloads none --no-cpu-baseline --sub-batches "" --steps 50 --warmup 5 > gpurun_out/r02c_bench_c2s8.json 2> gpurun_out/r02c_bench_c2s8.err; echo c2s8 $?
NB="--kernel-name-base demangled"
Q='--no-cpu-baseline --sub-batches "" --parity 0'
eval timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 400 --csv --log-file gpurun_out/r02c_launches_c2.csv python bench.py --workloads none $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_a.err"; echo A $?
eval timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 400 --csv --log-file gpurun_out/r02c_launches_c2s8.csv python bench.py --workload c2s8 --workloads none $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_b.err"; echo B $?
eval timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 600 --csv --log-file gpurun_out/r02c_launches_c5.csv python bench.py --workloads c5 $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_c.err"; echo C $?
tail -4 gpurun_out/r02c_pytest.log
cat gpurun_out/r02c_time_pool.log
