#!/bin/bash
# round 2, final single-GPU pass on the end-of-round build (run under gpurun): tests, smoke, default bench, launch
# lists, ncu captures; the files it leaves in gpurun_out/ are what profiles/r02_* were made from
set -o pipefail
timeout 900 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -30 > gpurun_out/r02z_pytest.log; echo pytest $?
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo smoke $?
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err ) 2> gpurun_out/r02z_bench_n1.time; echo bench $?
timeout 200 python tools/time_pool.py --json gpurun_out/r02z_time_pool.json > gpurun_out/r02z_time_pool.log 2>&1; echo pool $?
NB="--kernel-name-base demangled"
Q='--no-cpu-baseline --sub-batches "" --parity 0'
eval timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 400 --csv --log-file gpurun_out/r02z_launches_c2.csv python bench.py --workloads none $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_a.err"; echo A $?
eval timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 400 --csv --log-file gpurun_out/r02z_launches_c3.csv python bench.py --workload c3 --workloads none $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_b.err"; echo B $?
eval timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 700 --csv --log-file gpurun_out/r02z_launches_c5.csv python bench.py --workloads c5 $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_c.err"; echo C $?
timeout 200 ncu --set full --clock-control none --import-source on $NB -k regex:pool_ring -s 3 -c 1 -f -o gpurun_out/r02z_pool_ring python tools/time_pool.py --ring 1 --cases 1 > /dev/null 2> gpurun_out/ncu_f.err; echo ncuF $?
eval timeout 250 ncu --set full --clock-control none --import-source on $NB -k '"regex:tc_coarse_pair|tc_select|tc_maxima"' -s 12 -c 4 -f -o gpurun_out/r02z_tc_c2 python bench.py --workloads none $Q --steps 1 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_d.err"; echo ncuD $?
cat gpurun_out/r02z_bench_n1.time; tail -3 gpurun_out/r02z_pytest.log; cat gpurun_out/r02z_smoke.log | tail -2
