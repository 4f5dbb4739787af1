#!/bin/bash
# round 2, GPU pass D: tests, pool ring sweeps, hybrid + ingest numbers, launch lists
set -o pipefail
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -30 > gpurun_out/r02e_pytest.log; echo pytest $?
for cps in 4 2 1; do for ck in 11 22 44; do
  echo "== ctas_per_sm $cps chunk_kb $ck"
  ARCHI_POOL_CTAS_PER_SM=$cps ARCHI_POOL_CHUNK_KB=$ck timeout 100 python tools/time_pool.py --ring 1 --cases 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['B'], round(d['ring']['graph_ms']*1e3,1),'us', round(d['ring']['frac_hbm'],3))
"
done; done > gpurun_out/r02e_pool_sweep.log 2>&1
timeout 300 python bench.py --workloads c5 --no-cpu-baseline --sub-batches "" --parity 0 --steps 20 --warmup 5 > gpurun_out/r02e_bench_c5.json 2> gpurun_out/r02e_bench_c5.err; echo c5 $?
timeout 400 python tools/bench_ingest_driver.py > gpurun_out/r02e_ingest_driver.json 2> gpurun_out/r02e_ingest_driver.err; echo ingest $?
timeout 400 python tools/bench_ingest_driver.py --per-file > gpurun_out/r02e_ingest_driver_perfile.json 2>> gpurun_out/r02e_ingest_driver.err; echo ingestpf $?
timeout 400 python tools/bench_ingest_driver.py --files 2000 > gpurun_out/r02e_ingest_driver_2000.json 2>> gpurun_out/r02e_ingest_driver.err; echo ingest2000 $?
NB="--kernel-name-base demangled"
Q='--no-cpu-baseline --sub-batches "" --parity 0'
eval timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 400 --csv --log-file gpurun_out/r02e_launches_c2.csv python bench.py --workloads none $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_a.err"; echo A $?
eval timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 600 --csv --log-file gpurun_out/r02e_launches_c5.csv python bench.py --workloads c5 $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_c.err"; echo C $?
B="python bench.py --workloads none --no-cpu-baseline --sub-batches '' --steps 100 --warmup 5"
eval timeout 200 $B --workload c2s8 > gpurun_out/r02e_c2s8.json 2> /dev/null; echo c2s8 $?
eval timeout 200 $B > gpurun_out/r02e_c2.json 2>/dev/null; echo c2 $?
tail -4 gpurun_out/r02e_pytest.log
cat gpurun_out/r02e_pool_sweep.log
cat gpurun_out/r02e_ingest_driver*.json
for f in c2s8 c2; do python - <<P
import json
for l in open('gpurun_out/r02e_$f.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['ms_per_step'],4), round(d['value']), 'launch_ms', round(d['roofline']['launch_ms'],4), 'host', round(d.get('host_enqueue_ms_per_step',0),4), 'parity', d.get('parity_checked'), d.get('parity_failed'), 'unv', d.get('unverified_queries'), d['clocks']['reasons'])
P
done
