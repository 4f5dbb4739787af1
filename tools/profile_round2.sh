#!/bin/bash
# Round-2 profiling pass (run under gpurun on one B200):  bash tools/profile_round2.sh
#  A/B  launch lists (device time of every archi kernel launch) of the config-2 and config-3 bench commands
#  C/D  ncu --set full captures of the dominant kernel (the CTA-pair coarse scorer) for both configs
#  E    ncu --set full of the pool+normalise kernel and the posting-list hybrid kernels (config 5)
# Outputs land in gpurun_out/; the summaries under profiles/r02_* are made from them with tools/ncu_digest.py and
# tools/launch_shares.py.  (Mid-round pass; tools/profile_round2_final.sh is the pass on the end-of-round build,
# tools/profile_throttle.sh the A/B of the soft throttle.)
NB="--kernel-name-base demangled"
Q='--no-cpu-baseline --sub-batches "" --parity 0'
run() { eval "$@"; }
run timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 400 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --workloads none $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_a.err"; echo A $?
run timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k regex:archi -c 400 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --workload c3 --workloads none $Q --steps 2 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_b.err"; echo B $?
run timeout 250 ncu --set full --clock-control none --import-source on $NB -k regex:tc_coarse_pair -s 6 -c 2 -f -o gpurun_out/r02_tc_coarse_c2 python bench.py --workloads none $Q --steps 1 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_c.err"; echo C $?
run timeout 300 ncu --set full --clock-control none --import-source on $NB -k regex:tc_coarse_pair -s 2 -c 2 -f -o gpurun_out/r02_tc_coarse_c3 python bench.py --workload c3 --workloads none $Q --steps 1 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_d.err"; echo D $?
run timeout 300 ncu --set full --clock-control none $NB -k '"regex:pool_normalize|hyb_score|hyb_merge|hyb_collect|hyb_scatter|scan_topk"' -c 14 -f -o gpurun_out/r02_pool_hybrid python bench.py --workloads c5 $Q --steps 3 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_e.err"; echo E $?
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
