#!/bin/bash
# ncu captures behind profiles/r2b_* (GPU box, one GPU).  Numbers printed under ncu are not bench values.
set -x
N="ncu --set full --clock-control none --import-source on -f"
B="--steps 3 --warmup 3 --no-cpu-baseline"
$N -k regex:k_sell_infer_staged -s 8 -c 2 -o gpurun_out/r2b_c2_infer python bench.py --workload c2 $B > gpurun_out/r2b_ncu_c2_infer.log 2>&1
$N -k regex:k_sell_learn_live -s 2 -c 2 -o gpurun_out/r2b_c2_learn python bench.py --workload c2 $B > gpurun_out/r2b_ncu_c2_learn.log 2>&1
$N -k regex:k_cell_infer_staged -s 6 -c 1 -o gpurun_out/r2b_c3_infer python bench.py --workload c3 $B > gpurun_out/r2b_ncu_c3_infer.log 2>&1
$N -k regex:k_cell_learn_live -s 2 -c 2 -o gpurun_out/r2b_c3_learn python bench.py --workload c3 $B > gpurun_out/r2b_ncu_c3_learn.log 2>&1
$N -k regex:k_sell_infer_staged -s 2 -c 1 -o gpurun_out/r2b_c5_infer python bench.py --workload c5 $B > gpurun_out/r2b_ncu_c5_infer.log 2>&1
$N -k regex:k_sell_infer -s 2 -c 1 -o gpurun_out/r2b_c4_infer python bench.py --workload c4 $B > gpurun_out/r2b_ncu_c4_infer.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_c2_launches.csv python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_launches.log 2>&1
for f in c2_infer c2_learn c3_infer c3_learn c5_infer c4_infer; do
  ncu -i gpurun_out/r2b_$f.ncu-rep --page raw --csv > gpurun_out/r2b_${f}_raw.csv 2>/dev/null
  python profiles/ncu_extract.py gpurun_out/r2b_${f}_raw.csv > gpurun_out/r2b_${f}_ncu_summary.txt
done
rm -f gpurun_out/r2b_c5_infer.ncu-rep gpurun_out/r2b_c4_infer.ncu-rep gpurun_out/r2b_c3_*.ncu-rep gpurun_out/r2b_c2_learn.ncu-rep
g=tests/golden/graphs/biased_coin
DWG_TIMING=1 sampler_b200/dw gibbs -m $g/graph.meta -v $g/graph.variables -f $g/graph.factors -w $g/graph.weights -o /tmp -q -l 2000 -i 2000 --alpha 0.1 --diminish 0.995 --sample_evidence --reg_param 0 -c 1 > gpurun_out/r2b_c1_phases.log 2>&1
python scripts/cli_e2e.py --bundled > gpurun_out/r2b_c1.json 2>&1
