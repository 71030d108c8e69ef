#!/bin/bash
# Copy the outputs of tools/evidence_run.sh (merged back into gpurun_out/) into profiles/ under this round's names and
# regenerate the per-kernel summaries.  Run in the container, from the repo root, after the gpurun call.
set -e
R=${ROUND:-r02}; G=gpurun_out; P=profiles
for f in joint image_b128 text_b32 infer reference; do cp $G/r2_bench_${f}_final.json $P/${R}_bench_$f.json; done
for f in $G/parity_t*.json; do cp $f $P/${R}_$(basename $f); done
cp $G/r2_pytest_gpu_final.log $P/${R}_pytest_gpu.log
cp $G/r2_sanitizer_halo.log $P/${R}_compute_sanitizer_halo.txt
cp $G/r2_launches_final.csv $P/${R}_launches_final.csv
cp $G/r2_launches_infer_final.csv $P/${R}_launches_infer.csv
python tools/launch_report.py $P/${R}_launches_final.csv --traffic-json $P/${R}_conv_traffic.json > $P/${R}_launches_final_summary.txt
python tools/launch_report.py $P/${R}_launches_infer.csv > $P/${R}_launches_infer_summary.txt
PYTHONPATH=. python tools/tensor_pipe_report.py $G/r2_m4.csv 3 > $P/${R}_mixed4_tensor_pipe.csv
python tools/ncu_summary.py $G/r2_prof_conv_pair.ncu-rep > $P/${R}_ncu_conv_pair_mixed4e_b1_summary.txt
python tools/ncu_summary.py $G/r2_prof_halo_final.ncu-rep > $P/${R}_ncu_conv_halo_mixed4e_b2_summary.txt
cp $G/r2_halo_sweep_final.log $P/${R}_halo_sweep.txt
cp $G/r2_policy_sweep_final.log $P/${R}_policy_sweep.txt
tail -n 3 $P/${R}_launches_final_summary.txt; tail -n 4 $P/${R}_mixed4_tensor_pipe.csv; tail -n 2 $P/${R}_pytest_gpu.log
