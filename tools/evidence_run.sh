set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -n 3 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python bench.py --dump-launches gpurun_out/r2_launches_joint_final.json > gpurun_out/r2_bench_joint_final.json 2> gpurun_out/r2_bench_joint_final.err
timeout 200 python bench.py --model image --batch 128 > gpurun_out/r2_bench_image_b128_final.json 2> gpurun_out/r2_bench_image_final.err
timeout 200 python bench.py --model text --batch 32 > gpurun_out/r2_bench_text_b32_final.json 2> gpurun_out/r2_bench_text_final.err
timeout 200 python bench.py --mode infer > gpurun_out/r2_bench_infer_final.json 2> gpurun_out/r2_bench_infer_final.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 --cpu-budget-s 40 > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv python tools/step_for_ncu.py --batch 256 > gpurun_out/r2_ncu_final.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_infer_final.csv python tools/step_for_ncu.py --batch 256 --forward-only > gpurun_out/r2_ncu_infer_final.log 2>&1
timeout 300 ncu --kernel-name "regex:conv_bf16x3|conv3x3_halo" --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r2_m4.csv python tools/bench_conv.py --only Mixed_4 --reps 2 > gpurun_out/r2_m4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_bf16x3 -c 3 -o gpurun_out/r2_prof_conv_pair python tools/bench_conv.py --only "Mixed_4e b1" --reps 2 > gpurun_out/r2_prof_conv_pair.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -c 3 -o gpurun_out/r2_prof_halo_final python tools/bench_conv.py --only "Mixed_4e b2" --reps 2 > gpurun_out/r2_prof_halo_final.log 2>&1
timeout 200 python tools/bench_halo.py --json gpurun_out/r2_halo_sweep_final.json > gpurun_out/r2_halo_sweep_final.log 2>&1
timeout 400 python tools/policy_sweep.py gpurun_out/r2_launches_joint_final.json --json gpurun_out/r2_policy_sweep_final.json > gpurun_out/r2_policy_sweep_final.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_halo_gpu.py -q -k "3-14-32-64 or 2-28-16-32 or 1-56-192-64 or slices" > gpurun_out/r2_sanitizer_halo.log 2>&1; tail -n 5 gpurun_out/r2_sanitizer_halo.log
for f in joint image_b128 text_b32 infer reference; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2_bench_%s_final.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "sus", d.get("sustained") and round(d["sustained"]["value"]), d.get("parity") and (d["parity"]["logits_rel_l2"], d["parity"]["ok"]), d.get("cpu_baseline",{}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
