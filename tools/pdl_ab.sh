# A/B of the programmatic-dependent-launch policy (ds_dependent_launch; env DS_PDL = "main[,side]" overrides the engine's choice)
cd $GRAFT_REPO_ROOT
run() {  # name, pdl, bench args...
  name=$1; pdl=$2; shift; shift
  DS_PDL=$pdl timeout 200 python bench.py --no-cpu-baseline --no-kernel-pass --sustained-s 0 "$@" > gpurun_out/r2_pdl_${name}_$pdl.json 2> gpurun_out/r2_pdl_${name}_$pdl.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2_pdl_${name}_$pdl.json').read().strip().splitlines()[-1]);print('$name pdl=$pdl', round(d['ms_per_step'],3), round(d['value']))" || tail -n 5 gpurun_out/r2_pdl_${name}_$pdl.err
}
for pdl in ${PDL_MODES:-0 1 2 3 5 6 7 0}; do
  for w in ${PDL_WORKLOADS:-joint image text infer}; do
    case $w in
      joint) run joint $pdl;;
      image) run image $pdl --model image --batch 128;;
      image256) run image256 $pdl --model image --batch 256;;
      text) run text $pdl --model text --batch 32;;
      infer) run infer $pdl --mode infer;;
    esac
  done
done
