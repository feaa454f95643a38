#!/usr/bin/env bash
# First GPU call of round 2: validate the variants that were written without hardware access at the end of round 1
#  and A/B them.  Every step runs under its own timeout and logs to gpurun_out/r2_<step>.log, so a
# hang or crash in one variant does not take the others (or the box) down.
#   gpurun --timeout 1500 -- 'bash tools/round2_checks.sh'            # all steps (~15 min)
#   gpurun --timeout 600  -- 'bash tools/round2_checks.sh probe 2cta' # selected steps
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
steps=("$@")
[ ${#steps[@]} -eq 0 ] && steps=(probe 2cta persist napfused parfwd vgg convab bench)
run() {   # name, timeout seconds, command...
  local name=$1 t=$2; shift 2
  echo "=== $name ($(date +%T)) ==="
  timeout "$t" "$@" > "gpurun_out/r2_$name.log" 2>&1
  echo "rc=$? ; tail:"; tail -n 12 "gpurun_out/r2_$name.log"
}
bench_line() {  # tag, env assignments...
  local tag=$1; shift
  echo "--- bench $tag"
  env "$@" timeout 240 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2> "gpurun_out/r2_bench_$tag.err" \
    | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag', round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), 'images/s', d['cuda_graphs'])" \
    || tail -n 5 "gpurun_out/r2_bench_$tag.err"
}
for s in "${steps[@]}"; do
  case $s in
    probe)    run probe 90 python tests/halo_probe.py ;;
    2cta)     run 2cta 300 env SG_TEST_2CTA=1 python -m pytest tests/test_gpu_conv_2cta.py -q -s ;;
    persist)  run persist 1000 env SG_TEST_PERSIST=1 python -m pytest tests/test_gpu_conv_persist.py -q -s ;;
    napfused) run napfused 600 env SG_TEST_NAP_FUSED=1 python -m pytest tests/test_gpu_nap_fused.py -q -s ;;
    parfwd)   run parfwd 600 env SG_PARALLEL_FWD=1 python -m pytest tests/test_gpu_graph_step.py tests/test_gpu_train_step.py -q -k "graph or four_step or replayed or captured" ;;
    vgg)      run vgg 300 env SG_TEST_VGG=1 python -m pytest tests/test_gpu_vgg.py -q ;;
    convab)   run convab 900 python tools/conv_ab.py --configs base SG_CONV_PERSIST=1 SG_CONV_2CTA=1 SG_CONV_PERSIST=1,SG_CONV_2CTA=1 ;;
    bench)
      bench_line base SG_NOOP=1
      bench_line persist SG_CONV_PERSIST=1
      bench_line wpersist SG_WGRAD_PERSIST=1
      bench_line napfused SG_NAP_FUSED=1
      bench_line parfwd SG_PARALLEL_FWD=1
      bench_line 2cta SG_CONV_2CTA=1
      bench_line all SG_CONV_PERSIST=1 SG_WGRAD_PERSIST=1 SG_NAP_FUSED=1 SG_PARALLEL_FWD=1 ;;
    *) echo "unknown step $s" ;;
  esac
done
