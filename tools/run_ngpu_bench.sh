#!/usr/bin/env bash
# bench.py at the GPU counts given as arguments (one box): tools/run_ngpu_bench.sh 8 4
# gpurun --gpus 8 --timeout 900 -- 'bash tools/run_ngpu_bench.sh 8 4'
mkdir -p gpurun_out
for n in "$@"; do
  echo "=== N=$n ($(date +%T))"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29520 + n)) \
    bench.py --gpus "$n" --steps 20 --warmup 5 > "gpurun_out/scale_n$n.json" 2> "gpurun_out/scale_n$n.err"
  echo "rc=$?"
  grep -E "captured|timed" "gpurun_out/scale_n$n.err" | tr '\n' ' ' | cut -c 1-600; echo
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
for l in open('gpurun_out/scale_n%s.json' % n):
    if l.startswith('{'):
        d = json.loads(l)
        print('N=%s ms/step %.3f images/s %.1f e2e %.1f graphs %s' % (n, d['ms_per_step'], d['value'], d['e2e']['value'], d['cuda_graphs']))
PY
done
