#!/usr/bin/env bash
# single-GPU bench lines of a round: headline (cfg2, with CPU reference arm), cfg4, cfg5, VGG variant, reference on the GPU
tag=${1:-r02}
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
timeout 400 python bench.py --impl reference-gpu --steps 5 --warmup 3 > gpurun_out/${tag}_bench_reference_gpu.json 2> gpurun_out/${tag}_bench_reference_gpu.err; echo "reference-gpu rc=$?"
timeout 500 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.err; echo "cfg4 rc=$?"
timeout 500 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/${tag}_bench_cfg5.json 2> gpurun_out/${tag}_bench_cfg5.err; echo "cfg5 rc=$?"
timeout 500 python bench.py --vgg --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/${tag}_bench_vgg.json 2> gpurun_out/${tag}_bench_vgg.err; echo "vgg rc=$?"
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for name in ('cfg2', 'reference', 'reference_gpu', 'cfg4', 'cfg5', 'vgg'):
    try:
        d = [json.loads(l) for l in open('gpurun_out/%s_bench_%s.json' % (tag, name)) if l.startswith('{')][-1]
        print(name, {k: d.get(k) for k in ('value', 'ms_per_step', 'unavailable')}, (d.get('e2e') or {}).get('value'),
              (d.get('roofline') or {}).get('frac'), d.get('cpu_baseline'), d.get('dropin_eager'))
    except Exception as e:
        print(name, 'no line', e)
PY
