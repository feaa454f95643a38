timeout 600 python -m pytest tests/test_gpu_00_ops.py tests/test_gpu_04_modules.py tests/test_gpu_06_compact.py tests/test_gpu_07_train_step.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for c in cfg2 cfg4 cfg5; do timeout 200 python tools/hbm_kernels.py --config $c 2>&1 | grep -E "layout_fwd" ; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-dropin 2>gpurun_out/r2l.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['cuda_graphs'], d['kernels']['layout_fwd'], d['kernels']['layout_fwd_dense'])"
