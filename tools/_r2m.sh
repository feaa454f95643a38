timeout 600 python -m pytest tests/test_gpu_07_train_step.py -m gpu -q -x -p no:cacheprovider -k four_step 2>&1 | grep -E "^E|assert|passed|failed" | head -20
python - <<'PY'
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
y = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
for name, fn, nbytes in (('memset 1 GiB', lambda: x.zero_(), 1 << 30), ('fill_(3) 1 GiB', lambda: x.fill_(3), 1 << 30),
                         ('copy 1 GiB (read + write)', lambda: y.copy_(x), 2 << 30)):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print('%-28s %.3f ms  %.0f GB/s' % (name, ms, nbytes / ms / 1e6))
PY
for c in cfg2 cfg4; do timeout 200 python tools/hbm_kernels.py --config $c 2>&1 | grep -E "layout_fwd" ; done
