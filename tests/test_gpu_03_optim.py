"""optim.PackedAdam (csrc/adam.cu) vs torch.optim.Adam (the reference's optimizer, trainer.py:60) on the same
gradients: parameters and both moments after several steps, and the bf16 operands it keeps in step with the
masters (must equal a fresh pack of the updated master bit for bit)."""
import pytest
import torch
import torch.nn as nn

from scene_generation_b200 import functional as Fn
from scene_generation_b200.optim import PackedAdam

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    ps = {
        'linear': (r(300, 454) * 0.05, 's1'),                                   # operand rows of 454 at pitch 456 (scalar path)
        'conv_cl': ((r(64, 3, 3, 32) * 0.02).permute(0, 3, 1, 2), 's1'),         # stored [Cout][kh][kw][Cin]: float4 path
        'convT_cl': ((r(48, 3, 3, 96) * 0.02).permute(3, 0, 1, 2), 'T'),         # ConvTranspose2d weight (Cin, Cout, kh, kw)
        'big': ((r(256, 3, 3, 256) * 0.02).permute(0, 3, 1, 2), 's1'),           # 590 k elements: 18 chunks
        'cin1': ((r(64, 3, 3, 1) * 0.1).permute(0, 3, 1, 2), 's1'),              # Cin = 1 -> pitch 8
        'bias': (r(64) * 0.1, None),
        'odd': (r(70001) * 0.1, None),                                           # numel % 4 != 0, 3 chunks
        'nograd': (r(16, 16), None),
    }
    return ps


def test_packed_adam_matches_torch_adam_and_keeps_operands_current():
    spec = _params(1)
    mine = {k: nn.Parameter(v.clone().to(DEV)) for k, (v, _) in spec.items()}
    ref = {k: nn.Parameter(v.clone().to(DEV)) for k, (v, _) in spec.items()}
    for k in mine:      # .to() keeps the permuted (channels-last) storage order
        assert mine[k].stride() == spec[k][0].stride()
    opt = PackedAdam(list(mine.values()), lr=1e-4, betas=(0.5, 0.999))
    opt_ref = torch.optim.Adam(list(ref.values()), lr=1e-4, betas=(0.5, 0.999))
    kinds = {k: kind for k, (_, kind) in spec.items() if kind}
    wk0 = {k: Fn.packed_weights(mine[k], kind)[0] for k, kind in kinds.items()}     # first pack: the operands exist
    g = torch.Generator().manual_seed(7)
    for step in range(5):
        for k in mine:
            if k == 'nograd' or (k == 'bias' and step == 2):       # a parameter that skips a step keeps its own counter
                mine[k].grad = ref[k].grad = None
                continue
            gr = (torch.randn(spec[k][0].shape, generator=g) * (10.0 ** (step - 2))).to(DEV)
            gr = torch.empty_like(mine[k]).copy_(gr)               # same layout as the parameter
            mine[k].grad, ref[k].grad = gr, gr.clone()
        opt.step()
        opt_ref.step()
        for k in mine:
            assert torch.allclose(mine[k], ref[k], rtol=0, atol=3e-7), (step, k, (mine[k] - ref[k]).abs().max().item())
            if k == 'nograd':
                assert len(opt.state[mine[k]]) == 0
                continue
            sm, sr = opt.state[mine[k]], opt_ref.state[ref[k]]
            assert float(sm['step']) == float(sr['step']), (step, k)
            for key in ('exp_avg', 'exp_avg_sq'):      # fp32 rounding of one fused multiply-add, relative to the tensor's scale
                tol = 2e-6 * sr[key].abs().max().item()
                assert torch.allclose(sm[key], sr[key], rtol=1e-5, atol=tol), (step, k, key, (sm[key] - sr[key]).abs().max().item())
        for k, kind in kinds.items():
            wk, _ = Fn.packed_weights(mine[k], kind)
            assert wk is wk0[k], 'the operand was re-packed instead of being kept current by the optimizer'
            m3 = Fn.master3(mine[k].detach(), kind)
            C = m3.shape[2]
            assert torch.equal(wk[..., :C], m3.to(torch.bfloat16)), (step, k)
            assert (wk[..., C:] == 0).all(), (step, k)
    # a foreign in-place update moves the version counter: the operand is re-packed — into the SAME buffer (captured
    # CUDA graphs hold its address)
    with torch.no_grad():
        mine['linear'].mul_(0.5)
    wk, _ = Fn.packed_weights(mine['linear'], 's1')
    assert wk is wk0['linear']
    assert torch.equal(wk[..., :454], Fn.master3(mine['linear'].detach(), 's1').to(torch.bfloat16))
    # and the optimizer keeps maintaining it afterwards
    gr = torch.randn(300, 454, device=DEV)
    for k in mine:
        mine[k].grad = None
    mine['linear'].grad = gr
    opt.step()
    assert Fn.packed_weights(mine['linear'], 's1')[0] is wk0['linear']
    assert torch.equal(wk[..., :454], Fn.master3(mine['linear'].detach(), 's1').to(torch.bfloat16))


def test_packed_adam_state_dict_round_trip_with_torch_adam():
    """checkpoints: the state layout is torch.optim.Adam's (trainer.py:183-203 saves optimizer.state_dict())"""
    w = nn.Parameter(torch.randn(32, 16, device=DEV))
    w2 = nn.Parameter(w.detach().clone())
    a, b = PackedAdam([w], lr=1e-3, betas=(0.5, 0.999)), torch.optim.Adam([w2], lr=1e-3, betas=(0.5, 0.999))
    for _ in range(3):
        g = torch.randn(32, 16, device=DEV)
        w.grad, w2.grad = g, g.clone()
        a.step()
        b.step()
    sd = a.state_dict()
    assert set(sd['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
    c = torch.optim.Adam([w2], lr=1e-3, betas=(0.5, 0.999), capturable=True)
    c.load_state_dict(sd)                                    # a torch Adam can resume from it ...
    d = PackedAdam([w], lr=1e-3, betas=(0.5, 0.999))
    d.load_state_dict(b.state_dict())                        # ... and the other way round
    g = torch.randn(32, 16, device=DEV)
    w.grad, w2.grad = g, g.clone()
    d.step()
    c.step()
    assert torch.allclose(w, w2, rtol=0, atol=3e-7)
