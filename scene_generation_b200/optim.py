"""Adam for the masters of libsg_b200 networks (reference: torch.optim.Adam at trainer.py:60,80,106,133).

PackedAdam keeps torch.optim.Adam's constructor, param_groups and state layout ('step', 'exp_avg', 'exp_avg_sq' —
checkpoints written by the reference's Trainer.save_checkpoint load unchanged) but runs ONE hand-written
multi-tensor kernel (csrc/adam.cu) that, in the same pass, rewrites the bf16 tensor-core operand of every weight
it updates.  The per-step re-packing of 197.6 M masters (2 kernels per weight) disappears, and the operand cache
of functional.py never goes stale behind a fused update that does not bump Tensor._version."""
import ctypes

import torch

from . import _lib
from . import functional as Fn
from .ops import _stream


def _same_layout(a, b):
    """same element -> memory offset map (strides of size-1 dimensions are meaningless)"""
    return a.shape == b.shape and all(sa == sb or n == 1 for sa, sb, n in zip(a.stride(), b.stride(), a.shape))


class PackedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        # fused + capturable: the parent then keeps 'step' as an f32 device tensor (what the kernel reads) and its
        # state_dict / load_state_dict handle that placement
        super().__init__(params, lr=lr, betas=betas, eps=eps, fused=True, capturable=True)

    def load_state_dict(self, state_dict):
        """torch decides where 'step' lives from the flags stored IN the checkpoint: one written by the reference's
        plain Adam (trainer.py:183-203) leaves it on the CPU and switches the groups back to fused=False.  Put the
        counters on the device (the kernel reads them there) and keep this optimizer's own flags."""
        super().load_state_dict(state_dict)
        for group in self.param_groups:
            group['fused'], group['capturable'], group['foreach'] = True, True, None
            for p in group['params']:
                st = self.state.get(p)
                if st and 'step' in st:
                    st['step'] = torch.as_tensor(st['step'], dtype=torch.float32).to(p.device).reshape(())

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            if group['weight_decay'] != 0 or group['amsgrad'] or group['maximize']:
                raise NotImplementedError('PackedAdam implements the reference configuration: plain Adam')
            ps, gs, ms, vs, steps, wks, Cs, Cps, refreshed = [], [], [], [], [], [], [], [], []
            for p in group['params']:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse or p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError('PackedAdam needs dense f32 CUDA parameters')
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                m, v = st['exp_avg'], st['exp_avg_sq']
                if not _same_layout(g, p) or g.dtype != torch.float32:
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                if not (_same_layout(m, p) and _same_layout(v, p)):
                    raise RuntimeError('PackedAdam: optimizer state does not have the layout of its parameter')
                ps.append(p); gs.append(g); ms.append(m); vs.append(v); steps.append(st['step'])
                ent = Fn.operand_entry(p)
                if ent is not None and ent[2] is None:          # (version, wk, wt, weight, ...): only the fprop copy exists
                    wk = ent[1]
                    wks.append(wk.data_ptr()); Cs.append(ent[5]); Cps.append(wk.shape[-1])
                    refreshed.append(p)
                else:
                    if ent is not None:
                        Fn.invalidate_packed((p,))              # a transposed copy exists too: re-pack both next time
                    wks.append(0); Cs.append(1); Cps.append(1)
            n = len(ps)
            if n == 0:
                continue
            torch._foreach_add_(steps, 1.0)
            arr = lambda vals: (ctypes.c_void_p * n)(*vals)
            _lib.call('sg_adam_pack', n, arr([t.data_ptr() for t in ps]), arr([t.data_ptr() for t in gs]),
                      arr([t.data_ptr() for t in ms]), arr([t.data_ptr() for t in vs]), arr(wks),
                      arr([t.data_ptr() for t in steps]), (ctypes.c_longlong * n)(*[t.numel() for t in ps]),
                      (ctypes.c_int * n)(*Cs), (ctypes.c_int * n)(*Cps), float(group['lr']), float(group['betas'][0]),
                      float(group['betas'][1]), float(group['eps']), _stream())     # doubles across the ABI, like torch
            Fn.mark_operands_maintained(refreshed)
        return loss
