"""Scene-graph convolution (mirror of scene_generation/graph.py; kernels: sg_gconv_* + tcgen05 GEMMs)."""
import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .layers import build_mlp


def _init_weights(module):
    """graph.py:27-30."""
    if hasattr(module, 'weight') and isinstance(module, nn.Linear):
        nn.init.kaiming_normal_(module.weight)


class GraphIndex:
    """Per-batch incidence structure: the CSR of (triple, role) uses per object, built once on the
    host from the int64 edge list and reused by every layer's pooled scatter and its adjoints."""

    def __init__(self, edges, num_objs):
        self.edges = edges.contiguous()
        ptr, src = ops.build_incidence_csr(edges.detach().cpu().numpy(), num_objs)
        self.seg_ptr = torch.from_numpy(ptr).to(edges.device)
        self.seg_src = torch.from_numpy(src).to(edges.device)
        self.O = num_objs

    @classmethod
    def from_host(cls, edges, seg_ptr, seg_src, num_objs):
        """edges / CSR tensors already on the device (built by the loader from host data)."""
        self = cls.__new__(cls)
        self.edges, self.seg_ptr, self.seg_src, self.O = edges.contiguous(), seg_ptr, seg_src, num_objs
        return self


class GraphTripleConv(nn.Module):
    """graph.py:33-122: gather [s,p,o] -> net1 -> split -> per-object average -> net2."""

    def __init__(self, input_dim, attributes_dim=0, output_dim=None, hidden_dim=512, pooling='avg',
                 mlp_normalization='none'):
        super().__init__()
        output_dim = input_dim if output_dim is None else output_dim
        self.input_dim, self.output_dim, self.hidden_dim = input_dim, output_dim, hidden_dim
        assert pooling in ['sum', 'avg'], 'Invalid pooling "%s"' % pooling
        self.pooling = pooling
        self.net1 = build_mlp([3 * input_dim + 2 * attributes_dim, hidden_dim, 2 * hidden_dim + output_dim],
                              batch_norm=mlp_normalization)
        self.net1.apply(_init_weights)
        self.net2 = build_mlp([hidden_dim, hidden_dim, output_dim], batch_norm=mlp_normalization)
        self.net2.apply(_init_weights)

    def forward(self, obj_vecs, pred_vecs, edges, index=None):
        """edges: LongTensor (T,2); index: optional prebuilt GraphIndex for this batch."""
        if index is None:
            index = GraphIndex(edges, obj_vecs.size(0))
        H, Dout = self.hidden_dim, self.output_dim
        # the gather writes the bf16 operand of net1's first GEMM directly (columns padded to a multiple of 8 with zeros)
        cur_t = Fn.GatherFn.apply(obj_vecs, pred_vecs, index.edges, index.seg_ptr, index.seg_src, True)
        new_t = self.net1(cur_t)
        pooled, new_p = Fn.PoolFn.apply(new_t, index.edges, index.seg_ptr, index.seg_src, index.O, H, Dout,
                                        self.pooling == 'avg')
        return self.net2(pooled), new_p


class GraphTripleConvNet(nn.Module):
    """graph.py:125-147."""

    def __init__(self, input_dim, num_layers=5, hidden_dim=512, pooling='avg', mlp_normalization='none'):
        super().__init__()
        self.num_layers = num_layers
        self.gconvs = nn.ModuleList([GraphTripleConv(input_dim=input_dim, hidden_dim=hidden_dim, pooling=pooling,
                                                     mlp_normalization=mlp_normalization) for _ in range(num_layers)])

    def forward(self, obj_vecs, pred_vecs, edges, index=None):
        if index is None:
            index = GraphIndex(edges, obj_vecs.size(0))
        for gconv in self.gconvs:
            obj_vecs, pred_vecs = gconv(obj_vecs, pred_vecs, edges, index)
        return obj_vecs, pred_vecs
