"""Scene-graph -> image model (mirror of scene_generation/model.py: same constructor, same forward
signature and return tuple, same state_dict keys; all tensor math on libsg_b200 kernels)."""
import os

import torch
import torch.nn as nn

from .bilinear import crop_bbox_batch
from .generators import AppearanceEncoder, define_G, mask_net
from .graph import GraphIndex, GraphTripleConv, GraphTripleConvNet
from .layers import build_mlp
from .layout import COMPACT_CC, masks_to_layout
from .utils import VectorPool


class Model(nn.Module):
    """model.py:12-124.  Extra keyword-only options (defaults keep the reference behaviour of the
    in-container oracle): ``layout_dtype`` 'bf16' (channels-last operand, default) or 'f32' (reference
    NCHW tensors), ``align_corners`` for the three grid_sample call sites."""

    def __init__(self, vocab, image_size=(64, 64), embedding_dim=128, gconv_dim=128, gconv_hidden_dim=512,
                 gconv_pooling='avg', gconv_num_layers=5, mask_size=32, mlp_normalization='none',
                 appearance_normalization='', activation='', n_downsample_global=4, box_dim=128,
                 use_attributes=False, box_noise_dim=64, mask_noise_dim=64, pool_size=100, rep_size=32,
                 layout_dtype='bf16', align_corners=False):
        super().__init__()
        self.vocab = vocab
        self.image_size = image_size
        self.use_attributes = use_attributes
        self.box_noise_dim = box_noise_dim
        self.mask_noise_dim = mask_noise_dim
        self.object_size = 64
        self.fake_pool = None        # built below, once the vocabulary size is known
        self.pool_plan = None        # device index vector of VectorPool.plan() when the host half ran outside (graph replay)
        self.graph_branch_stream = None   # set by the Trainer (SG_PARALLEL_FWD=1): second branch of captured iterations
        self.noise_generator = None       # None: torch's default CUDA generator (model.py:149 uses the global RNG)
        self.layout_dtype = layout_dtype
        self.align_corners = align_corners

        self.num_objs = len(vocab['object_to_idx'])
        self.fake_pool = VectorPool(pool_size, max_rows=self.num_objs * pool_size)
        self.num_preds = len(vocab['pred_idx_to_name'])
        self.obj_embeddings = nn.Embedding(self.num_objs, embedding_dim)
        self.pred_embeddings = nn.Embedding(self.num_preds, embedding_dim)
        attributes_dim = vocab['num_attributes'] if use_attributes else 0
        if gconv_num_layers == 0:
            self.gconv = nn.Linear(embedding_dim, gconv_dim)
        else:
            self.gconv = GraphTripleConv(input_dim=embedding_dim, attributes_dim=attributes_dim, output_dim=gconv_dim,
                                         hidden_dim=gconv_hidden_dim, pooling=gconv_pooling,
                                         mlp_normalization=mlp_normalization)
        self.gconv_net = None
        if gconv_num_layers > 1:
            self.gconv_net = GraphTripleConvNet(input_dim=gconv_dim, hidden_dim=gconv_hidden_dim, pooling=gconv_pooling,
                                                num_layers=gconv_num_layers - 1, mlp_normalization=mlp_normalization)
        self.box_dim = box_dim
        self.box_net = build_mlp([box_dim, gconv_hidden_dim, 4], batch_norm=mlp_normalization)
        self.g_mask_dim = gconv_dim + mask_noise_dim
        self.mask_net = mask_net(self.g_mask_dim, mask_size)
        self.repr_input = self.g_mask_dim
        self.repr_net = build_mlp([self.repr_input, 64, rep_size], batch_norm=mlp_normalization)
        self.image_encoder = AppearanceEncoder(vocab=vocab, arch='C4-64-2,C4-128-2,C4-256-2',
                                               normalization=appearance_normalization, activation=activation,
                                               padding='valid', vecs_size=self.g_mask_dim)
        self.layout_to_image = define_G(self.num_objs + rep_size, 3, 64, n_downsample_global, 9, 'instance')
        # Channel-compacted layouts (csrc/compact.cu): Cc = 64 channels per image = S class slots + the
        # appearance vector + 3 spare channels for the discriminator's image slot.
        self.rep_size = rep_size
        self.compact_slots = (COMPACT_CC - 3 - rep_size) // 8 * 8
        self.compact_layout = os.environ.get('SG_LAYOUT_COMPACT', '1') != '0'

    def forward(self, gt_imgs, objs, triples, obj_to_img, boxes_gt=None, masks_gt=None, attributes=None,
                test_mode=False, use_gt_box=False, features=None):
        from . import ops
        ops.refresh_stream()
        O = objs.size(0)
        if self.graph_branch_stream is not None and not test_mode and features is None and self.training \
                and torch.cuda.is_current_stream_capturing():
            return self._forward_train_two_branches(gt_imgs, objs, triples, obj_to_img, boxes_gt, masks_gt, attributes)
        obj_vecs, pred_vecs = self.scene_graph_to_vectors(objs, triples, attributes)
        N = gt_imgs.size(0) if gt_imgs is not None else None
        plan = None if test_mode else self._compact_plan(objs, N)
        box_vecs, mask_vecs, scene_layout_vecs, wrong_layout_vecs = \
            self.create_components_vecs(gt_imgs, boxes_gt, obj_to_img, objs, obj_vecs, features, plan)
        boxes_pred = self.box_net(box_vecs)
        masks_pred = self.mask_net(mask_vecs, fused_sigmoid=True).squeeze(1)        # model.py:106-107
        H, W = self.image_size
        lay = dict(align_corners=self.align_corners, nhwc_bf16=self.layout_dtype == 'bf16', N=N,
                   cmap=None if plan is None else plan[1])
        if test_mode:
            boxes = boxes_gt if use_gt_box else boxes_pred
            masks = masks_gt if masks_gt is not None else masks_pred
            pred_layout = masks_to_layout(scene_layout_vecs, boxes, masks, obj_to_img, H, W, test_mode=True, **lay)
            return self.layout_to_image(pred_layout), boxes_pred, masks_pred, None, pred_layout, None
        # only the appearance channels of layout_vecs carry a gradient (one_hot_obj is a constant, model.py:165-168)
        c0 = self.num_objs if plan is None else self.compact_slots
        gt_layout = masks_to_layout(scene_layout_vecs, boxes_gt, masks_gt, obj_to_img, H, W,
                                    grad_channels=(c0, scene_layout_vecs.shape[1]), **lay)
        if self.layout_dtype == 'bf16':
            gt_layout._sg_grad_channels = (c0, scene_layout_vecs.shape[1])
        pred_layout = masks_to_layout(scene_layout_vecs, boxes_gt, masks_pred, obj_to_img, H, W, **lay)
        wrong_layout = masks_to_layout(wrong_layout_vecs, boxes_gt, masks_gt, obj_to_img, H, W, **lay)
        imgs_pred = self.layout_to_image(gt_layout)
        return imgs_pred, boxes_pred, masks_pred, gt_layout, pred_layout, wrong_layout

    def _forward_train_two_branches(self, gt_imgs, objs, triples, obj_to_img, boxes_gt, masks_gt, attributes):
        """EXPERIMENTAL (SG_PARALLEL_FWD=1, captured iterations only; not yet run on hardware).  In training the generator
        sees the GROUND-TRUTH boxes / masks and the appearance vectors of the image crops (model.py:119-121,158-170):
        nothing on that path depends on the graph network.  The chain gconv -> box_net / mask_net (~100 tiny GEMMs and five
        192-channel convs, forward and — since autograd replays each node on its forward stream — backward) therefore
        runs as a second branch of the captured graph next to crops -> encoder -> layouts -> generator.  Same arithmetic
        as forward(); the VectorPool / noise draws happen in the same order."""
        from . import ops
        N = gt_imgs.size(0)
        O = objs.size(0)
        H, W = self.image_size
        plan = self._compact_plan(objs, N)
        main, side = torch.cuda.current_stream(), self.graph_branch_stream
        side.wait_stream(main)
        with ops.on_stream(side):
            obj_vecs, pred_vecs = self.scene_graph_to_vectors(objs, triples, attributes)
            layout_noise = torch.randn((1, self.mask_noise_dim), dtype=obj_vecs.dtype, device=obj_vecs.device,
                                   generator=self.noise_generator).repeat((O, 1))
            mask_vecs = torch.cat([obj_vecs, layout_noise], dim=1)
            boxes_pred = self.box_net(obj_vecs)
            masks_pred = self.mask_net(mask_vecs, fused_sigmoid=True).squeeze(1)
        # main branch: appearance vectors of the crops -> layout vectors -> layouts -> generator
        crops = crop_bbox_batch(gt_imgs, boxes_gt, obj_to_img, self.object_size, align_corners=self.align_corners, operand=True)
        obj_repr = self.repr_net(self.image_encoder(crops))
        if plan is None:
            one_hot_obj = torch.zeros((O, self.num_objs), dtype=obj_repr.dtype, device=obj_repr.device)
            one_hot_obj = one_hot_obj.scatter_(1, objs.view(-1, 1).long(), 1.0)
        else:
            one_hot_obj = torch.zeros((O, self.compact_slots), dtype=obj_repr.dtype, device=obj_repr.device)
            one_hot_obj = one_hot_obj.scatter_(1, plan[0].view(-1, 1), 1.0)
        scene_layout_vecs = torch.cat([one_hot_obj, obj_repr], dim=1)
        wrong_objs_rep = self.fake_pool.query(objs, obj_repr, planned=self.pool_plan)
        wrong_layout_vecs = torch.cat([one_hot_obj, wrong_objs_rep], dim=1)
        lay = dict(align_corners=self.align_corners, nhwc_bf16=self.layout_dtype == 'bf16', N=N,
                   cmap=None if plan is None else plan[1])
        c0 = self.num_objs if plan is None else self.compact_slots
        gt_layout = masks_to_layout(scene_layout_vecs, boxes_gt, masks_gt, obj_to_img, H, W,
                                    grad_channels=(c0, scene_layout_vecs.shape[1]), **lay)
        if self.layout_dtype == 'bf16':
            gt_layout._sg_grad_channels = (c0, scene_layout_vecs.shape[1])
        wrong_layout = masks_to_layout(wrong_layout_vecs, boxes_gt, masks_gt, obj_to_img, H, W, **lay)
        imgs_pred = self.layout_to_image(gt_layout)
        main.wait_stream(side)
        ops.refresh_stream()
        pred_layout = masks_to_layout(scene_layout_vecs, boxes_gt, masks_pred, obj_to_img, H, W, **lay)
        return imgs_pred, boxes_pred, masks_pred, gt_layout, pred_layout, wrong_layout

    def scene_graph_to_vectors(self, objs, triples, attributes):
        """model.py:126-143."""
        s, p, o = triples[:, 0], triples[:, 1], triples[:, 2]
        edges = torch.stack([s, o], dim=1)
        obj_vecs = self.obj_embeddings(objs)
        pred_vecs = self.pred_embeddings(p)
        if self.use_attributes:
            obj_vecs = torch.cat([obj_vecs, attributes], dim=1)
        if isinstance(self.gconv, nn.Linear):
            from . import functional as Fn
            obj_vecs = Fn.linear(obj_vecs, self.gconv.weight, self.gconv.bias)
        else:
            meta = getattr(triples, '_sg_csr', None)      # (seg_ptr, seg_src) attached by the loader
            index = GraphIndex.from_host(edges, meta[0], meta[1], objs.size(0)) if meta is not None else \
                GraphIndex(edges, objs.size(0))
            obj_vecs, pred_vecs = self.gconv(obj_vecs, pred_vecs, edges, index)
            if self.gconv_net is not None:
                obj_vecs, pred_vecs = self.gconv_net(obj_vecs, pred_vecs, edges, index)
        return obj_vecs, pred_vecs

    def _compact_plan(self, objs, N):
        """(obj_slot (O,) int64, cmap (N, 64) int32) of the channel-compacted layouts, from the class-slot
        table the loader attached to ``objs`` (synthetic.HostMeta); None -> dense layouts."""
        meta = getattr(objs, '_sg_compact', None)
        S = self.compact_slots
        if meta is None or not self.compact_layout or self.layout_dtype != 'bf16' or S < 8 or N is None \
                or min(self.image_size) < 64:
            return None
        obj_slot, slot_cls, used = meta
        if used > S or slot_cls.shape[0] != N:
            return None
        tail = getattr(self, '_cmap_tail', None)
        if tail is None or tail.device != slot_cls.device:
            D, A = self.num_objs + self.rep_size, self.rep_size
            t = [self.num_objs + a for a in range(A)] + [D, D + 1, D + 2]      # appearance, then the image slot
            t += [-1] * (COMPACT_CC - S - len(t))
            tail = self._cmap_tail = torch.tensor(t, dtype=torch.int32, device=slot_cls.device)
        cmap = torch.cat([slot_cls[:, :S], tail.unsqueeze(0).expand(N, -1)], dim=1).contiguous()
        return obj_slot, cmap

    def create_components_vecs(self, imgs, boxes, obj_to_img, objs, obj_vecs, features, plan=None):
        """model.py:145-172.  With a compact plan the one-hot part of the layout vectors indexes the image's
        class slots instead of the vocabulary."""
        O = objs.size(0)
        layout_noise = torch.randn((1, self.mask_noise_dim), dtype=obj_vecs.dtype, device=obj_vecs.device,
                                   generator=self.noise_generator).repeat((O, 1))
        mask_vecs = torch.cat([obj_vecs, layout_noise], dim=1)
        if features is None:
            crops = crop_bbox_batch(imgs, boxes, obj_to_img, self.object_size, align_corners=self.align_corners, operand=True)
            obj_repr = self.repr_net(self.image_encoder(crops))
        else:
            obj_repr = self.repr_net(mask_vecs)
            rows = [i for i, f in enumerate(features) if f is not None]
            if rows:
                obj_repr = obj_repr.clone()
                obj_repr[rows] = torch.stack([features[i].to(obj_repr) for i in rows])
        if plan is None:
            one_hot_obj = torch.zeros((O, self.num_objs), dtype=obj_repr.dtype, device=obj_repr.device)
            one_hot_obj = one_hot_obj.scatter_(1, objs.view(-1, 1).long(), 1.0)
        else:
            one_hot_obj = torch.zeros((O, self.compact_slots), dtype=obj_repr.dtype, device=obj_repr.device)
            one_hot_obj = one_hot_obj.scatter_(1, plan[0].view(-1, 1), 1.0)
        layout_vecs = torch.cat([one_hot_obj, obj_repr], dim=1)
        wrong_objs_rep = self.fake_pool.query(objs, obj_repr, planned=self.pool_plan)
        wrong_layout_vecs = torch.cat([one_hot_obj, wrong_objs_rep], dim=1)
        return obj_vecs, mask_vecs, layout_vecs, wrong_layout_vecs
