"""Training step (mirror of scene_generation/trainer.py: same Trainer(args, vocab, checkpoint)
constructor, same train_* methods and checkpoint keys).  Differences, all outside the arithmetic:
loss terms stay on the device (no per-term .item()), the image discriminator takes (layout, image)
pairs instead of a materialised concat, and with torch.distributed initialised every optimizer step
is preceded by a flat-buffer NCCL all-reduce of that network's gradients (DDP semantics).

train_step() launches the ~1900 kernels of one iteration either eagerly (host-bound: ~31 ms of Python/launch
time per step against ~25 ms of GPU time) or — default — as FOUR captured CUDA graphs per batch geometry
(number of objects / triples, the use_gt coin): A = Model.forward + generator losses + backward, B = the three
discriminator forward/backward passes, Cg / Cd = Adam of the generator / of the discriminators.  The first step of
a geometry runs eagerly, the second is captured, later ones are replayed after copying the batch into the graphs'
static input buffers.  The gradient all-reduces of data-parallel runs are issued EAGERLY between the graphs (after
A: generator, after B: discriminators) in the same order on every rank whatever mix of eager / capturing /
replaying steps the ranks are in — no collective is ever captured, so ranks with different batch geometries cannot
deadlock — and overlap the next graph on NCCL's own stream.  The only per-step host work that survives is the
VectorPool replacement policy (python `random`, utils.py:62-90), whose index vector is an input of graph A."""
import collections
import os
import warnings

import torch
import torch.nn.functional as F

from . import ddp
from . import functional as Fn
from .discriminators import AcCropDiscriminator, define_D, define_mask_D
from .losses import GANLoss, get_gan_losses
from .model import Model
from .optim import PackedAdam
from .utils import LossManager

try:                                           # logging is optional plumbing (tensorboardX is not in the image)
    from tensorboardX import SummaryWriter
except Exception:                              # pragma: no cover
    SummaryWriter = None


def _adam(params, lr, betas):
    """Adam(lr, betas=(beta1, 0.999)) of trainer.py:60,80,106,133.  Default: the hand-written multi-tensor kernel that
    also refreshes the bf16 operands (optim.PackedAdam); SG_TORCH_ADAM=1: torch's fused Adam + per-step re-packing."""
    if os.environ.get('SG_TORCH_ADAM', '0') == '1':
        return torch.optim.Adam(params, lr=lr, betas=betas, fused=True, capturable=True)
    return PackedAdam(params, lr=lr, betas=betas)


class _BatchMeta:
    """The loader's host-side index structures of a batch (synthetic.HostMeta.attach tags the tensors with them)."""

    def __init__(self, ranges, seg_ptr, seg_src, obj_slot, slot_cls, slots_used, objs_host):
        self.ranges, self.seg_ptr, self.seg_src, self.obj_slot, self.slot_cls = ranges, seg_ptr, seg_src, obj_slot, slot_cls
        self.slots_used, self.objs_host = slots_used, objs_host

    @classmethod
    def of(cls, batch):
        objs, triples, obj_to_img = batch[1], batch[4], batch[5]
        ranges, csr = getattr(obj_to_img, '_sg_ranges', None), getattr(triples, '_sg_csr', None)
        compact, objs_host = getattr(objs, '_sg_compact', None), getattr(objs, '_sg_host', None)
        if ranges is None or csr is None or compact is None or objs_host is None:
            return None
        return cls(ranges, csr[0], csr[1], compact[0], compact[1], compact[2], objs_host)

    def tensors(self):
        return (self.ranges, self.seg_ptr, self.seg_src, self.obj_slot, self.slot_cls)

    def geometry(self):
        return tuple(tuple(t.shape) for t in self.tensors())

    def attach(self, batch, tensors):
        ranges, seg_ptr, seg_src, obj_slot, slot_cls = tensors
        batch[5]._sg_ranges = ranges
        batch[4]._sg_csr = (seg_ptr, seg_src)
        batch[1]._sg_host = self.objs_host
        batch[1]._sg_compact = (obj_slot, slot_cls, self.slots_used)
        return batch


def _batch_to_device(batch, device):
    """device copies of a host batch, metadata tags included (no-op for a batch already on the device)"""
    if all(t.is_cuda for t in batch):
        return batch
    meta = _BatchMeta.of(batch)
    out = tuple(t.to(device, non_blocking=True) for t in batch)
    if meta is not None:
        meta.attach(out, tuple(t.to(device, non_blocking=True) for t in meta.tensors()))
    return out


def _nvtx(name):
    """NVTX range (visible in Nsight Systems / ncu --nvtx); SG_NVTX=0 turns the ranges off"""
    if os.environ.get('SG_NVTX', '1') == '0':
        return _NullCtx()
    return torch.cuda.nvtx.range(name)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _detached(t):
    """t.detach() that keeps the loader / layout tags (channel map, gradient channels) of the tensor"""
    d = t.detach()
    for k, v in getattr(t, '__dict__', {}).items():
        if k.startswith('_sg_'):
            setattr(d, k, v)
    return d


class _StepGraph:
    """One captured training iteration: static input buffers (batch, index metadata, VectorPool plan), the four CUDA
    graphs (A, B, Cg, Cd), and the tensors they leave behind (outputs of Model.forward, the four LossManagers)."""

    def __init__(self, batch, meta, plan):
        dev = torch.device('cuda', torch.cuda.current_device())
        self.batch = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in batch)
        self.meta_tensors = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in meta.tensors())
        self.pool_idx = None if plan is None else torch.empty(plan.shape, dtype=plan.dtype, device=dev)
        self.slots_used = meta.slots_used
        self.graphs = None            # {'A','B','Cg','Cd'} -> torch.cuda.CUDAGraph
        self.out = self.losses = None
        self.launches = 0

    def load(self, batch, meta, plan):
        for dst, src in zip(self.batch + self.meta_tensors, tuple(batch) + meta.tensors()):
            dst.copy_(src, non_blocking=True)
        if plan is not None:
            # pinned staging block from torch's caching host allocator (it is not handed out again before this copy
            # has run), so the host does not wait for the previous iteration here
            self.pool_idx.copy_(plan.pin_memory(), non_blocking=True)
        # same compaction decision as at capture (part of the geometry key); the host class list feeds nothing in a replay
        meta_static = _BatchMeta(*self.meta_tensors, self.slots_used, meta.objs_host)
        meta_static.attach(self.batch, self.meta_tensors)


class Trainer:
    def __init__(self, args, vocab, checkpoint):
        self.vocab = vocab
        self.args = args
        self.num_obj = len(vocab['object_to_idx'])
        self.writer = SummaryWriter(args.output_dir) if SummaryWriter is not None else None
        self.gan_g_loss, self.gan_d_loss = get_gan_losses(args.gan_loss_type)
        self.init_generator(args, checkpoint)
        self.init_image_discriminator(args, checkpoint)
        self.init_obj_discriminator(args, checkpoint)
        self.init_mask_discriminator(args, checkpoint)
        # CUDA-graph replay of whole iterations (SG_CUDA_GRAPH=0 or args.cuda_graphs=False: eager launches)
        self.use_graphs = bool(getattr(args, 'cuda_graphs', True)) and os.environ.get('SG_CUDA_GRAPH', '1') != '0'
        # batch geometry -> 'warm' (seen once, ran eagerly) | _StepGraph, least recently used first; at most
        # SG_GRAPH_CACHE captured geometries stay alive (each pins its outputs in the shared pool)
        self._graphs = collections.OrderedDict()
        self._graph_cache_size = int(os.environ.get('SG_GRAPH_CACHE', getattr(args, 'graph_cache', 64)))
        self._graph_pool = None       # one private memory pool shared by all captured iterations (replayed one at a time)
        self._side_streams = None     # captured iterations run independent sub-steps as parallel graph branches
        self._branches = None         # the side streams while a capture is running, else None (eager: one stream)
        self._fwd_stream = None
        self._st = None               # tensors handed from one phase of the iteration to the next
        self.generator_losses = self.d_mask_losses = self.d_obj_losses = self.d_img_losses = None
        self.reducers = {}
        # data parallel (or args.flat_grads, which tests use on one GPU): flat gradient buffers per network
        if ddp.world_size() > 1 or getattr(args, 'flat_grads', False):
            for name, net in (('g', self.model), ('img', self.netD), ('obj', self.obj_discriminator),
                              ('mask', self.mask_discriminator)):
                if net is not None:
                    if ddp.world_size() > 1:
                        ddp.broadcast_parameters(net)
                    self.reducers[name] = ddp.FlatGradReducer(net)

    # ---- construction (trainer.py:30-134) ------------------------------------------------------
    def init_generator(self, args, checkpoint):
        if args.restore_from_checkpoint:
            model_kwargs = checkpoint['model_kwargs']
        else:
            keys = ('image_size', 'embedding_dim', 'gconv_dim', 'gconv_hidden_dim', 'gconv_num_layers',
                    'mlp_normalization', 'appearance_normalization', 'activation', 'mask_size', 'n_downsample_global',
                    'box_dim', 'use_attributes', 'box_noise_dim', 'mask_noise_dim', 'pool_size', 'rep_size')
            model_kwargs = dict(vocab=self.vocab, **{k: getattr(args, k) for k in keys})
            checkpoint['model_kwargs'] = model_kwargs
        extra = {k: getattr(args, k) for k in ('layout_dtype', 'align_corners') if hasattr(args, k)}
        self.model = model = Model(**model_kwargs, **extra).to('cuda')
        self.criterionVGG = None
        if getattr(args, 'vgg_features_weight', 0) > 0:
            # trainer.py:57.  The ImageNet weights cannot be downloaded here: args.vgg_weights names a torchvision
            # vgg19 state_dict file.  A seeded random VGG19 (same FLOPs, meaningless features) must be asked for.
            from .losses import VGGLoss
            weights = getattr(args, 'vgg_weights', None)
            if weights is None and not getattr(args, 'vgg_random_init', False):
                raise RuntimeError('vgg_features_weight = %g needs the pretrained VGG19 (--vgg_weights <torchvision vgg19 '
                                   'state_dict>); pass --vgg_features_weight 0 to train without the perceptual term or '
                                   '--vgg_random_init 1 for throughput runs' % args.vgg_features_weight)
            self.criterionVGG = VGGLoss(weights)
        self.criterionFeat = torch.nn.L1Loss()
        self.criterionGAN = GANLoss(use_lsgan=not args.no_lsgan)
        self.optimizer = _adam(model.parameters(), lr=args.learning_rate, betas=(args.beta1, 0.999))

    def init_obj_discriminator(self, args, checkpoint):
        self.obj_discriminator, self.optimizer_d_obj = None, None
        if args.d_obj_weight > 0:
            if args.restore_from_checkpoint:
                kw = checkpoint['d_obj_kwargs']
            else:
                kw = {'vocab': self.vocab, 'arch': args.d_obj_arch, 'normalization': args.d_normalization,
                      'activation': args.d_activation, 'padding': args.d_padding, 'object_size': args.crop_size}
                checkpoint['d_obj_kwargs'] = kw
            self.obj_discriminator = AcCropDiscriminator(**kw).to('cuda')
            self.obj_discriminator.align_corners = getattr(args, 'align_corners', False)
            self.obj_discriminator.train()
            self.optimizer_d_obj = _adam(self.obj_discriminator.parameters(), lr=args.learning_rate, betas=(args.beta1, 0.999))

    def init_mask_discriminator(self, args, checkpoint):
        self.mask_discriminator, self.optimizer_d_mask = None, None
        if args.d_mask_weight > 0:
            if args.restore_from_checkpoint:
                kw = checkpoint['d_mask_kwargs']
            else:
                kw = {'input_nc': 1, 'ndf': args.ndf_mask, 'n_layers_D': args.n_layers_D_mask, 'norm': args.norm_D_mask,
                      'use_sigmoid': args.no_lsgan, 'num_D': args.num_D_mask, 'num_objects': self.num_obj}
                checkpoint['d_mask_kwargs'] = kw
            self.mask_discriminator = define_mask_D(**kw).to('cuda')
            self.mask_discriminator.train()
            self.optimizer_d_mask = _adam(self.mask_discriminator.parameters(), lr=args.mask_learning_rate,
                                          betas=(args.beta1, 0.999))

    def init_image_discriminator(self, args, checkpoint):
        if args.d_img_weight == 0:
            self.netD, self.optimizer_d_img = None, None
            return
        if args.restore_from_checkpoint:
            kw = checkpoint['d_img_kwargs']
        else:
            kw = {'input_nc': self.num_obj + args.rep_size + args.output_nc, 'ndf': args.ndf, 'n_layers_D': args.n_layers_D,
                  'norm': args.norm_D, 'use_sigmoid': args.no_lsgan, 'num_D': args.num_D}
            checkpoint['d_img_kwargs'] = kw
        self.netD = define_D(**kw).to('cuda')
        self.netD.train()
        self.optimizer_d_img = _adam(list(self.netD.parameters()), lr=args.learning_rate, betas=(args.beta1, 0.999))

    # ---- checkpoint (trainer.py:136-203) -------------------------------------------------------
    def restore_checkpoint(self, checkpoint):
        self.release_graphs()         # captured Adam launches hold the addresses of the optimizer state they replace
        self.model.load_state_dict(checkpoint['model_state'])
        self.optimizer.load_state_dict(checkpoint['optim_state'])
        for net, opt, k in ((self.obj_discriminator, self.optimizer_d_obj, 'd_obj'),
                            (self.mask_discriminator, self.optimizer_d_mask, 'd_mask'),
                            (self.netD, self.optimizer_d_img, 'd_img')):
            if net is not None:
                net.load_state_dict(checkpoint[k + '_state'])
                opt.load_state_dict(checkpoint[k + '_optim_state'])

    def save_checkpoint(self, checkpoint, t, args, epoch, train_results=None, val_results=None):
        for net, opt, k in ((self.obj_discriminator, self.optimizer_d_obj, 'd_obj'),
                            (self.mask_discriminator, self.optimizer_d_mask, 'd_mask'),
                            (self.netD, self.optimizer_d_img, 'd_img')):
            if net is not None:
                checkpoint[k + '_state'] = net.state_dict()
                checkpoint[k + '_optim_state'] = opt.state_dict()
        checkpoint['model_state'] = self.model.state_dict()
        checkpoint['optim_state'] = self.optimizer.state_dict()
        checkpoint.setdefault('counters', {})
        checkpoint['counters']['t'] = t
        checkpoint['counters']['epoch'] = epoch
        path = os.path.join(args.output_dir, '%s_with_model.pt' % args.checkpoint_name)
        if ddp.rank() == 0:
            os.makedirs(args.output_dir, exist_ok=True)
            torch.save(checkpoint, path)
        return path

    # ---- the four sub-steps (trainer.py:205-325) -----------------------------------------------
    # Each train_* method computes its losses and gradients (zero_grad + backward); the optimizer updates are separate
    # (_update) so that data-parallel runs can all-reduce in between.  train_generator / train_*_discriminator keep the
    # reference's one-call semantics (backward AND update) unless a phased iteration is running (self._phased).
    _phased = False

    def _backward(self, name, optimizer, losses):
        from . import ops
        ops.refresh_stream()
        direct = name in self.reducers
        if direct:
            self.reducers[name].zero()            # .grad tensors are views of one flat buffer
            Fn.DIRECT_GRADS[0] = set()            # weight / bias gradients are written straight into those views
        else:
            # None gradients: backward then hands its fresh tensors over instead of adding them into zero-filled ones
            # (257 add kernels per iteration).  A parameter without a gradient in a step (box_net when use_gt is False)
            # is skipped by Adam, where the reference's pytorch-1.0 zero_grad — and the data-parallel flat buffers —
            # apply its momentum-only update; DESIGN.md lists the difference.
            optimizer.zero_grad(set_to_none=True)
        try:
            losses.total_loss.backward()
        finally:
            if direct:
                Fn.DIRECT_GRADS[0] = None
        # drop the autograd graph now: a loss kept for logging would keep this iteration's AccumulateGrad nodes (bound
        # to the stream they were created on) alive into the next iteration — fatal for a CUDA graph capture
        losses.total_loss = losses.total_loss.detach()
        if not self._phased:
            self._allreduce((name,))
            self._update(name, optimizer)

    def _allreduce(self, names, async_op=False):
        """data parallel: average the flat gradient buffers of these networks over the ranks (eager NCCL calls, issued in
        the same order on every rank).  async_op: returns the work handles; the caller waits before the update."""
        works = []
        for name in names:
            r = self.reducers.get(name)
            if r is not None:
                w = r.allreduce(async_op=async_op)
                if w is not None:
                    works.append(w)
        return works

    def _update(self, name, optimizer):
        """Adam step of one network"""
        from . import ops
        ops.refresh_stream()
        optimizer.step()
        if not isinstance(optimizer, PackedAdam):
            # torch's fused Adam does not bump Tensor._version: tell the bf16 operand cache which masters changed
            # (PackedAdam rewrites the operands itself)
            Fn.invalidate_packed(p for group in optimizer.param_groups for p in group['params'])

    def _one_hot(self, objs, like):
        oh = torch.zeros((objs.numel(), self.num_obj), dtype=like.dtype, device=like.device)
        return oh.scatter_(1, objs.view(-1, 1).long(), 1.0)

    def train_generator(self, imgs, imgs_pred, masks, masks_pred, layout, objs, boxes, boxes_pred, obj_to_img, use_gt):
        from . import ops
        args = self.args
        self.generator_losses = gl = LossManager()
        if use_gt:
            if args.l1_pixel_loss_weight > 0:
                gl.add_loss(F.l1_loss(imgs_pred, imgs), 'L1_pixel_loss', args.l1_pixel_loss_weight)
            gl.add_loss(F.mse_loss(boxes_pred, boxes), 'bbox_pred', args.bbox_pred_loss_weight)
        # The G step back-propagates THROUGH the discriminators; the gradients it would deposit in their
        # parameters (trainer.py:262) are cleared by every D step's zero_grad before use, so the D weights are
        # frozen for this graph and their wgrad kernels are skipped.
        d_nets = [n for n in (self.obj_discriminator, self.mask_discriminator, self.netD) if n is not None]
        for net in d_nets:
            for p in net.parameters():
                p.requires_grad_(False)
        # Inside a captured iteration the loss branches that only share imgs_pred / masks_pred — the VGG term, the object
        # discriminator, the mask discriminator — run on side streams next to the image discriminator (forward here,
        # backward through autograd's stream semantics); their loss terms join the total in the reference's order.
        br = self._branches
        main = torch.cuda.current_stream()
        if br:
            for st in br:
                st.wait_stream(main)

        def on(i):
            return ops.on_stream(br[i]) if br else _NullCtx()
        try:
            terms = {}
            if self.criterionVGG is not None:                       # trainer.py:218-221 (with and without use_gt)
                with on(2):
                    terms['g_vgg'] = self.criterionVGG(imgs_pred, imgs)
            with on(0):
                scores_fake, ac_loss, _ = self.obj_discriminator(imgs_pred, objs, boxes, obj_to_img)
                terms['ac_loss'] = ac_loss
                terms['g_gan_obj_loss'] = self.gan_g_loss(scores_fake)
            if self.mask_discriminator is not None:
                with on(1):
                    scores_fake = self.mask_discriminator(masks_pred.unsqueeze(1), objs)
                    terms['g_gan_mask_obj_loss'] = self.criterionGAN(scores_fake, True)
                    if args.d_mask_features_weight > 0:
                        with torch.no_grad():
                            scores_real = self.mask_discriminator(masks.unsqueeze(1), objs)
                        terms['g_mask_features_loss'] = self.calculate_features_loss(scores_fake, scores_real)
            if self.netD is not None:
                with torch.no_grad():       # only used detached (trainer.py:246,339)
                    pred_real = self.netD.forward_pair(layout, imgs)
                img_pred_fake = self.netD.forward_pair(layout, imgs_pred)
                terms['g_gan_img_loss'] = self.criterionGAN(img_pred_fake, True)
                if args.d_img_features_weight > 0:
                    terms['g_gan_features_loss_img'] = self.calculate_features_loss(img_pred_fake, pred_real)
            if br:
                for st in br:
                    main.wait_stream(st)
                ops.refresh_stream()
            weights = (('g_vgg', args.vgg_features_weight), ('ac_loss', args.ac_loss_weight),
                       ('g_gan_obj_loss', args.d_obj_weight), ('g_gan_mask_obj_loss', args.d_mask_weight),
                       ('g_mask_features_loss', args.d_mask_features_weight), ('g_gan_img_loss', args.d_img_weight),
                       ('g_gan_features_loss_img', args.d_img_features_weight))
            for name, wgt in weights:
                if name in terms:
                    gl.add_loss(terms[name], name, wgt)
            gl._terms['total_loss'] = gl.total_loss.detach()
            self._backward('g', self.optimizer, gl)
        finally:
            for net in d_nets:
                for p in net.parameters():
                    p.requires_grad_(True)

    def train_obj_discriminator(self, imgs, imgs_pred, objs, boxes, boxes_pred, obj_to_img):
        if self.obj_discriminator is None:
            return
        self.d_obj_losses = dl = LossManager()
        scores_fake, ac_fake, self.d_fake_crops = self.obj_discriminator(imgs_pred, objs, boxes_pred, obj_to_img)
        scores_real, ac_real, self.d_real_crops = self.obj_discriminator(imgs, objs, boxes, obj_to_img)
        dl.add_loss(self.gan_d_loss(scores_real, scores_fake), 'd_obj_gan_loss', 0.5)
        dl.add_loss(ac_real, 'd_ac_loss_real')
        dl.add_loss(ac_fake, 'd_ac_loss_fake')
        self._backward('obj', self.optimizer_d_obj, dl)

    def train_mask_discriminator(self, masks, masks_pred, objs):
        if self.mask_discriminator is None:
            return
        self.d_mask_losses = dl = LossManager()
        if self.args.norm_D_mask == 'instance' and not masks_pred.requires_grad:
            # fake and real masks in one batched pass: every layer acts per sample (conv, InstanceNorm, LeakyReLU)
            O = objs.numel()
            both = self.mask_discriminator(torch.cat([masks_pred.unsqueeze(1).float(), masks.unsqueeze(1).float()]),
                                           torch.cat([objs, objs]))
            scores_fake = [[f[:O] for f in col] for col in both]
            scores_real = [[f[O:] for f in col] for col in both]
        else:
            scores_fake = self.mask_discriminator(masks_pred.unsqueeze(1), objs)
            scores_real = self.mask_discriminator(masks.unsqueeze(1), objs)
        dl.add_loss(self.criterionGAN(scores_fake, False), 'fake_loss', 0.5)
        dl.add_loss(self.criterionGAN(scores_real, True), 'real_loss', 0.5)
        self._backward('mask', self.optimizer_d_mask, dl)

    def train_image_discriminator(self, imgs, imgs_pred, layout, layout_wrong):
        if self.netD is None:
            return
        self.d_img_losses = dl = LossManager()
        alpha = 0.25
        # the three discriminate() calls of trainer.py:309-319 as one batched pass (per-sample layers: exact)
        pairs = [(layout, imgs_pred.detach()), (layout_wrong, imgs), (layout, imgs)]
        if self.args.norm_D == 'instance':
            fake, wrong, real = self.netD.forward_pairs(pairs)
        else:
            fake, wrong, real = (self.discriminate(l, i) for l, i in pairs)
        dl.add_loss(self.criterionGAN(fake, False), 'fake_image_loss', alpha)
        dl.add_loss(self.criterionGAN(wrong, False), 'wrong_texture_loss', alpha)
        dl.add_loss(self.criterionGAN(real, True), 'd_img_gan_real_loss', 0.5)
        self._backward('img', self.optimizer_d_img, dl)

    def discriminate(self, input_label, test_image):
        return self.netD.forward_pair(input_label, test_image)

    def calculate_features_loss(self, pred_fake, pred_real):
        """trainer.py:331-340."""
        loss = 0
        fw, dw = 4.0 / len(pred_fake[0]), 1.0 / len(pred_fake)
        for i in range(len(pred_fake)):
            for j in range(len(pred_fake[i]) - 1):
                loss = loss + dw * fw * self.criterionFeat(pred_fake[i][j].float(), pred_real[i][j].detach().float())
        return loss

    # ---- one iteration of train.py:190-215 in four phases ------------------------------------------
    def train_step(self, batch, use_gt=True, graph=None):
        """One iteration of train.py:190-215 on a collated batch.  The batch tensors may live on the device or
        (graph replay) in pinned host memory — they are copied into the captured iteration's input buffers.
        graph: None -> self.use_graphs."""
        use_graph = self.use_graphs if graph is None else graph
        if use_graph and torch.is_grad_enabled():
            return self._train_step_graphed(batch, use_gt)
        return self._train_step_eager(batch, use_gt)

    def _d_nets(self):
        return [(n, o) for n, o in (('mask', self.optimizer_d_mask), ('obj', self.optimizer_d_obj), ('img', self.optimizer_d_img))
                if o is not None]

    def _phase_a(self, batch, use_gt):
        """Model.forward (model.py:94-124) + generator losses + backward (trainer.py:205-262 without the update)"""
        with _nvtx('sg.phase_a: forward + generator backward'):
            return self._phase_a_impl(batch, use_gt)

    def _phase_a_impl(self, batch, use_gt):
        imgs, objs, boxes, masks, triples, obj_to_img, triple_to_img, attributes = batch
        if not use_gt:
            attributes = torch.zeros_like(attributes)
        out = self.model(imgs, objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks, attributes=attributes)
        imgs_pred, boxes_pred, masks_pred, layout, layout_pred, layout_wrong = out
        self.train_generator(imgs, imgs_pred, masks, masks_pred, layout, objs, boxes, boxes_pred, obj_to_img, use_gt)
        self._st = (batch, out)
        return out

    def _phase_b(self):
        """the three discriminator passes (train.py:207-215): losses + backward, no update.  They only share read-only
        inputs, and none of them reads the generator's weights: inside a capture they are parallel graph branches."""
        with _nvtx('sg.phase_b: discriminator passes'):
            self._phase_b_impl()

    def _phase_b_impl(self):
        from . import ops
        (imgs, objs, boxes, masks, triples, obj_to_img, _, _), out = self._st
        imgs_pred, boxes_pred, masks_pred, layout, layout_pred, layout_wrong = out
        masks_fake, imgs_fake = masks_pred.detach(), imgs_pred.detach()
        # the reference hands layout.detach() / layout_wrong.detach() to the image discriminator (train.py:211-215)
        lay, lay_wrong = _detached(layout), _detached(layout_wrong)
        br = self._branches
        if br:
            main = torch.cuda.current_stream()
            for st in br[:2]:
                st.wait_stream(main)
            with ops.on_stream(br[0]):
                self.train_mask_discriminator(masks, masks_fake, objs)
            with ops.on_stream(br[1]):
                self.train_obj_discriminator(imgs, imgs_fake, objs, boxes, boxes.detach(), obj_to_img)
            self.train_image_discriminator(imgs, imgs_fake, lay, lay_wrong)
            for st in br[:2]:
                main.wait_stream(st)
            ops.refresh_stream()
        else:
            self.train_mask_discriminator(masks, masks_fake, objs)
            self.train_obj_discriminator(imgs, imgs_fake, objs, boxes, boxes.detach(), obj_to_img)
            self.train_image_discriminator(imgs, imgs_fake, lay, lay_wrong)

    def _phase_cd(self):
        with _nvtx('sg.phase_cd: discriminator Adam'):
            for name, opt in self._d_nets():
                self._update(name, opt)

    def _train_step_eager(self, batch, use_gt):
        """eager launches on the current stream; data parallel: the generator's all-reduce runs on NCCL's stream next
        to the discriminator passes"""
        batch = _batch_to_device(batch, 'cuda')
        self._phased = True
        try:
            out = self._phase_a(batch, use_gt)
            works = self._allreduce(('g',), async_op=True)
            self._phase_b()
            for w in works:
                w.wait()
            self._update('g', self.optimizer)
            self._allreduce([n for n, _ in self._d_nets()])
            self._phase_cd()
        finally:
            self._phased = False
            self._st = None
        return out

    # ---- captured iterations ---------------------------------------------------------------------
    def _train_step_graphed(self, batch, use_gt):
        from . import _lib, ops
        meta = _BatchMeta.of(batch)
        if meta is None:              # no loader metadata: Model.forward would have to sync for it -> not capturable
            return self._train_step_eager(batch, use_gt)
        compact = meta.slots_used <= self.model.compact_slots
        key = (bool(use_gt), compact, tuple((tuple(t.shape), t.dtype) for t in batch)) + meta.geometry()
        ent = self._graphs.get(key)
        if ent is None:               # first sight of this geometry: eager (also warms every lazy initialisation)
            self._graphs[key] = 'warm'
            self._evict()
            # detached like the outputs of a replay: a caller holding on to this iteration's autograd graph would
            # keep its AccumulateGrad nodes (bound to the eager stream) alive into the capture
            return tuple(o.detach() if torch.is_tensor(o) else o for o in self._train_step_eager(batch, use_gt))
        self._graphs.move_to_end(key)
        pool = self.model.fake_pool
        plan = None
        if pool.pool_size > 0:        # host half of the VectorPool, exactly once per iteration
            if pool.store is None:
                return self._train_step_eager(batch, use_gt)
            plan = pool.plan(meta.objs_host)
        if ent == 'warm' and not self._capturable():
            return tuple(o.detach() if torch.is_tensor(o) else o for o in self._train_step_eager(batch, use_gt))
        if ent == 'warm':
            ent = _StepGraph(batch, meta, plan)
            ent.load(batch, meta, plan)
            try:
                self._capture(ent, use_gt)
            except Exception as e:          # capture executes nothing on the device: the eager path can still run this step
                warnings.warn('CUDA graph capture of the training step failed (%s: %s); continuing with eager launches'
                              % (type(e).__name__, e))
                self.use_graphs = False
                self._graphs.clear()
                # a capture that ended in an error leaves torch's default CUDA generator in capture mode (its epilogue
                # never ran): draw the mask noise from a generator of our own from here on
                gen = torch.Generator(device=ent.batch[0].device)
                gen.manual_seed(torch.initial_seed() + 1)
                self.model.noise_generator = gen
                ops.refresh_stream()
                Fn.clear_weight_cache()
                self.model.pool_plan = None if plan is None else ent.pool_idx
                try:
                    return self._train_step_eager(ent.batch, use_gt)
                finally:
                    self.model.pool_plan = None
            self._graphs[key] = ent
            self._evict()
        else:
            ent.load(batch, meta, plan)
        # replay: A | all-reduce(generator) on NCCL's stream | B | Cg on a side stream behind the all-reduce |
        # all-reduce(discriminators) | Cd.  The same order of collectives as _train_step_eager.
        g = ent.graphs
        main = torch.cuda.current_stream()
        side = self._side_streams[2]
        g['A'].replay()
        works = self._allreduce(('g',), async_op=True)
        if not works:
            side.wait_stream(main)    # Cg needs A's gradients only (with an all-reduce it waits for that instead)
        g['B'].replay()
        with torch.cuda.stream(side):
            for w in works:
                w.wait()              # stream-side wait: `side` waits for NCCL's stream
            g['Cg'].replay()
        self._allreduce([n for n, _ in self._d_nets()])
        g['Cd'].replay()
        main.wait_stream(side)
        _lib.add_launch_count(ent.launches)
        Fn.drop_unmaintained()        # the replay updated masters behind the version counters of the bf16 operand cache
        self.generator_losses, self.d_mask_losses, self.d_obj_losses, self.d_img_losses = ent.losses
        return ent.out

    def _evict(self):
        """keep at most _graph_cache_size captured geometries (least recently used go first)"""
        captured = [k for k, v in self._graphs.items() if not isinstance(v, str)]
        while len(captured) > max(1, self._graph_cache_size):
            del self._graphs[captured.pop(0)]
        warm = [k for k, v in self._graphs.items() if isinstance(v, str)]
        while len(warm) > 4096:
            del self._graphs[warm.pop(0)]

    def release_graphs(self):
        """forget every captured iteration (and the memory pool they share); the next train_step starts over with eager
        sightings."""
        self._graphs.clear()
        self._graph_pool = None
        self.generator_losses = self.d_mask_losses = self.d_obj_losses = self.d_img_losses = None

    def _capturable(self):
        """Adam creates a parameter's state at its first gradient and skips parameters without one (like the
        reference's optimizer).  A capture would freeze a "no state yet" skip into the graph: wait until every trainable
        parameter has had a gradient once (box_net gets its first one in the first use_gt iteration, train.py:195)."""
        for opt in (self.optimizer, self.optimizer_d_obj, self.optimizer_d_mask, self.optimizer_d_img):
            if opt is None:
                continue
            for group in opt.param_groups:
                for p in group['params']:
                    if p.requires_grad and len(opt.state.get(p, ())) == 0:
                        return False
        return True

    def _materialize_optimizer_state(self):
        """Adam creates a parameter's moments lazily at its first gradient (torch/optim/adam.py, _init_group).  Inside a
        capture that creation would be recorded — and the moments re-zeroed by every replay — so parameters that have
        not had a gradient yet get their state here."""
        for opt in (self.optimizer, self.optimizer_d_obj, self.optimizer_d_mask, self.optimizer_d_img):
            if opt is None:
                continue
            for group in opt.param_groups:
                for p in group['params']:
                    if p.requires_grad and len(opt.state[p]) == 0:
                        st = opt.state[p]
                        st['step'] = torch.zeros((), dtype=torch.float32, device=p.device)
                        st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                        st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)

    def _capture(self, ent, use_gt):
        """capture the four graphs of one batch geometry into the shared pool (nothing executes; no collective is
        captured).  thread_local capture mode: NCCL's watchdog thread may query its events meanwhile."""
        from . import _lib, ops
        self._materialize_optimizer_state()
        if self._side_streams is None:
            self._side_streams = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream())
        Fn.drop_unmaintained()        # operands no optimizer keeps current are re-packed at their first use inside the graph
        self.model.pool_plan = ent.pool_idx
        parallel = os.environ.get('SG_PARALLEL_BRANCHES', '1') != '0'
        if parallel and os.environ.get('SG_PARALLEL_FWD', '1') != '0':     # Model._forward_train_two_branches
            if self._fwd_stream is None:
                self._fwd_stream = torch.cuda.Stream()
            self.model.graph_branch_stream = self._fwd_stream
        ops.CACHE_STREAM[0] = False   # backward nodes of a side branch run on that branch's stream
        mode = os.environ.get('SG_GRAPH_CAPTURE_MODE', 'thread_local')
        torch.cuda.synchronize()      # nothing of an earlier iteration (or of NCCL) in flight while capturing
        graphs = {}
        n0 = _lib.launch_count()
        self._phased = True
        self._branches = self._side_streams if parallel else None

        def cap(name, fn):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._graph_pool, capture_error_mode=mode):
                r = fn()
            if self._graph_pool is None:
                self._graph_pool = g.pool()
            graphs[name] = g
            return r
        try:
            out = cap('A', lambda: self._phase_a(ent.batch, use_gt))
            ent.out = tuple(o.detach() if torch.is_tensor(o) else o for o in out)
            cap('B', self._phase_b)
            cap('Cg', lambda: self._update('g', self.optimizer))
            cap('Cd', self._phase_cd)
            del out
        finally:
            self._phased = False
            self._branches = None
            self._st = None
            self.model.pool_plan = None
            self.model.graph_branch_stream = None
            ops.CACHE_STREAM[0] = True
            ops.refresh_stream()      # the cached stream handle is the capture stream
            Fn.drop_unmaintained()
        ent.launches = _lib.launch_count() - n0
        ent.losses = (self.generator_losses, self.d_mask_losses, self.d_obj_losses, self.d_img_losses)
        ent.graphs = graphs

    def write_losses(self, checkpoint, t):
        print('t = %d / %d' % (t, self.args.num_iterations))
        for tag, lm in (('G', self.generator_losses), ('D_obj', getattr(self, 'd_obj_losses', None)),
                        ('D_mask', getattr(self, 'd_mask_losses', None)), ('D_img', getattr(self, 'd_img_losses', None))):
            if lm is None:
                continue
            for name, val in lm.items():
                print(' %s [%s]: %.4f' % (tag, name, val))
                if self.writer is not None:
                    self.writer.add_scalar('%s/%s' % (tag, name), val, int(t / self.args.print_every))
