"""Training step (mirror of scene_generation/trainer.py: same Trainer(args, vocab, checkpoint)
constructor, same train_* methods and checkpoint keys).  Differences, all outside the arithmetic:
loss terms stay on the device (no per-term .item()), the image discriminator takes (layout, image)
pairs instead of a materialised concat, and with torch.distributed initialised every optimizer step
is preceded by a flat-buffer NCCL all-reduce of that network's gradients (DDP semantics)."""
import os

import torch
import torch.nn.functional as F

from . import ddp
from . import functional as Fn
from .discriminators import AcCropDiscriminator, define_D, define_mask_D
from .losses import GANLoss, get_gan_losses
from .model import Model
from .utils import LossManager

try:                                           # logging is optional plumbing (tensorboardX is not in the image)
    from tensorboardX import SummaryWriter
except Exception:                              # pragma: no cover
    SummaryWriter = None


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
        self.reducers = {}
        if ddp.world_size() > 1:
            for name, net in (('g', self.model), ('img', self.netD), ('obj', self.obj_discriminator),
                              ('mask', self.mask_discriminator)):
                if net is not None:
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
        if getattr(args, 'vgg_features_weight', 0) > 0:
            raise NotImplementedError('VGG perceptual loss needs pretrained weights (no network): run with '
                                      '--vgg_features_weight 0 (SURVEY.md §8f-2)')
        self.criterionVGG = None
        self.criterionFeat = torch.nn.L1Loss()
        self.criterionGAN = GANLoss(use_lsgan=not args.no_lsgan)
        self.optimizer = torch.optim.Adam(model.parameters(), lr=args.learning_rate, betas=(args.beta1, 0.999), fused=True)

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
            self.optimizer_d_obj = torch.optim.Adam(self.obj_discriminator.parameters(), lr=args.learning_rate,
                                                    betas=(args.beta1, 0.999), fused=True)

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
            self.optimizer_d_mask = torch.optim.Adam(self.mask_discriminator.parameters(), lr=args.mask_learning_rate,
                                                     betas=(args.beta1, 0.999), fused=True)

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
        self.optimizer_d_img = torch.optim.Adam(list(self.netD.parameters()), lr=args.learning_rate,
                                                betas=(args.beta1, 0.999), fused=True)

    # ---- checkpoint (trainer.py:136-203) -------------------------------------------------------
    def restore_checkpoint(self, checkpoint):
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
    def _step(self, name, optimizer, losses):
        from . import ops
        ops.refresh_stream()
        if name in self.reducers:
            self.reducers[name].zero()            # .grad tensors are views of one flat buffer
        else:
            optimizer.zero_grad(set_to_none=True)
        losses.total_loss.backward()
        if name in self.reducers:
            self.reducers[name].allreduce()
        optimizer.step()

    def _one_hot(self, objs, like):
        oh = torch.zeros((objs.numel(), self.num_obj), dtype=like.dtype, device=like.device)
        return oh.scatter_(1, objs.view(-1, 1).long(), 1.0)

    def train_generator(self, imgs, imgs_pred, masks, masks_pred, layout, objs, boxes, boxes_pred, obj_to_img, use_gt):
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
        try:
            scores_fake, ac_loss, _ = self.obj_discriminator(imgs_pred, objs, boxes, obj_to_img)
            gl.add_loss(ac_loss, 'ac_loss', args.ac_loss_weight)
            gl.add_loss(self.gan_g_loss(scores_fake), 'g_gan_obj_loss', args.d_obj_weight)
            if self.mask_discriminator is not None:
                scores_fake = self.mask_discriminator(masks_pred.unsqueeze(1), objs)
                gl.add_loss(self.criterionGAN(scores_fake, True), 'g_gan_mask_obj_loss', args.d_mask_weight)
                if args.d_mask_features_weight > 0:
                    with torch.no_grad():
                        scores_real = self.mask_discriminator(masks.unsqueeze(1), objs)
                    gl.add_loss(self.calculate_features_loss(scores_fake, scores_real), 'g_mask_features_loss',
                                args.d_mask_features_weight)
            if self.netD is not None:
                with torch.no_grad():       # only used detached (trainer.py:246,339)
                    pred_real = self.netD.forward_pair(layout, imgs)
                img_pred_fake = self.netD.forward_pair(layout, imgs_pred)
                gl.add_loss(self.criterionGAN(img_pred_fake, True), 'g_gan_img_loss', args.d_img_weight)
                if args.d_img_features_weight > 0:
                    gl.add_loss(self.calculate_features_loss(img_pred_fake, pred_real), 'g_gan_features_loss_img',
                                args.d_img_features_weight)
            gl._terms['total_loss'] = gl.total_loss.detach()
            self._step('g', self.optimizer, gl)
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
        self._step('obj', self.optimizer_d_obj, dl)

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
        self._step('mask', self.optimizer_d_mask, dl)

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
        self._step('img', self.optimizer_d_img, dl)

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

    def train_step(self, batch, use_gt=True):
        """One iteration of train.py:190-215 on a collated batch (tensors already on the device)."""
        imgs, objs, boxes, masks, triples, obj_to_img, triple_to_img, attributes = batch
        if not use_gt:
            attributes = torch.zeros_like(attributes)
        Fn.ARENA.begin_step(imgs.device)        # zero-initialised scratch of this iteration (one fill per step)
        out = self.model(imgs, objs, triples, obj_to_img, boxes_gt=boxes, masks_gt=masks, attributes=attributes)
        imgs_pred, boxes_pred, masks_pred, layout, layout_pred, layout_wrong = out
        self.train_generator(imgs, imgs_pred, masks, masks_pred, layout, objs, boxes, boxes_pred, obj_to_img, use_gt)
        self.train_mask_discriminator(masks, masks_pred.detach(), objs)
        self.train_obj_discriminator(imgs, imgs_pred.detach(), objs, boxes, boxes.detach(), obj_to_img)
        self.train_image_discriminator(imgs, imgs_pred.detach(), layout, layout_wrong)
        return out

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
