"""Flag set of the training driver (same names and defaults as scene_generation/args.py:10-113 for every
flag the hot path reads), plus the B200-specific switches."""
import argparse

from .utils import bool_flag, int_tuple

parser = argparse.ArgumentParser()
parser.add_argument('--batch_size', default=12, type=int)
parser.add_argument('--num_iterations', default=1000000, type=int)
parser.add_argument('--learning_rate', default=1e-4, type=float)
parser.add_argument('--mask_learning_rate', default=1e-5, type=float)
parser.add_argument('--image_size', default='128,128', type=int_tuple)
parser.add_argument('--min_objects_per_image', default=3, type=int)
parser.add_argument('--max_objects_per_image', default=8, type=int)
# generator
parser.add_argument('--mask_size', default=32, type=int)
parser.add_argument('--embedding_dim', default=128, type=int)
parser.add_argument('--gconv_dim', default=128, type=int)
parser.add_argument('--gconv_hidden_dim', default=512, type=int)
parser.add_argument('--gconv_num_layers', default=5, type=int)
parser.add_argument('--mlp_normalization', default='none', type=str)
parser.add_argument('--activation', default='leakyrelu-0.2')
parser.add_argument('--pool_size', default=100, type=int)
parser.add_argument('--output_nc', default=3, type=int)
parser.add_argument('--n_downsample_global', default=4, type=int)
parser.add_argument('--box_dim', default=128, type=int)
parser.add_argument('--use_attributes', default=True, type=bool_flag)
parser.add_argument('--beta1', default=0.5, type=float)
parser.add_argument('--box_noise_dim', default=64, type=int)
parser.add_argument('--mask_noise_dim', default=64, type=int)
parser.add_argument('--rep_size', default=32, type=int)
parser.add_argument('--appearance_normalization', default='batch')
# generator losses
parser.add_argument('--l1_pixel_loss_weight', default=.0, type=float)
parser.add_argument('--bbox_pred_loss_weight', default=10, type=float)
parser.add_argument('--vgg_features_weight', default=10.0, type=float)  # args.py:73; needs the ImageNet VGG19 weights:
parser.add_argument('--vgg_weights', default=None, type=str)             # a torchvision vgg19 state_dict file (nothing is downloaded)
parser.add_argument('--vgg_random_init', default=False, type=bool_flag)  # throughput runs / tests only: seeded random VGG19
parser.add_argument('--d_img_weight', default=1.0, type=float)
parser.add_argument('--d_img_features_weight', default=10.0, type=float)
parser.add_argument('--d_mask_weight', default=1.0, type=float)
parser.add_argument('--d_mask_features_weight', default=10.0, type=float)
parser.add_argument('--d_obj_weight', default=0.1, type=float)
parser.add_argument('--ac_loss_weight', default=0.1, type=float)
# image discriminator
parser.add_argument('--ndf', default=64, type=int)
parser.add_argument('--num_D', default=2, type=int)
parser.add_argument('--norm_D', default='instance', type=str)
parser.add_argument('--n_layers_D', default=3, type=int)
parser.add_argument('--no_lsgan', default=False, type=bool_flag)
# mask discriminator
parser.add_argument('--ndf_mask', default=64, type=int)
parser.add_argument('--num_D_mask', default=1, type=int)
parser.add_argument('--norm_D_mask', default='instance', type=str)
parser.add_argument('--n_layers_D_mask', default=2, type=int)
# object discriminator
parser.add_argument('--gan_loss_type', default='gan')
parser.add_argument('--d_normalization', default='batch')
parser.add_argument('--d_padding', default='valid')
parser.add_argument('--d_activation', default='leakyrelu-0.2')
parser.add_argument('--d_obj_arch', default='C4-64-2,C4-128-2,C4-256-2')
parser.add_argument('--crop_size', default=32, type=int)
# output
parser.add_argument('--print_every', default=100, type=int)
parser.add_argument('--checkpoint_every', default=10000, type=int)
parser.add_argument('--output_dir', default='output')
parser.add_argument('--checkpoint_name', default='checkpoint')
parser.add_argument('--restore_from_checkpoint', default=False, type=bool_flag)
# B200 additions
parser.add_argument('--layout_dtype', default='bf16', choices=['bf16', 'f32'])
parser.add_argument('--align_corners', default=False, type=bool_flag)
parser.add_argument('--num_objs', default=172, type=int)   # synthetic vocabulary size (COCO-Stuff: 172)
parser.add_argument('--cuda_graphs', default=True, type=bool_flag)   # replay captured iterations (trainer.py)
parser.add_argument('--graph_cache', default=64, type=int)           # captured batch geometries kept alive


def get_args(argv=None):
    return parser.parse_args(argv)


def default_args(**over):
    """the reference's defaults for offline runs (tests, bench.py): the VGG19 term is off unless asked for, because
    its pretrained weights cannot be downloaded here (pass vgg_features_weight=10 with vgg_weights=... or
    vgg_random_init=True to include it)"""
    a = parser.parse_args([])
    a.vgg_features_weight = 0.0
    for k, v in over.items():
        setattr(a, k, v)
    return a
