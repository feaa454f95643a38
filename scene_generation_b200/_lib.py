"""ctypes binding of libsg_b200.so (the C ABI declared in include/sg_b200.h).

The library is built in-tree by :func:`build` (nvcc, sm_100a only) and loaded lazily.  There is
no fallback: if the shared object is missing or an entry point fails, a RuntimeError is raised.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, 'libsg_b200.so')
SOURCES = ['runtime.cu', 'layout.cu', 'graph.cu', 'crop.cu', 'conv_tc.cu', 'elementwise.cu', 'smallconv.cu', 'compact.cu', 'adam.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

SG_MAX_TAPS = 64
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4

c_int, c_long, c_float, c_void_p = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p


class Tap(ctypes.Structure):
    _fields_ = [('dh', ctypes.c_int16), ('dw', ctypes.c_int16), ('plane', ctypes.c_int16), ('wtap', ctypes.c_int16)]


class Phase(ctypes.Structure):
    _fields_ = [('tap_begin', c_int), ('ntaps', c_int), ('oh_off', c_int), ('ow_off', c_int)]


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ('x', c_void_p), ('x_N', c_int), ('x_P', c_int), ('x_H', c_int), ('x_W', c_int), ('x_C', c_int),
        ('w', c_void_p), ('w_Cout', c_int), ('w_taps', c_int), ('w_C', c_int),
        ('y', c_void_p), ('y_dtype', c_int),
        ('y_os_img', c_long), ('y_os_h', c_long), ('y_os_w', c_long), ('y_os_c', c_long),
        ('Hout', c_int), ('Wout', c_int), ('oh_mul', c_int), ('ow_mul', c_int),
        ('in_h0', c_int), ('in_w0', c_int),
        ('nphases', c_int), ('phases', Phase * 4),
        ('ntaps', c_int), ('taps', Tap * SG_MAX_TAPS),
        ('bias', c_void_p), ('act', c_int), ('slope', c_float), ('stats', c_void_p), ('stats_slots', c_int),
        ('w_img_rows', c_int), ('w_row0', c_int),
        ('w_mn', c_int), ('w_rows', c_int), ('w_col0', c_int),
    ]


class NapDesc(ctypes.Structure):
    _fields_ = [
        ('src', c_void_p), ('N', c_int), ('H', c_int), ('W', c_int), ('C', c_int),
        ('scale', c_void_p), ('shift', c_void_p), ('act', c_int), ('slope', c_float),
        ('res', c_void_p), ('res_os_img', c_long), ('res_os_h', c_long), ('res_os_w', c_long),
        ('up', c_int), ('pad', c_int), ('pad_mode', c_int), ('planes', c_int),
    ]


class WTap(ctypes.Structure):
    _fields_ = [('dha', ctypes.c_int16), ('dwa', ctypes.c_int16), ('pa', ctypes.c_int16),
                ('dhb', ctypes.c_int16), ('dwb', ctypes.c_int16), ('pb', ctypes.c_int16),
                ('wtap', ctypes.c_int16), ('pad', ctypes.c_int16)]


class WgradDesc(ctypes.Structure):
    _fields_ = [
        ('dy', c_void_p), ('N', c_int), ('dy_P', c_int), ('dy_H', c_int), ('dy_W', c_int), ('dy_C', c_int),
        ('x', c_void_p), ('x_P', c_int), ('x_H', c_int), ('x_W', c_int), ('x_C', c_int),
        ('Hred', c_int), ('Wred', c_int),
        ('dw', c_void_p), ('Cout', c_int), ('Cin', c_int), ('w_taps', c_int), ('dw_C', c_int),
        ('ntaps', c_int), ('taps', WTap * SG_MAX_TAPS),
        ('ksplit', c_int), ('per_image', c_int), ('ws', c_void_p), ('ws_floats', c_long),
    ]


def build(verbose=False, force=False):
    """Compile every CUDA source of the package into libsg_b200.so for sm_100a (cross-compiles
    without a GPU).  Skips the build when the library is newer than all sources."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')] + \
        [os.path.join(os.path.dirname(_HERE), 'include', 'sg_b200.h')]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB_PATH] + srcs
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None

# name -> argtypes (restype is int unless listed in _RESTYPES)
_P = c_void_p
_SIGS = {
    'sg_masks_to_layout_fwd': [_P, _P, _P, c_int, _P] + [c_int] * 9 + [_P, _P],
    'sg_masks_to_layout_bwd': [_P, _P, _P, c_int, _P] + [c_int] * 9 + [_P, c_int, c_int, _P, _P, _P, c_long, _P],
    'sg_masks_to_layout_test': [_P, _P, _P, c_int, _P] + [c_int] * 9 + [_P, _P, _P],
    'sg_gconv_gather_fwd': [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_gconv_pool_fwd': [_P, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_gconv_pool_bwd': [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_gconv_gather_bwd': [_P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P],
    'sg_crop_bbox_fwd': [_P, _P, _P] + [c_int] * 10 + [_P, _P],
    'sg_crop_bbox_bwd': [_P, _P] + [c_int] * 10 + [_P, _P, _P],
    'sg_conv_tc': [ctypes.POINTER(ConvDesc), _P],
    'sg_conv_stats_slots': [ctypes.POINTER(ConvDesc), ctypes.POINTER(c_int)],
    'sg_cast_pad_bf16': [_P, c_long, c_int, c_long, c_int, _P, c_float, _P, _P],
    'sg_pack_weight': [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P],
    'sg_pack_weight_cmap': [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P, _P, _P],
    'sg_wgrad_cmap_scatter': [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_norm_finalize': [_P, c_int, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, _P, c_float, _P, _P, _P, _P, _P],
    'sg_norm_act_pad_fwd': [ctypes.POINTER(NapDesc), _P, _P],
    'sg_norm_act_pad_bwd': [ctypes.POINTER(NapDesc), _P, _P, _P, c_int, c_float, _P, c_int, _P, _P, _P],
    'sg_act_bwd_nchw': [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_nchw_to_nhwc': [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_nhwc_to_nchw': [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_concat_cond': [_P, c_long, c_int, c_int, c_int, _P, c_int, _P, _P],
    'sg_slice_channels': [_P, c_long, c_int, c_int, _P, _P],
    'sg_avgpool3x3s2_fwd': [_P, c_int, c_int, c_int, c_int, _P, _P],
    'sg_avgpool3x3s2_bwd': [_P, c_int, c_int, c_int, c_int, _P, _P],
    'sg_maxpool2x2_fwd': [_P, c_int, c_int, c_int, c_int, _P, _P],
    'sg_maxpool2x2_bwd': [_P, _P, c_int, c_int, c_int, c_int, _P, _P],
    'sg_gap_fwd': [_P, c_int, c_int, c_int, _P, _P],
    'sg_gap_bwd': [_P, c_int, c_int, c_int, _P, _P],
    'sg_colsum_bf16': [_P, c_long, c_int, c_int, _P, _P, c_long, _P],
    'sg_norm_act_pad_bwd_parts': [c_int, c_int, c_int, c_int],
    'sg_adam_pack': [c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                     ctypes.c_double, _P],
    'sg_wgrad_tc': [ctypes.POINTER(WgradDesc), _P],
    'sg_probe_shifted_desc': [_P, _P, _P, c_int, _P],
    'sg_probe_mma_rate': [c_int, c_int, c_int, _P, _P],
    'sg_im2col_dz': [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_dgrad_small_cout': [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'sg_wgrad_small_cout': [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_long, _P],
}
_RESTYPES = {'sg_last_error': ctypes.c_char_p, 'sg_version': ctypes.c_char_p, 'sg_arch': c_int,
             'sg_launch_count': ctypes.c_ulonglong, 'sg_reset_launch_count': None, 'sg_add_launch_count': None}


def declared_symbols():
    """Every function name declared in include/sg_b200.h (parsed from the header)."""
    import re
    hdr = open(os.path.join(os.path.dirname(_HERE), 'include', 'sg_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(sg_[a-z0-9_]+)\s*\(', hdr)))


def lib():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libsg_b200.so is not built (run `python -c "import __graft_entry__ as g; g.build()"`); '
                               'there is no fallback path')
        h = ctypes.CDLL(LIB_PATH)
        for name, rt in _RESTYPES.items():
            getattr(h, name).restype = rt
            getattr(h, name).argtypes = [ctypes.c_ulonglong] if name == 'sg_add_launch_count' else []
        for name, args in _SIGS.items():
            fn = getattr(h, name)
            fn.restype = c_int
            fn.argtypes = args
        _lib = h
    return _lib


def register(name, argtypes):
    _SIGS[name] = argtypes
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes


def check(code, what=''):
    if code != 0:
        raise RuntimeError('libsg_b200 %s failed (%d): %s' % (what, code, lib().sg_last_error().decode()))


def call(name, *args):
    check(getattr(lib(), name)(*args), name)


def launch_count():
    return int(lib().sg_launch_count())


def reset_launch_count():
    lib().sg_reset_launch_count()


def add_launch_count(n):
    """account launches that a CUDA graph replay performed without passing through the entry points"""
    lib().sg_add_launch_count(int(n))
