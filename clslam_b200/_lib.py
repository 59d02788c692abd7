"""ctypes binding of include/clslam_b200.h (the C-ABI drop-in boundary).

The library is required: there is no CPU or PyTorch fallback.  ``load()`` raises if
``libclslam_b200.so`` has not been built (``python -m clslam_b200.build``).
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_longlong, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / 'libclslam_b200.so'

ACT_NONE, ACT_RELU, ACT_ELU, ACT_SIGMOID = 0, 1, 2, 3
PAD_ZERO, PAD_REFLECT, PAD_FULL = 0, 1, 2
TMA_IN_HALO_VALID = 1          # clslam_conv_tma.flags
TMA_SPLITK_CLUSTER = 2         # clslam_conv_tma.flags
TMA_SPLITK_COUNTERS = 256      # ints behind clslam_conv_tma.splitk_counters
IN_PLAIN, IN_UPCAT = 0, 1
DEPTH_INV, DEPTH_MIN_ONLY, DEPTH_MIN_MAX = 0, 1, 2
N_LOSS_SLOTS = 24
ADAM_NO_TICK = -1


class ConvDesc(Structure):
    _fields_ = [(n, c_int) for n in ('n', 'hi', 'wi', 'ci', 'ho', 'wo', 'co', 'kh', 'kw', 'stride', 'pad',
                                     'pad_mode', 'in_mode', 'c1', 'h1', 'w1', 'act')]


class Planes(Structure):
    """clslam_planes: bf16 hi/lo activation planes [2][n][h+2][w+2][c_total]"""
    _fields_ = [('base', c_void_p), ('plane_stride', c_longlong), ('n', c_int), ('h', c_int), ('w', c_int),
                ('c_total', c_int), ('c_off', c_int), ('c', c_int)]


class ConvTma(Structure):
    """clslam_conv_tma"""
    _fields_ = [('inp', Planes), ('pad_mode', c_int), ('cout', c_int), ('w_hi', c_void_p), ('w_lo', c_void_p),
                ('scale', c_void_p), ('shift', c_void_p), ('act', c_int), ('residual', POINTER(Planes)), ('gate', c_void_p), ('gate_act', c_int), ('y', c_void_p),
                ('out', POINTER(Planes)), ('out_up', c_int), ('out_halo', c_int), ('taps', c_int), ('flags', c_int),
                ('splitk_counters', c_void_p), ('sm_budget', c_int)]


class PhotoArgs(Structure):
    _fields_ = [
        ('b', c_int), ('h', c_int), ('w', c_int), ('n_scales', c_int),
        ('depth_mode', c_int), ('min_depth', c_float), ('max_depth', c_float),
        ('target', c_void_p), ('sources', c_void_p), ('sources_rgba', c_void_p), ('inv_k', c_void_p), ('proj', c_void_p),
        ('identity', c_void_p), ('sample_w', c_void_p),
        ('disp', c_void_p * 4), ('noise', c_void_p * 4), ('depth', c_void_p * 4), ('warped', c_void_p * 4),
        ('argmin', c_void_p * 4), ('loss_partials', c_void_p), ('ddisp_full', c_void_p * 4),
        ('dproj_partials', c_void_p), ('dyn_mask', c_void_p), ('reproj_w', c_void_p),
    ]


class LossArgs(Structure):
    _fields_ = [
        ('b', c_int), ('b_global', c_int), ('h', c_int), ('w', c_int), ('n_scales', c_int), ('n_tiles', c_int),
        ('has_sample0', c_int), ('disparity_smoothness', c_float), ('velocity_scale', c_float),
        ('loss_partials', c_void_p), ('sample_w', c_void_p), ('sample_w_global', c_void_p),
        ('disp', c_void_p * 4), ('color', c_void_p * 4), ('translation', c_void_p), ('rel_dist', c_void_p),
        ('losses', c_void_p), ('smooth_grad', c_void_p), ('reproj_w', c_void_p),
    ]


# name -> (restype, argtypes); must list every function declared in include/clslam_b200.h
SIGNATURES = {
    'clslam_mask_prepare': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'clslam_smooth_masked': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    'clslam_smooth_masked_apply': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'clslam_last_error': (c_char_p, []),
    'clslam_abi_version': (c_int, []),
    'clslam_launch_count': (c_longlong, []),
    'clslam_set_device': (c_int, [c_int]),
    'clslam_event_create': (c_void_p, []),
    'clslam_event_destroy': (c_int, [c_void_p]),
    'clslam_event_record': (c_int, [c_void_p, c_void_p]),
    'clslam_event_elapsed_ms': (c_int, [c_void_p, c_void_p, POINTER(c_float)]),
    'clslam_disp_to_depth': (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_float, c_float, c_void_p]),
    'clslam_stage_images': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'clslam_set_conv_mode': (c_int, [c_int]),
    'clslam_get_conv_mode': (c_int, []),
    'clslam_debug_tc_stamps': (c_int, [c_void_p]),
    'clslam_conv_fwd_workspace_bytes': (c_longlong, [POINTER(ConvDesc)]),
    'clslam_conv_fwd': (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_longlong, c_void_p]),
    'clslam_split_bf16': (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    'clslam_conv_fwd_bf16w': (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    'clslam_conv3x3_tma': (c_int, [POINTER(ConvTma), c_void_p]),
    'clslam_conv3x3_tma_workspace_bytes': (c_longlong, [POINTER(ConvTma)]),
    'clslam_conv3x3_tma_ws': (c_int, [POINTER(ConvTma), c_void_p, c_longlong, c_void_p]),
    'clslam_dgrad3x3_tma_ws': (c_int, [POINTER(ConvTma), c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, POINTER(Planes),
                                       c_void_p, c_longlong, c_void_p]),
    'clslam_debug_tma_flags': (c_int, [c_int]),
    'clslam_debug_tma_stamps': (c_int, [c_void_p]),
    'clslam_planes_s2d_from_f32': (c_int, [c_void_p, POINTER(Planes), c_int, c_int, c_int, c_int, c_void_p]),
    'clslam_stage_images_s2d': (c_int, [c_void_p, c_void_p, POINTER(Planes), c_int, c_int, c_int, c_int, c_void_p]),
    'clslam_wgrad3x3_tma_workspace_bytes': (c_longlong, [c_int, c_int, c_int, c_int, c_int]),
    'clslam_wgrad3x3_tma': (c_int, [POINTER(Planes), c_int, POINTER(Planes), c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                    c_void_p]),
    'clslam_wt_flip': (c_int, [c_void_p, c_int, c_int, c_void_p, c_longlong, c_void_p, c_longlong, c_void_p]),
    'clslam_dgrad_fold': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                  POINTER(Planes), c_void_p]),
    'clslam_conv_fwd_emit': (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, POINTER(Planes), c_int, c_int, c_void_p, c_longlong, c_void_p]),
    'clslam_maxpool3x3s2_emit': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, POINTER(Planes), c_int, c_void_p]),
    'clslam_planes_from_f32': (c_int, [c_void_p, POINTER(Planes), c_int, c_int, c_void_p]),
    'clslam_maxpool3x3s2': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'clslam_dispconv_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'clslam_conv_dgrad_workspace_bytes': (c_longlong, [POINTER(ConvDesc)]),
    'clslam_conv_dgrad': (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_longlong, c_void_p]),
    'clslam_conv_wgrad_workspace_bytes': (c_longlong, [POINTER(ConvDesc)]),
    'clslam_conv_wgrad': (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_longlong, c_void_p]),
    'clslam_dispconv_wgrad_workspace_bytes': (c_longlong, [c_int, c_int, c_int, c_int]),
    'clslam_dispconv_wgrad': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_void_p, c_longlong, c_void_p]),
    'clslam_dispconv_dgrad_act': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                          c_int, c_void_p]),
    'clslam_pose_head_fwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    'clslam_pose_head_bwd': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                     c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'clslam_photo_tiles': (c_int, [c_int, c_int]),
    'clslam_pack_rgba': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'clslam_photo_identity': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'clslam_photo_fwd': (c_int, [POINTER(PhotoArgs), c_void_p]),
    'clslam_photo_bwd': (c_int, [POINTER(PhotoArgs), c_void_p]),
    'clslam_loss_finalize': (c_int, [POINTER(LossArgs), c_void_p]),
    'clslam_disp_grad': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                 c_void_p]),
    'clslam_resample_u8': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    'clslam_u8_to_f32_chw': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'clslam_color_jitter_workspace_bytes': (c_longlong, [c_int]),
    'clslam_color_jitter_op': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p, c_longlong, c_void_p]),
    'clslam_dispconv_dgrad_act_emit': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, POINTER(Planes), c_int, c_int, c_int,
                                               c_int, c_void_p]),
    'clslam_conv_dgrad_ex_workspace_bytes': (c_longlong, [POINTER(ConvDesc), c_int]),
    'clslam_conv_dgrad_ex': (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                     c_longlong, c_void_p]),
    'clslam_bn_workspace_bytes': (c_longlong, [c_int, c_int]),
    'clslam_bn_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_int, c_int, c_int,
                              c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    'clslam_bn_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    'clslam_maxpool3x3s2_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'clslam_add3': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    'clslam_adam_tick': (c_int, [c_void_p, c_void_p]),
    'clslam_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_double, c_double, c_double,
                                 c_double, c_int, c_void_p, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the C-ABI library; loud failure when it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -m clslam_b200.build` '
                           '(the adaptation path has no CPU / PyTorch fallback)')
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.clslam_abi_version() != 1:
        raise RuntimeError('libclslam_b200.so ABI version mismatch; rebuild')
    _lib = lib
    return lib


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = load().clslam_last_error()
        raise RuntimeError(f'clslam_b200 {what} failed (rc={rc}): {msg.decode() if msg else "?"}')
