"""ctypes binding of libe3b.so (the C ABI declared in include/e3b.h).

The product path has NO fallback: if the shared library is missing or a call fails this module
raises.  Build the library with ``python -m elektronn3_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libe3b.so')

c_int, c_i32, c_i64, c_float, c_void_p = ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class ConvArgs(ctypes.Structure):
    """struct e3b_conv_args (include/e3b.h)"""
    _fields_ = [
        ('src0', c_void_p), ('C0', c_i32),
        ('src1', c_void_p), ('C1', c_i32),
        ('N', c_i32), ('D', c_i32), ('H', c_i32), ('W', c_i32),
        ('D1', c_i32), ('H1', c_i32), ('W1', c_i32),
        ('off1_d', c_i32), ('off1_h', c_i32), ('off1_w', c_i32),
        ('kd', c_i32), ('kh', c_i32), ('kw', c_i32),
        ('pd', c_i32), ('ph', c_i32), ('pw', c_i32),
        ('wpk', c_void_p),
        ('bias', c_void_p), ('n_bias', c_i32),
        ('n_total', c_i32),
        ('dst0', c_void_p), ('Cd0', c_i32),
        ('dst1', c_void_p), ('Cd1', c_i32),
        ('relu', c_i32), ('half_out', c_i32),
        ('out_scale', c_void_p),
        ('w_unscale', c_void_p),
        ('stats', c_void_p), ('stats_channels', c_i32),
        ('scatter', c_i32), ('sd', c_i32), ('sh', c_i32), ('sw', c_i32),
        ('Ds', c_i32), ('Hs', c_i32), ('Ws', c_i32),
        ('force_tz', c_i32),
        ('variant', c_i32),
    ]


class WgradArgs(ctypes.Structure):
    """struct e3b_wgrad_args"""
    _fields_ = [
        ('src0', c_void_p), ('C0', c_i32),
        ('src1', c_void_p), ('C1', c_i32),
        ('N', c_i32), ('D', c_i32), ('H', c_i32), ('W', c_i32),
        ('D1', c_i32), ('H1', c_i32), ('W1', c_i32),
        ('off1_d', c_i32), ('off1_h', c_i32), ('off1_w', c_i32),
        ('dy', c_void_p), ('Co', c_i32),
        ('kd', c_i32), ('kh', c_i32), ('kw', c_i32),
        ('pd', c_i32), ('ph', c_i32), ('pw', c_i32),
        ('dw', c_void_p), ('layout', c_i32), ('up_taps', c_i32), ('up_co', c_i32),
        ('workspace', c_void_p),
        ('dy_unscale', c_void_p),
        ('defer_reduce', c_i32),
    ]


class NormBwdArgs(ctypes.Structure):
    """struct e3b_norm_bwd_args"""
    _fields_ = [
        ('y', c_void_p), ('scale', c_void_p), ('shift', c_void_p),
        ('g0', c_void_p), ('g1', c_void_p), ('gp', c_void_p),
        ('pool_idx', c_void_p),
        ('N', c_i32), ('C', c_i32), ('D', c_i32), ('H', c_i32), ('W', c_i32),
        ('pk_d', c_i32), ('pk_h', c_i32), ('pk_w', c_i32),
        ('mode', c_i32), ('G', c_i32), ('eps', c_float),
        ('gamma', c_void_p), ('mean', c_void_p), ('rstd', c_void_p),
        ('fwd_stats', c_void_p),
        ('sums', c_void_p),
        ('amax', c_void_p), ('dy_scale', c_void_p),
        ('m1', c_void_p), ('m2', c_void_p),
        ('dgamma', c_void_p), ('dbeta', c_void_p), ('dbias', c_void_p),
        ('dy', c_void_p), ('s2d', c_i32), ('sd', c_i32), ('sh', c_i32), ('sw', c_i32),
        ('relu', c_i32),
        ('g1_crop', c_i32),
        ('g1_od', c_i32), ('g1_oh', c_i32), ('g1_ow', c_i32), ('g1_D', c_i32), ('g1_H', c_i32), ('g1_W', c_i32),
        ('act_slope', c_float),
        ('act_slope_dev', c_void_p), ('slope_sums', c_void_p), ('dslope', c_void_p),
    ]


class PackJob(ctypes.Structure):
    """struct e3b_pack_job"""
    _fields_ = [
        ('w', c_void_p), ('scale', c_void_p), ('wscale', c_void_p), ('dst', c_void_p),
        ('mode', c_i32), ('C0', c_i32), ('C1', c_i32), ('Co', c_i32), ('kd', c_i32), ('kh', c_i32), ('kw', c_i32),
    ]


class WsJob(ctypes.Structure):
    """struct e3b_ws_job"""
    _fields_ = [('w', c_void_p), ('n', c_i64)]


class HeadArgs(ctypes.Structure):
    """struct e3b_head_args"""
    _fields_ = [
        ('a', c_void_p), ('N', c_i32), ('C', c_i32), ('D', c_i32), ('H', c_i32), ('W', c_i32),
        ('w', c_void_p), ('b', c_void_p), ('Co', c_i32),
        ('out_mode', c_i32),
        ('dst', c_void_p), ('Dd', c_i32), ('Hd', c_i32), ('Wd', c_i32),
        ('c0_d', c_i32), ('c0_h', c_i32), ('c0_w', c_i32),
        ('cn_d', c_i32), ('cn_h', c_i32), ('cn_w', c_i32),
        ('dst_origin', c_void_p),
        ('dst_single', c_i32),
        ('flip', c_i32),
        ('accumulate', c_i32), ('acc_scale', c_float),
        ('use_threshold', c_i32), ('threshold', c_float),
        ('round_half', c_i32),
    ]


# name -> (restype, argtypes); every symbol include/e3b.h declares
SIGNATURES = {
    'e3b_version': (c_int, []),
    'e3b_last_error': (ctypes.c_char_p, []),
    'e3b_launch_count': (c_i64, []),
    'e3b_pack_ncdhw': (c_int, [c_void_p, c_void_p] + [c_int] * 11 + [c_void_p]),
    'e3b_unpack_qp': (c_int, [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]),
    'e3b_gather_tiles': (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [c_void_p]),
    'e3b_packed_weight_floats': (c_i64, [c_int] * 7),
    'e3b_pack_weights': (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    'e3b_pack_job_table_bytes': (c_i64, [c_int]),
    'e3b_pack_jobs_fill': (c_int, [ctypes.POINTER(PackJob), c_int, c_void_p, ctypes.POINTER(c_i64)]),
    'e3b_pack_weights_batched': (c_int, [c_void_p, c_int, c_i64, c_void_p]),
    'e3b_weight_scales': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'e3b_conv': (c_int, [ctypes.POINTER(ConvArgs), c_void_p]),
    'e3b_conv_variant': (c_int, [c_int] * 7),
    'e3b_debug_zs_read': (c_int, [c_void_p, c_int]),
    'e3b_debug_zs_prof': (c_int, [c_void_p, c_int]),
    'e3b_debug_fused_prof': (c_int, [ctypes.POINTER(ctypes.c_ulonglong)]),
    'e3b_debug_conv_counters': (c_int, [c_void_p, c_int]),
    'e3b_wgrad_workspace_floats': (c_i64, [ctypes.POINTER(WgradArgs)]),
    'e3b_wgrad': (c_int, [ctypes.POINTER(WgradArgs), c_void_p]),
    'e3b_wgrad_reduce_batched': (c_int, [ctypes.POINTER(WgradArgs), c_int, c_void_p]),
    'e3b_norm_finalize': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_i64, c_void_p, c_void_p, c_float,
                                  c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'e3b_norm_act': (c_int, [c_void_p] * 6 + [c_int] * 9 + [c_float, c_void_p, c_int, c_void_p]),
    'e3b_norm_bwd_fused': (c_int, [ctypes.POINTER(NormBwdArgs), c_void_p]),
    'e3b_norm_bwd_reduce': (c_int, [ctypes.POINTER(NormBwdArgs), c_void_p]),
    'e3b_norm_bwd_finalize': (c_int, [ctypes.POINTER(NormBwdArgs), c_void_p]),
    'e3b_norm_bwd_apply': (c_int, [ctypes.POINTER(NormBwdArgs), c_void_p]),
    'e3b_add_qh': (c_int, [c_void_p] * 3 + [c_int] * 11 + [c_void_p]),
    'e3b_upsample_qh': (c_int, [c_void_p] * 2 + [c_int] * 18 + [c_void_p]),
    'e3b_upsample_bwd_qp': (c_int, [c_void_p] * 2 + [c_int] * 18 + [c_void_p]),
    'e3b_residual_add': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_i64, c_void_p]),
    'e3b_qp_axpy': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_i64, c_void_p]),
    'e3b_head': (c_int, [ctypes.POINTER(HeadArgs), c_void_p]),
    'e3b_prob_argmax': (c_int, [c_void_p, c_void_p, c_int, c_int, c_i64, c_int, c_float, c_void_p]),
    'e3b_head_bwd': (c_int, [c_void_p] * 7 + [c_int] * 6 + [c_void_p]),
    'e3b_dice_fwd': (c_int, [c_void_p] * 4 + [c_int, c_int, c_int, c_i64, c_int, c_float, c_float] + [c_void_p] * 4),
    'e3b_dice_bwd': (c_int, [c_void_p] * 6 + [c_int, c_int, c_i64, c_int, c_void_p]),
}

_lib = None


def lib():
    """The loaded shared library.  Raises if it has not been built (no CPU / torch fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: the CUDA extension has not been built. Run '
                '`python -m elektronn3_b200.build`. elektronn3_b200 has no CPU or PyTorch fallback.')
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().e3b_last_error()
        raise RuntimeError(f'libe3b {what}: {msg.decode() if msg else "error"}')


def launch_count():
    return int(lib().e3b_launch_count())
