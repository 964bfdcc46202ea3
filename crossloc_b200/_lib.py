"""ctypes binding of libcrossloc_b200.so (the C ABI declared in include/crossloc_b200.h).

There is no fallback: if the library is missing or a symbol is absent the import fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_C', 'libcrossloc_b200.so')

_c = ctypes
_f32p = _c.POINTER(_c.c_float)
_f64p = _c.POINTER(_c.c_double)
_i32p = _c.POINTER(_c.c_int32)

# name -> (restype, argtypes); mirrors include/crossloc_b200.h one to one
SIGNATURES = {
    'cl_version': (_c.c_char_p, []),
    'cl_last_error': (_c.c_char_p, []),
    'cl_dsac_forward_rgb': (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_float,
        _c.c_void_p, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_int,
        _c.c_uint64, _c.c_uint32, _c.c_uint32, _c.c_int,
        _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_dsac_backward_rgb': (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_float,
        _c.c_void_p, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_int,
        _c.c_uint64, _c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_dsac_timing': (_c.c_int, [_c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_release_workspaces': (_c.c_int, []),
    'cl_conv_igemm': (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _i32p, _c.c_int,
        _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_void_p]),
    'cl_conv_igemm_fp4': (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _i32p,
        _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_pack_conv_fp4': (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_void_p, _c.c_void_p,
                                    _c.c_void_p]),
    'cl_conv_wgrad_pf': (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                    _c.c_int, _i32p, _i32p, _c.c_int, _c.c_float, _c.c_void_p, _c.c_int, _c.c_void_p,
                                    _c.c_void_p]),
    'cl_gn_backward': (_c.c_int, [
        _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_float, _c.c_int, _c.c_int, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _i32p,
        _i32p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p,
        _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_gn_backward_fp4': (_c.c_int, [
        _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_float, _c.c_int, _c.c_int, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_void_p), _i32p,
        _i32p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p,
        _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p]),
    'cl_head_backward': (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
                                    _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_nchw_to_pf': (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                 _c.c_void_p]),
    'cl_pf_to_nchw': (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int,
                                 _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_pow2_scale': (_c.c_int, [_c.c_void_p, _c.c_int64, _c.c_float, _c.c_void_p, _c.c_void_p]),
    'cl_pack_filter': (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _i32p,
                                  _i32p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    'cl_gn_apply': (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_float, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p]),
    'cl_gn_apply_fp4': (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_float, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p,
        _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
        _c.c_void_p, _c.c_void_p]),
    'cl_pf_groupnorm': (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_float,
        _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p]),
    'cl_stem_forward': (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p,
        _c.c_void_p, _c.c_void_p, _c.c_float, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p]),
    'cl_head_forward': (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p,
        _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p]),
    'cl_frames_to_nchw': (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                     _c.c_void_p]),
    'cl_duc_head_forward': (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
        _c.c_void_p, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_float, _c.c_float,
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p]),
    # input frames (csrc/frames.cu)
    'cl_png_info': (_c.c_int, [_c.c_void_p, _c.c_size_t, _i32p, _i32p, _i32p]),
    'cl_decode_png': (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_int, _c.c_int]),
    'cl_decode_png_batch': (_c.c_int, [_c.POINTER(_c.c_void_p), _c.POINTER(_c.c_size_t), _c.c_int, _c.c_void_p, _c.c_int,
                                       _c.c_int, _c.c_int]),
    'cl_resize_frames': (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p,
                                    _c.c_size_t, _c.c_void_p]),
    'cl_resize_workspace_bytes': (_c.c_size_t, [_c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    'cl_resize_coeffs': (_c.c_int, [_c.c_int, _c.c_int, _i32p, _i32p, _i32p]),
    # whole-network runtime (csrc/net.cu); the descriptor structures live in crossloc_b200/net.py
    'cl_net_create': (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    'cl_net_update': (_c.c_int, [_c.c_void_p]),
    'cl_net_forward': (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p]),
    'cl_net_forward_frames': (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                         _c.c_void_p, _c.c_void_p]),
    'cl_net_wait_fork': (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    'cl_net_output_shape': (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_net_buffers': (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                  _c.c_void_p]),
    'cl_net_profile': (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,
                                  _c.c_void_p, _c.c_void_p, _c.c_void_p]),
    'cl_net_destroy': (None, [_c.c_void_p]),
}

_lib = None


def load():
    """Load the shared library and bind every declared entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'crossloc_b200: %s is missing -- build it with `python -m crossloc_b200.build` '
            '(or __graft_entry__.build()); there is no CPU or PyTorch fallback for this path.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise RuntimeError('crossloc_b200: ' + load().cl_last_error().decode())
