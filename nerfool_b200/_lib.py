"""ctypes binding of libnerfool_b200.so (the C ABI declared in include/nerfool_b200.h).

This is the stub a maintainer of the reference would add next to ibrnet/projection.py (see
INTEGRATION.md).  There is no fallback: if the shared library is missing or a call fails, a RuntimeError
is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NFB_LIB_PATH') or os.path.join(_HERE, 'libnerfool_b200.so')     # NFB_LIB_PATH: experiment builds (build.py)

_P = c_void_p
_I = c_int

# name -> argtypes (all return int status unless listed in _SPECIAL)
_SIGNATURES = {
    'nfb_coarse_depths': [_I, _I, c_float, c_float, _I, _P, _P, _P],
    'nfb_project_gather_fwd': [_I] * 7 + [_P] * 11,
    'nfb_project_gather_bwd': [_I] * 7 + [_P] * 9,
    'nfb_project_grid_bwd': [_I] * 7 + [_P] * 10,
    'nfb_ibrnet_view_fwd': [_I] * 4 + [_P] * 3 + [_I] * 4 + [_P] * 10 + [_I, _P],
    'nfb_ibrnet_ray_fwd': [_I, _I, _P, _P, _P, _P, _P, _P, _I, _P],
    'nfb_ibrnet_ray_bwd': [_I, _I, _P, _P, _P, _P, _P, _P, _I, _P],
    'nfb_ibrnet_view_bwd': [_I] * 4 + [_P] * 3 + [_I] * 4 + [_P] * 14 + [_I, _P],
    'nfb_ibrnet_ray_wgrad': [_I, _I, _P, _P, _P, _P, _P, _P, _P],
    'nfb_ibrnet_view_wgrad': [_I] * 4 + [_P] * 3 + [_I] * 4 + [_P] * 15,
    'nfb_composite_fwd': [_I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    'nfb_composite_bwd': [_I, _I, _I] + [_P] * 8,
    'nfb_sample_pdf': [_I, _I, _I, _P, _P, _P, _I, _P, _P, _P],
    'nfb_fine_depths': [_I, _I, _I, _I, _P, _P, _P, _I, _P, _P],
    'nfb_gnt_fwd': [_I] * 5 + [_P] * 8 + [ctypes.c_size_t, _I, _P],
    'nfb_gnt_bwd': [_I] * 5 + [_P] * 10 + [ctypes.c_size_t, _P],
    'nfb_gnt_fwd_save': [_I] * 5 + [_P] * 8 + [ctypes.c_size_t, _P],
    'nfb_gnt_bwd_saved': [_I] * 5 + [_P] * 10 + [ctypes.c_size_t, _P],
    'nfb_forward_warp': [_I, _I] + [_P] * 6 + [_I] + [_P] * 4 + [_I, _P],
}
EXPORTS = ['nfb_version', 'nfb_last_error_string', 'nfb_ibrnet_param_offset', 'nfb_view_stash_bytes', 'nfb_ray_stash_bytes',
           'nfb_gnt_param_floats', 'nfb_gnt_param_offset', 'nfb_gnt_workspace_bytes', 'nfb_gnt_bwd_workspace_bytes'] + list(_SIGNATURES)

_lib = None

# Arithmetic of the view-stage dense layers (include/nerfool_b200.h: NfbPrecision).
PRECISIONS = {'fp32': 0, 'bf16x3': 1, 'bf16': 2}
_precision = PRECISIONS[os.environ.get('NFB_PRECISION', 'bf16x3')]


def set_precision(name: str) -> None:
    """'fp32' (CUDA-core FMA), 'bf16x3' (tcgen05, split operands, fp32-equivalent; default) or 'bf16'."""
    global _precision
    if name not in PRECISIONS:
        raise ValueError(f'precision must be one of {sorted(PRECISIONS)}, got {name!r}')
    _precision = PRECISIONS[name]


def get_precision() -> str:
    return {v: k for k, v in PRECISIONS.items()}[_precision]


def precision_code() -> int:
    return _precision


# Largest activation stash (GiB, per render level) the fused backward may allocate; beyond it the backward recomputes
# the forward instead (include/nerfool_b200.h: nfb_view_stash_bytes).  0 disables the stash.
STASH_MAX_GIB = float(os.environ.get('NFB_STASH_MAX_GIB', '48'))


def stash_bytes(N: int, V: int) -> int:
    """Size of the activation stash for N samples x V views, or 0 when it is disabled / too large / fp32 mode."""
    if _precision == PRECISIONS['fp32'] or STASH_MAX_GIB <= 0:
        return 0
    n = int(load().nfb_view_stash_bytes(int(N), int(V)))
    return n if n <= STASH_MAX_GIB * 2 ** 30 else 0


def ray_stash_bytes(R: int, S: int) -> int:
    """Size of the ray-stage activation stash (0 when disabled, fp32 mode or S > 128)."""
    if _precision == PRECISIONS['fp32'] or STASH_MAX_GIB <= 0:
        return 0
    n = int(load().nfb_ray_stash_bytes(int(R), int(S)))
    return n if n <= STASH_MAX_GIB * 2 ** 30 else 0


def load():
    """Load (once) and return the ctypes library handle."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -m nerfool_b200.build` '
            '(nvcc, sm_100a).  nerfool_b200 has no CPU / PyTorch fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    lib.nfb_version.restype = c_int
    lib.nfb_version.argtypes = []
    lib.nfb_last_error_string.restype = c_char_p
    lib.nfb_last_error_string.argtypes = []
    lib.nfb_ibrnet_param_offset.restype = c_int
    lib.nfb_ibrnet_param_offset.argtypes = [c_char_p]
    lib.nfb_view_stash_bytes.restype = ctypes.c_size_t
    lib.nfb_view_stash_bytes.argtypes = [c_int, c_int]
    lib.nfb_ray_stash_bytes.restype = ctypes.c_size_t
    lib.nfb_ray_stash_bytes.argtypes = [c_int, c_int]
    lib.nfb_gnt_param_floats.restype = c_int
    lib.nfb_gnt_param_floats.argtypes = [c_int]
    lib.nfb_gnt_param_offset.restype = c_int
    lib.nfb_gnt_param_offset.argtypes = [c_int, c_char_p]
    lib.nfb_gnt_workspace_bytes.restype = ctypes.c_size_t
    lib.nfb_gnt_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.nfb_gnt_bwd_workspace_bytes.restype = ctypes.c_size_t
    lib.nfb_gnt_bwd_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


PARAM_FLOATS = 20136  # NFB_IBRNET_PARAM_FLOATS (include/nerfool_b200.h)
LAUNCHES = 0          # kernels launched through the C ABI since import (every entry point = one launch)
_profile = None       # when a dict: entry-point name -> list of (start_event, end_event)


def profile_start():
    """Start recording a CUDA-event pair around every C-ABI call (on the launching stream)."""
    global _profile
    _profile = {}


def profile_stop():
    """Stop recording; returns {entry point: [ms per launch, ...]} (synchronises the device)."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    return {k: [a.elapsed_time(b) for a, b in v] for k, v in (rec or {}).items()}


def call(name, *args):
    """Invoke an entry point; raise RuntimeError with the library's message on a non-zero status."""
    global LAUNCHES
    lib = load()
    fn = getattr(lib, name)
    if _profile is not None:
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        _profile.setdefault(name, []).append((a, b))
    else:
        rc = fn(*args)
    LAUNCHES += 1
    if rc != 0:
        msg = lib.nfb_last_error_string().decode('utf-8', 'replace')
        raise RuntimeError(f'{name} failed ({rc}): {msg}')


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('nerfool_b200 ops need CUDA tensors (no CPU fallback); got a tensor on ' + str(t.device))


def f32c(t):
    """fp32 + contiguous (no copy when already so)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
