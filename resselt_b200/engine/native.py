"""ctypes binding of the C ABI declared in include/resselt_b200.h.

The shared library is built in-tree (resselt_b200/csrc/libresselt_b200.so) by ``build_library`` /
``__graft_entry__.build``.  There is no CPU implementation behind this module: if the library is
missing or no sm_100 device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import threading

CSRC_DIR = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'csrc'))
BRINGUP = bool(os.environ.get('RSB_BRINGUP'))  # bring-up flavour: separate library, never the one the product loads
LIB_PATH = os.path.join(CSRC_DIR, 'libresselt_b200_bringup.so' if BRINGUP else 'libresselt_b200.so')
SOURCES = ('conv_tc.cu', 'conv_rs.cu', 'conv_pair.cu', 'conv_lk.cu', 'conv_direct.cu', 'dat_ops.cu', 'winattn_tc.cu', 'rt_ops.cu', 'plan.cu')
HEADERS = ('kernels.cuh', 'ptx.cuh', os.path.join('..', '..', 'include', 'resselt_b200.h'))
NVCC_FLAGS = ('-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC')
OBJ_DIR = os.path.join(CSRC_DIR, 'build')  # git-ignored object files, one per source

# enums (keep in sync with include/resselt_b200.h)
F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_MISH, ACT_LRELU, ACT_PRELU, ACT_SIGMOID, ACT_GELU = range(7)
COMB_NONE, COMB_SPAB_GATE, COMB_MUL, COMB_AXPY = range(4)
EXTERNAL_INPUT, EXTERNAL_OUTPUT, NO_BUFFER = -1, -2, -3
OP_LAYERNORM, OP_DWCONV3, OP_WINATTN, OP_CHANATTN, OP_AIM, OP_DYSAMPLE, OP_RMSNORM, OP_UNSHUFFLE_POOL, OP_SE_SHUFFLE = 1, 2, 3, 4, 5, 6, 7, 8, 9
OP_CHAN_GATE, OP_CHAN_AFFINE = 10, 11

EXPORTED_SYMBOLS = (
    'rsb_version', 'rsb_last_error', 'rsb_device_count', 'rsb_plan_create', 'rsb_plan_destroy',
    'rsb_plan_add_buffer', 'rsb_plan_add_conv', 'rsb_plan_add_groupnorm', 'rsb_plan_add_op', 'rsb_plan_finalize',
    'rsb_plan_num_ops', 'rsb_plan_launches_per_forward', 'rsb_plan_flops', 'rsb_plan_workspace_bytes',
    'rsb_plan_forward', 'rsb_plan_forward_ops', 'rsb_plan_read_buffer',
    'rsb_plan_num_direct_convs', 'rsb_plan_op_info', 'rsb_kernel_name', 'rsb_plan_set_nvtx', 'rsb_abi_struct_size', 'rsb_plan_set_base_divisor',
)


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'resselt_b200 native error {code}: {message}')
        self.code = code


class ConvDesc(C.Structure):
    _fields_ = [
        ('src_buf', C.c_int32), ('src_ch_off', C.c_int32), ('cin', C.c_int32),
        ('dst_buf', C.c_int32), ('dst_ch_off', C.c_int32), ('cout', C.c_int32),
        ('kh', C.c_int32), ('kw', C.c_int32),
        ('weight', C.POINTER(C.c_float)), ('bias', C.POINTER(C.c_float)),
        ('act', C.c_int32), ('act_param', C.c_float), ('act_slopes', C.POINTER(C.c_float)),
        ('combine', C.c_int32),
        ('res1_buf', C.c_int32), ('res1_ch_off', C.c_int32),
        ('res2_buf', C.c_int32), ('res2_ch_off', C.c_int32),
        ('alpha', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float),
        ('in_mean', C.c_float * 4), ('in_scale', C.c_float),
        ('ps', C.c_int32), ('add_base', C.c_int32), ('out_scale', C.c_float), ('out_mean', C.c_float * 4),
        ('src_upsample2', C.c_int32),
        ('dst_ps', C.c_int32), ('dst2_buf', C.c_int32), ('dst2_ch_off', C.c_int32), ('split_ch', C.c_int32),
        ('dst_phase', C.c_int32), ('pad_t', C.c_int32), ('pad_l', C.c_int32),
        ('border_bias', C.POINTER(C.c_float)),
        ('ln_fold', C.c_int32), ('ln_stats_buf', C.c_int32),
        ('ln_out', C.c_int32), ('ln_out_buf', C.c_int32), ('ln_eps', C.c_float),
    ]


class GroupNormDesc(C.Structure):
    _fields_ = [
        ('src_buf', C.c_int32), ('src_ch_off', C.c_int32),
        ('dst_buf', C.c_int32), ('dst_ch_off', C.c_int32),
        ('channels', C.c_int32), ('groups', C.c_int32), ('eps', C.c_float),
        ('gamma', C.POINTER(C.c_float)), ('beta', C.POINTER(C.c_float)),
        ('skip_buf', C.c_int32), ('skip_ch_off', C.c_int32),
    ]


class OpDesc(C.Structure):
    _fields_ = [
        ('kind', C.c_int32),
        ('src_buf', C.c_int32), ('src_ch_off', C.c_int32),
        ('src2_buf', C.c_int32), ('src2_ch_off', C.c_int32),
        ('dst_buf', C.c_int32), ('dst_ch_off', C.c_int32),
        ('channels', C.c_int32),
        ('i', C.c_int32 * 8), ('f', C.c_float * 4),
        ('w', C.POINTER(C.c_float) * 8), ('wn', C.c_int64 * 8),
    ]


class OpInfo(C.Structure):
    _fields_ = [
        ('kind', C.c_int32), ('kernel', C.c_int32), ('fused_next', C.c_int32), ('launches', C.c_int32),
        ('flops', C.c_double), ('bytes', C.c_double),
    ]


def _deps_mtime() -> float:
    return max(os.path.getmtime(os.path.join(CSRC_DIR, h)) for h in HEADERS if os.path.exists(os.path.join(CSRC_DIR, h)))


def library_is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC_DIR, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into one shared library (nvcc cross-compiles without a GPU).

    One object per source, compiled in parallel and re-used while neither the source nor a header changed; RSB_BRINGUP=1 in
    the environment of the BUILD adds -DRSB_BRINGUP (the kernel-selection environment switches, see kernels.cuh)."""
    if not force and not library_is_stale():
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libresselt_b200.so')
    flags = list(NVCC_FLAGS) + (['-DRSB_BRINGUP'] if BRINGUP else [])
    tag = 'bringup' if BRINGUP else 'product'
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr = _deps_mtime()
    jobs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, f'{os.path.splitext(src)[0]}.{tag}.o')
        fresh = (not force and os.path.exists(obj)
                 and os.path.getmtime(obj) >= max(hdr, os.path.getmtime(os.path.join(CSRC_DIR, src))))
        jobs.append((src, obj, fresh))
    procs = [(src, subprocess.Popen([nvcc, *flags, '-c', src, '-o', obj], cwd=CSRC_DIR, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
             for src, obj, fresh in jobs if not fresh]
    log, failed = '', []
    for src, proc in procs:
        out, _ = proc.communicate()
        log += out
        if proc.returncode != 0:
            failed.append(src)
    if failed:
        raise RuntimeError(f'nvcc failed on {failed}:\n' + log)
    link = subprocess.run([nvcc, '-shared', '-o', LIB_PATH + '.tmp', *[obj for _, obj, _ in jobs], '-ldl'], cwd=CSRC_DIR, capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError('link failed:\n' + link.stdout + link.stderr)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    if verbose:
        print(log + link.stdout + link.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load the native library (never falls back to anything else)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                '(resselt_b200 has no CPU or PyTorch fallback path)'
            )
        L = C.CDLL(LIB_PATH)
        L.rsb_version.restype = C.c_int
        L.rsb_last_error.restype = C.c_char_p
        L.rsb_device_count.restype = C.c_int
        L.rsb_plan_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.rsb_plan_destroy.argtypes = [C.c_void_p]
        L.rsb_plan_add_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.rsb_plan_add_conv.argtypes = [C.c_void_p, C.POINTER(ConvDesc)]
        L.rsb_plan_add_groupnorm.argtypes = [C.c_void_p, C.POINTER(GroupNormDesc)]
        L.rsb_plan_add_op.argtypes = [C.c_void_p, C.POINTER(OpDesc)]
        L.rsb_plan_finalize.argtypes = [C.c_void_p, C.c_int]
        L.rsb_plan_num_ops.argtypes = [C.c_void_p]
        L.rsb_plan_launches_per_forward.argtypes = [C.c_void_p]
        L.rsb_plan_flops.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.rsb_plan_workspace_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.rsb_plan_forward.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
            C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
        ]
        L.rsb_plan_forward_ops.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
            C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int,
        ]
        L.rsb_plan_read_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.rsb_plan_num_direct_convs.argtypes = [C.c_void_p]
        L.rsb_plan_set_nvtx.argtypes = [C.c_void_p, C.c_int]
        L.rsb_abi_struct_size.argtypes = [C.c_int]
        L.rsb_plan_set_base_divisor.argtypes = [C.c_void_p, C.c_int]
        L.rsb_plan_op_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(OpInfo)]
        L.rsb_kernel_name.argtypes = [C.c_int]
        L.rsb_kernel_name.restype = C.c_char_p
        for name in EXPORTED_SYMBOLS:
            getattr(L, name)  # AttributeError if the library does not export the declared ABI
        for which, struct in enumerate((ConvDesc, GroupNormDesc, OpDesc, OpInfo)):
            if L.rsb_abi_struct_size(which) != C.sizeof(struct):
                raise RuntimeError(f'{struct.__name__}: ctypes layout is {C.sizeof(struct)} bytes, the library was built with {L.rsb_abi_struct_size(which)}')
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise NativeError(code, (lib().rsb_last_error() or b'').decode('utf-8', 'replace'))
