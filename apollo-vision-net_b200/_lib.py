"""ctypes binding of libmsda_b200.so (the C ABI declared in include/msda_b200.h).

The library is built in-tree by :func:`build` (``nvcc`` for sm_100a only) and loaded lazily.
There is no CPU or PyTorch fallback: if the library is missing every operator raises.
"""
import ctypes
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
# APOLLO_B200_LIB: load another build of the same ABI (kernel experiments, tools/dev_variants.sh)
LIB_PATH = os.environ.get('APOLLO_B200_LIB') or os.path.join(_HERE, 'libmsda_b200.so')
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'msda_b200.h')
SOURCES = ['abi.cu', 'msda_fwd.cu', 'msda_bwd.cu', 'point_sampling.cu', 'fused.cu', 'coarse_scatter.cu', 'rowops.cu', 'bev_prep.cu', 'wgrad.cu', 'mha.cu', 'dcnv3.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC']

F32, F16, BF16 = 0, 1, 2

_c_int, _c_vp, _c_f, _c_i64 = ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_int64

_SIGNATURES = {
    'msda_abi_version': (_c_int, []),
    'msda_last_error': (ctypes.c_char_p, []),
    'msda_launch_count': (_c_i64, []),
    'msda_fwd': (_c_int, [_c_vp] * 6 + [_c_int] * 10 + [_c_vp]),
    'msda_bwd': (_c_int, [_c_vp] * 9 + [_c_int] * 10 + [_c_vp]),
    'msda_host_scratch_bytes': (_c_i64, [_c_int] * 10),
    'msda_fwd_host': (_c_int, [_c_vp] * 6 + [_c_int] * 9 + [_c_vp, _c_i64, _c_vp]),
    'msda_fwd_bwd_host': (_c_int, [_c_vp] * 10 + [_c_int] * 9 + [_c_vp, _c_i64, _c_vp]),
    'bev_point_sampling': (_c_int, [_c_vp] * 3 + [_c_f, _c_f] + [_c_int] * 4 + [_c_vp] * 6),
    'sca_fwd': (_c_int, [_c_vp] * 10 + [_c_int] * 12 + [_c_i64, _c_i64, _c_vp]),
    'sca_bwd': (_c_int, [_c_vp] * 12 + [_c_int] * 12 + [_c_i64, _c_i64, _c_int, _c_vp, _c_vp, _c_int, _c_int,
                         _c_vp, _c_vp, _c_vp, _c_vp]),
    'sca_coarse_workspace_bytes': (_c_i64, [_c_int] * 7),
    'bev_hit_lists': (_c_int, [_c_vp, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    'tsa_fwd': (_c_int, [_c_vp] * 7 + [_c_int] * 9 + [_c_f, _c_int, _c_int, _c_i64, _c_i64, _c_vp]),
    'tsa_bwd': (_c_int, [_c_vp] * 10 + [_c_int] * 9 + [_c_f, _c_int, _c_int, _c_i64, _c_i64, _c_int, _c_vp, _c_vp]),
    'grad_amax_scale': (_c_int, [_c_vp, _c_i64, _c_int, _c_f, _c_vp, _c_vp]),
    'grad_amax_scale_zero': (_c_int, [_c_vp, _c_i64, _c_int, _c_f, _c_vp, _c_vp, _c_i64, _c_vp]),
    'unscale_cast': (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_vp, _c_int, _c_i64, _c_i64, _c_vp, _c_vp, _c_int, _c_vp, _c_vp]),
    'unscale_cast_strided': (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_vp, _c_int, _c_i64, _c_i64, _c_vp, _c_vp, _c_int, _c_vp, _c_i64, _c_vp]),
    'rowops_workspace_rows': (_c_int, []),
    'ln_fwd': (_c_int, [_c_vp] * 6 + [_c_i64, _c_int, _c_f, _c_int, _c_vp]),
    'ln_bwd': (_c_int, [_c_vp] * 8 + [_c_i64, _c_int, _c_int, _c_vp]),
    'colsum': (_c_int, [_c_vp] * 3 + [_c_i64, _c_int, _c_int, _c_int, _c_vp]),
    'bev_flatten_level': (_c_int, [_c_vp] * 4 + [_c_int] * 4 + [_c_i64, _c_i64, _c_int, _c_vp]),
    'bev_rotate_nearest': (_c_int, [_c_vp] * 5 + [_c_int] * 5 + [_c_vp]),
    'linear_wgrad_workspace_floats': (_c_i64, [_c_int, _c_int]),
    'linear_wgrad': (_c_int, [_c_vp] * 5 + [_c_i64, _c_int, _c_int, _c_int, _c_vp]),
    'relu_bwd_colsum': (_c_int, [_c_vp] * 5 + [_c_i64, _c_int, _c_int, _c_int, _c_f, _c_vp]),
    'ln_residual_dropout_fwd': (_c_int, [_c_vp] * 8 + [_c_i64, _c_int, _c_f, _c_int, _c_vp, _c_vp, ctypes.c_uint32, _c_f, _c_vp]),
    'ln_bwd_dxsum_dropout': (_c_int, [_c_vp] * 9 + [_c_i64, _c_int, _c_int, _c_vp, ctypes.c_uint32, _c_f, _c_vp]),
    'relu_dropout_fwd': (_c_int, [_c_vp, _c_i64, _c_int, _c_vp, _c_vp, ctypes.c_uint32, _c_f, _c_vp]),
    'dropout_keep_mask': (_c_int, [_c_vp, _c_i64, _c_int, _c_vp, ctypes.c_uint32, _c_f, _c_vp]),
    'ln_residual_fwd': (_c_int, [_c_vp] * 8 + [_c_i64, _c_int, _c_f, _c_int, _c_vp]),
    'ln_bwd_dxsum': (_c_int, [_c_vp] * 8 + [_c_i64, _c_int, _c_int, _c_vp]),
    'mha_impl': (_c_int, [_c_int] * 5),
    'mha_pack_mask': (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp]),
    'mha_fwd': (_c_int, [_c_vp] * 5 + [_c_i64] * 4 + [_c_vp] + [_c_int] * 4 + [_c_i64] * 3 + [_c_int, _c_f, _c_int, _c_int,
                         _c_vp, _c_vp, ctypes.c_uint32, _c_f, _c_vp]),
    'mha_bwd': (_c_int, [_c_vp] * 10 + [_c_i64] * 7 + [_c_vp, _c_vp] + [_c_int] * 4 + [_c_i64] * 3 +
                [_c_int, _c_f, _c_int, _c_int, _c_vp, ctypes.c_uint32, _c_f, _c_vp]),
    'dcnv3_scratch_floats': (_c_i64, [_c_int] * 7),
    'dcnv3_fwd': (_c_int, [_c_vp] * 5 + [_c_int] * 15 + [_c_f, _c_int, _c_vp]),
    'dcnv3_bwd': (_c_int, [_c_vp] * 8 + [_c_int] * 15 + [_c_f, _c_int, _c_vp]),
    'mha_keep_mask': (_c_int, [_c_vp, _c_int, _c_int, _c_vp, ctypes.c_uint32, _c_f, _c_vp]),
}

_lib = None
ABI_VERSION = 8          # must equal MSDA_ABI_VERSION of include/msda_b200.h and the loaded library


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libmsda_b200.so with nvcc (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = [os.path.join(CSRC, s[:-3] + '.o') for s in SOURCES]

    def compile_one(args):
        src, obj = args
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get('MSDA_NVCC_FLAGS', '').split() + [
            '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        return r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(compile_one, zip(SOURCES, objs)))
    if verbose:
        sys.stderr.write('\n'.join(logs))
    r = subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB_PATH]
                       + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    for o in objs:
        os.remove(o)
    global _lib
    _lib = None
    return LIB_PATH


def lib():
    """The loaded library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: the CUDA library has not been built. Run '
                '`python -c "import __graft_entry__ as g; g.build()"` at the repo root. '
                'There is no CPU fallback.')
        l = ctypes.CDLL(LIB_PATH)
        # the .so is a build artefact that travels outside version control: refuse a binary whose
        # argument lists may differ from the ctypes signatures below, and say so when it is older
        # than its sources (the caller decides whether to rebuild: build() does)
        l.msda_abi_version.restype = _c_int
        l.msda_abi_version.argtypes = []
        found = int(l.msda_abi_version())
        if found != ABI_VERSION:
            raise RuntimeError(
                f'{LIB_PATH} has ABI version {found}, this package needs {ABI_VERSION}: stale binary. '
                'Rebuild with `python -c "import __graft_entry__ as g; g.build()"`.')
        if not os.environ.get('APOLLO_B200_LIB') and _stale():
            import warnings
            warnings.warn(f'{LIB_PATH} is older than its sources under {CSRC}; rebuild with build()')
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().msda_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{what} failed (code {rc}): {msg}')


def launch_count():
    return int(lib().msda_launch_count())


class KernelTimer:
    """Optional per-kernel CUDA-event timing on the launching stream (used by bench.py for the
    live roofline measurement).  While installed, every C-ABI launch is bracketed by two events."""

    def __init__(self):
        self.records = {}

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, evs in self.records.items():
            ms = [s.elapsed_time(e) for s, e in evs]
            out[name] = dict(launches=len(ms), mean_us=1e3 * sum(ms) / len(ms), min_us=1e3 * min(ms))
        return out


_timer = None


def set_timer(timer):
    global _timer
    _timer = timer


# Tracing: APOLLO_B200_NVTX=1 wraps every C-ABI launch and every encoder / decoder block
# in an NVTX range, so a timeline (Nsight Systems, or ncu --nvtx filters) shows the reference's module
# names above our kernels.  Off by default: a range costs two host calls per launch.
NVTX = os.environ.get('APOLLO_B200_NVTX', '0') == '1'


class nvtx_range:
    """``with nvtx_range('BEVFormerLayer.cross_attn'): ...`` -- a no-op unless APOLLO_B200_NVTX=1."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_pop()


def call(name, *args):
    """Invoke C entry point `name`; raises on a non-zero return code."""
    fn = getattr(lib(), name)
    if NVTX:
        with nvtx_range('msda::' + name):
            rc = fn(*args)
        check(rc, name)
        return
    if _timer is None:
        rc = fn(*args)
    else:
        import torch
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        _timer.records.setdefault(name, []).append((s, e))
    check(rc, name)


def header_symbols():
    """Function names declared in include/msda_b200.h (used by the ABI export test)."""
    import re
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b([a-z_0-9]+)\s*\(', text)) - {'defined'})


def header_abi_version():
    """MSDA_ABI_VERSION as include/msda_b200.h declares it."""
    import re
    return int(re.search(r'#define\s+MSDA_ABI_VERSION\s+(\d+)', open(HEADER).read()).group(1))
