"""ctypes binding of `libdiffsg_b200.so` (the C-ABI declared in include/diffsg_b200.h).

The library is built in-tree by `build_library()` (nvcc, sm_100a) and loaded lazily.
There is no fallback: if the shared object is missing or a call fails, a
`DiffsgError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libdiffsg_b200.so"
INCLUDE_DIR = PKG_DIR.parent / "include"
SOURCES = ("diffsg.cu", "side_kernels.cu", "unet_tc.cu", "train_tc.cu")
# Tensor-core engine geometry (diffsg_b200/csrc/unet_tc.cuh): K columns per operand chunk, TMEM columns per
# accumulator region (= widest vector), co-resident CTAs per SM.
TC_VARIANT = dict(chunk=64, aslots=2, region=128, ctas=2)
NVCC_FLAGS = ("-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared")
EXTRA_FLAGS = tuple(filter(None, os.environ.get("DIFFSG_NVCC_FLAGS", "").split()))   # experiments: -DDIFFSG_TC_TIMING ...

ABI_VERSION = 3
OP_GEMM, OP_LNSW, OP_PUSH, OP_POP = 1, 2, 3, 4
F_ACC, F_TIME, F_NOBIAS = 1, 2, 4
BUF_COND, N_BUF = 4, 4
ENGINE_SIMT, ENGINE_TC = 0, 1


class DiffsgError(RuntimeError):
    pass


class Op(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("kind", "src", "dst", "K", "N", "flags", "w_off", "b_off", "t_off", "dcol", "ldw", "pad_")]


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("abi_version", "input_dim", "cond_dim", "max_width", "n_skip", "skip_floats", "tt_stride",
                 "tt_rows", "in_buf", "out_buf", "device")] + [("reserved", C.c_int32 * 5)]


class TcProgramC(C.Structure):
    _fields_ = [("stages", C.c_void_p), ("chunks", C.c_void_p), ("epis", C.c_void_p), ("skip_widths", C.c_void_p),
                ("n_stages", C.c_int32), ("n_chunks", C.c_int32), ("n_epi", C.c_int32), ("n_skip", C.c_int32),
                ("nterms", C.c_int32), ("tt_stride", C.c_int32), ("reserved", C.c_int32 * 2)]


class SampleArgs(C.Structure):
    _fields_ = [("cond_dev", C.c_void_p), ("y_dev", C.c_void_p), ("noise_dev", C.c_void_p),
                ("rec_y_dev", C.c_void_p), ("rec_eps_dev", C.c_void_p), ("stat_ws_dev", C.c_void_p),
                ("coef_host", C.c_void_p), ("B", C.c_int64), ("T", C.c_int32), ("norm_steps", C.c_int32),
                ("omega", C.c_float), ("pad_", C.c_uint32), ("philox_seed", C.c_uint64),
                ("philox_offset", C.c_uint64)]


class Mat(C.Structure):
    """diffsg_mat / diffsg_mat_out: cat(p0[rows, k0], p1[rows, k1]) along columns."""
    _fields_ = [("p0", C.c_void_p), ("p1", C.c_void_p), ("k0", C.c_int32), ("k1", C.c_int32)]


class TlinFwdArgs(C.Structure):
    _fields_ = [("a", Mat), ("w", C.c_void_p), ("bias", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("mean", C.c_void_p), ("rstd", C.c_void_p), ("a2", Mat), ("w2", C.c_void_p), ("bias2", C.c_void_p),
                ("add", C.c_void_p), ("gadd", C.c_void_p), ("gidx", C.c_void_p), ("y", C.c_void_p), ("B", C.c_int64),
                ("N", C.c_int32), ("gadd_ld", C.c_int32)]


class TlinDgradArgs(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("w", C.c_void_p), ("x", Mat), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("mean", C.c_void_p), ("rstd", C.c_void_p), ("dres", Mat), ("dx", Mat), ("dgamma", C.c_void_p),
                ("dbeta", C.c_void_p), ("B", C.c_int64), ("N", C.c_int32), ("K", C.c_int32)]


class TlinWgradArgs(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("a", Mat), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mean", C.c_void_p),
                ("rstd", C.c_void_p), ("gidx", C.c_void_p), ("dw", C.c_void_p), ("dbias", C.c_void_p),
                ("dgadd", C.c_void_p), ("B", C.c_int64), ("N", C.c_int32), ("gadd_rows", C.c_int32), ("dgadd_ld", C.c_int32),
                ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/diffsg_b200.h declares
_P, _I32, _I64, _F, _D, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_uint64
SYMBOLS = {
    "diffsg_last_error": (C.c_char_p, []),
    "diffsg_abi_version": (C.c_int, []),
    "diffsg_plan_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(Op), _I32, C.POINTER(_I32), C.POINTER(_P)]),
    "diffsg_plan_destroy": (C.c_int, [_P]),
    "diffsg_plan_set_weights": (C.c_int, [_P, _P, C.c_size_t, _P, _I32]),
    "diffsg_unet_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _P]),
    "diffsg_sample": (C.c_int, [_P, C.POINTER(SampleArgs), _P]),
    "diffsg_launch_count": (C.c_int64, [C.c_int]),
    "diffsg_philox_normal": (C.c_int, [_P, _I64, _I32, _I32, _U64, _U64, _P]),
    "diffsg_ema_update": (C.c_int, [_P, _P, _I64, _D, _I32, _P]),
    "diffsg_ema_update_multi": (C.c_int, [_P, _P, _P, _I32, _I64, _D, _I32, _P]),
    "diffsg_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "diffsg_mlp_forward": (C.c_int, [_P, _P, _I64, _I32, _I32, C.POINTER(_I32), C.POINTER(_I32), _I32, _I32, _P, _P]),
    "diffsg_minmax": (C.c_int, [_P, _I64, _I32, _I32, _I32, _P, _P]),
    "diffsg_objective_msr": (C.c_int, [_P, _P, _P, _F, _P, _P, _I64, _I32, _P]),
    "diffsg_rate_msr": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "diffsg_decode_nu": (C.c_int, [_P, _P, _F, _F, _F, _P, _I64, _I32, _P]),
    "diffsg_rate_nu": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "diffsg_decode_co": (C.c_int, [_P, _P, _I64, _I32, _P]),
    "diffsg_cost_co": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "diffsg_plan_attach_tc": (C.c_int, [_P, C.POINTER(TcProgramC)]),
    "diffsg_plan_set_tc_weights": (C.c_int, [_P, _P, _P, C.c_size_t, _P, C.c_size_t, _P, _I32, _P, _I32, _I64]),
    "diffsg_plan_status": (C.c_int, [_P, C.POINTER(_I32), _I32, _P]),
    "diffsg_plan_set_engine": (C.c_int, [_P, _I32]),
    "diffsg_lnsw_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "diffsg_lnsw_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _P]),
    "diffsg_tlin_forward": (C.c_int, [C.POINTER(TlinFwdArgs), _P]),
    "diffsg_tlin_dgrad": (C.c_int, [C.POINTER(TlinDgradArgs), _P]),
    "diffsg_tlin_wgrad": (C.c_int, [C.POINTER(TlinWgradArgs), _P]),
    "diffsg_tlin_backward": (C.c_int, [C.POINTER(TlinDgradArgs), _I32, C.POINTER(TlinWgradArgs), _I32, _P]),
    "diffsg_plan_query": (C.c_int, [_P, _I32]),
    "diffsg_sample_steps": (C.c_int, [_P, C.POINTER(SampleArgs), _I32, _I32, _I32, _P]),
    "diffsg_sample_renorm": (C.c_int, [_P, _P, _P, _I64, _I64, _P]),
    "diffsg_debug_tc_gemm": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, C.c_uint32, C.c_uint32, _I32, _P]),
}

_lib = None


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise DiffsgError("nvcc not found; cannot build libdiffsg_b200.so")


HASH_PATH = PKG_DIR / "libdiffsg_b200.srchash"      # digest of the sources the in-tree .so was built from


def _source_digest() -> str:
    import hashlib
    h = hashlib.sha256()
    deps = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))) + [INCLUDE_DIR / "diffsg_b200.h"]
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    h.update(" ".join(NVCC_FLAGS + EXTRA_FLAGS).encode())
    return h.hexdigest()


def _stale() -> bool:
    """True if the .so is missing or was built from other sources / flags (content digest, not mtimes: the snapshot
    that ships the tree to a GPU box does not have to preserve them)."""
    if not LIB_PATH.exists() or not HASH_PATH.exists():
        return True
    return HASH_PATH.read_text().strip() != _source_digest()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into diffsg_b200/libdiffsg_b200.so (in-tree).

    Safe under torchrun: ranks serialise on a lock file, the first one builds into a temporary file and renames it
    into place (atomic), the others find a fresh library when they get the lock."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl
    with open(str(LIB_PATH) + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH
            srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
            tmp = LIB_PATH.with_name(f".{LIB_PATH.name}.{os.getpid()}.tmp")
            cmd = [nvcc_path(), *NVCC_FLAGS, *EXTRA_FLAGS, "-I", str(INCLUDE_DIR), *srcs, "-o", str(tmp)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                tmp.unlink(missing_ok=True)
                raise DiffsgError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
            os.replace(tmp, LIB_PATH)
            HASH_PATH.write_text(_source_digest() + "\n")
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def load():
    """Load the shared library (building it if sources are newer) and bind all symbols."""
    global _lib
    if _lib is not None:
        return _lib
    if _stale():        # missing, or older than a source / the header: rebuild (needs nvcc; a stale binary is never loaded silently)
        try:
            build_library()
        except DiffsgError as e:
            if not LIB_PATH.exists():
                raise DiffsgError(f"{LIB_PATH} is missing and could not be built: {e}") from e
            raise DiffsgError(f"{LIB_PATH} is older than its sources and could not be rebuilt: {e}") from e
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    got = lib.diffsg_abi_version()
    if got != ABI_VERSION:
        raise DiffsgError(f"libdiffsg_b200.so ABI {got} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().diffsg_last_error()
        raise DiffsgError(f"{what or 'diffsg call'} failed (rc={rc}): {msg.decode() if msg else '?'}")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count(reset: bool = False) -> int:
    return int(load().diffsg_launch_count(1 if reset else 0))
