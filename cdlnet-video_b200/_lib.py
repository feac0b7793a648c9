"""Build and load libcdl_b200.so (the C-ABI library declared in include/cdl_b200.h) with ctypes.

The library is built IN-TREE with nvcc for sm_100a only; there is no Triton / torch.compile /
CPU fallback.  `load()` raises if the shared object is missing or fails to load.
"""
import ctypes
import os
import shutil
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("CDL_LIB_PATH") or os.path.join(HERE, "libcdl_b200.so")     # CDL_LIB_PATH: an alternate build (e.g. -DCDL_TC_PROFILE)
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ldl"]

SYMBOLS = [
    "cdl_abi_version", "cdl_status_string", "cdl_plan_create", "cdl_plan_destroy", "cdl_plan_layout",
    "cdl_plan_workspace_bytes", "cdl_plan_precision", "cdl_set_weights", "cdl_reduce_sums",
    "cdl_mean_from_sums", "cdl_center_pad", "cdl_preprocess", "cdl_analysis_step", "cdl_synthesis_step",
    "cdl_forward", "cdl_postprocess", "cdl_denoise", "cdl_plan_host_workspace_bytes", "cdl_denoise_host",
    "cdl_plan_launch_count", "cdl_plan_code_bytes", "cdl_code_export", "cdl_code_import", "cdl_plan_set_rearm",
    "cdl_plan_reduce_workspace_bytes", "cdl_plan_step_workspace_bytes", "cdl_plan_host_workspace_bytes_noz",
    "cdl_comm_unique_id", "cdl_comm_create", "cdl_comm_destroy", "cdl_comm_allreduce_f64", "cdl_halo_bytes",
    "cdl_halo_exchange", "cdl_analysis_step_halo", "cdl_halo_add", "cdl_forward_sharded",
    "cdl_analysis_step_csr", "cdl_preprocess_noisy", "cdl_nle_mad_workspace_bytes", "cdl_nle_mad",
]


class CdlDesc(ctypes.Structure):
    """mirror of cdl_desc_t (include/cdl_b200.h)"""
    _fields_ = [("ndim", ctypes.c_int32), ("N", ctypes.c_int32), ("C", ctypes.c_int32), ("M", ctypes.c_int32),
                ("K", ctypes.c_int32), ("dims", ctypes.c_int32 * 3), ("P", ctypes.c_int32 * 3), ("s", ctypes.c_int32),
                ("has_mask", ctypes.c_int32), ("precision", ctypes.c_int32), ("halo_front", ctypes.c_int32),
                ("halo_back", ctypes.c_int32), ("device", ctypes.c_int32)]


class CdlLayout(ctypes.Structure):
    """mirror of cdl_layout_t"""
    _fields_ = [("pad", ctypes.c_int32 * 6), ("fine", ctypes.c_int32 * 3), ("coarse", ctypes.c_int32 * 3)]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "cdl_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> cdlnet-video_b200/libcdl_b200.so"""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcdl_b200.so")
    tmp = LIB_PATH + ".tmp%d" % os.getpid()
    extra = ["-DCDL_TC_PROFILE"] if os.environ.get("CDL_TC_PROFILE") else []      # per-role cycle counters (dev aid)
    extra += os.environ.get("CDL_NVCC_DEFS", "").split()                              # e.g. -DCDL_ANA_PF=1 (tuning experiments)
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", INCLUDE, "-o", tmp, *sources()]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library and declare the prototypes.  Raises if it is absent (no fallback)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU or PyTorch fallback for the CUDA path.")
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
        P = ctypes.POINTER
        lib.cdl_abi_version.restype = i32
        lib.cdl_status_string.restype = cp
        lib.cdl_status_string.argtypes = [i32]
        lib.cdl_plan_create.restype = i32
        lib.cdl_plan_create.argtypes = [P(vp), P(CdlDesc)]
        lib.cdl_plan_destroy.restype = None
        lib.cdl_plan_destroy.argtypes = [vp]
        lib.cdl_plan_layout.restype = i32
        lib.cdl_plan_layout.argtypes = [vp, P(CdlLayout)]
        for name in ("cdl_plan_workspace_bytes", "cdl_plan_host_workspace_bytes", "cdl_plan_code_bytes",
                     "cdl_plan_reduce_workspace_bytes", "cdl_plan_step_workspace_bytes", "cdl_plan_host_workspace_bytes_noz"):
            getattr(lib, name).restype = i32
            getattr(lib, name).argtypes = [vp, P(ctypes.c_size_t)]
        lib.cdl_plan_precision.restype = i32
        lib.cdl_plan_precision.argtypes = [vp]
        lib.cdl_plan_set_rearm.restype = i32
        lib.cdl_plan_set_rearm.argtypes = [vp, i32]
        lib.cdl_plan_launch_count.restype = i32
        lib.cdl_plan_launch_count.argtypes = [vp, P(ctypes.c_uint64)]
        lib.cdl_set_weights.restype = i32
        lib.cdl_set_weights.argtypes = [vp, P(vp), P(vp), vp, vp]
        lib.cdl_reduce_sums.restype = i32
        lib.cdl_reduce_sums.argtypes = [vp] * 6
        lib.cdl_mean_from_sums.restype = i32
        lib.cdl_mean_from_sums.argtypes = [vp] * 4
        lib.cdl_center_pad.restype = i32
        lib.cdl_center_pad.argtypes = [vp] * 7
        lib.cdl_preprocess.restype = i32
        lib.cdl_preprocess.argtypes = [vp] * 8
        lib.cdl_analysis_step.restype = i32
        lib.cdl_analysis_step.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
        lib.cdl_synthesis_step.restype = i32
        lib.cdl_synthesis_step.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp]
        lib.cdl_forward.restype = i32
        lib.cdl_forward.argtypes = [vp] * 8
        for name in ("cdl_code_export", "cdl_code_import"):
            getattr(lib, name).restype = i32
            getattr(lib, name).argtypes = [vp] * 4
        lib.cdl_postprocess.restype = i32
        lib.cdl_postprocess.argtypes = [vp] * 5
        lib.cdl_denoise.restype = i32
        lib.cdl_denoise.argtypes = [vp] * 8
        lib.cdl_denoise_host.restype = i32
        lib.cdl_denoise_host.argtypes = [vp] * 8
        lib.cdl_comm_unique_id.restype = i32
        lib.cdl_comm_unique_id.argtypes = [vp]
        lib.cdl_comm_create.restype = i32
        lib.cdl_comm_create.argtypes = [P(vp), vp, i32, i32, i32]
        lib.cdl_comm_destroy.restype = None
        lib.cdl_comm_destroy.argtypes = [vp]
        lib.cdl_comm_allreduce_f64.restype = i32
        lib.cdl_comm_allreduce_f64.argtypes = [vp, vp, ctypes.c_size_t, vp]
        lib.cdl_halo_bytes.restype = i32
        lib.cdl_halo_bytes.argtypes = [vp, P(ctypes.c_size_t)]
        lib.cdl_halo_exchange.restype = i32
        lib.cdl_halo_exchange.argtypes = [vp] * 6
        lib.cdl_analysis_step_halo.restype = i32
        lib.cdl_analysis_step_halo.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.cdl_halo_add.restype = i32
        lib.cdl_halo_add.argtypes = [vp] * 6
        lib.cdl_forward_sharded.restype = i32
        lib.cdl_forward_sharded.argtypes = [vp] * 9
        lib.cdl_analysis_step_csr.restype = i32
        lib.cdl_analysis_step_csr.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.cdl_preprocess_noisy.restype = i32
        lib.cdl_preprocess_noisy.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
        lib.cdl_nle_mad_workspace_bytes.restype = i32
        lib.cdl_nle_mad_workspace_bytes.argtypes = [i32, i32, i32, i32, P(ctypes.c_size_t)]
        lib.cdl_nle_mad.restype = i32
        lib.cdl_nle_mad.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
        if lib.cdl_abi_version() != 1:
            raise RuntimeError("libcdl_b200.so ABI version mismatch")
        _lib = lib
        return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().cdl_status_string(rc).decode()
        raise RuntimeError(f"libcdl_b200 {what} failed: {msg} (status {rc})")
