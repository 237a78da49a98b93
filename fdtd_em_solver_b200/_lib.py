"""ctypes binding of libfdtd2d.so (include/fdtd2d.h) and its in-tree build.

The library is the only compute path.  If it is missing and cannot be built,
or no CUDA device is present, callers get an exception -- there is no CPU
fallback (the numpy restatement in oracle/ is test infrastructure only).
"""
import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC_DIR = os.path.join(HERE, "csrc")
INCLUDE_DIR = os.path.join(ROOT, "include")
# experiment hooks (tools/sweep.py): alternative builds side by side, e.g.
#   FDTD_LIB_SUFFIX=_cs FDTD_NVCC_EXTRA='-DFDTD_ST_MOD=".cs"'
LIB_PATH = os.path.join(HERE, "libfdtd2d%s.so" % os.environ.get("FDTD_LIB_SUFFIX", ""))
NVCC_EXTRA = os.environ.get("FDTD_NVCC_EXTRA", "").split()
OBJ_DIR = os.path.join(HERE, "_obj%s" % os.environ.get("FDTD_LIB_SUFFIX", ""))     # git-ignored, not needed on the GPU box

F64, F32, F32X = 0, 1, 2       # F32X: fp32 hi + bf16 lo field storage, float64 coefficients and arithmetic
SRC_POINT, SRC_RIGHT, SRC_LEFT, SRC_UP, SRC_DOWN = 0, 1, 2, 3, 4
KERNEL_EXACT, KERNEL_STREAM = 0, 1
DIRECTION_KIND = {None: SRC_POINT, "right": SRC_RIGHT, "left": SRC_LEFT, "up": SRC_UP, "down": SRC_DOWN}

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared"]


class Config(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int64), ("row_begin", C.c_int64), ("row_end", C.c_int64),
                ("pitch", C.c_int64), ("dtype", C.c_int32), ("device", C.c_int32), ("reflect", C.c_int32 * 4),
                ("kernel", C.c_int32), ("tile_rows", C.c_int32), ("vec", C.c_int32), ("block_warps", C.c_int32),
                ("courant", C.c_double), ("one_minus_courant", C.c_double)]


class Source(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("row", C.c_int64), ("col", C.c_int64)]


class Shape(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("a", C.c_int64 * 6), ("eps", C.c_double), ("mu", C.c_double)]


SHAPE_RECT, SHAPE_QUARTER, SHAPE_CONVEX = 0, 1, 2


class FramePlan(C.Structure):
    _fields_ = [("k_tile_rows", C.c_int32), ("k_tile_cols", C.c_int32), ("k_load_cols", C.c_int32),
                ("k_margin_cols", C.c_int32), ("k_tiles_y", C.c_int32), ("k_tiles_x", C.c_int32),
                ("k_row_lo", C.c_int64), ("k_row_hi", C.c_int64), ("frame_tile_rows", C.c_int32),
                ("frame_tile_cols", C.c_int32), ("frame_tiles_y", C.c_int32), ("frame_tiles_x", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("cell_updates", C.c_int64), ("kernel_launches", C.c_int64),
                ("halo_bytes_sent", C.c_int64), ("last_run_ms", C.c_double)]


# name -> (restype, argtypes); must list every symbol include/fdtd2d.h declares
_P = C.c_void_p
_D = C.POINTER(C.c_double)
SIGNATURES = {
    "fdtd_default_pitch": (C.c_int64, [C.c_int64, C.c_int32]),
    "fdtd_local_rows": (C.c_int64, [C.POINTER(Config)]),
    "fdtd_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "fdtd_destroy": (C.c_int, [_P]),
    "fdtd_last_error": (C.c_char_p, [_P]),
    "fdtd_bind_buffers": (C.c_int, [_P] * 9),
    "fdtd_plan_frame": (C.c_int, [C.POINTER(Config), C.c_int32, C.c_int32, _P, C.POINTER(FramePlan), _P, _P]),
    "fdtd_bind_scratch": (C.c_int, [_P, _P, _P, _P]),
    "fdtd_set_temporal_blocking": (C.c_int, [_P, C.c_int32]),
    "fdtd_set_resident": (C.c_int, [_P, C.c_int32, C.c_int64]),
    "fdtd_resident_active": (C.c_int, [_P]),
    "fdtd_set_frame_mode": (C.c_int, [_P, C.c_int32]),
    "fdtd_set_kstep_variant": (C.c_int, [_P, C.c_int32]),
    "fdtd_set_host_row_origin": (C.c_int, [_P, C.c_int64]),
    "fdtd_upload_coeffs": (C.c_int, [_P, _P, _P, C.c_int64]),
    "fdtd_paint_coeffs": (C.c_int, [_P, C.c_int32, C.POINTER(Shape), C.c_double, C.c_double, C.c_double]),
    "fdtd_fill_synthetic_coeffs": (C.c_int, [_P, C.c_uint64, C.c_double, C.c_double, C.c_double]),
    "fdtd_set_fields": (C.c_int, [_P, _P, _P, _P]),
    "fdtd_get_fields": (C.c_int, [_P, _P, _P, _P]),
    "fdtd_zero_fields": (C.c_int, [_P]),
    "fdtd_set_sources": (C.c_int, [_P, C.c_int32, C.POINTER(Source), C.c_int64, _P, _P, _P, _P]),
    "fdtd_set_reflectors": (C.c_int, [_P, C.c_int32, _P]),
    "fdtd_set_probes": (C.c_int, [_P, C.c_int32, _P]),
    "fdtd_probe_fetch": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "fdtd_probe_clear": (C.c_int, [_P]),
    "fdtd_step": (C.c_int, [_P, C.c_int64]),
    "fdtd_run": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]),
    "fdtd_run_enqueue": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "fdtd_run_streamed": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _P, C.c_int64]),
    "fdtd_view_begin": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "fdtd_view_end": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fdtd_set_tuning": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "fdtd_get_tuning": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "fdtd_sync": (C.c_int, [_P]),
    "fdtd_kernel_timing": (C.c_int, [_P, C.c_int32]),
    "fdtd_kernel_timing_fetch": (C.c_int, [_P, C.POINTER(C.c_int32), _D, _D, _D, C.POINTER(C.c_int64)]),
    "fdtd_rescan_coeffs": (C.c_int, [_P]),
    "fdtd_uniform_mu": (C.c_int, [_P]),
    "fdtd_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "fdtd_current_generation": (C.c_int, [_P]),
    "fdtd_comm_unique_id": (C.c_int, [_P]),
    "fdtd_comm_init": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "fdtd_halo_exchange": (C.c_int, [_P]),
    "fdtd_halo_rows": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int64)]),
    "fdtd_local_rows_halo": (C.c_int64, [C.POINTER(Config), C.c_int32]),
    "fdtd_set_halo_depth": (C.c_int, [_P, C.c_int32]),
    "fdtd_plan_edge": (C.c_int, [C.POINTER(Config), C.c_int32, C.c_int32, C.c_int32, _P, _P, C.POINTER(FramePlan), _P,
                                 C.POINTER(C.c_int32), _P, C.c_int32]),
    "fdtd_halo_blocks": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int64)]),
}


# The library is linked from one host translation unit and the kernel families, each compiled on its own (several of
# them once per dtype and slice of K): (source, defines, headers it really depends on, tag).  csrc/launch.h is the
# interface between them.
_KERNEL_HEADERS = {"k_step.cu": ["kernels.cuh"], "k_stepk.cu": ["kernels.cuh"], "k_resident.cu": ["kernels.cuh", "resident.cuh"],
                   "k_edge.cu": ["kernels.cuh", "resident.cuh", "edge.cuh"]}


def parts():
    out = [("fdtd2d.cu", [], ["kernels.cuh", "resident.cuh", "edge.cuh"], "host"),
           ("k_step.cu", [], _KERNEL_HEADERS["k_step.cu"], "step"),
           ("k_resident.cu", [], _KERNEL_HEADERS["k_resident.cu"], "resident")]
    for t in ("double", "float"):
        for lo, hi in ((2, 4), (5, 6), (7, 8)):
            out.append(("k_stepk.cu", ["-DFDTD_PART_T=" + t, "-DFDTD_K_LO=%d" % lo, "-DFDTD_K_HI=%d" % hi],
                        _KERNEL_HEADERS["k_stepk.cu"], "stepk_%s_%d_%d" % (t, lo, hi)))
        for lo, hi in ((2, 3), (4, 5), (6, 6), (7, 7), (8, 8)):
            out.append(("k_edge.cu", ["-DFDTD_PART_T=" + t, "-DFDTD_K_LO=%d" % lo, "-DFDTD_K_HI=%d" % hi],
                        _KERNEL_HEADERS["k_edge.cu"], "edge_%s_%d_%d" % (t, lo, hi)))
    for lo, hi in ((1, 3), (4, 6), (7, 8)):
        out.append(("k_split.cu", ["-DFDTD_K_LO=%d" % lo, "-DFDTD_K_HI=%d" % hi], _KERNEL_HEADERS["k_edge.cu"], "split_%d_%d" % (lo, hi)))
    return out


def sources():
    return sorted({os.path.join(SRC_DIR, src) for src, _, _, _ in parts()})


def _stale():
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR)] + [os.path.join(INCLUDE_DIR, "fdtd2d.h")]
    return any(os.path.getmtime(d) > built for d in deps)


def _object_for(nvcc, part, force, verbose):
    """Compile one part into OBJ_DIR, named by the hash of everything that determines it; reuse it if it is there."""
    import hashlib
    src, defines, headers, tag = part
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + NVCC_EXTRA + defines
    digest = hashlib.sha256(" ".join(flags).encode())
    for dep in [os.path.join(SRC_DIR, src), os.path.join(SRC_DIR, "launch.h"), os.path.join(INCLUDE_DIR, "fdtd2d.h")] + \
            [os.path.join(SRC_DIR, h) for h in headers]:
        with open(dep, "rb") as fh:
            digest.update(fh.read())
    obj = os.path.join(OBJ_DIR, "%s-%s.o" % (tag, digest.hexdigest()[:16]))
    if os.path.isfile(obj) and not force:
        return obj, ""
    for old in os.listdir(OBJ_DIR):                       # one object per part: older versions of this one go
        if old.startswith(tag + "-") and old.endswith(".o"):
            os.remove(os.path.join(OBJ_DIR, old))
    tmp = "%s.tmp.%d" % (obj, os.getpid())
    cmd = [nvcc] + flags + (["-Xptxas=-v"] if verbose else []) + ["-I", INCLUDE_DIR, "-c", os.path.join(SRC_DIR, src), "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed on %s (%s):\n%s%s" % (src, " ".join(defines), res.stdout, res.stderr))
    os.replace(tmp, obj)
    return obj, res.stderr


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> fdtd_em_solver_b200/libfdtd2d.so (in tree).

    The parts compile in parallel into content-addressed objects (an edit to one kernel family recompiles that family);
    force=True ignores the object cache.  Safe under torchrun: the build runs under an exclusive file lock (ranks that
    lose the race find a fresh library when they get the lock and return), and the link writes to a per-process
    temporary that is renamed into place, so no process can ever map a half-written library."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libfdtd2d.so (and there is no CPU fallback)")
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():          # another rank built it while this one waited for the lock
                return LIB_PATH
            os.makedirs(OBJ_DIR, exist_ok=True)
            jobs = int(os.environ.get("FDTD_BUILD_JOBS", "0")) or max(1, min(os.cpu_count() or 1, 16))
            with ThreadPoolExecutor(jobs) as pool:
                done = list(pool.map(lambda part: _object_for(nvcc, part, force, verbose), parts()))
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [o for o, _ in done] +
                                 ["-o", tmp], capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print("".join(log for _, log in done))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None


def load():
    """Load (building first if the sources are newer) and type the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        build()
    except RuntimeError as exc:
        # A library older than its sources may not match SIGNATURES below: never load it silently.  It is used only
        # when no compiler exists at all (a deployment box that received a prebuilt library), and loudly.
        if not os.path.isfile(LIB_PATH) or "nvcc not found" not in str(exc):
            raise
        import warnings
        warnings.warn("libfdtd2d.so is older than its sources and nvcc is missing: loading the prebuilt library as is",
                      RuntimeWarning)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the .so lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class FdtdError(RuntimeError):
    pass


def check(lib, handle, rc):
    if rc == 0:
        return
    msg = lib.fdtd_last_error(handle)
    msg = msg.decode() if msg else ""
    if rc == -4:
        raise IndexError(msg)              # the reference raises IndexError for a short pulse table (App. A-Q10)
    raise FdtdError("libfdtd2d error %d: %s" % (rc, msg))
