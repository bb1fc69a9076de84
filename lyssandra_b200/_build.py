"""In-tree build of the CUDA library (sm_100a only): nvcc -> lyssandra_b200/liblyssa_b200.so.

The .so is git-ignored but travels to the GPU box with the repo snapshot.  There is no
fallback: if nvcc is missing or the build fails this raises."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "liblyssa_b200.so")

SOURCES = ["runtime.cu", "gemm.cu", "bomp_generic.cu", "bomp_fast.cu", "corr_gemm_tc.cu", "bomp_fused.cu", "bomp.cu", "ksvd.cu", "ksvd_sweep.cu", "ksvd_exact.cu", "odl.cu", "odl_gemm_tc.cu", "comm.cu", "spm.cu", "dsift.cu", "thresh.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--threads", "0"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build liblyssa_b200.so (no fallback path exists)")


def _fingerprint():
    """Hash of the sources + flags.  runtime.cu is compiled with -DLYS_FINGERPRINT=<this>, so the library itself says
    what it was built from (lys_build_fingerprint); no side file has to travel with it."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(_PKG), "include", "lyssa_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())      # names, not absolute paths: the tree is built here and run elsewhere
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, bringup: bool = False, variant: str = "", defines=()) -> str:
    """Product library, or with ``bringup`` a separate liblyssa_b200_bringup.so compiled with -DLYS_BRINGUP
    (in-kernel phase timers and their debug hooks), or with ``variant``/``defines`` an experiment build
    liblyssa_b200_<variant>.so; the non-product builds are loaded through LYSSA_B200_LIB by scripts only."""
    if bringup or variant:
        name = variant or "bringup"
        flags = list(defines) + (["-DLYS_BRINGUP"] if bringup else [])
        return _build_to(os.path.join(_PKG, "liblyssa_b200_%s.so" % name), os.path.join(_PKG, "build_%s" % name), flags, verbose)
    fp = _fingerprint()

    def fresh():
        if not os.path.isfile(LIB_PATH):
            return False
        with open(LIB_PATH, "rb") as fh:                 # read, not dlopen: a stale handle would be cached by the loader
            return ("LYSFP:" + fp).encode() in fh.read()

    if not force and fresh():
        return LIB_PATH
    # one builder at a time (the ranks of a torchrun launch all arrive here); the others wait on the lock and then
    # find a fresh library.  The .so is linked under a temporary name and renamed, so a reader never sees half a file.
    import fcntl
    with open(os.path.join(_PKG, ".build_lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or not fresh():
                tmp = LIB_PATH + ".tmp.%d" % os.getpid()
                _build_to(tmp, os.path.join(_PKG, "build"), [], verbose, fingerprint=fp)
                os.replace(tmp, LIB_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def _build_to(lib_path, objdir, extra_flags, verbose, fingerprint=None):
    nvcc = _nvcc()
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra_flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        if fingerprint and src == "runtime.cu":
            cmd.insert(1, '-DLYS_FINGERPRINT="%s"' % fingerprint)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose and out:
            sys.stderr.write(out)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", lib_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s" % res.stdout)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, bringup="--bringup" in sys.argv))
