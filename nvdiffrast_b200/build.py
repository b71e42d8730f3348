"""Build libb200dr.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m nvdiffrast_b200.build [--force]

The shared library is written next to this file (nvdiffrast_b200/libb200dr.so): git-ignored, but it travels to the
GPU box with the gpurun snapshot.  No JIT cache, no torch extension machinery: plain nvcc -shared.
"""
import contextlib
import fcntl
import os
import subprocess
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_DIR)
SO_PATH = os.path.join(_DIR, "libb200dr.so")
SOURCES = sorted(f for f in os.listdir(os.path.join(_DIR, "csrc")) if f.endswith(".cu"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.access(c, os.X_OK) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(_DIR, "csrc", f) for f in os.listdir(os.path.join(_DIR, "csrc")) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(_ROOT, "include", "b200dr.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defs):
    """Tuning build: the same sources with extra -D flags, written to nvdiffrast_b200/variants/libb200dr_<name>.so
    (select it with the B200DR_LIB environment variable)."""
    out = os.path.join(_DIR, "variants", "libb200dr_%s.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    return build(force=True, defs=list(defs), out=out, bdir=os.path.join(_DIR, "csrc", "_build", "var_" + name))


@contextlib.contextmanager
def _build_lock(path):
    """Exclusive inter-process lock: the ranks of one node import this package at the same moment, and each of them would start
    the same nvcc jobs into the same object files (round 2: a 2-rank run loaded a library that the other rank was still linking)."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def build(force=False, verbose=False, defs=(), out=None, bdir=None):
    if not force and not needs_build():
        return SO_PATH
    out = out or SO_PATH
    bdir = bdir or os.path.join(_DIR, "csrc", "_build")
    os.makedirs(bdir, exist_ok=True)
    with _build_lock(os.path.join(bdir, ".lock")):
        if not force and out == SO_PATH and not needs_build():      # another process built it while this one waited for the lock
            return SO_PATH
        return _build_locked(verbose, defs, out, bdir)


def _build_locked(verbose, defs, out, bdir):
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + list(defs) + ["-I", os.path.join(_ROOT, "include"), "-I", os.path.join(_DIR, "csrc"),
                                        "-c", os.path.join(_DIR, "csrc", src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, pr in procs:
        text, _ = pr.communicate()
        log.append("== %s ==\n%s" % (src, text))
        failed |= pr.returncode != 0
    with open(os.path.join(bdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see %s" % os.path.join(bdir, "build.log"))
    # link to a temporary name and rename: a process that loads the library never sees a partially written file
    tmp = "%s.tmp.%d" % (out, os.getpid())
    subprocess.check_call([_nvcc(), "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
