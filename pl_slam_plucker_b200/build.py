"""Builds pl_slam_plucker_b200/libplba.so (CUDA kernels + C ABI + scene generator) for sm_100a, in-tree."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libplba.so")
SCENE_SO = os.path.join(HERE, "libplba_scene.so")      # the synthetic-scene generator alone (host code): what the CPU reference arm loads
SOURCES = ["plba_api.cu", "scene_gen.cpp"]


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the LBA library is CUDA-only and cannot be built without the CUDA toolkit")


def is_stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    # every file under csrc/ (kernels, warp kernels, tracker, shims) plus the public header: a stale binary must never be measured
    deps = [os.path.join(d, f) for d, _, fs in os.walk(CSRC) for f in fs] + [os.path.join(HERE, "..", "include", "plba.h")]
    return any(os.path.getmtime(f) > t for f in deps)


def build_scene(force=False):
    """g++ build of the scene generator alone: bench.py --impl reference must not map the product library."""
    src = os.path.join(CSRC, "scene_gen.cpp")
    hdr = os.path.join(HERE, "..", "include", "plba.h")
    if force or not os.path.exists(SCENE_SO) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(SCENE_SO):
        subprocess.check_call(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SCENE_SO, src])
    return SCENE_SO


def build(force=False, verbose=False, out=None, defines=()):
    global SO
    if out is None:
        build_scene(force)
    if out is None and not force and not is_stale():
        return SO
    target = out or SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared", "-diag-suppress", "39", "-o", target] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lgomp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    # the image exports CXX=/opt/gcc/bin/g++ which lacks parts of the toolchain; nvcc must use the system g++
    env = dict(os.environ)
    subprocess.check_call(cmd + ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else cmd, env=env)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else None, defines=defs))
