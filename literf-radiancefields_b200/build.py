"""Build liblrf_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a with nvcc.  No JIT cache: the .so
lands in literf-radiancefields_b200/_build/ and travels with the repo snapshot to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liblrf_b200.so")
SOURCES = ["lrf_kernels.cu", "lrf_capi.cu"]
HEADERS = ["lrf_device.cuh", "lrf_kernels.cuh", "lrf_fast.cuh", "lrf_wray.cuh", "lrf_staged.cuh", "lrf_tv.cuh", "lrf_lean.cuh", os.path.join("..", "..", "include", "lrf.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",   # IEEE fp32, no contraction: parity with the reference's x86 build
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-Xptxas", "-v",
] + [("-D" + d) for d in os.environ.get("LRF_DEFINES", "").split() if d] + [
]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """kernel-variant experiments (tools/variants.py): same sources with -D knobs -> _build/variants/liblrf_b200_<name>.so"""
    vdir = os.path.join(OUT_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    lib = os.path.join(vdir, "liblrf_b200_%s.so" % name)
    cmd = [NVCC] + FLAGS + ["-D%s" % d for d in defines] + ["-o", lib] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(lib + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed for variant %s (see %s.log)" % (name, lib))
    return lib


CAMERA_SRC = os.path.join(HERE, "host", "harness", "lrf_camera.cpp")
CAMERA_LIB = os.path.join(OUT_DIR, "liblrf_camera.so")


def build_camera():
    """the harness's host-only camera helpers (LiteMath restatements; not part of the product library)"""
    if os.path.exists(CAMERA_LIB) and os.path.getmtime(CAMERA_LIB) >= os.path.getmtime(CAMERA_SRC):
        return CAMERA_LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", CAMERA_LIB, CAMERA_SRC])
    return CAMERA_LIB


def build(force=False, verbose=False):
    build_camera()
    if not force and not stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [NVCC] + FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(OUT_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or r.returncode:
        sys.stderr.write(log)
    if r.returncode:
        raise RuntimeError("nvcc failed building liblrf_b200.so (see _build/build.log)")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
