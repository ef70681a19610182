"""Builds probqa_b200/lib/libPqaCore.so (the C-ABI shared library: host engine + sm_100a kernels) in-tree with nvcc.

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libPqaCore.so")
SOURCES = ["pqa_kernels.cu", "pqa_eval_staged.cu", "pqa_train_sort.cu", "pqa_engine.cu", "pqa_maint.cu", "pqa_group.cu", "pqa_cabi.cu", "pqa_errors.cpp"]
HEADERS = ["pqa_kernels.cuh", "pqa_device.cuh", "pqa_engine.h", "pqa_group.h", "pqa_errors.h",
           os.path.join("..", "..", "include", "PqaCInterop.h"), os.path.join("..", "..", "include", "PqaB200Ext.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-Xptxas", "-v"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    objdir = os.path.join(LIB_DIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [_nvcc()] + NVCC_FLAGS + ["-x", "cu", "-c", src, "-o", obj]
            log = open(obj + ".log", "w")
            procs.append((s, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log, obj))
    failed = []
    for s, p, log, obj in procs:
        rc = p.wait()
        log.close()
        if rc != 0:
            failed.append((s, open(obj + ".log").read()))
        elif verbose:
            sys.stdout.write(open(obj + ".log").read())
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join("== %s ==\n%s" % f for f in failed))
    if force or procs or _stale(LIB_PATH, objs):
        subprocess.check_call([_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-Xcompiler", "-fPIC"])
    return LIB_PATH


CLIENT_SRC = os.path.join(os.path.dirname(_HERE), "clients", "pqa_client.cpp")
CLIENT_PATH = os.path.join(LIB_DIR, "pqa_client")


def build_client(force=False):
    """clients/pqa_client.cpp (the reference's PqaClient learner loop over the C ABI) -> probqa_b200/lib/pqa_client."""
    build()
    hdr = os.path.join(os.path.dirname(_HERE), "include", "PqaCInterop.h")
    if force or _stale(CLIENT_PATH, [CLIENT_SRC, hdr, LIB_PATH]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", CLIENT_SRC, "-o", CLIENT_PATH, "-L" + LIB_DIR, "-lPqaCore",
                               "-Wl,-rpath,$ORIGIN", "-lpthread"])
    return CLIENT_PATH


LAUNCHER_SRC = os.path.join(os.path.dirname(_HERE), "clients", "pqa_shard_launcher.cpp")
LAUNCHER_PATH = os.path.join(LIB_DIR, "pqa_shard_launcher")


def build_launcher(force=False):
    """clients/pqa_shard_launcher.cpp (one process per GPU over the C ABI, no Python / torch) -> probqa_b200/lib/."""
    build()
    hdrs = [os.path.join(os.path.dirname(_HERE), "include", h) for h in ("PqaCInterop.h", "PqaB200Ext.h")]
    if force or _stale(LAUNCHER_PATH, [LAUNCHER_SRC, LIB_PATH] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", LAUNCHER_SRC, "-o", LAUNCHER_PATH, "-L" + LIB_DIR, "-lPqaCore",
                               "-Wl,-rpath,$ORIGIN", "-lpthread"])
    return LAUNCHER_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_client(force="--force" in sys.argv))
    print(build_launcher(force="--force" in sys.argv))
