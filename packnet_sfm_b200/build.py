"""Build libpacknet_b200.so in-tree with nvcc for sm_100a (no torch extension machinery, no JIT cache).

`python -m packnet_sfm_b200.build` or packnet_sfm_b200.build.build_library().  The .so is git-ignored
but travels to the GPU box with the gpurun snapshot.  cudart is linked statically and the driver API
(cuTensorMapEncodeTiled) is resolved at run time through cudaGetDriverEntryPoint, so the library
loads on a machine without a GPU driver (the CPU test tier checks its exported symbols)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpacknet_b200.so")
SOURCES = ["api.cu", "loss_kernels.cu", "conv_engine.cu", "layer_kernels.cu", "fold_kernels.cu", "frame_kernels.cu", "pack_kernels.cu", "optim_kernels.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (looked at %s)" % cand)
    return cand


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "packnet_b200.h"))
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs, procs = [], []
    for src in sources:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc, "-c", path, "-o", obj, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                   "-I", os.path.join(ROOT, "include"), "-I", CSRC] + ARCH
            if verbose:
                cmd += ["-Xptxas", "-v"]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write("[nvcc %s]\n%s\n" % (src, out))
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-cudart", "static", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
