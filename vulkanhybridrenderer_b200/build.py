"""Builds the in-tree native libraries with nvcc / g++ (no JIT cache: the .so files travel with the repo snapshot).

  vulkanhybridrenderer_b200/libvhr_b200.so   CUDA kernels + C-ABI (include/vhr_b200.h), sm_100a only
  vulkanhybridrenderer_b200/libvhr_host.so   C++ host mirror of the reference's RenderGraph / HybridRenderPath
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OBJ = os.path.join(HERE, "_obj")
LIB_CUDA = os.path.join(HERE, "libvhr_b200.so")
LIB_HOST = os.path.join(HERE, "libvhr_host.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"),
]
GXX = "g++"   # the system compiler; $CXX in this image points at a wrapper without libstdc++ specs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r.stdout + r.stderr


def build_cuda(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "vhr_b200.h"))
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    objs, jobs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for out in ex.map(_run, jobs):
            if verbose:
                print(out)
    if jobs or _stale(LIB_CUDA, objs):
        _run([NVCC, "-shared", "-o", LIB_CUDA] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])    # -ldl: NVTX loads its injection library lazily
    return LIB_CUDA


def build_host(force=False):
    if not os.path.isdir(HOST):
        return None
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))
    if not srcs:
        return None
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + [os.path.join(ROOT, "include", "vhr_b200.h")]
    if force or _stale(LIB_HOST, deps + [LIB_CUDA]):
        # the reference's vendored third-party headers (stb_image.h), used where they lie: include path only, nothing copied
        ref_deps = os.path.join(os.environ.get("VHR_REFERENCE_ROOT", "/root/reference"), "dependencies")
        extra = ["-I", ref_deps] if os.path.isfile(os.path.join(ref_deps, "stb", "stb_image.h")) else []
        _run([GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include")] + extra + ["-o", LIB_HOST] + srcs +
             ["-L", HERE, "-l:libvhr_b200.so", "-lz", "-Wl,-rpath,$ORIGIN"])
    return LIB_HOST


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB_CUDA)
