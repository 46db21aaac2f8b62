"""In-tree build of libb200q.so with nvcc for sm_100a (no torch headers: seconds per file, not minutes).
Each translation unit is compiled to its own object (in parallel, only when stale), then linked."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.environ.get("B200Q_LIB") or os.path.join(PKG, "libb200q.so")   # B200Q_LIB: A/B runs of two builds on one box
SOURCES = ["api.cu", "unpack.cu", "gemv_generic.cu", "gemv_mma.cu", "gemv_stream.cu", "gemv_imma.cu",
           "gemm_tcgen05.cu", "decode_chain.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "--compiler-options", "-fPIC"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return hs + [os.path.join(PKG, "..", "include", "b200q.h")]


def _newer(path, deps):
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    for src in SOURCES:
        obj = os.path.join(OBJ, src[:-3] + ".o")
        if force or _newer(obj, [os.path.join(CSRC, src)] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for log in ex.map(compile_one, jobs):
                if verbose:
                    print(log)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if jobs or _newer(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
