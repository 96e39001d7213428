"""Build libsixdgs.so in-tree with nvcc for sm_100a (no torch / pybind dependency: plain C ABI).

Usage: python 6dgs_b200/csrc/build.py [--force] [--verbose]
raygen.cu / normals.cu carry discrete decisions that must round like the torch expressions they
restate, so they are compiled with -fmad=false; the GEMM-shaped files keep FMA contraction.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libsixdgs.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
SOURCES = {
    "api.cu": [],
    "raygen.cu": ["-fmad=false"],
    "normals.cu": ["-fmad=false"],
    "knn_grid.cu": ["-fmad=false"],
    "features.cu": [],
    "features_tc.cu": [],
    "features_x2.cu": [],
    "score_simt.cu": [],
    "score_tc.cu": [],
    "score_tc_mq.cu": [],
    "topk.cu": [],
    "pose.cu": ["-fmad=false"],
}


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(HERE)):
        if name.endswith((".cu", ".cuh", ".py")):
            h.update(open(os.path.join(HERE, name), "rb").read())
    h.update(open(os.path.join(HERE, "..", "..", "include", "sixdgs.h"), "rb").read())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = os.path.join(HERE, "build", "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [_nvcc(), *ARCH, *COMMON, *extra, "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc(), *ARCH, "-shared", "-o", OUT, *objs, "-lcudart"]
    subprocess.check_call(link)
    open(stamp, "w").write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
