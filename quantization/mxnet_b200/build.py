"""In-tree build of libfq_b200.so (hand-written sm_100a CUDA + the C ABI of include/fq.h).

    python -m quantization.mxnet_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is written next to this file so that it travels
with the tree; nothing is installed into site-packages.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfq_b200.so")
SOURCES = ["fq_api.cu", "fq_range.cu", "fq_quant.cu", "fq_fused.cu", "fq_calib.cu", "fq_wino.cu", "fq_stats.cu", "fq_foldbwd.cu", "fq_nccl.cu", "fq_qconv_mma.cu"]
HEADERS = [os.path.join(CSRC, "fq_common.cuh"), os.path.join(CSRC, "fq_fused.cuh"),
           os.path.join(ROOT, "include", "fq.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # parity: no FMA contraction between round() and * scale
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-extended-lambda",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libfq_b200.so cannot be built")
    return exe


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
        objs.append(obj)
    # libnccl is resolved with dlsym at run time (csrc/fq_nccl.cu): -ldl, no -lnccl
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
