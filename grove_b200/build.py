"""Build libgrove_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so must travel with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgrove_b200.so")
# test-only: the round-1 mma.sync forward attention kernels, an independent cross-check of the tcgen05 kernels (include/grove_b200_legacy.h)
LIB_LEGACY = os.path.join(HERE, "libgrove_b200_legacy.so")
LEGACY_SOURCES = ["lib.cu", "attention.cu"]
SOURCES = ["lib.cu", "gemm_tcgen05.cu", "encoder_ops.cu", "attention_tc.cu", "attention_win_tc.cu", "decoder_ops.cu", "decoder_fused.cu", "box_ops.cu",
           "backward_ops.cu", "attention_bwd.cu", "attention_bwd_tc.cu", "attention_win_bwd_tc.cu", "decoder_bwd.cu", "preprocess.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC, "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("GROVE_NVCC_EXTRA", "").split()


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    hdrs = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tmem_ldst.cuh"), os.path.join(CSRC, "mma_sync.cuh"), os.path.join(ROOT, "include", "grove_b200.h"),
            os.path.join(ROOT, "include", "grove_b200_legacy.h")]
    objs, jobs = [], []
    for src in SOURCES + [s_ for s_ in LEGACY_SOURCES if s_ not in SOURCES]:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "_build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((src, [NVCC] + FLAGS + ["-c", s, "-o", o], o))

    def compile_one(job):
        src, cmd, o = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(o + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        return src, r

    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:   # one nvcc per translation unit, in parallel
            results = list(pool.map(compile_one, jobs))
        for src, r in results:
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
        failed = [src for src, r in results if r.returncode]
        if failed:
            raise RuntimeError(f"nvcc failed on {', '.join(failed)}")
    def obj_of(src):
        return os.path.join(HERE, "_build", src.replace(".cu", ".o"))
    for lib, srcs in ((LIB, SOURCES), (LIB_LEGACY, LEGACY_SOURCES)):
        lobjs = [obj_of(s_) for s_ in srcs]
        if force or _stale(lib, lobjs):
            cmd = [NVCC, "-shared", "-o", lib] + lobjs + ["-gencode", "arch=compute_100a,code=sm_100a"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
