"""Build the C-ABI CUDA library IN-TREE (sm_100a only), once per 16-bit operand format:

    flipped_vqa_b200/libfvqa.so        fp16 operands (default; the reference's dtype, meets the full-depth parity bound)
    flipped_vqa_b200/libfvqa_bf16.so   bf16 operands (-DFVQA_BF16; selected with FVQA_DTYPE=bf16)

    python -m flipped_vqa_b200.build [--force] [-v]

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with the
repo snapshot. No JIT cache, no torch extension machinery: plain `nvcc -shared`.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libfvqa.so")
VARIANTS = {"fp16": ("libfvqa.so", []), "bf16": ("libfvqa_bf16.so", ["-DFVQA_BF16"])}
SOURCES = ["api.cu", "elementwise.cu", "embed.cu", "heads.cu", "gemm_tcgen05.cu", "gemm_skinny.cu", "attention.cu", "attention_tc.cu", "attention_tc_long.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tmap.h"), os.path.join(CSRC, "attention.h"), os.path.join(CSRC, "attention_tc.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "fvqa.h"), os.path.join(os.path.dirname(HERE), "include", "fvqa_debug.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v", "-DNDEBUG"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, force: bool, variant: str = "fp16") -> str:
    obj = os.path.join(OBJ, variant, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if force or _stale(obj, [path] + HEADERS):
        cmd = [NVCC] + FLAGS + VARIANTS[variant][1] + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    return obj


def lib_path(variant: str = "fp16") -> str:
    return os.path.join(HERE, VARIANTS[variant][0])


def build(force: bool = False, verbose: bool = False, variants=("fp16", "bf16")) -> str:
    """Compile and link every variant; returns the path of the default (fp16) library."""
    for v in variants:
        os.makedirs(os.path.join(OBJ, v), exist_ok=True)
    jobs = [(s, v) for v in variants for s in SOURCES]
    with cf.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(jobs))) as ex:
        done = list(ex.map(lambda j: _compile(j[0], force, j[1]), jobs))
    for v in variants:
        objs = [o for o, (_, jv) in zip(done, jobs) if jv == v]
        lib = lib_path(v)
        if force or _stale(lib, objs):
            cmd = [NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    objs = [o for o, (_, jv) in zip(done, jobs) if jv == variants[0]]
    if verbose:
        for o in objs:
            with open(o + ".log") as f:
                for line in f:
                    if "registers" in line or "spill" in line or "Compiling entry" in line:
                        print(line.rstrip())
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
