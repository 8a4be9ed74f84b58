"""In-tree build of libtextboost_b200.so and libtextboost_b200_bf16.so (sm_100a only) with plain nvcc.

The built libraries live in textboost_b200/lib/: git-ignored, but they travel to the GPU box with the
gpurun snapshot.  No JIT cache, no torch cpp_extension: the C ABI has no torch types in it.  The two
libraries are the same sources compiled for the two precision policies (precision.py): fp16 storage by
default, bf16 storage with -DTB_BF16.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libtextboost_b200.so")
# policy -> (library file, object directory, extra nvcc flags)
VARIANTS = {
    "fp16": (LIB, OBJDIR, []),
    "bf16": (os.path.join(LIBDIR, "libtextboost_b200_bf16.so"), os.path.join(HERE, "build_bf16"), ["-DTB_BF16"]),
}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths, extra=()) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, policies=("fp16", "bf16")) -> str:
    """Build the library of every policy in `policies`; returns the path of the fp16 one (or the first built)."""
    with ThreadPoolExecutor(max_workers=len(policies)) as ex:
        libs = list(ex.map(lambda p: _build_variant(p, force, verbose), policies))
    return LIB if "fp16" in policies else libs[0]


def _build_variant(policy: str, force: bool, verbose: bool) -> str:
    LIB, OBJDIR, extra = VARIANTS[policy]  # noqa: N806 (shadow the module-level fp16 paths)
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "textboost_b200.h"))
    srcs = _sources()
    stamp_path = os.path.join(OBJDIR, "stamp")
    objs = []
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s[:-3] + ".o")
        dig = _digest([src] + headers, extra)
        digf = obj + ".sha"
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(digf)
                and open(digf).read() == dig):
            continue
        jobs.append((src, obj, dig, digf))

    def compile_one(job):
        src, obj, dig, digf = job
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        with open(digf, "w") as f:
            f.write(dig)
        return src, log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(compile_one, jobs):
                if verbose:
                    print(f"== {src}\n{log}")
    link_dig = _digest(objs, extra)
    if force or jobs or not os.path.exists(LIB) or not os.path.exists(stamp_path) \
            or open(stamp_path).read() != link_dig:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
        with open(stamp_path, "w") as f:
            f.write(link_dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
