"""In-tree build of the CUDA extension (libproxsdp_b200.so) for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libproxsdp_b200.so")
# translation units (compiled in parallel, linked into one shared library)
SOURCES = ["solver.cu", "runtime.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp"))]
    deps += [os.path.join(HERE, "..", "include", f) for f in ("proxsdp_b200.h", "proxsdp_b200_types.h")]
    return deps


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps() if os.path.exists(d))


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    if src.endswith("runtime.cu"):      # depends only on its own headers
        deps = [os.path.join(CSRC, f) for f in ("runtime.cu", "runtime.cuh", "common.cuh")]
        deps.append(os.path.join(HERE, "..", "include", "proxsdp_b200_types.h"))
    else:
        deps = _deps()
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_extension(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if not force and not _obj_stale(os.path.join(CSRC, src), obj):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("PROXSDP_B200_NVCC_EXTRA", "").split() + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        for _, log in results:
            print(log)
    return LIB


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))
