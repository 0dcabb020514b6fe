"""In-tree build of the CUDA library (sm_100a only).  `python -m jpc_b200._build` or
`__graft_entry__.build()`.  The resulting jpc_b200/libjpc_b200.so is git-ignored but travels to
the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjpc_b200.so")
SOURCES = ["api.cu", "tc_gemm.cu", "bpc.cu", "optim.cu", "plan.cu", "reg.cu", "call.cu"]
HEADERS = ["common.cuh", "simt_gemm.cuh", "tc_gemm.cuh", "tc_staged.cuh", "warp_gemm.cuh", "elementwise.cuh", "host.cuh", "chain_ffwd.cuh",
           os.path.join("..", "..", "include", "jpc_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build jpc_b200's CUDA library")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False, defines=()) -> str:
    """Compiles every translation unit for sm_100a (in parallel) and links libjpc_b200.so."""
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objdir = os.path.join(HERE, "csrc", "_obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    extra = (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        results = list(pool.map(compile_one, srcs))
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + [o for o, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB)  # new inode: a process that had the old binary mapped can dlopen the new one
    if verbose:
        for _, err in results:
            print(err)
    return LIB


XLA_LIB = os.path.join(HERE, "libjpc_b200_xla.so")


def xla_ffi_include() -> str | None:
    """Directory holding xla/ffi/api/ffi.h (shipped inside jaxlib), or None when jaxlib is not installed."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("jaxlib")
    except (ImportError, ValueError):
        spec = None
    if spec is None or not spec.submodule_search_locations:
        return None
    inc = os.path.join(list(spec.submodule_search_locations)[0], "include")
    return inc if os.path.exists(os.path.join(inc, "xla", "ffi", "api", "ffi.h")) else None


def build_xla_shim(include_dir: str | None = None) -> str | None:
    """Compiles csrc/xla_ffi_shim.cc (the jax.ffi handlers over jpcb_call) into libjpc_b200_xla.so next to the library.
    Needs jaxlib's headers: returns None -- and builds nothing -- when they are absent (as in this image)."""
    inc = include_dir or xla_ffi_include()
    if inc is None:
        return None
    build()
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", inc, "-I", "/usr/local/cuda/include",
           os.path.join(CSRC, "xla_ffi_shim.cc"), "-L", HERE, "-l:libjpc_b200.so", "-Wl,-rpath,$ORIGIN", "-o", XLA_LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the XLA FFI shim failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return XLA_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
    print(build_xla_shim() or "XLA FFI shim: jaxlib headers not found, not built")
