"""In-tree build of libavconnector_b200.so (sm_100a only) with nvcc.

The shared library is the drop-in boundary (include/avconnector_b200.h).  It links cudart statically
and resolves the one driver entry point it needs (cuTensorMapEncodeTiled) at run time, so it can be
dlopen'ed on a box without a GPU for the symbol-export test; every compute entry point fails loudly
there instead of falling back.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libavconnector_b200.so"
SOURCES = ["avc_capi.cu", "avc_mc.cu", "gather.cu", "proj_gemm.cu", "splice.cu", "elementwise.cu"]
HEADERS = [CSRC / "avc_kernels.h", CSRC / "avc_ptx.cuh", PKG_DIR.parent / "include" / "avconnector_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the connector has no non-CUDA build")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in [CSRC / s for s in SOURCES] + HEADERS:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link lib/libavconnector_b200.so."""
    LIB_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "build.stamp"
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = LIB_DIR / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out, file=sys.stderr)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
            "-cudart", "static", "-o", str(LIB_PATH), *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    stamp.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
