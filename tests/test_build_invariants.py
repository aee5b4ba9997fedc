"""Resource limits the kernels rely on, read from the built library with cuobjdump (no GPU needed).

With AVC_BIAS_IN_GEMM=0 the data-parallel dW GEMM (comm warps, 320 threads, one CTA per SM, all of the SM's shared
memory) waits inside the kernel for the bias sums, which `colsum_kernel` computes on another stream WHILE the GEMM runs.
That only works if a colsum CTA fits next to a GEMM CTA on every SM: zero shared memory, and registers of both within
the 64 K file.  A compiler or code change that breaks this would turn that step into a 20 s timeout, so it is checked at
build time.  (By default the bias gradients come out of the GEMM launch itself and nothing runs beside it.)
The fused step's text-row splice (`splice_light_kernel`) runs beside the FORWARD GEMM the same way: its static shared
memory must fit into what the GEMM CTA leaves free.
"""
import re
import shutil
import subprocess

import pytest


def resource_usage(avc):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "--dump-resource-usage", str(avc._lib.lib_path())], capture_output=True, text=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", out):
        res[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)))
    return res


def alloc(regs_per_thread, threads):
    return (regs_per_thread + 7) // 8 * 8 * threads  # registers are allocated in units of 8 per thread


def test_colsum_fits_next_to_the_fused_gemm(avc):
    res = resource_usage(avc)
    if not res:
        pytest.skip("cuobjdump printed no resource usage")
    colsum = [v for k, v in res.items() if "colsum_kernel" in k]
    assert len(colsum) == 1
    assert colsum[0]["shared"] == 0, "colsum must not use static shared memory (the GEMM CTA owns all of it)"
    # gemm_kernel<MODE=1 (NT), OUT=1 (fp32), CG=2, MT, COMM != 0>: mangled ...ILi1ELi1ELi2ELi<MT>ELi<COMM>E / ELin<-COMM>E
    fused = {k: v for k, v in res.items() if re.search(r"gemm_kernelILi1ELi1ELi2ELi[12]EL(i[1-9]|in\d)", k)}
    assert len(fused) >= 10, sorted(res)
    for name, v in fused.items():
        total = alloc(v["reg"], 320) + alloc(colsum[0]["reg"], 256)
        assert total <= 65536, f"{name}: {v['reg']} regs x 320 + colsum {colsum[0]['reg']} x 256 = {total} > 64 K"
        assert v["stack"] <= 64, f"{name}: spills ({v['stack']} bytes of stack)"


def test_light_splice_fits_next_to_the_forward_gemm(avc):
    res = resource_usage(avc)
    if not res:
        pytest.skip("cuobjdump printed no resource usage")
    light = [v for k, v in res.items() if "splice_light_kernel" in k]
    assert len(light) >= 1
    # forward GEMM CTA: 6 x 32 KB stages + 32 KB epilogue staging + 256 B barriers + 1 KB alignment slack of dynamic
    # shared memory, plus the 1 KB the driver reserves per CTA; an SM has 228 KB
    gemm_smem = 6 * 32768 + 32768 + 256 + 1024 + 1024
    # forward kernels the fused step can launch: TN, any output type, CTA pairs, MT = 1, no comm, ACT = 0
    fwd = {k: v for k, v in res.items() if re.search(r"gemm_kernelILi0ELi[012]ELi2ELi1ELi0ELi0E", k)}
    assert len(fwd) == 3, sorted(res)
    for v in light:
        # cuobjdump's SHARED already contains the 1 KB per-CTA reservation of the small kernel
        assert gemm_smem + v["shared"] <= 228 * 1024, f"light splice uses {v['shared']} B of shared memory"
        for name, g in fwd.items():
            assert alloc(g["reg"], 192) + alloc(v["reg"], 256) <= 65536, name
