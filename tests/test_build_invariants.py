"""Resource limits the kernels rely on, read from the built library with cuobjdump (no GPU needed).

The data-parallel dW GEMM (comm warps, 320 threads, one CTA per SM, all of the SM's shared memory) must fit the
register file without spills.  (The bias gradients come out of the GEMM launch itself; with AVC_BIAS_IN_GEMM=0 the
stand-alone bias-sum kernel runs BEFORE the fused launch on the same stream, so the two never have to share an SM.)
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


def test_fused_allreduce_kernels_fit_one_cta_per_sm(avc):
    """The dW GEMM with the comm warps runs 320 threads per CTA, one CTA per SM: registers must fit the 64 K file and
    nothing may spill (the wide multimem loads of the last round are the register-hungriest variant)."""
    res = resource_usage(avc)
    if not res:
        pytest.skip("cuobjdump printed no resource usage")
    # gemm_kernel<MODE=1 (NT), OUT=1 (fp32), CG=2, MT, COMM != 0>: mangled ...ILi1ELi1ELi2ELi<MT>ELi<COMM>E / ELin<-COMM>E
    fused = {k: v for k, v in res.items() if re.search(r"gemm_kernelILi1ELi1ELi2ELi[12]EL(i[1-9]|in\d)", k)}
    assert len(fused) >= 10, sorted(res)
    for name, v in fused.items():
        assert alloc(v["reg"], 320) <= 65536, f"{name}: {v['reg']} regs x 320 threads exceed the register file"
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
