"""CPU: the C-ABI shared library builds, loads without a GPU, and exports exactly what include/*.h declares."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "avconnector_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"AVC_API\s+[\w\s\*]+?\b(avc_\w+)\s*\(", text)))


def test_header_declares_the_path_entry_points():
    syms = declared_symbols()
    for need in ("avc_gather_fwd", "avc_proj_fwd", "avc_proj_bwd_dw", "avc_colsum", "avc_splice_fwd",
                 "avc_splice_bwd", "avc_pack_weight", "avc_row_resample", "avc_device_check", "avc_last_error",
                 "avc_proj_bwd_dw_db", "avc_proj_bwd_dw_db_allreduce", "avc_proj_bwd_dx", "avc_gather_bwd",
                 "avc_cast_bf16"):
        assert need in syms


def test_library_exports_every_declared_symbol(avc):
    lib = ctypes.CDLL(str(avc._lib.lib_path()))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in {HEADER.name} but not exported"
    assert sorted(avc._lib.EXPORTS) == declared_symbols()
    out = subprocess.run(["nm", "-D", "--defined-only", str(avc._lib.lib_path())], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (avc_\w+)", out)))
    assert exported == declared_symbols(), "library exports symbols the header does not declare (or vice versa)"


def test_library_is_sm100a_native_code(avc):
    """tcgen05 / TMA must be in the SASS (UTC*MMA, LDTM, UTMALDG/UTMASTG, UBLKCP), and no legacy HMMA."""
    r = subprocess.run(["cuobjdump", "-sass", str(avc._lib.lib_path())], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = r.stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_no_gpu_means_loud_failure_not_fallback(avc):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = avc._lib
    assert L.load().avc_abi_version() == L.AVC_ABI_VERSION
    with pytest.raises(L.ConnectorError):
        L.require_device(0)
    with pytest.raises(L.ConnectorError):
        avc.fused_connector(torch.zeros(1, 4, 8), None, torch.zeros(16, 8), torch.zeros(16), None, None,
                            avc.FusePlan(modality="audio"))
    with pytest.raises(L.ConnectorError):
        avc.ModalityConnector(8, 16, device="cpu")(torch.zeros(1, 2, 8))


def test_product_never_imports_the_oracle():
    for p in (ROOT / "audio-visual-llm_b200").rglob("*.py"):
        assert "oracle" not in p.read_text().replace("oracle/", ""), f"{p} references the oracle"


def test_plain_c_program_links_against_the_abi(avc, tmp_path):
    """The boundary is a C ABI: a C99 translation unit that only includes include/avconnector_b200.h compiles,
    links against the shared library and runs (without a GPU the device check reports an error, it does not crash)."""
    import shutil

    if shutil.which("gcc") is None:
        pytest.skip("gcc unavailable")
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "avconnector_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  avc_feat f; avc_mat m; avc_splice s; avc_bias_grad bg; (void)f; (void)m; (void)s; (void)bg;\n"
        '  printf("abi=%d ws=%zu\\n", avc_abi_version(), avc_colsum_workspace_bytes(4096));\n'
        "  int rc = avc_device_check(0);\n"
        '  printf("device_check=%d msg=%s\\n", rc, avc_last_error());\n'
        "  return 0;\n}\n")
    exe = tmp_path / "abi"
    lib_dir = avc._lib.lib_path().parent
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                        "-L", str(lib_dir), "-lavconnector_b200", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "abi=3" in out.stdout and "device_check=" in out.stdout
