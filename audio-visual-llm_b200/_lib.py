"""ctypes binding of libavconnector_b200.so (the C ABI in include/avconnector_b200.h).

This is the only way the host layer reaches the GPU kernels.  There is no fallback: if the library is
missing or the device is not sm_100 every call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional, Sequence

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libavconnector_b200.so"
_lib: Optional[C.CDLL] = None

AVC_ABI_VERSION = 3
EXPORTS = (
    "avc_abi_version", "avc_last_error", "avc_device_check", "avc_gather_fwd", "avc_proj_fwd",
    "avc_proj_bwd_dw", "avc_colsum_workspace_bytes", "avc_colsum", "avc_pack_weight", "avc_splice_fwd",
    "avc_splice_bwd", "avc_row_resample", "avc_sumsq_workspace_bytes", "avc_sumsq", "avc_adamw_step",
    "avc_gelu_fwd", "avc_gelu_bwd", "avc_pack_weight_t",
    "avc_comm_flag_bytes", "avc_comm_alloc", "avc_comm_free", "avc_comm_export", "avc_comm_open", "avc_comm_close",
    "avc_proj_bwd_dw_allreduce", "avc_comm_signal_extra", "avc_colsum_comm", "avc_colsum_workspace_header_bytes",
    "avc_mc_supported", "avc_mc_padded_bytes", "avc_mc_create", "avc_mc_import", "avc_mc_add_device",
    "avc_mc_bucket_alloc", "avc_mc_bucket_free",
    "avc_proj_bwd_dw_db", "avc_proj_bwd_dw_db_allreduce", "avc_proj_bwd_dx", "avc_gather_bwd", "avc_cast_bf16",
    "avc_debug_gemm_profile", "avc_proj_bwd_dw_plan",
)

# AVC_DTYPE_* codes of include/avconnector_b200.h
DTYPE_CODES = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}
BIAS_COLS = 64  # columns of the token-present operand of the bias work items (avc_bias_grad.present)


def dtype_code(dt: torch.dtype) -> int:
    try:
        return DTYPE_CODES[dt]
    except KeyError:
        raise ConnectorError(f"dtype {dt} is not supported by the B200 connector (bf16, fp16 or fp32)") from None


class AvcFeat(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p), ("batch_stride", C.c_int64), ("frame_stride", C.c_int64),
        ("frames", C.c_int32), ("dim", C.c_int32), ("stack", C.c_int32), ("repeat", C.c_int32),
        ("valid_frames", C.c_void_p),
    ]


class AvcMat(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64), ("row_stride", C.c_int64),
        ("batches", C.c_int64), ("batch_stride", C.c_int64),
    ]


class AvcSplice(C.Structure):
    _fields_ = [
        ("input_ids", C.c_void_p), ("placeholder_id", C.c_int64), ("pad_id", C.c_int64),
        ("batch", C.c_int32), ("seq", C.c_int32), ("hidden", C.c_int32), ("tokens_per_sample", C.c_int32),
        ("tok_offset", C.c_void_p), ("embed_table", C.c_void_p), ("vocab", C.c_int64),
        ("attention_mask", C.c_void_p), ("mask_mode", C.c_int32), ("label_mode", C.c_int32),
        ("labels_in", C.c_void_p), ("label_len", C.c_int32), ("elem_size", C.c_int32),
        ("labels_out", C.c_void_p), ("status", C.c_void_p), ("av_rows_in_place", C.c_int32), ("reserved", C.c_int32),
    ]


COMM_MAX_WORLD = 8


class AvcComm(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("epoch", C.c_uint32), ("reserved", C.c_uint32),
        ("bucket", C.c_void_p * COMM_MAX_WORLD), ("flags", C.c_void_p * COMM_MAX_WORLD), ("status", C.c_void_p),
        ("timeout_ns", C.c_uint64), ("bucket_bytes", C.c_uint64), ("mc_bucket", C.c_void_p),
    ]


class AvcBiasGrad(C.Structure):
    _fields_ = [("present", C.POINTER(AvcMat)), ("out0", C.c_void_p), ("out1", C.c_void_p), ("alpha0", C.c_float),
                ("alpha1", C.c_float)]


class AvcMcBucket(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("mc_ptr", C.c_void_p), ("bytes", C.c_uint64), ("mem_handle", C.c_uint64),
                ("mc_handle", C.c_uint64)]


class ConnectorError(RuntimeError):
    """Raised for every non-zero status of the C ABI (message from avc_last_error())."""


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """dlopen the in-tree shared library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ConnectorError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The connector has no PyTorch/CPU fallback.")
    lib = C.CDLL(str(_LIB_PATH))
    lib.avc_abi_version.restype = C.c_int
    lib.avc_last_error.restype = C.c_char_p
    lib.avc_colsum_workspace_bytes.restype = C.c_size_t
    lib.avc_colsum_workspace_bytes.argtypes = [C.c_int32]
    lib.avc_colsum_workspace_header_bytes.restype = C.c_size_t
    lib.avc_comm_flag_bytes.restype = C.c_size_t
    lib.avc_comm_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    lib.avc_comm_free.argtypes = [C.c_void_p]
    lib.avc_comm_export.argtypes = [C.c_void_p, C.c_void_p]
    lib.avc_comm_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.avc_comm_close.argtypes = [C.c_void_p]
    if lib.avc_abi_version() != AVC_ABI_VERSION:
        raise ConnectorError(f"ABI version mismatch: library {lib.avc_abi_version()} != binding {AVC_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise ConnectorError(f"avconnector_b200 error {rc}: {load().avc_last_error().decode()}")


def require_device(index: int = 0) -> None:
    """Replaces the reference's `"cuda" if torch.cuda.is_available() else "cpu"` (clip_whisper_model.py:91)."""
    check(load().avc_device_check(index))


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def mat(t: torch.Tensor) -> AvcMat:
    """2-D [rows, cols] or 3-D [batches, rows, cols] tensor (cols contiguous) -> avc_mat."""
    if t.dim() == 2:
        if t.stride(1) != 1:
            raise ValueError("matrix columns must be contiguous")
        return AvcMat(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), 1, t.shape[0] * t.stride(0))
    if t.dim() == 3:
        if t.stride(2) != 1:
            raise ValueError("matrix columns must be contiguous")
        return AvcMat(t.data_ptr(), t.shape[1], t.shape[2], t.stride(1), t.shape[0], t.stride(0))
    raise ValueError("expected a 2-D or 3-D tensor")


def feat(t: Optional[torch.Tensor], stack: int, valid: Optional[torch.Tensor] = None,
         repeat: int = 1) -> Optional[AvcFeat]:
    """[batch, frames, dim] bf16 features (any batch/frame stride, dim contiguous) -> avc_feat."""
    if t is None:
        return None
    if t.dim() != 3 or t.stride(2) != 1 or t.dtype != torch.bfloat16:
        raise ValueError("features must be bf16 [batch, frames, dim] with contiguous dim")
    if valid is not None and (valid.dtype != torch.int32 or not valid.is_contiguous()):
        raise ValueError("valid_frames must be a contiguous int32 tensor")
    return AvcFeat(t.data_ptr(), t.stride(0), t.stride(1), t.shape[1], t.shape[2], stack, repeat, _ptr(valid))


def _mat_array(ms: Sequence[AvcMat]):
    return (AvcMat * len(ms))(*ms)


# ------------------------------------------------------------------------------------------ kernels
def gather_fwd(audio, video, ka: int, kv: int, batch: int, tokens_per_sample: int, a_out: torch.Tensor,
               row_flags: Optional[torch.Tensor] = None, tok_offset: Optional[torch.Tensor] = None,
               audio_valid=None, video_valid=None, audio_repeat: int = 1, video_repeat: int = 1) -> None:
    fa, fv = feat(audio, ka, audio_valid, audio_repeat), feat(video, kv, video_valid, video_repeat)
    check(load().avc_gather_fwd(
        C.byref(fa) if fa is not None else None, C.byref(fv) if fv is not None else None, C.c_int32(batch),
        C.c_void_p(_ptr(tok_offset)), C.c_int32(tokens_per_sample), C.c_int64(a_out.shape[0]),
        C.c_void_p(a_out.data_ptr()), C.c_int64(a_out.stride(0)), C.c_void_p(_ptr(row_flags)), stream_ptr()))


def proj_fwd(a_segs: Sequence[torch.Tensor], w_segs: Sequence[torch.Tensor], y: torch.Tensor,
             bias0: Optional[torch.Tensor] = None, bias1: Optional[torch.Tensor] = None,
             row_flags: Optional[torch.Tensor] = None, flag_rows0: int = 1 << 30, flag_rows1: int = 1 << 30,
             act: int = 0, bias_scale0: float = 1.0, bias_scale1: float = 1.0) -> None:
    check(load().avc_proj_fwd(
        C.c_int32(len(a_segs)), _mat_array([mat(t) for t in a_segs]), _mat_array([mat(t) for t in w_segs]),
        C.byref(mat(y)), C.c_int32(dtype_code(y.dtype)), C.c_void_p(_ptr(bias0)),
        C.c_void_p(_ptr(bias1)), C.c_float(bias_scale0), C.c_float(bias_scale1), C.c_void_p(_ptr(row_flags)),
        C.c_int32(flag_rows0), C.c_int32(flag_rows1), C.c_int32(act), stream_ptr()))


_PRESENT_CACHE: dict = {}


def present_operand(batch: int, rows: int, device, row_flags: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Token-present operand of the bias work items (`avc_bias_grad.present`): bf16 [batch, rows, 64] whose column
    0 / 1 is 1 where the row carries an audio / video token.  Dense streams: all ones (a cached constant);
    `row_flags` (uint8 [batch * rows], bit 0 / 1 as written by `avc_gather_fwd`): per-row bits."""
    device = torch.device(device)
    if row_flags is None:
        key = (batch, rows, device)
        f = _PRESENT_CACHE.get(key)
        if f is None:
            f = torch.zeros(batch, rows, BIAS_COLS, dtype=torch.bfloat16, device=device)
            f[:, :, :2] = 1
            if len(_PRESENT_CACHE) >= 8:
                _PRESENT_CACHE.pop(next(iter(_PRESENT_CACHE)))
            _PRESENT_CACHE[key] = f
        return f
    f = torch.zeros(batch, rows, BIAS_COLS, dtype=torch.bfloat16, device=device)
    fl = row_flags.view(batch, rows)
    f[:, :, 0] = (fl & 1).to(torch.bfloat16)
    f[:, :, 1] = ((fl >> 1) & 1).to(torch.bfloat16)
    return f


def _bias_grad(present: torch.Tensor, out0: Optional[torch.Tensor], out1: Optional[torch.Tensor], alpha0: float,
               alpha1: float):
    if present.dtype != torch.bfloat16 or present.dim() != 3 or present.shape[2] != BIAS_COLS:
        raise ValueError(f"token-present operand must be bf16 [batch, rows, {BIAS_COLS}]")
    m = mat(present)
    b = AvcBiasGrad(C.pointer(m), _ptr(out0), _ptr(out1), alpha0, alpha1)
    b._keep = (m, present, out0, out1)
    return b


_DW_PLAN: dict = {}
_DW_WORKSPACE: dict = {}


def dw_plan(dy: torch.Tensor, x_segs: Sequence[torch.Tensor], with_bias: bool):
    """(reduction slices, workspace bytes) the dW launch would use for these operands (`avc_proj_bwd_dw_plan`)."""
    import os

    key = (tuple(dy.shape), tuple(tuple(x.shape) for x in x_segs), bool(with_bias), dy.device,
           os.environ.get("AVC_GEMM_KSPLIT"), os.environ.get("AVC_GEMM_MAX_WORKERS"))
    hit = _DW_PLAN.get(key)
    if hit is None:
        splits, nbytes = C.c_int32(1), C.c_size_t(0)
        check(load().avc_proj_bwd_dw_plan(C.byref(mat(dy)), C.c_int32(len(x_segs)), _mat_array([mat(t) for t in x_segs]),
                                          C.c_int32(1 if with_bias else 0), C.byref(splits), C.byref(nbytes)))
        hit = (int(splits.value), int(nbytes.value))
        if len(_DW_PLAN) > 64:
            _DW_PLAN.clear()
        _DW_PLAN[key] = hit
    return hit


def dw_workspace(device, nbytes: int) -> Optional[torch.Tensor]:
    """Per-device scratch of the split-reduction dW launch: zero-filled when (re)allocated, the launches leave its
    arrival counters zero.  Calls that share it must be ordered on one stream (they are: the autograd backward)."""
    if nbytes <= 0:
        return None
    device = torch.device(device)
    ws = _DW_WORKSPACE.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _DW_WORKSPACE[device] = ws
    return ws


def proj_bwd_dw(dy: torch.Tensor, x_segs: Sequence[torch.Tensor], dw_segs: Sequence[torch.Tensor],
                alpha: Sequence[float], dy_row_base: int = 0, max_sms: int = 0, bias=None,
                workspace: Optional[torch.Tensor] = None, split: bool = True) -> None:
    """bias = (present, db0, db1, alpha0, alpha1): also produce the bias gradients inside the same launch.
    With few output tiles the launch splits the row reduction over the idle CTA pairs (`split`; scratch from
    `dw_workspace` unless `workspace` is given)."""
    al = (C.c_float * len(x_segs))(*alpha)
    if workspace is None and split and max_sms == 0:
        workspace = dw_workspace(dy.device, dw_plan(dy, x_segs, bias is not None)[1])
    bg = _bias_grad(*bias) if bias is not None else None
    check(load().avc_proj_bwd_dw_db(
        C.byref(mat(dy)), C.c_int32(dy_row_base), C.c_int32(len(x_segs)), _mat_array([mat(t) for t in x_segs]),
        _mat_array([mat(t) for t in dw_segs]), al, C.byref(bg) if bg is not None else None,
        C.c_void_p(_ptr(workspace)), C.c_size_t(0 if workspace is None else workspace.numel() * workspace.element_size()),
        C.c_int32(max_sms), stream_ptr()))


def proj_bwd_dx(dy_segs: Sequence[torch.Tensor], wt_segs: Sequence[torch.Tensor], dx: torch.Tensor) -> None:
    """dX = sum_s dY_s . W_s with wt_segs[s] = W_s^T (bf16 [K_in, H_s] from pack_weight_t)."""
    check(load().avc_proj_bwd_dx(_mat_array([mat(t) for t in dy_segs]), C.c_int32(len(dy_segs)),
                                 _mat_array([mat(t) for t in wt_segs]), C.byref(mat(dx)),
                                 C.c_int32(dtype_code(dx.dtype)), stream_ptr()))


def gather_bwd(da: torch.Tensor, col_off: int, d_feat: torch.Tensor, stack: int, repeat: int, batch: int,
               tokens_per_sample: int, tok_offset: Optional[torch.Tensor] = None,
               valid: Optional[torch.Tensor] = None) -> None:
    """d_feat [B, T, D] (bf16 or fp32, dense) <- column segment [col_off, col_off + stack * D) of dA [M, K] (bf16)."""
    if da.dtype != torch.bfloat16 or da.dim() != 2 or da.stride(1) != 1:
        raise ValueError("dA must be a bf16 matrix with contiguous columns")
    if d_feat.dim() != 3 or d_feat.stride(2) != 1 or d_feat.dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("d_feat must be bf16 / fp32 [batch, frames, dim] with contiguous dim")
    f = AvcFeat(d_feat.data_ptr(), d_feat.stride(0), d_feat.stride(1), d_feat.shape[1], d_feat.shape[2], stack, repeat,
                _ptr(valid))
    check(load().avc_gather_bwd(C.c_void_p(da.data_ptr()), C.c_int64(da.stride(0)), C.c_int64(col_off), C.byref(f),
                                C.c_int32(dtype_code(d_feat.dtype)), C.c_int32(batch), C.c_void_p(_ptr(tok_offset)),
                                C.c_int32(tokens_per_sample), stream_ptr()))


def cast_bf16(src: torch.Tensor, dst: torch.Tensor, alpha: float = 1.0) -> None:
    """dst_bf16[r, c] = bf16(alpha * src[r, c]) for 2-D bf16 / fp16 / fp32 `src` (columns contiguous)."""
    if src.dim() != 2 or dst.dim() != 2 or src.stride(1) != 1 or dst.stride(1) != 1 or dst.dtype != torch.bfloat16:
        raise ValueError("cast_bf16 needs 2-D matrices with contiguous columns and a bf16 destination")
    check(load().avc_cast_bf16(C.c_void_p(src.data_ptr()), C.c_int32(dtype_code(src.dtype)), C.c_int64(src.stride(0)),
                               C.c_void_p(dst.data_ptr()), C.c_int64(dst.stride(0)), C.c_int64(src.shape[0]),
                               C.c_int64(src.shape[1]), C.c_float(alpha), stream_ptr()))


# ------------------------------------------------------------------------------------------ peer memory (data parallel)
class _DeviceBlock:
    """A raw device allocation exposed through __cuda_array_interface__ so torch can alias it without a copy."""

    def __init__(self, ptr: int, numel: int, typestr: str):
        self.ptr = ptr
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


def comm_alloc(nbytes: int) -> int:
    """Zero-filled cudaMalloc block on the current device that other processes can map (CUDA IPC)."""
    p = C.c_void_p()
    check(load().avc_comm_alloc(C.c_size_t(nbytes), C.byref(p)))
    return int(p.value)


def comm_free(ptr: int) -> None:
    check(load().avc_comm_free(C.c_void_p(ptr)))


def comm_export(ptr: int) -> bytes:
    h = C.create_string_buffer(64)
    check(load().avc_comm_export(C.c_void_p(ptr), h))
    return h.raw


def comm_open(handle: bytes) -> int:
    p = C.c_void_p()
    h = C.create_string_buffer(handle, 64)
    check(load().avc_comm_open(h, C.byref(p)))
    return int(p.value)


def comm_close(ptr: int) -> None:
    check(load().avc_comm_close(C.c_void_p(ptr)))


# NVSwitch multicast bucket (optional transport of the fused all-reduce)
def mc_supported(device: int) -> bool:
    v = C.c_int32(0)
    check(load().avc_mc_supported(C.c_int32(device), C.byref(v)))
    return bool(v.value)


def mc_padded_bytes(world: int, min_bytes: int) -> int:
    v = C.c_uint64(0)
    check(load().avc_mc_padded_bytes(C.c_int32(world), C.c_uint64(min_bytes), C.byref(v)))
    return int(v.value)


def mc_create(world: int, padded_bytes: int):
    """Rank 0: the multicast object and a POSIX fd of it to hand to the other ranks (SCM_RIGHTS)."""
    h, fd = C.c_uint64(0), C.c_int32(-1)
    check(load().avc_mc_create(C.c_int32(world), C.c_uint64(padded_bytes), C.byref(h), C.byref(fd)))
    return int(h.value), int(fd.value)


def mc_import(fd: int) -> int:
    h = C.c_uint64(0)
    check(load().avc_mc_import(C.c_int32(fd), C.byref(h)))
    return int(h.value)


def mc_add_device(handle: int) -> None:
    check(load().avc_mc_add_device(C.c_uint64(handle)))


def mc_bucket_alloc(handle: int, padded_bytes: int) -> AvcMcBucket:
    b = AvcMcBucket()
    check(load().avc_mc_bucket_alloc(C.c_uint64(handle), C.c_uint64(padded_bytes), C.byref(b)))
    return b


def mc_bucket_free(b: AvcMcBucket) -> None:
    check(load().avc_mc_bucket_free(C.byref(b)))


def debug_gemm_profile(buf: Optional[torch.Tensor]) -> None:
    """Per-CTA cycle counters of the next projector GEMM launches into `buf` (int64 [148, 8], zeroed); None: off."""
    check(load().avc_debug_gemm_profile(C.c_void_p(_ptr(buf))))


def comm_flag_bytes() -> int:
    return int(load().avc_comm_flag_bytes())


def as_tensor(ptr: int, numel: int, dtype: torch.dtype, device) -> torch.Tensor:
    """torch view of `numel` elements at device address `ptr` (no copy; the caller keeps the block alive)."""
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DeviceBlock(ptr, numel, typestr), device=device)


def proj_bwd_dw_allreduce(dy: torch.Tensor, x_segs: Sequence[torch.Tensor], dw_segs: Sequence[torch.Tensor],
                          alpha: Sequence[float], comm: AvcComm, extra0: Optional[torch.Tensor] = None,
                          extra1: Optional[torch.Tensor] = None, dy_row_base: int = 0, max_sms: int = 0,
                          bias=None) -> None:
    """proj_bwd_dw whose launch also all-reduces (sums) the dW segments and the extra ranges over the ranks.
    bias = (present, db0, db1, alpha0, alpha1): the launch produces the bias gradients itself (they are its extra
    ranges; extra0 / extra1 are ignored)."""
    al = (C.c_float * len(x_segs))(*alpha)
    if bias is not None:
        bg = _bias_grad(*bias)
        check(load().avc_proj_bwd_dw_db_allreduce(
            C.byref(mat(dy)), C.c_int32(dy_row_base), C.c_int32(len(x_segs)), _mat_array([mat(t) for t in x_segs]),
            _mat_array([mat(t) for t in dw_segs]), al, C.byref(bg), C.byref(comm), C.c_int32(max_sms), stream_ptr()))
        return
    check(load().avc_proj_bwd_dw_allreduce(
        C.byref(mat(dy)), C.c_int32(dy_row_base), C.c_int32(len(x_segs)), _mat_array([mat(t) for t in x_segs]),
        _mat_array([mat(t) for t in dw_segs]), al, C.byref(comm), C.c_void_p(_ptr(extra0)),
        C.c_int64(0 if extra0 is None else extra0.numel()), C.c_void_p(_ptr(extra1)),
        C.c_int64(0 if extra1 is None else extra1.numel()), C.c_int32(max_sms), stream_ptr()))


def comm_signal_extra(comm: AvcComm, extra0_len: int, extra1_len: int) -> None:
    check(load().avc_comm_signal_extra(C.byref(comm), C.c_int64(extra0_len), C.c_int64(extra1_len), stream_ptr()))


def colsum_workspace(cols: int, device) -> torch.Tensor:
    """Scratch for `colsum`: partial sums behind a zero-initialised header of arrival counters (the kernel leaves the
    header zero, so one workspace serves any number of stream-ordered calls)."""
    lib = load()
    ws = torch.empty(lib.avc_colsum_workspace_bytes(cols) // 4, dtype=torch.float32, device=device)
    ws[: lib.avc_colsum_workspace_header_bytes() // 4].zero_()
    return ws


def colsum(dy: torch.Tensor, out0: Optional[torch.Tensor], out1: Optional[torch.Tensor], workspace: torch.Tensor,
           row_flags: Optional[torch.Tensor] = None, flag_rows0: int = 1 << 30, flag_rows1: int = 1 << 30,
           alpha0: float = 1.0, alpha1: float = 1.0, dy_row_base: int = 0, sum_rows: Optional[int] = None,
           comm: Optional[AvcComm] = None) -> None:
    """comm: also flag out0 / out1 ready for that epoch's fused all-reduce (`proj_bwd_dw_allreduce` extras)."""
    m = mat(dy)
    args = (C.byref(m), C.c_int32(dy_row_base), C.c_int32(m.rows - dy_row_base if sum_rows is None else sum_rows),
            C.c_void_p(_ptr(row_flags)), C.c_int32(flag_rows0), C.c_int32(flag_rows1),
            C.c_float(alpha0), C.c_float(alpha1), C.c_void_p(_ptr(out0)), C.c_void_p(_ptr(out1)),
            C.c_void_p(workspace.data_ptr()))
    if comm is None:
        check(load().avc_colsum(*args, stream_ptr()))
    else:
        check(load().avc_colsum_comm(*args, C.byref(comm), stream_ptr()))


def pack_weight(src: torch.Tensor, dst: torch.Tensor, alpha: float = 1.0) -> None:
    """dst_bf16[r, c] = bf16(alpha * src_f32[r, c]); dst may be a column slice of a wider matrix."""
    check(load().avc_pack_weight(
        C.c_void_p(src.data_ptr()), C.c_int64(src.stride(0)), C.c_void_p(dst.data_ptr()), C.c_int64(dst.stride(0)),
        C.c_int64(src.shape[0]), C.c_int64(src.shape[1]), C.c_float(alpha), stream_ptr()))


def make_splice(input_ids: torch.Tensor, placeholder_id: int, pad_id: int, hidden: int, tokens_per_sample: int = 0,
                tok_offset=None, embed_table=None, attention_mask=None, mask_mode: int = 0, label_mode: int = 0,
                labels_in=None, labels_out=None, status=None, elem_size: int = 2,
                av_rows_in_place: bool = False) -> AvcSplice:
    b, s = input_ids.shape
    sp = AvcSplice(
        input_ids.data_ptr(), placeholder_id, pad_id, b, s, hidden, tokens_per_sample, _ptr(tok_offset),
        _ptr(embed_table), 0 if embed_table is None else embed_table.shape[0], _ptr(attention_mask), mask_mode,
        label_mode, _ptr(labels_in), 0 if labels_in is None else labels_in.shape[1], elem_size, _ptr(labels_out),
        _ptr(status), 1 if av_rows_in_place else 0, 0)
    # the struct only holds raw pointers: keep the tensors alive as long as the descriptor is
    sp._keep = (input_ids, tok_offset, embed_table, attention_mask, labels_in, labels_out, status)
    return sp


def splice_fwd(s: AvcSplice, y: Optional[torch.Tensor], inputs_embeds: torch.Tensor) -> None:
    check(load().avc_splice_fwd(C.byref(s), C.c_void_p(_ptr(y)), C.c_void_p(inputs_embeds.data_ptr()), stream_ptr()))


def splice_bwd(s: AvcSplice, d_inputs_embeds: torch.Tensor, dy: torch.Tensor) -> None:
    check(load().avc_splice_bwd(C.byref(s), C.c_void_p(d_inputs_embeds.data_ptr()), C.c_void_p(dy.data_ptr()),
                                stream_ptr()))


def row_resample(x: torch.Tensor, out: torch.Tensor, row_ptr: torch.Tensor, col_idx: torch.Tensor,
                 weight: torch.Tensor) -> None:
    """out[b, i, :] = sum_t weight[t] * x[b, col_idx[t], :] over CSR row i; x [B, S, H], out [B, L, H] contiguous."""
    if not (x.is_contiguous() and out.is_contiguous()) or x.dtype != out.dtype:
        raise ValueError("row_resample needs contiguous tensors of one dtype")
    check(load().avc_row_resample(
        C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int32(dtype_code(x.dtype)), C.c_int32(x.shape[0]),
        C.c_int32(x.shape[1]), C.c_int32(out.shape[1]), C.c_int32(x.shape[2]), C.c_void_p(row_ptr.data_ptr()),
        C.c_void_p(col_idx.data_ptr()), C.c_void_p(weight.data_ptr()), stream_ptr()))


def sumsq_workspace(device) -> torch.Tensor:
    lib = load()
    lib.avc_sumsq_workspace_bytes.restype = C.c_size_t
    return torch.empty(lib.avc_sumsq_workspace_bytes() // 4, dtype=torch.float32, device=device)


def sumsq(x: torch.Tensor, out: torch.Tensor, workspace: torch.Tensor, accumulate: bool = False) -> None:
    """out[0] (+)= sum(x^2) over a contiguous fp32 tensor, deterministic."""
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("sumsq needs a contiguous fp32 tensor")
    check(load().avc_sumsq(C.c_void_p(x.data_ptr()), C.c_int64(x.numel()), C.c_void_p(out.data_ptr()),
                           C.c_void_p(workspace.data_ptr()), C.c_int32(1 if accumulate else 0), stream_ptr()))


def adamw_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, *, lr: float,
               beta1: float, beta2: float, eps: float, weight_decay: float, step: int,
               grad_scale: Optional[torch.Tensor] = None, clip_sumsq: Optional[torch.Tensor] = None,
               max_norm: float = 0.0, packed: Optional[torch.Tensor] = None, packed_alpha: float = 1.0) -> None:
    rows, cols = (param.shape[0], param.shape[1]) if param.dim() == 2 else (1, param.numel())
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("adamw_step needs contiguous fp32 tensors")
    check(load().avc_adamw_step(
        C.c_void_p(param.data_ptr()), C.c_void_p(grad.data_ptr()), C.c_void_p(exp_avg.data_ptr()),
        C.c_void_p(exp_avg_sq.data_ptr()), C.c_int64(rows), C.c_int64(cols), C.c_float(lr), C.c_float(beta1),
        C.c_float(beta2), C.c_float(eps), C.c_float(weight_decay), C.c_int32(step), C.c_void_p(_ptr(grad_scale)),
        C.c_void_p(_ptr(clip_sumsq)), C.c_float(max_norm), C.c_void_p(_ptr(packed)), C.c_int64(0 if packed is None else packed.stride(0)), C.c_float(packed_alpha),
        stream_ptr()))


def gelu_fwd(z: torch.Tensor, out: torch.Tensor, row_flags: Optional[torch.Tensor] = None, flag_bit: int = 1) -> None:
    """out[r] = flag(r) ? gelu(z[r]) : 0 on bf16 [rows, cols] matrices (columns contiguous)."""
    check(load().avc_gelu_fwd(C.c_void_p(z.data_ptr()), C.c_int64(z.stride(0)), C.c_void_p(out.data_ptr()),
                              C.c_int64(out.stride(0)), C.c_int64(z.shape[0]), C.c_int64(z.shape[1]),
                              C.c_void_p(_ptr(row_flags)), C.c_int32(flag_bit), stream_ptr()))


def gelu_bwd(dh: torch.Tensor, z: torch.Tensor, out: torch.Tensor, row_flags: Optional[torch.Tensor] = None,
             flag_bit: int = 1) -> None:
    """out[r] = flag(r) ? dh[r] * gelu'(z[r]) : 0."""
    check(load().avc_gelu_bwd(C.c_void_p(dh.data_ptr()), C.c_int64(dh.stride(0)), C.c_void_p(z.data_ptr()),
                              C.c_int64(z.stride(0)), C.c_void_p(out.data_ptr()), C.c_int64(out.stride(0)),
                              C.c_int64(z.shape[0]), C.c_int64(z.shape[1]), C.c_void_p(_ptr(row_flags)),
                              C.c_int32(flag_bit), stream_ptr()))


def pack_weight_t(src: torch.Tensor, dst: torch.Tensor, alpha: float = 1.0) -> None:
    """dst_bf16[c, r] = bf16(alpha * src_f32[r, c])."""
    check(load().avc_pack_weight_t(C.c_void_p(src.data_ptr()), C.c_int64(src.stride(0)), C.c_void_p(dst.data_ptr()),
                                   C.c_int64(dst.stride(0)), C.c_int64(src.shape[0]), C.c_int64(src.shape[1]),
                                   C.c_float(alpha), stream_ptr()))
