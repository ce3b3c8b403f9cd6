"""Device-side plumbing shared by the Python mirror: pointer/stream extraction and argument checks.
PyTorch is used for allocation and streams only."""
from __future__ import annotations

import torch

from . import _lib


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.device.type != "cuda":
        raise _lib.CometB200Error(
            f"{name} is on {t.device}: comet_pose_estimation_b200 only runs on a CUDA (sm_100a) device and has no "
            "CPU fallback"
        )


def f32(t: torch.Tensor) -> torch.Tensor:
    """float32 view of ``t`` (cast only when the caller handed another dtype, e.g. bf16 encoder output)."""
    return t if t.dtype == torch.float32 else t.float()


def f32c(t: torch.Tensor) -> torch.Tensor:
    return f32(t).contiguous()


def inner_contig(t: torch.Tensor) -> torch.Tensor:
    """float32 tensor whose last dimension is contiguous; outer dimensions may be arbitrary strides
    (the kernels take element strides, so permuted views such as track_feats.permute(0,2,1,3) cost nothing)."""
    t = f32(t)
    if t.dim() and t.shape[-1] > 1 and t.stride(-1) != 1:
        t = t.contiguous()
    return t


def prec_mode() -> int:
    """SURVEY 8(b) threading note: honour an active autocast context (bf16 rounding points of the reference)."""
    if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        return _lib.PREC_BF16_AUTOCAST
    return _lib.PREC_F32


def pad_mode(padding_mode: str) -> int:
    if padding_mode == "zeros":
        return _lib.PAD_ZEROS
    if padding_mode == "border":
        return _lib.PAD_BORDER
    raise ValueError(f"padding_mode {padding_mode!r} is not supported (the reference path uses 'zeros' and 'border')")


def require_no_grad(*tensors: torch.Tensor) -> None:
    """The kernels are forward-only (the reference runs the tracker frozen under ``torch.no_grad()``,
    E2Epose2.py:176, train_util.py:311-318).  Outputs are written into fresh buffers without an autograd node, so a
    call that autograd would have to differentiate through is refused instead of silently dropping the gradient."""
    if torch.is_grad_enabled():
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.requires_grad:
                raise RuntimeError(
                    "comet_pose_estimation_b200 kernels are forward-only: an input requires grad while autograd is "
                    "enabled. Run the tracker under torch.no_grad() (as COMET.forward_all does) or detach the inputs.")
