"""Thin tensor-level wrappers over the C-ABI ViT building blocks (include/clipguide_b200.h, "CLIP ViT").
They only marshal torch tensors into pointers/sizes; all arithmetic is in csrc/vit_*.cu."""
import torch

from clip_diffusion_b200 import _lib


def gemm_bf16_tn(a, b, epilogue, bias=None, out=None, aux=None, pos=None, g2=0, m=None):
    """acc[M,N] = a[M,K] @ b[N,K]^T with a fused epilogue (see CG_EPI_* in the header).  a, b: bf16 row-major
    (row stride may exceed K).  ``out``/``aux`` must be preallocated for in-place / auxiliary epilogues."""
    _lib.require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(-1) == 1 and b.stride(-1) == 1
    M = a.shape[0] if m is None else m
    K = a.shape[1]
    N = b.shape[0]
    assert b.shape[1] == K
    if out is None:
        dt = torch.float32 if epilogue in (_lib.EPI_F32,) else torch.bfloat16
        assert epilogue != _lib.EPI_PATCH_POS_F32, "the patch epilogue writes into a preallocated token buffer"
        if epilogue == _lib.EPI_BIAS_RESID_F32:
            dt = torch.float32
        out = torch.empty(M, N, device=a.device, dtype=dt)
    if epilogue == _lib.EPI_BIAS_QGELU_BF16 and aux is None:
        aux = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    _lib.call(
        "cg_gemm_bf16_tn", _lib.ptr(a), _lib.ptr(b), M, N, K, a.stride(0), b.stride(0), epilogue, _lib.ptr(bias), _lib.ptr(out),
        _lib.ptr(aux), out.stride(0), _lib.ptr(pos), g2,
    )
    return (out, aux) if epilogue == _lib.EPI_BIAS_QGELU_BF16 else out
