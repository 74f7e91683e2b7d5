"""The fused QKV-projection + cross-view-attention kernel (bf16 mode, LayerNorm folded; multiview_mpl.py:48-64 behind norm1) in
isolation through the C-ABI test hook, against an fp64 reference built from the same rounded operands:
    q|k|v = rstd * (bf16(x) W'^T - mu * colsum(W')) + b',  softmax(q k^T * hd^-0.5) v  over the V views of a pose, per head."""
import numpy as np
import pytest
import torch

from openmpl_b200 import _lib

pytestmark = pytest.mark.gpu
H = 8


def _pad256(m):
    return (m + 255) // 256 * 256


def _al(n):
    return (n + 255) // 256 * 256


@pytest.mark.parametrize("V,poses", [(4, 64), (4, 77), (2, 300), (8, 40), (4, 5000), (2, 9000), (8, 3000),
                                     # pose-aligned row tiling: 30 (V = 3, 5, 6) or 28 (V = 7) rows per 32-lane quarter
                                     (3, 50), (5, 77), (6, 41), (7, 33), (5, 1), (3, 7001), (5, 4000), (6, 3000), (7, 2500)])
@pytest.mark.parametrize("HD", [136, 68])
def test_fused_qkv_attention_against_fp64(V, poses, HD):
    """HD = 136: one head per 416-column tile (hm_0, D = 1088); HD = 68: two heads per tile (the "chosen" architecture, D = 544,
    whose K = 544 ends in half a K block that TMA zero-fills)."""
    L = _lib.lib()
    D = H * HD
    tiles = H if HD == 136 else H // 2
    M = poses * V
    g = torch.Generator(device="cuda").manual_seed(V * 1000 + poses + HD)
    x = torch.randn(M, D, device="cuda", generator=g) * 0.8 + 0.05 * torch.randn(M, 1, device="cuda", generator=g)
    gamma = 1.0 + 0.1 * torch.randn(D, device="cuda", generator=g)
    beta = 0.1 * torch.randn(D, device="cuda", generator=g)
    W = (torch.randn(3 * D, D, device="cuda", generator=g) / np.sqrt(D)).contiguous()
    b = torch.randn(3 * D, device="cuda", generator=g) * 0.5
    xb = x.to(torch.bfloat16).contiguous()
    slots = 4
    Mp = _pad256(M)
    stats = torch.zeros(slots, Mp, 2, device="cuda")
    h = D // 2
    stats[0, :M, 0] = x[:, :h].sum(1); stats[0, :M, 1] = (x[:, :h] ** 2).sum(1)
    stats[3, :M, 0] = x[:, h:].sum(1); stats[3, :M, 1] = (x[:, h:] ** 2).sum(1)
    eps, scale = 1e-6, HD ** -0.5
    att = torch.full((M, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    scratch = torch.empty(_al(tiles * 416 * D * 2) + 2 * _al(tiles * 416 * 4), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(L.mpl_test_qkv_attn(xb.data_ptr(), W.data_ptr(), b.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(),
                                   slots, eps, scale, att.data_ptr(), M, D, H, V, scratch.data_ptr(), scratch.numel(), stream))
    torch.cuda.synchronize()
    # reference from the same rounded operands: W' = bf16(W gamma) (q rows additionally carry scale * log2 e in the kernel: a
    # different rounding of the same number, inside the tolerance), raw rows in bf16, folded bias in fp64
    s1 = stats[:, :M, 0].double().sum(0); s2 = stats[:, :M, 1].double().sum(0)
    mu = s1 / D
    rstd = 1.0 / torch.sqrt((s2 / D - mu * mu).clamp_min(0) + eps)
    Wf = (W * gamma).to(torch.bfloat16).double()
    biasf = b.double() + W.double() @ beta.double()
    qkv = rstd[:, None] * (xb.double() @ Wf.T - mu[:, None] * Wf.sum(1)[None, :]) + biasf
    qkv = qkv.to(torch.bfloat16).double()                      # the kernel keeps q, k, v as bf16 before the attention
    q, k, v = (qkv[:, i * D:(i + 1) * D].reshape(poses, V, H, HD).permute(0, 2, 1, 3) for i in range(3))
    p = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(M, D)
    assert torch.isfinite(att.float()).all()
    err = (att.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1.2e-2, f"V={V} poses={poses}: {err:.3e}"
    # and against the true LayerNorm -> Linear -> attention in fp64 (bf16 operand rounding only)
    ln = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), eps)
    t = ln @ W.double().T + b.double()
    q, k, v = (t[:, i * D:(i + 1) * D].reshape(poses, V, H, HD).permute(0, 2, 1, 3) for i in range(3))
    true = (torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1) @ v).permute(0, 2, 1, 3).reshape(M, D)
    err_true = (att.double() - true).abs().max().item() / true.abs().max().item()
    assert err_true <= 3e-2, f"vs fp64 LayerNorm + Linear + attention: {err_true:.3e}"


def test_fused_qkv_attention_rejects_other_shapes():
    L = _lib.lib()
    x = torch.zeros(1024, device="cuda")
    args = (x.data_ptr(),) * 6
    assert L.mpl_test_qkv_attn(*args, 2, 1e-6, 0.1, x.data_ptr(), 12, 600, 8, 4, x.data_ptr(), 4096, None) == _lib.MPL_ERR_UNSUPPORTED
    assert L.mpl_test_qkv_attn(*args, 2, 1e-6, 0.1, x.data_ptr(), 12, 476, 7, 4, x.data_ptr(), 4096, None) == _lib.MPL_ERR_UNSUPPORTED
    assert L.mpl_test_qkv_attn(*args, 2, 1e-6, 0.1, x.data_ptr(), 18, 1088, 8, 9, x.data_ptr(), 4096, None) == _lib.MPL_ERR_UNSUPPORTED
    assert L.mpl_test_qkv_attn(*args, 2, 1e-6, 0.1, x.data_ptr(), 12, 1088, 8, 1, x.data_ptr(), 4096, None) == _lib.MPL_ERR_UNSUPPORTED
