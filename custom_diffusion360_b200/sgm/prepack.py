"""Weight pre-packing: reference parameter layouts -> the layouts the kernels consume.

Done once at load time (state-dict keys and shapes stay the reference's; packed copies live beside
them).  Pure index shuffles and dtype casts — torch is used here as a tensor container only.
"""
from __future__ import annotations

import torch

from .. import ops


def pack_conv3x3(w: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d weight [Cout, Cin, 3, 3] -> bf16 [Cout, 9*Cin] with k = (ky*3+kx)*Cin + c,
    the K order of the implicit-GEMM taps (include/cd360.h, cd360_gemm_bf16 conv mode)."""
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    return w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous().to(torch.bfloat16)


def pack_conv3x3_padded(w: torch.Tensor, cin_pad: int, cout_pad: int) -> torch.Tensor:
    """Same, zero-padding Cin / Cout (e.g. the 4-channel input conv run via im2col, K = 36 -> 64;
    the 4-channel output conv, N = 4 -> 8)."""
    cout, cin = w.shape[:2]
    wp = torch.zeros(cout_pad, cin_pad, 3, 3, dtype=w.dtype, device=w.device)
    wp[:cout, :cin] = w
    return pack_conv3x3(wp)


def pack_conv3x3_im2col(w: torch.Tensor, kpad: int) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> bf16 [Cout, kpad], k = (ky*3+kx)*Cin + c, zero padded (small Cin)."""
    cout, cin = w.shape[:2]
    flat = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    out = torch.zeros(cout, kpad, dtype=torch.bfloat16, device=w.device)
    out[:, : 9 * cin] = flat.to(torch.bfloat16)
    return out


def pack_geglu(w: torch.Tensor, bias: torch.Tensor):
    """GEGLU.proj weight [8c, c] / bias [8c] (attention.py:92: first half x, second half gate)
    -> rows interleaved in blocks so that every N tile of the GEMM holds [x-block | gate-block]."""
    n = w.shape[0]
    blk = ops.geglu_pack_block(n) if w.is_cuda else _geglu_block(n)
    half = n // 2
    nb = half // blk
    idx = torch.arange(n, device=w.device).view(2, nb, blk).permute(1, 0, 2).reshape(-1)
    return w[idx].contiguous().to(torch.bfloat16), bias[idx].contiguous().float()


def _geglu_block(n: int) -> int:
    if n % 256 == 0:
        return 128
    if n % 128 == 0:
        return 64
    raise ValueError(f"GEGLU width {n} not a multiple of 128")


def pack_conv3x3_bwd(w: torch.Tensor, cout_pad: int = 0) -> torch.Tensor:
    """Weight pack of the DATA gradient of a stride-1 pad-1 3x3 conv (training step): dX is the
    same convolution of dY with the taps flipped and the channel roles swapped,
    W'[ci, co, ky, kx] = W[co, ci, 2-ky, 2-kx] -> bf16 [Cin, 9*Cout'] in the implicit-GEMM K order.
    cout_pad zero-pads the (small) Cout of the UNet's 4-channel output conv to a full 64-wide K chunk."""
    cout, cin = w.shape[:2]
    wt = w.flip(2, 3).permute(1, 0, 2, 3)
    if cout_pad > cout:
        wp = torch.zeros(cin, cout_pad, 3, 3, dtype=w.dtype, device=w.device)
        wp[:, :cout] = wt
        wt = wp
    return pack_conv3x3(wt)


def transposed(w: torch.Tensor) -> torch.Tensor:
    """[N, K] -> contiguous bf16 [K, N]: the `w` operand of dX = dY W through cd360_gemm_bf16
    (which computes A W'^T, so W' = W^T)."""
    return w.detach().t().contiguous().to(torch.bfloat16)
