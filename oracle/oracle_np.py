"""Second, independent restatement of forward_blu (torch-CPU conv2d + numpy), used ONLY to
cross-validate oracle/qvrcnn_oracle.c in tests.  TEST INFRASTRUCTURE, not the product.

Differs from the C oracle on purpose: it starts from the in-memory Model (not the file image),
computes convolutions with torch.nn.functional.conv2d in float64 (exact: |acc| < 2^53) and
re-derives the fp32 materialisation with numpy float32 casts.
Reference citations as in the C oracle.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _conv_acc(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    """int conv, cross-correlation, pad (k-1)/2, stride 1 (inference/cnn.cu:44-49). a:[C,H,W]."""
    k = w.shape[-1]
    t = F.conv2d(torch.from_numpy(a.astype(np.float64))[None], torch.from_numpy(w.astype(np.float64)),
                 padding=(k - 1) // 2)
    acc = t[0].numpy()
    assert np.all(np.abs(acc) < 2.0 ** 52)
    return np.rint(acc).astype(np.int64)


def _u_fp32(acc: np.ndarray, b: np.ndarray) -> np.ndarray:
    """u = fl32(acc) + fl32(b) in fp32 (inference/cnn.cu:104,148-155)."""
    return acc.astype(np.float32) + b.astype(np.float32)[:, None, None]


def _blu(u: np.ndarray, blu: int, mul: int, shift: int) -> np.ndarray:
    """inference/mat.cu:262-303."""
    bias = (1 << (shift - 1)) // mul
    ui = u.astype(np.int64)                       # (int)temp, truncation; values are integral anyway
    prod = ((ui + bias) * mul) & 0xFFFFFFFF        # int32 wrap
    prod = np.where(prod >= 1 << 31, prod - (1 << 32), prod)
    q = (prod >> shift) & 0xFF                     # arithmetic shift, then (char) truncation
    q = np.where(q >= 128, q - 256, q)
    out = np.where(u > np.float32(blu), 127, np.where(u < 0, 0, q))
    return out.astype(np.int8)


def forward_frame(model, x: np.ndarray):
    """x: u8 [H,W] -> dict(rec, a1, a2, a3, u4)."""
    xp = (x.astype(np.int32) - 128).astype(np.int8)[None]            # cnn.cu:450
    q = model.qparams
    a1 = _blu(_u_fp32(_conv_acc(xp, model.w[0]), model.b[0]), *q[0])
    a2 = np.concatenate([_blu(_u_fp32(_conv_acc(a1, model.w[1]), model.b[1]), *q[1]),
                         _blu(_u_fp32(_conv_acc(a1, model.w[2]), model.b[2]), *q[2])])
    a3 = np.concatenate([_blu(_u_fp32(_conv_acc(a2, model.w[3]), model.b[3]), *q[3]),
                         _blu(_u_fp32(_conv_acc(a2, model.w[4]), model.b[4]), *q[4])])
    u4 = _u_fp32(_conv_acc(a3, model.w[5]), model.b[5])[0].astype(np.int64)
    _, mul, shift = q[5]
    t = (u4 * mul + (1 << (shift - 1))) & 0xFFFFFFFF                   # cnn.cu:512-516, int32 wrap
    t = np.where(t >= 1 << 31, t - (1 << 32), t) >> shift
    r = (x.astype(np.int64) + t) & 0xFFFF                              # (short) cnn.cu:517
    r = np.where(r >= 1 << 15, r - (1 << 16), r)
    rec = np.clip(r, 0, 255).astype(np.uint8)
    return dict(rec=rec, a1=a1, a2=a2, a3=a3, u4=u4.astype(np.int32))
