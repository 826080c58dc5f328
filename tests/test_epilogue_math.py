"""The fused kernel's fast requantiser (qv_fused.cu: requant_store<FAST>) is an algebraic rewrite of the
reference's BLU formula (inference/mat.cu:286-291).  These CPU tests prove the rewrite on the whole
accumulator range for every shipped parameter row and, with hypothesis, on random rows that satisfy the
precondition the host checks at upload (fused_upload: mkq)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from qcnn_gpu_b200.host import formats


def reference_blu(u, blu, mul, sh):
    """mat.cu:268,286-291 with int32 wrap and the (char) store."""
    rb = (1 << (sh - 1)) // mul
    prod = ((u + rb) * mul) & 0xFFFFFFFF
    prod = np.where(prod >= 1 << 31, prod - (1 << 32), prod)
    q = (prod >> sh) & 0xFF
    q = np.where(q >= 128, q - 256, q)
    return np.where(u > blu, 127, np.where(u < 0, 0, q))


def fast_precondition(blu, mul, sh):
    rb = (1 << (sh - 1)) // mul
    top = (blu + rb) * mul
    return sh <= 24 and mul < (1 << sh) and top < (1 << 31) and (top >> sh) == 127 and blu + rb < (1 << 30)


def fast_blu(acc, bias, blu, mul, sh):
    """What the kernel computes: t = relu(min(acc + (b + rb), blu + rb)); q = byte 3 of t * (mul << (24 - sh))."""
    rb = (1 << (sh - 1)) // mul
    t = np.maximum(np.minimum(acc + (bias + rb), blu + rb), 0)
    p = (t * (mul << (24 - sh)))
    assert p.max() < 1 << 32
    return (p >> 24) & 0xFF


def test_fast_requant_equals_reference_for_every_shipped_row():
    for qp, rows in formats.SHIPPED_QPARAMS.items():
        for blu, mul, sh in rows[:5]:
            assert fast_precondition(blu, mul, sh), (qp, blu, mul, sh)
            u = np.arange(-(1 << 17), blu + (1 << 17), dtype=np.int64)
            for bias in (0, -4000, 123456):
                assert np.array_equal(fast_blu(u - bias, bias, blu, mul, sh), reference_blu(u, blu, mul, sh)), (qp, blu, mul, sh)
            # and at the far ends of the exact-integer envelope
            far = np.array([-(1 << 24), -(1 << 24) + 1, (1 << 24) - 1, 1 << 24], dtype=np.int64)
            assert np.array_equal(fast_blu(far, 0, blu, mul, sh), reference_blu(far, blu, mul, sh))


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 24), st.data())
def test_fast_requant_equals_reference_whenever_the_precondition_holds(sh, data):
    """Rows constructed to satisfy the precondition: ((blu + rb) * mul) >> sh == 127."""
    mul = data.draw(st.integers(1, (1 << sh) - 1))
    rb = (1 << (sh - 1)) // mul
    lo = -(-(127 << sh) // mul)                     # smallest t with (t*mul)>>sh == 127
    hi = -(-(128 << sh) // mul) - 1                 # largest
    if hi < lo or lo - rb < 0 or hi * mul >= 1 << 31 or hi >= 1 << 30:
        return
    blu = data.draw(st.integers(max(lo - rb, 0), hi - rb))
    bias = data.draw(st.integers(-(1 << 23), 1 << 23))
    assert fast_precondition(blu, mul, sh)
    rng = np.random.default_rng(blu ^ mul)
    u = np.concatenate([rng.integers(-(1 << 24), 1 << 24, 4096), np.arange(-64, 64), np.arange(blu - 64, blu + 64)]).astype(np.int64)
    assert np.array_equal(fast_blu(u - bias, bias, blu, mul, sh), reference_blu(u, blu, mul, sh))


def test_precondition_rejects_rows_the_rewrite_cannot_handle():
    assert not fast_precondition(5000, 100, 12)          # BLU(blu) = 122
    assert not fast_precondition(8000, 100, 12)          # 195: wraps as char
    assert not fast_precondition(4526, 115 << 14, 26)    # shift > 24
    assert not fast_precondition(0, 5, 24)               # the shipped (unused) C4 row of QP22
