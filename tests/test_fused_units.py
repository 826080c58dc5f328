"""How the fused kernel's launches are dealt out to the persistent CTAs, checked on the CPU (no GPU needed).

`qv_debug_fused_units` runs the host's plan (`plan_units`, qv_fused.cu) and reads it back through the same `unit_geo()`
the kernel's three warp roles call, so what is checked here is what the CTAs will do: every (frame, strip column, row) of
the launch belongs to exactly one work unit, a unit never crosses a column, no CTA gets much more than its share, and the
row line is only chosen when it is the cheaper plan.  A hole or an overlap here would be a wrong (or racy) pixel on the GPU.
"""
import ctypes as C
import random

import numpy as np
import pytest

from qcnn_gpu_b200 import api

WT, PIPE = 120, 14          # output columns per strip column, pipeline fill iterations per unit (qv_fused.cu)


def units_of(sm, n, h, w, row0=0, row1=None, window=False, line=True):
    L = api.lib()
    L.qv_debug_fused_units.argtypes = [C.c_int] * 8 + [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    L.qv_debug_fused_units.restype = C.c_int
    row1 = h if row1 is None else row1
    cnt, grid = C.c_size_t(0), C.c_int(0)
    assert L.qv_debug_fused_units(sm, n, h, w, row0, row1, int(window), int(line), None, C.byref(cnt), C.byref(grid)) == 0
    u = np.zeros((cnt.value, 5), np.int32)
    cap = C.c_size_t(cnt.value)
    assert L.qv_debug_fused_units(sm, n, h, w, row0, row1, int(window), int(line), u.ctypes.data, C.byref(cap), C.byref(grid)) == 0
    assert cap.value == cnt.value
    return u, grid.value


def check_cover(u, grid, n, h, w, row0, row1):
    nstrips = (w + WT - 1) // WT
    cover = np.zeros((n, nstrips, h), np.int32)
    for cta, f, s, y0, y1 in u:
        assert 0 <= cta < grid and 0 <= f < n and 0 <= s < nstrips and row0 <= y0 < y1 <= row1, (cta, f, s, y0, y1)
        cover[f, s, y0:y1] += 1
    assert (cover[:, :, row0:row1] == 1).all(), "a row of some strip column is not covered exactly once"
    assert cover[:, :, :row0].sum() == 0 and cover[:, :, row1:].sum() == 0
    # iterations per CTA (rows + one pipeline fill per unit)
    it = np.zeros(grid, np.int64)
    for cta, f, s, y0, y1 in u:
        it[cta] += (y1 - y0) + PIPE
    return it


def test_bench_workload_uses_the_row_line_and_every_sm():
    u, grid = units_of(148, 64, 1080, 1920)
    it = check_cover(u, grid, 64, 1080, 1920, 0, 1080)
    assert grid == 148
    ideal = 64 * 16 * 1080 / 148
    assert it.max() <= 7600 and it.min() >= ideal          # 7585: 7473 rows + 8 fills; equal segments: 7658
    u2, grid2 = units_of(148, 64, 1080, 1920, line=False)
    it2 = check_cover(u2, grid2, 64, 1080, 1920, 0, 1080)
    assert it2.max() == 7 * (1080 + PIPE) and it.max() < it2.max()


def test_neighbouring_ctas_work_on_neighbouring_columns():
    """The line's positions are permuted so that CTA b and b + 1 are on adjacent strip columns of one frame at the same
    time (their halo columns and output sectors then meet in L2)."""
    u, grid = units_of(148, 64, 1080, 1920)
    first = {}
    for cta, f, s, y0, y1 in u:                               # units arrive in CTA-major order of pieces: keep each CTA's first full column
        if y0 == 0 and y1 == 1080 and cta not in first:
            first[cta] = f * 16 + s
    near = sum(1 for b in range(147) if b in first and b + 1 in first and abs(first[b + 1] - first[b]) == 1)
    assert near >= 120, near


@pytest.mark.parametrize("n,h,w", [(1, 240, 416), (8, 480, 832), (1, 1080, 1920), (1, 4320, 7680), (240, 2160, 3840), (3, 17, 121), (1, 16, 120), (5, 1, 1)])
def test_named_geometries(n, h, w):
    for line in (True, False):
        u, grid = units_of(148, n, h, w, line=line)
        it = check_cover(u, grid, n, h, w, 0, h)
        assert grid <= 148
    # the plan with the line allowed is never the costlier one
    a = check_cover(*units_of(148, n, h, w, line=True), n, h, w, 0, h).max()
    b = check_cover(*units_of(148, n, h, w, line=False), n, h, w, 0, h).max()
    assert a <= b


def test_random_geometries_whole_frames_and_row_windows():
    rng = random.Random(20260)
    for _ in range(300):
        sm = rng.choice([1, 2, 7, 64, 132, 148, 160])
        h, w = rng.randint(1, 700), rng.randint(1, 1500)
        n = rng.randint(1, 9)
        u, grid = units_of(sm, n, h, w, line=rng.random() < 0.7)
        it = check_cover(u, grid, n, h, w, 0, h)
        assert grid <= sm
        # a row window of one frame, as the strip entry points launch it
        r0 = rng.randint(0, h - 1)
        r1 = rng.randint(r0 + 1, h)
        u, grid = units_of(sm, 1, h, w, r0, r1, window=True)
        it = check_cover(u, grid, 1, h, w, r0, r1)
        assert grid <= sm
        # the row line gives every CTA the same number of rows (the last one may have fewer) plus its fills
        rows = np.zeros(grid, np.int64)
        for cta, f, s, y0, y1 in u:
            rows[cta] += y1 - y0
        assert rows.max() - rows[:-1].min() <= 0 if grid > 1 else True


def test_bad_geometry_is_refused():
    L = api.lib()
    L.qv_debug_fused_units.argtypes = [C.c_int] * 8 + [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    L.qv_debug_fused_units.restype = C.c_int
    cnt, grid = C.c_size_t(0), C.c_int(0)
    for args in [(0, 1, 16, 16, 0, 16, 0, 1), (148, 0, 16, 16, 0, 16, 0, 1), (148, 1, 16, 16, 8, 8, 0, 1), (148, 1, 16, 16, 0, 17, 0, 1),
                 (148, 2, 16, 16, 0, 16, 1, 1)]:
        assert L.qv_debug_fused_units(*args, None, C.byref(cnt), C.byref(grid)) != 0, args
