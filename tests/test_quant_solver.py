"""the quant-parameter solver (training/quantization.py:5-64) and the float -> int8 model
quantiser re-hosted in C++.  Golden vectors were produced by importing the reference's own module
(tests/golden/make_quant_solver_golden.py); doubles are compared bit for bit."""
import json
import os

import numpy as np
import pytest

from qcnn_gpu_b200 import api
from qcnn_gpu_b200.host import formats

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quant_solver_golden.json")))


def test_solver_matches_reference_bit_for_bit():
    assert len(GOLD) >= 44
    for case in GOLD:
        stepw = [float.fromhex(v) for v in case["stepw"]]
        blu = [float.fromhex(v) for v in case["blu"]]
        want = np.array([[float.fromhex(v) for v in r] for r in case["rows"]])
        got = api.solve_quant_params(stepw, blu)
        assert got.tobytes() == want.tobytes()
        # the defining property (quantization.py:5-14): the BLU bound maps just above 127
        for l in range(5):
            blu_q, mul, sh = got[l, 3], got[l, 4], int(got[l, 5])
            assert 127 <= int((blu_q + (1 << (sh - 1)) // int(mul)) * mul) >> sh <= 128


def test_solver_output_round_trips_through_both_file_formats(tmp_path):
    case = GOLD[2]
    rows = api.solve_quant_params([float.fromhex(v) for v in case["stepw"]], [float.fromhex(v) for v in case["blu"]])
    c = tmp_path / "quant_params_cpp_32.data"
    api.write_quant_params_cpp(str(c), rows)
    assert c.stat().st_size == 288
    want = [[int(r[3]), int(r[4]), int(r[5])] for r in rows]
    assert api.read_quant_params(str(c)).tolist() == want
    p = tmp_path / "quant_params32.data"
    formats.write_quant_params_pickle(str(p), [[r[0], r[1], r[2], r[3], r[4], int(r[5])] for r in rows.tolist()])
    assert api.read_quant_params(str(p)).tolist() == want
    with pytest.raises(api.QVError):
        api.solve_quant_params([0.01] * 5 + [0.0], [0.1] * 6)


def test_quantize_layer_matches_numpy_formula():
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((16, 48, 3, 3)) * 0.4).astype(np.float32)
    w.flat[:5] = [0.0125, -0.0125, 0.0375, 5.0, -5.0]        # exact ties (half-to-even) and clipping
    b = (rng.standard_normal(16) * 0.05).astype(np.float32)
    stepw, ratio = 0.025, 365.1238183116871
    wq, bq = api.quantize_layer(w, b, stepw, ratio)
    assert np.array_equal(wq, np.clip(np.around(w.astype(np.float64) / stepw), -128, 127).astype(np.int8))   # training/model.py:167
    assert np.array_equal(bq, np.around(b.astype(np.float64) * ratio / stepw).astype(np.int32))              # quantization.py:104
    assert wq.flat[3] == 127 and wq.flat[4] == -128
