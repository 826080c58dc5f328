"""File formats either side of the path: quant-param files, static model files, YUV, PSNR.
CPU only: exercises the C++ readers through the C ABI (no compute calls) against Python."""
import os
import pickle
import struct

import numpy as np
import pytest

from qcnn_gpu_b200 import api
from qcnn_gpu_b200.host import formats, synth

REF_TRAINING = "/root/reference/training"


def test_abi_exports_every_declared_symbol():
    L = api.lib()
    for name in api.ABI_SYMBOLS:
        assert hasattr(L, name), name
    # and every QV_API declaration in the header is in that list
    hdr = open(os.path.join(os.path.dirname(api._PKG), "include", "qvrcnn_b200.h")).read()
    import re
    declared = set(re.findall(r"QV_API\s+[\w\s\*]+?\b(qv_\w+)\s*\(", hdr))
    assert declared == set(api.ABI_SYMBOLS), declared ^ set(api.ABI_SYMBOLS)
    assert b"sm_100a" in L.qv_version()


@pytest.mark.parametrize("qp", [22, 27, 32, 37])
def test_quant_params_pickle_and_cpp(tmp_path, qp):
    rows = formats.qparams_rows_from_table(qp)
    p = tmp_path / ("quant_params%d.data" % qp)
    formats.write_quant_params_pickle(str(p), rows)
    got = api.read_quant_params(str(p))
    assert got.tolist() == [list(t) for t in formats.SHIPPED_QPARAMS[qp]]
    c = tmp_path / ("quant_params_cpp_%d.data" % qp)
    formats.write_quant_params_cpp(str(c), rows)
    assert api.read_quant_params(str(c)).tolist() == got.tolist()


@pytest.mark.parametrize("proto", [2, 3, 4])
def test_quant_params_other_pickle_protocols(tmp_path, proto):
    rows = [[0.5, 255, np.float64(0.1), np.float64(1000 + i), 31.0, 11] for i in range(6)]
    p = tmp_path / "q.data"
    with open(p, "wb") as fp:
        pickle.dump(rows, fp, protocol=proto)
    got = api.read_quant_params(str(p))
    assert got[:, 0].tolist() == [1000 + i for i in range(6)] and set(got[:, 1]) == {31} and set(got[:, 2]) == {11}


@pytest.mark.skipif(not os.path.isdir(REF_TRAINING), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("qp", [22, 27, 32, 37])
def test_shipped_pickles_decode_to_the_recorded_table(qp):
    """The four files the reference ships (training/quant_params{22,27,32,37}.data) through the
    C++ mini-unpickler == python pickle == the table hard-wired in formats.SHIPPED_QPARAMS."""
    path = os.path.join(REF_TRAINING, "quant_params%d.data" % qp)
    got = api.read_quant_params(path)
    py = formats.load_quant_params_pickle(path)
    assert got.tolist() == [[int(r[3]), int(r[4]), int(r[5])] for r in py]
    assert got.tolist() == [list(t) for t in formats.SHIPPED_QPARAMS[qp]]


def test_quant_params_errors(tmp_path):
    with pytest.raises(api.QVError) as e:
        api.read_quant_params(str(tmp_path / "missing.data"))
    assert e.value.code == -2
    bad = tmp_path / "bad.data"
    bad.write_bytes(b"\x80\x03]q\x00.")          # empty list
    with pytest.raises(api.QVError):
        api.read_quant_params(str(bad))
    trunc = tmp_path / "trunc.data"
    good = tmp_path / "good.data"
    formats.write_quant_params_pickle(str(good), formats.qparams_rows_from_table(32))
    trunc.write_bytes(good.read_bytes()[:200])
    with pytest.raises(api.QVError):
        api.read_quant_params(str(trunc))


def test_model_file_sizes_and_roundtrip(models):
    assert formats.MODEL_FILE_SIZE == 60028 and formats.MODEL_FILE_SIZE_HWCN == 55228
    m = models[32]
    img = formats.write_model_vect_c(m)
    back = formats.read_model_vect_c(img)
    for a, b in zip(m.w, back.w):
        assert np.array_equal(a, b)
    for a, b in zip(m.b, back.b):
        assert np.array_equal(a, b)
    assert [tuple(q) for q in back.qparams] == [tuple(q) for q in m.qparams]
    # padded lanes of C1 (1 -> 4 channels) are zero, as HWCN2NCHW_VECT_C_CPU leaves them (mat.cu:108)
    c1 = np.frombuffer(img[:6400], np.int8).reshape(64, 1, 5, 5, 4)
    assert not c1[..., 1:].any()


def test_hwcn_to_vect_c_converter_matches_python(tmp_path, models):
    """qv_convert_model_hwcn_to_vect_c == model_qfp_HWCN2NCHW_VECT_C (inference/qvrcnn.cu:558-585):
    index map out[k][c>>2][r][s][c&3] = in[r][s][c][k] (inference/mat.cu:109-117)."""
    m = models[27]
    fin, fout = tmp_path / "hwcn.data", tmp_path / "vect_c.data"
    fin.write_bytes(formats.write_model_hwcn(m))
    api.convert_model_hwcn_to_vect_c(str(fin), str(fout))
    assert fout.read_bytes() == formats.write_model_vect_c(m)
    # spot-check the reference's literal index formula on C2_2
    cin, cout, k = formats.LAYERS[2]
    hw = np.frombuffer(formats.write_model_hwcn(m), np.int8)
    off_h = sum(l[0] * l[1] * l[2] * l[2] + 4 * l[1] + 12 for l in formats.LAYERS[:2])
    off_v = sum(formats.vect_c_wsize(*l) + 4 * l[1] + 12 for l in formats.LAYERS[:2])
    vc = np.frombuffer(fout.read_bytes(), np.int8)
    rng = np.random.default_rng(0)
    for _ in range(200):
        i, j, r, s = rng.integers(cout), rng.integers(cin), rng.integers(k), rng.integers(k)
        a = vc[off_v + i * (k * k * 16 * 4) + (j >> 2) * (k * k * 4) + r * (k * 4) + s * 4 + (j & 3)]
        b = hw[off_h + r * k * cin * cout + s * cin * cout + j * cout + i]
        assert a == b == m.w[2][i, j, r, s]
    with pytest.raises(api.QVError):
        api.convert_model_hwcn_to_vect_c(str(tmp_path / "nope"), str(fout))


@pytest.mark.parametrize("qp", [27, 32])
def test_hwcn_to_vect_c_converter_matches_the_reference_converter(tmp_path, qp):
    """Golden produced by the reference's OWN model_qfp_HWCN2NCHW_VECT_C (inference/qvrcnn.cu:558-585 ->
    HWCN2NCHW_VECT_C_CPU, inference/mat.cu:97-119), compiled unmodified into oracle/_ref/qcnn_ref_convert and run by
    tests/golden/make_converter_golden.py: same HWCN file in, byte-identical NCHW_VECT_C file out."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_converter_qp%d.npz" % qp))
    fin, fout = tmp_path / "hwcn.data", tmp_path / "vect_c.data"
    fin.write_bytes(g["hwcn"].tobytes())
    api.convert_model_hwcn_to_vect_c(str(fin), str(fout))
    assert fout.read_bytes() == g["vect_c"].tobytes()
    # and the python-side writers used by every other test agree with the reference's bytes
    m = formats.read_model_hwcn(g["hwcn"].tobytes())
    assert formats.write_model_vect_c(m) == g["vect_c"].tobytes()


def test_yuv_io_and_psnr(tmp_path):
    anchor, ori = synth.make_frames(7, 3, 18, 34)
    # a "real" 4:2:0 file: Y then non-zero chroma the reader must skip (inference/yuv_data.cpp:32-38)
    p = tmp_path / "a.yuv"
    with open(p, "wb") as fp:
        for f in range(3):
            fp.write(anchor[f].tobytes())
            fp.write(bytes([0x55]) * (18 * 34 // 2))
    d = api.VRCNNData(3, 18, 34)
    o = tmp_path / "o.yuv"
    formats.write_yuv420_luma(str(o), ori)
    d.read_data(str(o), str(p))
    assert np.array_equal(d.input, anchor) and np.array_equal(d.ori, ori)
    psnr_py, sse = formats.psnr(anchor, ori)
    assert d.psnr(d.input) == pytest.approx(psnr_py, abs=1e-12)
    assert api.psnr_from_sse(sse, anchor.size) == d.psnr(d.input)   # bit-identical: exact integer partial sums
    d.recon[:] = anchor
    r = tmp_path / "r.yuv"
    d.save_recon_as(str(r))
    raw = r.read_bytes()
    assert len(raw) == 3 * (18 * 34 * 3 // 2)
    assert raw[:18 * 34] == anchor[0].tobytes() and not any(raw[18 * 34:18 * 34 * 3 // 2])   # zero chroma
    d1 = api.VRCNNData(1, 18, 34)
    d1.read_frame(str(o), str(p), 2)
    assert np.array_equal(d1.input[0], anchor[2]) and np.array_equal(d1.ori[0], ori[2])
    with pytest.raises(api.QVError):
        d.read_data(str(tmp_path / "missing.yuv"), str(p))
    short = api.VRCNNData(4, 18, 34)
    with pytest.raises(api.QVError):
        short.read_data(str(o), str(p))


def test_no_gpu_fails_loudly():
    """Without a CUDA device the product refuses to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.QVError) as e:
        api.QVRCNN(0, 1, 1, 16, 16)
    assert e.value.code == -3 and "no CUDA device" in str(e.value)


def test_host_alloc_falls_back_to_malloc_without_a_gpu():
    """qv_host_alloc / qv_host_free (the vrcnn_data shim's frame buffers): page-locked on a GPU box, plain malloc here."""
    import ctypes
    L = api.lib()
    p = L.qv_host_alloc(1 << 20)
    assert p
    ctypes.memset(p, 0x5A, 1 << 20)
    assert ctypes.string_at(p + (1 << 20) - 4, 4) == b"\x5a" * 4
    L.qv_host_free(p)
    L.qv_host_free(None)
    q = L.qv_host_alloc(0)
    assert q
    L.qv_host_free(q)
