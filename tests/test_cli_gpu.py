"""The reference's driver surface (inference/kernel.cu:74-138, run_all.bat): `qcnn_gpu <ori.yuv>
<anchor_prefix> <H> <W>` on YUV 4:2:0 files, report lines, log.txt / recon_psnr.data appends."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from qcnn_gpu_b200.host import formats, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "qcnn_gpu_b200", "qcnn_gpu")


def _write_yuv(path, luma, chroma_value):
    with open(path, "wb") as fp:
        for f in range(luma.shape[0]):
            fp.write(luma[f].tobytes())
            fp.write(bytes([chroma_value]) * (luma.shape[1] * luma.shape[2] // 2))


@pytest.mark.parametrize("frames,gpus", [(1, 1), (3, 1)])
def test_cli_matches_oracle_report(tmp_path, models, frames, gpus):
    from oracle import oracle
    qp, h, w = 37, 64, 112
    anchor, ori = synth.make_frames(0xC0FFEE + 11, frames, h, w)
    _write_yuv(tmp_path / "ori.yuv", ori, 0x80)
    _write_yuv(tmp_path / ("anchor_Q%d.yuv" % qp), anchor, 0x33)
    (tmp_path / ("model_%d.data" % qp)).write_bytes(formats.write_model_vect_c(models[qp]))
    assert os.path.exists(CLI), "qcnn_gpu CLI not built"
    p = subprocess.run([CLI, "ori.yuv", "anchor_", str(h), str(w), "--model", "model_%d.data", "--qp", str(qp),
                        "--frames", str(frames), "--gpus", str(gpus), "--save-recon", "recon.yuv"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    want = oracle.OracleModel(formats.write_model_vect_c(models[qp])).forward_blu(anchor)
    before = float(re.search(r"before net:PSNR=([0-9.]+)", p.stdout).group(1))
    after = float(re.search(r"after quantized net:PSNR=([0-9.]+)", p.stdout).group(1))
    assert before == pytest.approx(round(oracle.psnr(anchor, ori)[0], 3), abs=1.1e-3)
    assert after == pytest.approx(round(oracle.psnr(want, ori)[0], 3), abs=1.1e-3)
    # recon file: Y plane + zeroed chroma (inference/yuv_data.cpp:119-125), luma identical to the oracle
    raw = np.frombuffer((tmp_path / "recon.yuv").read_bytes(), np.uint8).reshape(frames, h * w * 3 // 2)
    assert np.array_equal(raw[:, :h * w].reshape(frames, h, w), want)
    assert not raw[:, h * w:].any()
    # appended artefacts of kernel.cu:107-115
    assert "after quantized net:PSNR=" in (tmp_path / "log.txt").read_text()
    (psnr2,) = struct.unpack("<d", (tmp_path / "recon_psnr.data").read_bytes())
    assert psnr2 == oracle.psnr(want, ori)[0]


def test_cli_missing_model_exits_like_the_reference(tmp_path):
    anchor, ori = synth.make_frames(1, 1, 16, 16)
    _write_yuv(tmp_path / "ori.yuv", ori, 0)
    _write_yuv(tmp_path / "a_Q22.yuv", anchor, 0)
    p = subprocess.run([CLI, "ori.yuv", "a_", "16", "16", "--model", "nope_%d.data"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert p.returncode == 1 and "cannot open model file." in p.stdout        # inference/qvrcnn.cu:50-54


@pytest.mark.parametrize("impl", ["fused", "layered"])
def test_stream_yuv_equals_in_memory_path(tmp_path, models, impl):
    """qv_stream_yuv (files in, file out, reader / GPU / writer overlapped) against the in-memory path:
    same reconstruction file bytes as save_recon_as, same exact SSEs, frame ranges writable independently, and
    chunking that does not divide the frame count."""
    from oracle import oracle
    from qcnn_gpu_b200 import api
    qp, h, w, frames = 27, 72, 250, 11
    anchor, ori = synth.make_frames(0xC0FFEE + 12, frames, h, w)
    _write_yuv(tmp_path / "ori.yuv", ori, 0x80)
    _write_yuv(tmp_path / "anchor.yuv", anchor, 0x33)
    net = api.QVRCNN(0, 4, 1, h, w)                       # chunks of 4: 4 + 4 + 3
    net.load_static_para_mem(formats.write_model_vect_c(models[qp]))
    net.set_impl(api.IMPL_FUSED if impl == "fused" else api.IMPL_LAYERED)
    want = net.forward_frames_host(anchor)
    assert np.array_equal(want[0], oracle.OracleModel(formats.write_model_vect_c(models[qp])).forward_blu(anchor[0:1])[0])
    sb, sa = net.stream_yuv(str(tmp_path / "anchor.yuv"), str(tmp_path / "ori.yuv"), str(tmp_path / "recon.yuv"), 0, frames)
    assert sb == int(((anchor.astype(np.int64) - ori) ** 2).sum())
    assert sa == int(((want.astype(np.int64) - ori) ** 2).sum())
    raw = np.frombuffer((tmp_path / "recon.yuv").read_bytes(), np.uint8).reshape(frames, h * w * 3 // 2)
    assert np.array_equal(raw[:, :h * w].reshape(frames, h, w), want)
    assert not raw[:, h * w:].any()
    # a frame range in the middle, written into a second file at its own offset; no original -> no SSE
    assert net.stream_yuv(str(tmp_path / "anchor.yuv"), None, str(tmp_path / "part.yuv"), 5, 3) == (0, 0)
    raw2 = np.frombuffer((tmp_path / "part.yuv").read_bytes(), np.uint8)
    assert raw2.size == 8 * (h * w * 3 // 2)
    assert np.array_equal(raw2.reshape(8, -1)[5:8, :h * w].reshape(3, h, w), want[5:8])
    # asking for more frames than the file holds is an I/O error, not a hang
    with pytest.raises(api.QVError, match="short read"):
        net.stream_yuv(str(tmp_path / "anchor.yuv"), None, None, 9, 5)


def test_cli_stream_mode(tmp_path, models):
    from oracle import oracle
    qp, h, w, frames = 32, 64, 112, 5
    anchor, ori = synth.make_frames(0xC0FFEE + 13, frames, h, w)
    _write_yuv(tmp_path / "ori.yuv", ori, 0x80)
    _write_yuv(tmp_path / ("anchor_Q%d.yuv" % qp), anchor, 0x33)
    (tmp_path / ("model_%d.data" % qp)).write_bytes(formats.write_model_vect_c(models[qp]))
    p = subprocess.run([CLI, "ori.yuv", "anchor_", str(h), str(w), "--model", "model_%d.data", "--qp", str(qp),
                        "--frames", str(frames), "--stream", "--save-recon", "recon.yuv"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    want = oracle.OracleModel(formats.write_model_vect_c(models[qp])).forward_blu(anchor)
    after = float(re.search(r"after quantized net:PSNR=([0-9.]+)", p.stdout).group(1))
    assert after == pytest.approx(round(oracle.psnr(want, ori)[0], 3), abs=1.1e-3)
    raw = np.frombuffer((tmp_path / "recon.yuv").read_bytes(), np.uint8).reshape(frames, h * w * 3 // 2)
    assert np.array_equal(raw[:, :h * w].reshape(frames, h, w), want)
    (psnr2,) = struct.unpack("<d", (tmp_path / "recon_psnr.data").read_bytes())
    assert psnr2 == oracle.psnr(want, ori)[0]


REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "qcnn_ref_driver_on_shim")


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="oracle/_ref/qcnn_ref_driver_on_shim not built (needs /root/reference at build time)")
def test_reference_driver_compiled_against_the_shim(tmp_path, models):
    """The reference's OWN driver text -- testqvrcnn / run_all / main, inference/kernel.cu:74-138, extracted verbatim at
    build time by oracle/ref_witness/Makefile -- compiled against this repository's qvrcnn.cuh / yuv_data.h and linked with
    libqvrcnn_b200.so: same argv, same per-frame sequence (load_data, forward_blu, cudaMemcpy of I1.x_rec), same report
    and appended files, results equal to the oracle's."""
    from oracle import oracle
    qp, h, w = 22, 72, 136                                  # run_all only loops qp = 22 (kernel.cu:122), FRAME = 1
    anchor, ori = synth.make_frames(0xC0FFEE + 14, 1, h, w)
    _write_yuv(tmp_path / "ori.yuv", ori, 0x80)
    _write_yuv(tmp_path / ("anchor_Q%d.yuv" % qp), anchor, 0x33)
    image = formats.write_model_vect_c(models[qp])
    (tmp_path / ("qvrcnn_nchw_vect_c_8bit_qfp_%d.data" % qp)).write_bytes(image)
    p = subprocess.run([REF_DRIVER, "ori.yuv", "anchor_", str(h), str(w)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    want = oracle.OracleModel(image).forward_blu(anchor)
    before = float(re.search(r"before net:PSNR=([0-9.]+)", p.stdout).group(1))
    after = float(re.search(r"after quantized net:PSNR=([0-9.]+)", p.stdout).group(1))
    assert before == pytest.approx(round(oracle.psnr(anchor, ori)[0], 3), abs=1.1e-3)
    assert after == pytest.approx(round(oracle.psnr(want, ori)[0], 3), abs=1.1e-3)
    (psnr2,) = struct.unpack("<d", (tmp_path / "recon_psnr.data").read_bytes())
    assert psnr2 == oracle.psnr(want, ori)[0]                # bit-identical frame => bit-identical double
    assert "after quantized net:PSNR=" in (tmp_path / "log.txt").read_text()
