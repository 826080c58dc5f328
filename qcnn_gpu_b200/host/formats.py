"""On-disk formats either side of the QVRCNN hot path (Python tooling mirror).

The product's own readers are C++ (qcnn_gpu_b200/csrc/qv_formats.cpp, reached through the
C ABI); this module exists so tests and bench.py can *write* the files those readers
consume and cross-check what they parsed.

Formats (paths relative to /root/reference/):
  * static model file, NCHW_VECT_C flavour -- read by CovLayer::load_static_para
    (inference/cnn.cu:90-112), record order inference/qvrcnn.cu:55-60
  * static model file, HWCN flavour -- input of layer_qfp_HWCN2NCHW_VECT_C
    (inference/qvrcnn.cu:535-557)
  * quant_params<QP>.data pickle / quant_params_cpp_<QP>.data raw doubles
    (training/quantization.py:90-96)
  * YUV 4:2:0 8-bit planar luma I/O (inference/yuv_data.cpp:15-42,113-128)
"""
from __future__ import annotations

import pickle
import struct
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

# (Cin, Cout, ksize) for C1, C2_1, C2_2, C3_1, C3_2, C4 -- inference/qvrcnn.cu:11-18
LAYERS = ((1, 64, 5), (64, 32, 3), (64, 16, 5), (48, 16, 3), (48, 32, 1), (48, 1, 3))
LAYER_NAMES = ("C1", "C2_1", "C2_2", "C3_1", "C3_2", "C4")

# blu_q / mul / shift per layer as shipped in training/quant_params{22,27,32,37}.data
# (decoded from the pickles; tests/test_formats.py re-checks them against the shipped files
# whenever /root/reference is present).
SHIPPED_QPARAMS = {
    22: ((3849, 271, 13), (4526, 115, 12), (5312, 49, 11), (18635, 7, 10), (4923, 53, 11), (0, 5, 24)),
    27: ((8390, 31, 11), (6600, 79, 12), (13622, 153, 14), (6983, 299, 14), (6020, 347, 14), (0, 1, 12)),
    32: ((10354, 6431, 19), (10629, 49, 12), (10523, 99, 13), (7426, 281, 14), (4526, 115, 12), (0, 13, 15)),
    37: ((11512, 723, 16), (10182, 205, 14), (11030, 189, 14), (13154, 635, 16), (7580, 551, 15), (0, 7, 13)),
}


@dataclass
class Model:
    """Six layers of plain [K][C][R][S] int8 weights + int32 biases + (blu, mul, shift)."""

    w: List[np.ndarray] = field(default_factory=list)      # int8  [K,C,R,S]
    b: List[np.ndarray] = field(default_factory=list)      # int32 [K]
    qparams: List[Sequence[int]] = field(default_factory=list)  # (blu, mul, shift)

    def check(self) -> None:
        assert len(self.w) == len(self.b) == len(self.qparams) == 6
        for (cin, cout, k), w, b in zip(LAYERS, self.w, self.b):
            assert w.dtype == np.int8 and w.shape == (cout, cin, k, k), (w.dtype, w.shape)
            assert b.dtype == np.int32 and b.shape == (cout,)


def vect_c_wsize(cin: int, cout: int, k: int) -> int:
    """wSize of inference/cnn.cu:24."""
    return k * k * ((cin + 3) // 4) * 4 * cout


MODEL_FILE_SIZE = sum(vect_c_wsize(*l) + 4 * l[1] + 12 for l in LAYERS)          # 60 028
MODEL_FILE_SIZE_HWCN = sum(l[0] * l[1] * l[2] * l[2] + 4 * l[1] + 12 for l in LAYERS)  # 55 228


def pack_weights_vect_c(w: np.ndarray) -> bytes:
    """[K,C,R,S] -> int8 w[K][ceil(C/4)][R][S][4] with zero-padded lanes (inference/mat.cu:108-117)."""
    K, C, R, S = w.shape
    c4 = (C + 3) // 4
    out = np.zeros((K, c4 * 4, R, S), np.int8)
    out[:, :C] = w
    out = out.reshape(K, c4, 4, R, S).transpose(0, 1, 3, 4, 2)
    return np.ascontiguousarray(out).tobytes()


def unpack_weights_vect_c(buf: bytes, cin: int, cout: int, k: int) -> np.ndarray:
    c4 = (cin + 3) // 4
    a = np.frombuffer(buf, np.int8).reshape(cout, c4, k, k, 4).transpose(0, 1, 4, 2, 3)
    return np.ascontiguousarray(a.reshape(cout, c4 * 4, k, k)[:, :cin])


def write_model_vect_c(model: Model) -> bytes:
    """Serialise in the layout CovLayer::load_static_para reads (inference/cnn.cu:99-103)."""
    model.check()
    out = bytearray()
    for w, b, q in zip(model.w, model.b, model.qparams):
        out += pack_weights_vect_c(w)
        out += b.astype("<i4").tobytes()
        out += struct.pack("<3i", *[int(v) for v in q])
    assert len(out) == MODEL_FILE_SIZE
    return bytes(out)


def read_model_vect_c(buf: bytes) -> Model:
    assert len(buf) == MODEL_FILE_SIZE, len(buf)
    m, off = Model(), 0
    for cin, cout, k in LAYERS:
        n = vect_c_wsize(cin, cout, k)
        m.w.append(unpack_weights_vect_c(buf[off:off + n], cin, cout, k)); off += n
        m.b.append(np.frombuffer(buf[off:off + 4 * cout], "<i4").astype(np.int32)); off += 4 * cout
        m.qparams.append(struct.unpack("<3i", buf[off:off + 12])); off += 12
    return m


def write_model_hwcn(model: Model) -> bytes:
    """TF weight order [R][S][C][K] per layer (inference/qvrcnn.cu:542-555)."""
    model.check()
    out = bytearray()
    for w, b, q in zip(model.w, model.b, model.qparams):
        out += np.ascontiguousarray(w.transpose(2, 3, 1, 0)).tobytes()
        out += b.astype("<i4").tobytes()
        out += struct.pack("<3i", *[int(v) for v in q])
    assert len(out) == MODEL_FILE_SIZE_HWCN
    return bytes(out)


def read_model_hwcn(buf: bytes) -> Model:
    assert len(buf) == MODEL_FILE_SIZE_HWCN, len(buf)
    m, off = Model(), 0
    for cin, cout, k in LAYERS:
        n = cin * cout * k * k
        a = np.frombuffer(buf[off:off + n], np.int8).reshape(k, k, cin, cout); off += n
        m.w.append(np.ascontiguousarray(a.transpose(3, 2, 0, 1)))
        m.b.append(np.frombuffer(buf[off:off + 4 * cout], "<i4").astype(np.int32)); off += 4 * cout
        m.qparams.append(struct.unpack("<3i", buf[off:off + 12])); off += 12
    return m


def load_quant_params_pickle(path: str) -> List[List[float]]:
    """6 rows [stepw, ratio, blu_adj, blu_q, mul, shift] (training/quantization.py:90-91)."""
    with open(path, "rb") as fp:
        rows = pickle.load(fp)
    return [[float(v) for v in r] for r in rows]


def write_quant_params_pickle(path: str, rows) -> None:
    """Same object shape the reference pickles: list of lists mixing numpy f8 scalars and
    python ints, protocol 3 (what quantNsave produced under Python 3.5-3.7)."""
    obj = []
    for r in rows:
        stepw, ratio, blu_adj, blu_q, mul, shift = r
        obj.append([np.float64(stepw), ratio if isinstance(ratio, int) else np.float64(ratio),
                    blu_adj if isinstance(blu_adj, int) else np.float64(blu_adj),
                    blu_q if isinstance(blu_q, int) else np.float64(blu_q),
                    np.float64(mul), int(shift)])
    with open(path, "wb") as fp:
        pickle.dump(obj, fp, protocol=3)


def write_quant_params_cpp(path: str, rows) -> None:
    """quant_params_cpp_<QP>.data: six rows of struct.pack('6d') (training/quantization.py:93-96)."""
    with open(path, "wb") as fp:
        for r in rows:
            fp.write(struct.pack("<6d", *[float(v) for v in r]))


def qparams_rows_from_table(qp: int):
    """Rows in the pickle's shape carrying the shipped integer triples (stepw/ratio/blu_adj are
    not consumed by inference -- inference/cnn.cu:101-103 -- so they are placeholders here)."""
    rows = []
    for i, (blu, mul, shift) in enumerate(SHIPPED_QPARAMS[qp]):
        last = i == 5
        rows.append([0.01, 255 if i == 0 else 255.0, 0 if last else 0.1,
                     0 if last else float(blu), float(mul), int(shift)])
    return rows


# ---- YUV 4:2:0 8-bit planar -------------------------------------------------------------

def write_yuv420_luma(path: str, luma: np.ndarray) -> None:
    """frames x H x W u8 luma -> Y plane + H*W/2 zero bytes per frame
    (the byte layout of vrcnn_data::save_recon_as, inference/yuv_data.cpp:119-125)."""
    assert luma.dtype == np.uint8 and luma.ndim == 3
    f, h, w = luma.shape
    uv = bytes(h * w // 2)
    with open(path, "wb") as fp:
        for i in range(f):
            fp.write(luma[i].tobytes())
            fp.write(uv)


def read_yuv420_luma(path: str, frames: int, h: int, w: int) -> np.ndarray:
    """Luma of the first `frames` frames (vrcnn_data::read_data, inference/yuv_data.cpp:32-38)."""
    out = np.empty((frames, h, w), np.uint8)
    with open(path, "rb") as fp:
        for i in range(frames):
            out[i] = np.frombuffer(fp.read(h * w), np.uint8).reshape(h, w)
            fp.seek(h * w // 2, 1)
    return out


def psnr(data: np.ndarray, ori: np.ndarray):
    """vrcnn_data::psnr (inference/yuv_data.cpp:87-97). Returns (psnr, sse)."""
    d = data.astype(np.int64) - ori.astype(np.int64)
    sse = int((d * d).sum())
    mse = float(sse) / data.size
    return (10.0 * np.log10(65025.0 / mse) if mse > 0 else float("inf")), sse
