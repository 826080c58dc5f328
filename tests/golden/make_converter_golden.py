"""Golden vectors for the HWCN -> NCHW_VECT_C model converter, produced by the REFERENCE'S OWN code:
oracle/_ref/qcnn_ref_convert is the unmodified model_qfp_HWCN2NCHW_VECT_C (inference/qvrcnn.cu:558-585,
inference/mat.cu:97-119) compiled by oracle/ref_witness/Makefile (host-only code, runs without a GPU).

    python tests/golden/make_converter_golden.py      # in the build container, where /root/reference is mounted

Writes tests/golden/ref_converter_qp{27,32}.npz: the HWCN file that went in (synthetic int8 model, seeded) and the
NCHW_VECT_C file the reference wrote.  tests/test_formats.py replays them against qv_convert_model_hwcn_to_vect_c and
the HWCN loader."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qcnn_gpu_b200.host import formats, synth  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "qcnn_ref_convert")


def main():
    for qp in (27, 32):
        m = synth.make_model(0xC0FFEE + qp, qp)
        hwcn = formats.write_model_hwcn(m)
        with tempfile.TemporaryDirectory() as td:
            open(os.path.join(td, "hwcn_%d.data" % qp), "wb").write(hwcn)
            subprocess.check_call([BIN, os.path.join(td, "hwcn_%d.data"), os.path.join(td, "vect_c_%d.data"), str(qp)])
            vect_c = open(os.path.join(td, "vect_c_%d.data" % qp), "rb").read()
        out = os.path.join(ROOT, "tests", "golden", "ref_converter_qp%d.npz" % qp)
        np.savez_compressed(out, hwcn=np.frombuffer(hwcn, np.uint8), vect_c=np.frombuffer(vect_c, np.uint8), qp=qp,
                            produced_by="oracle/_ref/qcnn_ref_convert = reference model_qfp_HWCN2NCHW_VECT_C, unmodified")
        print(out, len(hwcn), "->", len(vect_c), "bytes; equals our python writer:", vect_c == formats.write_model_vect_c(m))


if __name__ == "__main__":
    main()
