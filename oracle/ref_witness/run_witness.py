"""Runs the reference witness (oracle/_ref/qcnn_ref_witness: the UNMODIFIED reference sources + cuDNN) on
a B200 with synthetic models / frames, compares its reconstructed luma with the CPU oracle and writes the
golden fixtures tests/golden/ref_witness_qp<QP>.npz.  TEST INFRASTRUCTURE; run on the GPU box:
    python oracle/ref_witness/run_witness.py gpurun_out/witness
The fixtures pin the oracle against outputs of the reference itself (the reference ships no
golden vectors).  Inputs are NOT stored: they are rebuilt from the seeds recorded in each fixture."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle                                  # noqa: E402
from qcnn_gpu_b200.host import formats, synth             # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "qcnn_ref_witness")
CASES = [(37, 2, 240, 416), (22, 1, 120, 208), (27, 1, 120, 208), (32, 1, 120, 208), (32, 1, 33, 61)]


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    ok_all = True
    for qp, frames, h, w in CASES:
        model = synth.make_model(0xC0FFEE + qp, qp)
        seed = 0xC0FFEE + 7
        anchor, _ = synth.make_frames(seed, frames, h, w)
        with tempfile.TemporaryDirectory() as td:
            mf, fi, fo = os.path.join(td, "m.data"), os.path.join(td, "in.luma"), os.path.join(td, "out.luma")
            open(mf, "wb").write(formats.write_model_vect_c(model))
            anchor.tofile(fi)
            p = subprocess.run([BIN, mf, str(h), str(w), str(frames), fi, fo], capture_output=True, text=True, timeout=300)
            print("qp%d %dx%dx%d: rc=%d %s %s" % (qp, frames, w, h, p.returncode, p.stdout.strip()[-200:], p.stderr.strip()[-300:]))
            if p.returncode != 0 or not os.path.exists(fo):
                ok_all = False
                continue
            rec = np.fromfile(fo, np.uint8).reshape(frames, h, w)
        want = oracle.OracleModel(formats.write_model_vect_c(model)).forward_blu(anchor)
        same = bool(np.array_equal(rec, want))
        print("   reference(cuDNN on this GPU) == CPU oracle: %s (%d differing pixels)" % (same, int((rec != want).sum())))
        ok_all &= same
        np.savez_compressed(os.path.join(outdir, "ref_witness_qp%d_%dx%dx%d.npz" % (qp, frames, h, w)), recon=rec,
                            qp=qp, frames=frames, h=h, w=w, frame_seed=seed, model_seed=0xC0FFEE + qp)
    print("WITNESS", "ALL EQUAL" if ok_all else "MISMATCH OR FAILURE")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/witness"))
