"""Oracle pinning against the reference itself.  tests/golden/ref_witness_*.npz hold reconstructed luma
produced by the UNMODIFIED reference sources (inference/{mat,cnn,qvrcnn}.cu + cuDNN 9.10, compiled by
oracle/ref_witness/Makefile) run on a B200 by oracle/ref_witness/run_witness.py; inputs are rebuilt
from the seeds stored in each fixture.  The CPU oracle must reproduce them bit for bit (CPU test), and
so must both CUDA implementations (GPU test)."""
import glob
import os

import numpy as np
import pytest

from qcnn_gpu_b200.host import formats, synth

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_witness_*.npz")))


def _case(path):
    z = np.load(path)
    qp, frames, h, w = int(z["qp"]), int(z["frames"]), int(z["h"]), int(z["w"])
    model = synth.make_model(int(z["model_seed"]), qp)
    anchor, _ = synth.make_frames(int(z["frame_seed"]), frames, h, w)
    return model, anchor, z["recon"]


def test_fixtures_present():
    assert len(FIXTURES) >= 5


@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_oracle_reproduces_reference_output(path):
    from oracle import oracle
    model, anchor, recon = _case(path)
    got = oracle.OracleModel(formats.write_model_vect_c(model)).forward_blu(anchor)
    assert np.array_equal(got, recon)
    assert (recon != anchor).any()          # the fixture is not the identity


@pytest.mark.gpu
@pytest.mark.parametrize("impl", [1, 2], ids=["layered", "fused"])
@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_cuda_paths_reproduce_reference_output(path, impl):
    from qcnn_gpu_b200 import api
    model, anchor, recon = _case(path)
    n, h, w = anchor.shape
    net = api.QVRCNN(0, n, 1, h, w)
    net.load_static_para_mem(formats.write_model_vect_c(model))
    net.set_impl(impl)
    net.load_data(anchor)
    net.forward_blu()
    assert np.array_equal(net.get_recon(), recon)
