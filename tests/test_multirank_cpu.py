"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups.  The per-rank compute is the CPU
oracle, so what is tested is exactly what the multi-GPU driver adds: frame sharding, strip windows,
neighbour halo exchange and the int64 SSE all-reduce feeding the PSNR report."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qcnn_gpu_b200.host import formats, shard, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    oracle.set_num_threads(2)
    qp = 27
    model = synth.make_model(0xC0FFEE + qp, qp)
    om = oracle.OracleModel(formats.write_model_vect_c(model))
    if mode == "frames":
        frames, h, w = 5, 24, 40
        anchor, ori = synth.make_frames(21, frames, h, w)
        f0, nf = shard.split(frames, rank, world)
        rec = om.forward_blu(anchor[f0:f0 + nf]) if nf else np.zeros((0, h, w), np.uint8)
        _, sse = formats.psnr(rec, ori[f0:f0 + nf]) if nf else (0.0, 0)
        psnr, gsse, n = shard.psnr_from_global_sse(sse, rec.size, dist)
        np.savez(os.path.join(outdir, "r%d.npz" % rank), rec=rec, f0=f0, psnr=psnr, sse=gsse, n=n)
    else:
        h, w = 47, 36
        anchor, ori = synth.make_frames(22, 1, h, w)
        y0, y1, r0, r1 = shard.strip_window(h, rank, world)
        own = torch.from_numpy(anchor[0, y0:y1].copy())          # each rank holds ONLY its own rows
        win = shard.exchange_halos(own, y0, y1, h, rank, world, dist, lambda s: torch.zeros(s, dtype=torch.uint8))
        assert np.array_equal(win.numpy(), anchor[0, r0:r1])       # halos arrived intact
        rec = om.forward_blu(win.numpy()[None])[0][y0 - r0:y0 - r0 + (y1 - y0)]
        _, sse = formats.psnr(rec, ori[0, y0:y1])
        psnr, gsse, n = shard.psnr_from_global_sse(sse, rec.size, dist)
        np.savez(os.path.join(outdir, "r%d.npz" % rank), rec=rec, f0=y0, psnr=psnr, sse=gsse, n=n)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, mode, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    return [np.load(os.path.join(tmp_path, "r%d.npz" % r)) for r in range(world)]


def test_strips_shorter_than_the_halo_are_refused():
    """A strip must hold at least the 6 halo rows its neighbours read: otherwise the halo would reach into the strip
    after next, which no exchange carries (silent size mismatch / hang before)."""
    assert shard.strip_window(47, 1, 7) == (7, 14, 1, 20)          # 47 // 7 = 6 rows: just enough
    with pytest.raises(ValueError, match="shorter than the 6-row halo"):
        shard.strip_window(47, 0, 8)                               # 47 // 8 = 5 rows
    with pytest.raises(ValueError, match="shorter than the 6-row halo"):
        shard.strip_window(5, 0, 2)
    assert shard.strip_window(5, 0, 1) == (0, 5, 0, 5)             # a single rank has no neighbours


def test_split_covers_everything():
    for n in (0, 1, 5, 64, 240):
        for world in (1, 2, 3, 8):
            got = [shard.split(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (a, ca), (b, _) in zip(got, got[1:]):
                assert a + ca == b
    assert shard.strip_window(4320, 3, 8) == (1620, 2160, 1614, 2166)
    assert shard.strip_window(4320, 0, 8) == (0, 540, 0, 546)
    assert shard.strip_window(4320, 7, 8) == (3780, 4320, 3774, 4320)


@pytest.mark.parametrize("world", [2, 3])
def test_frame_sharded_equals_single_process(world, tmp_path):
    from oracle import oracle
    res = _run(world, "frames", tmp_path)
    qp = 27
    om = oracle.OracleModel(formats.write_model_vect_c(synth.make_model(0xC0FFEE + qp, qp)))
    anchor, ori = synth.make_frames(21, 5, 24, 40)
    want = om.forward_blu(anchor)
    got = np.concatenate([r["rec"] for r in res])
    assert np.array_equal(got, want)
    p, sse = formats.psnr(want, ori)
    for r in res:                                  # every rank holds the same, exact, global report
        assert int(r["sse"]) == sse and int(r["n"]) == want.size and float(r["psnr"]) == p


@pytest.mark.parametrize("world", [2, 3])
def test_strip_partition_with_halo_exchange_equals_whole_frame(world, tmp_path):
    from oracle import oracle
    res = _run(world, "strips", tmp_path)
    qp = 27
    om = oracle.OracleModel(formats.write_model_vect_c(synth.make_model(0xC0FFEE + qp, qp)))
    anchor, ori = synth.make_frames(22, 1, 47, 36)
    want = om.forward_blu(anchor)[0]
    got = np.concatenate([r["rec"] for r in res])
    assert np.array_equal(got, want)
    p, sse = formats.psnr(want, ori[0])
    for r in res:
        assert int(r["sse"]) == sse and float(r["psnr"]) == p
