"""Multi-GPU driver for the QVRCNN pass: one process per GPU (torchrun), NCCL for the two exchanges
the path has -- neighbour input-halo exchange in strip mode and the int64 SSE all-reduce of the
PSNR report.  Compute goes through the C ABI (qv_forward_frames_device / qv_forward_rows_device).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m qcnn_gpu_b200.host.multi_gpu \\
      --mode frames --qp 27 --frames 240 --height 2160 --width 3840        # BASELINE config 4
  torchrun ... -m qcnn_gpu_b200.host.multi_gpu --mode strips --qp 22 --height 4320 --width 7680   # config 5

--check additionally computes the whole job on rank 0 alone and asserts the N-GPU result is bit-identical.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from qcnn_gpu_b200 import api
    from qcnn_gpu_b200.host import formats, shard, synth

    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["frames", "strips"], default="frames")
    ap.add_argument("--qp", type=int, default=27)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--uniq", type=int, default=4, help="distinct synthetic frames per rank (repeated)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    from qcnn_gpu_b200.host import numa
    numa.bind_to_gpu(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h, w = args.height, args.width
    model = synth.make_model(0xC0FFEE + args.qp, args.qp)
    image = formats.write_model_vect_c(model)
    stream = torch.cuda.current_stream()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.mode == "frames":
        f0, nf = shard.split(args.frames, rank, world)
        uniq = max(1, min(args.uniq, nf))
        a, o = synth.make_frames(0xC0FFEE + 4, uniq, h, w, first_frame=f0 % 1024)
        reps = (nf + uniq - 1) // uniq
        d_in = torch.from_numpy(np.tile(a, (reps, 1, 1))[:nf]).cuda()
        d_ori = torch.from_numpy(np.tile(o, (reps, 1, 1))[:nf]).cuda()
        d_out = torch.empty_like(d_in)
        net = api.QVRCNN(local, max(1, min(nf, 8)), 1, h, w)
        net.load_static_para_mem(image)
        run = lambda: net.forward_frames_device(d_in.data_ptr(), d_out.data_ptr(), nf, stream.cuda_stream)
        npx_local = nf * h * w
    else:
        y0, y1, r0, r1 = shard.strip_window(h, rank, world)
        a, o = synth.make_frames(0xC0FFEE + 5, 1, h, w)          # every rank could read its rows from a file instead
        own = torch.from_numpy(a[0, y0:y1].copy()).cuda()          # only this rank's rows live on this GPU
        d_ori = torch.from_numpy(o[0, y0:y1].copy()).cuda()
        d_out = torch.empty_like(own)
        net = api.QVRCNN(local, 1, 1, r1 - r0, w)
        net.load_static_para_mem(image)
        win = [None]

        def run():
            # one neighbour exchange of 6 input rows each way (NVLink P2P via NCCL send/recv), then compute
            win[0] = shard.exchange_halos(own, y0, y1, h, rank, world, dist,
                                          lambda s: torch.empty(s, dtype=torch.uint8, device="cuda")) if world > 1 else own
            net.forward_rows_device(win[0].data_ptr(), h, r0, r1 - r0, d_out.data_ptr(), y0, y1, stream.cuda_stream)
        npx_local = (y1 - y0) * w

    for _ in range(2):
        run()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        run()
    e1.record(stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    api.sse_device(d_out.data_ptr(), d_ori.data_ptr(), npx_local, acc.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    psnr, sse, n = shard.psnr_from_global_sse(int(acc.item()), npx_local, dist if world > 1 else None, device="cuda")
    result = {"mode": args.mode, "qp": args.qp, "n_gpus": world, "height": h, "width": w,
              "frames": args.frames if args.mode == "frames" else 1,
              "Mpixel_per_s": n * args.steps / (float(ms.item()) * 1e-3) / 1e6, "ms_per_step": float(ms.item()) / args.steps,
              "after_quantized_net_PSNR": psnr, "sse": sse}
    if args.check:
        # gather every rank's output on rank 0 and compare with rank 0 computing everything alone
        outs = [None] * world
        if world > 1:
            dist.all_gather_object(outs, d_out.cpu().numpy())
        else:
            outs = [d_out.cpu().numpy()]
        if rank == 0:
            got = np.concatenate(outs)
            if args.mode == "frames":
                full = np.concatenate([np.tile(synth.make_frames(0xC0FFEE + 4, max(1, min(args.uniq, shard.split(args.frames, r, world)[1])), h, w,
                                                first_frame=shard.split(args.frames, r, world)[0] % 1024)[0],
                                               ((shard.split(args.frames, r, world)[1] + args.uniq - 1) // max(1, args.uniq) + 1, 1, 1))[:shard.split(args.frames, r, world)[1]]
                                       for r in range(world)])
                solo = api.QVRCNN(local, 4, 1, h, w)
            else:
                full = a
                got = got[None]
                solo = api.QVRCNN(local, 1, 1, h, w)
            solo.load_static_para_mem(image)
            want = solo.forward_frames_host(full)
            result["bit_identical_to_1gpu"] = bool(np.array_equal(got.reshape(want.shape), want))
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
