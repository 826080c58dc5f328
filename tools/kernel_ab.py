#!/usr/bin/env python
"""Same-box A/B of fused-kernel builds: python tools/kernel_ab.py lib_a.so lib_b.so ... [--steps 40] [--reps 2] [--sustained 0]
Raw ctypes on the handful of entry points every build of the library has had (so a round-1 library can be compared with
today's); config 3 (QP 32, 64 x 1920x1080, device resident), CUDA events on the launching stream, variants interleaved
rep by rep.  Prints min / median / mean step time per variant; --sustained N adds N back-to-back steps (power-capped)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qcnn_gpu_b200.host import formats, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--sustained", type=int, default=0)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    args = ap.parse_args()
    qp, F, H, W = 32, args.frames, args.height, args.width
    image = formats.write_model_vect_c(synth.make_model(0xC0FFEE + qp, qp))
    a, _ = synth.make_frames(0xC0FFEE + 3, min(8, F), H, W)
    d_in = torch.from_numpy(np.tile(a, ((F + 7) // 8, 1, 1))[:F]).cuda()
    outs = {}
    nets = {}
    for p in args.libs:
        # "lib.so@KEY=VAL": the same library with an environment switch that is read when the model is uploaded
        path, _, env = p.partition("@")
        if env:
            os.environ[env.split("=")[0]] = env.split("=")[1]
        L = C.CDLL(os.path.abspath(path))
        L.qv_last_error.restype = C.c_char_p
        L.qv_create.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_void_p)]
        L.qv_load_static_para_mem.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.qv_forward_frames_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        h = C.c_void_p()
        assert L.qv_create(0, 8, 1, H, W, C.byref(h)) == 0, L.qv_last_error()
        assert L.qv_load_static_para_mem(h, image, len(image)) == 0, L.qv_last_error()
        if env:
            del os.environ[env.split("=")[0]]
        nets[p] = (L, h)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()

    def run(p, n, d_out):
        L, h = nets[p]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record(st)
        for i in range(n):
            rc = L.qv_forward_frames_device(h, d_in.data_ptr(), d_out.data_ptr(), F, st.cuda_stream)
            assert rc == 0, L.qv_last_error()
            evs[i + 1].record(st)
        st.synchronize()
        return [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]

    ref = None
    for rep in range(args.reps):
        for p in args.libs:
            d_out = torch.empty_like(d_in)
            run(p, 3, d_out)
            torch.cuda.synchronize()
            import time
            time.sleep(1.0)                      # let the board cool off the previous variant's burst
            ms = run(p, args.steps, d_out)
            if ref is None:
                ref = d_out.clone()
            same = bool(torch.equal(ref, d_out))
            line = "%-40s rep %d: min %.3f  median %.3f  mean %.3f ms  (first 5: %s)  same_output=%s" % (
                os.path.basename(p)[:40], rep, min(ms), float(np.median(ms)), float(np.mean(ms)), " ".join("%.2f" % x for x in ms[:5]), same)
            if args.sustained:
                ms2 = run(p, args.sustained, d_out)
                line += "  | sustained %d steps: mean %.3f, last quarter %.3f ms" % (args.sustained, float(np.mean(ms2)), float(np.mean(ms2[-len(ms2) // 4:])))
            print(line, flush=True)


if __name__ == "__main__":
    main()
