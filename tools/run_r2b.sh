# round 2, second GPU call (1 GPU): widened GPU suite, f3 witness, kernel A/B (round-1 kernel, HEAD, load_in fast path, wait hints), PROF counters
T=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
( cd /tmp && python - <<'PY'
import os, sys, subprocess
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
root = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
from qcnn_gpu_b200.host import formats, synth
for qp, h, w in ((32, 120, 208), (22, 64, 96)):
    open("/tmp/m.data", "wb").write(formats.write_model_vect_c(synth.make_model(0xC0FFEE + qp, qp)))
    synth.make_frames(0xC0FFEE + 50, 1, h, w)[0].tofile("/tmp/in.luma")
    for run in range(2):
        p = subprocess.run([os.path.join(root, "oracle/_ref/qcnn_ref_forward"), "/tmp/m.data", str(h), str(w), "/tmp/in.luma"], capture_output=True, text=True, timeout=300)
        print("== qp %d %dx%d run %d rc=%d" % (qp, w, h, run, p.returncode)); print(p.stdout[-3000:]); print(p.stderr[-500:])
PY
) > gpurun_out/${T}_f3_witness.log 2>&1
timeout 600 python tools/kernel_ab.py tools/bin/lib_r1.so tools/bin/lib_head.so tools/bin/lib_x1.so tools/bin/lib_h500.so tools/bin/lib_h100.so --steps 30 --reps 3 > gpurun_out/${T}_kernel_ab.log 2>&1
QV_FUSED_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sustained --no-extra-configs 2>&1 | grep -E "fused (profile|trace|stamps|mma-side)" | head -40 > gpurun_out/${T}_prof.log
tail -3 gpurun_out/${T}_tests.log; tail -12 gpurun_out/${T}_f3_witness.log; cat gpurun_out/${T}_kernel_ab.log | cut -c1-200
