// Reference witness driver (TEST INFRASTRUCTURE): links the reference's own, unmodified qvrcnn / layer
// / kernel objects (compiled from /root/reference/inference by the Makefile beside this file) and runs
// the exact per-frame sequence of testqvrcnn (inference/kernel.cu:86-97) on files we hand it:
//     qvrcnn net(0, 1, 1, H, W); net.load_static_para(model); per frame: load_data, forward_blu,
//     cudaDeviceSynchronize, cudaMemcpy(recon <- net.I1.x_rec)
// usage: qcnn_ref_witness <model.data> <H> <W> <frames> <in.luma> <out.luma> [reps]
// With reps > 1 the frame loop is repeated and each repetition is timed with the reference's own
// timer scope (inference/kernel.cu:89-101: H2D + forward_blu + sync + D2H for all frames).
// in/out are raw u8 luma, frames*H*W bytes.  Exit code 0 on success; the reference's own check() macro
// prints and exit(1)s on any CUDA / cuDNN failure (inference/cnn.cuh:8-15).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "qvrcnn.cuh"

int main(int argc, char **argv)
{
    if (argc != 7 && argc != 8) { fprintf(stderr, "usage: %s model H W frames in.luma out.luma\n", argv[0]); return 2; }
    const int H = atoi(argv[2]), W = atoi(argv[3]), frames = atoi(argv[4]);
    const size_t hw = (size_t)H * W;
    std::vector<unsigned char> in(hw * frames), out(hw * frames);
    FILE *fp = fopen(argv[5], "rb");
    if (!fp || fread(in.data(), 1, in.size(), fp) != in.size()) { fprintf(stderr, "cannot read %s\n", argv[5]); return 2; }
    fclose(fp);
    qvrcnn net(0, 1, 1, H, W);
    net.load_static_para(argv[1]);
    const int reps = argc == 8 ? atoi(argv[7]) : 1;
    for (int rep = 0; rep < reps; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < frames; ++i) {
            net.load_data(in.data() + i * hw);
            net.forward_blu();
            cudaDeviceSynchronize();
            cudaMemcpy(out.data() + i * hw, (datatype *)net.I1.x_rec, hw, cudaMemcpyDeviceToHost);
        }
        auto t1 = std::chrono::steady_clock::now();
        printf("time_us:%lld\n", (long long)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count());
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 3; }
    fp = fopen(argv[6], "wb");
    fwrite(out.data(), 1, out.size(), fp);
    fclose(fp);
    printf("witness ok: %d frame(s) %dx%d\n", frames, W, H);
    return 0;
}
