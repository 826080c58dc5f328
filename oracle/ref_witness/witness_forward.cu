// Reference witness for SURVEY 8 f3 (TEST INFRASTRUCTURE): is qvrcnn::forward() -- the dynamic-step operator,
// inference/qvrcnn.cu:82-167 -- a defined function of (model file, frame) when the model comes from load_static_para,
// the only loader whose file format ships?  forward() reads CovLayer::step_w (set only by load_para, inference/cnn.cu:78)
// and C1.step_y (set only by quantize_out / quantize_out_fix, cnn.cu:175,185; forward() calls quantize_out_static, cnn.cu:189-194,
// which sets neither) and feeds them through insert_w / insert_y (qvrcnn.cu:104-105) into adjustBasic (qvrcnn.cu:336-349), which
// rescales the biases of every later layer by them.  Neither the constructors (cnn.cu:3-12, qvrcnn.cu:4-29) nor
// load_static_para (cnn.cu:90-112) initialise those members.
// The witness builds the reference's UNMODIFIED qvrcnn object twice, by placement new into storage pre-filled with two
// different byte patterns, runs load_static_para + load_data + forward() on the same model and frame, and compares.
// usage: qcnn_ref_forward <model.data> <H> <W> <in.luma>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "qvrcnn.cuh"

static unsigned long long fnv(const std::vector<unsigned char> &v)
{
    unsigned long long h = 1469598103934665603ull;
    for (unsigned char c : v) { h ^= c; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv)
{
    if (argc != 5) { fprintf(stderr, "usage: %s model H W in.luma\n", argv[0]); return 2; }
    const int H = atoi(argv[2]), W = atoi(argv[3]);
    const size_t hw = (size_t)H * W;
    std::vector<unsigned char> in(hw);
    FILE *fp = fopen(argv[4], "rb");
    if (!fp || fread(in.data(), 1, hw, fp) != hw) { fprintf(stderr, "cannot read %s\n", argv[4]); return 2; }
    fclose(fp);
    const unsigned char fills[3] = {0x00, 0x5A, 0x01};
    std::vector<unsigned char> out[3], blu[3];
    for (int t = 0; t < 3; ++t) {
        void *mem = aligned_alloc(64, (sizeof(qvrcnn) + 63) / 64 * 64);
        memset(mem, fills[t], sizeof(qvrcnn));
        qvrcnn *net = new (mem) qvrcnn(0, 1, 1, H, W);
        net->load_static_para(argv[1]);
        printf("fill 0x%02X: before forward(): C1.step_w=%d C1.step_y=%d C2_1.step_w=%d C3_1.step_w=%d C4.step_w=%d steps.stepw[0]=%d steps.stepy[0]=%d\n",
               fills[t], net->C1.step_w, net->C1.step_y, net->C2_1.step_w, net->C3_1.step_w, net->C4.step_w, net->steps.stepw[0], net->steps.stepy[0]);
        out[t].resize(hw); blu[t].resize(hw);
        net->load_data(in.data());
        net->forward();
        cudaDeviceSynchronize();
        cudaMemcpy(out[t].data(), (datatype *)net->I1.x_rec, hw, cudaMemcpyDeviceToHost);
        printf("fill 0x%02X: forward()     x_rec fnv1a=%016llx  steps.stepw={%d,%d,%d,%d} steps.stepy={%d,%d,%d,%d}\n", fills[t], fnv(out[t]),
               net->steps.stepw[0], net->steps.stepw[1], net->steps.stepw[2], net->steps.stepw[3],
               net->steps.stepy[0], net->steps.stepy[1], net->steps.stepy[2], net->steps.stepy[3]);
        net->load_data(in.data());
        net->forward_blu();
        cudaDeviceSynchronize();
        cudaMemcpy(blu[t].data(), (datatype *)net->I1.x_rec, hw, cudaMemcpyDeviceToHost);
        printf("fill 0x%02X: forward_blu() x_rec fnv1a=%016llx\n", fills[t], fnv(blu[t]));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) printf("fill 0x%02X: CUDA error after the run: %s\n", fills[t], cudaGetErrorString(e));
        net->~qvrcnn();
        free(mem);
    }
    size_t d01 = 0, d02 = 0, b01 = 0;
    for (size_t i = 0; i < hw; ++i) { d01 += out[0][i] != out[1][i]; d02 += out[0][i] != out[2][i]; b01 += blu[0][i] != blu[1][i] || blu[0][i] != blu[2][i]; }
    printf("forward():     pixels differing between fill 0x00 and 0x5A: %zu of %zu ; between 0x00 and 0x01: %zu\n", d01, hw, d02);
    printf("forward_blu(): pixels differing between the fills: %zu of %zu\n", b01, hw);
    printf(d01 || d02 ? "VERDICT: forward() after load_static_para depends on uninitialised members -- not a defined operator\n"
                      : "VERDICT: forward() gave the same frame for all fills\n");
    return 0;
}
