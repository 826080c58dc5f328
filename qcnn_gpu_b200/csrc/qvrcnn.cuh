// Header-only C++ shim: class qvrcnn with the reference's name, constructor signature and hot-path
// methods (inference/qvrcnn.cuh:25-59), implemented over the C ABI of libqvrcnn_b200.so, so that a
// driver written against the reference (inference/kernel.cu:74-116) compiles against this header.
// Error behaviour is the reference's: print and exit(1) (inference/cnn.cuh:8-15, qvrcnn.cu:50-54).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "../../include/qvrcnn_b200.h"
#include "yuv_data.h"

class qvrcnn {
public:
    qvrcnn(int gpu_num, int batch, int channel, int height, int width)   // inference/qvrcnn.cu:4-29
        : batch(batch), channel(channel), height(height), width(width)
    {
        check(qv_create(gpu_num, batch, channel, height, width, &net));
        void *x = nullptr, *rec = nullptr;
        check(qv_device_buffers(net, &x, &rec));
        I1.x = x; I1.x_rec = rec;
        I1.batch = batch; I1.height = height; I1.width = width; I1.inChannel = channel;
    }
    int load_static_para(const char *filename)                           // inference/qvrcnn.cu:47-63
    {
        int rc = qv_load_static_para(net, filename);
        if (rc == QV_ERR_IO) { printf("cannot open model file.\n"); exit(1); }
        check(rc);
        return 0;
    }
    int load_quant_params(const char *filename) { check(qv_load_quant_params(net, filename)); return 0; }
    int load_data(datatype *input) { check(qv_load_data(net, input)); return 0; }    // qvrcnn.cu:64-68
    int forward_blu(void) { check(qv_forward_blu(net)); return 0; }                    // qvrcnn.cu:168-242
    int get_recon(datatype *out) { check(qv_get_recon(net, out)); return 0; }          // kernel.cu:96
    int forward_frames(const datatype *in, datatype *out, int n) { check(qv_forward_frames_host(net, in, out, n)); return 0; }
    ~qvrcnn() { qv_destroy(net); }                                                    // qvrcnn.cu:331-335
    qvrcnn(const qvrcnn &) = delete;
    qvrcnn &operator=(const qvrcnn &) = delete;

    // The reference leaves its members public and its driver reads I1.x_rec (kernel.cu:96).
    struct InputLayerView { int batch, height, width, inChannel; void *x, *x_rec; } I1;
    int batch, channel, height, width;
    qv_net *net = nullptr;

private:
    static void check(int status)                                        // inference/cnn.cuh:8-15
    {
        if (status != 0) {
            printf("qvrcnn returned none 0: %s\n", qv_last_error());
            exit(1);
        }
    }
};
