// Header-only C++ shim: class vrcnn_data with the reference's name, constructor, public fields and
// methods (inference/yuv_data.h:11-27, inference/yuv_data.cpp), implemented over the C ABI.
// Error behaviour is the reference's: print and exit(1) (inference/yuv_data.cpp:19-31).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cmath>

#include "../../include/qvrcnn_b200.h"

typedef unsigned char datatype;   // inference/yuv_data.h:9
typedef char restype;             // inference/yuv_data.h:10

class vrcnn_data {
public:
    vrcnn_data(int frame, int height, int width)                      // yuv_data.cpp:3-14
        : frame(frame), h(height), w(width), nSize(frame * height * width), xSize(0)
    {
        // page-locked (qv_host_alloc): the driver's per-frame load_data / cudaMemcpy of x_rec become plain DMA
        ori = static_cast<datatype *>(qv_host_alloc((size_t)nSize));
        input = static_cast<datatype *>(qv_host_alloc((size_t)nSize));
        recon = static_cast<datatype *>(qv_host_alloc((size_t)nSize));
        if (!ori || !input || !recon) { printf("vrcnn_data: out of memory\n"); exit(1); }
    }
    int read_data(const char *orifile, const char *inputfile)          // yuv_data.cpp:15-42
    {
        if (qv_yuv_read_luma(orifile, frame, h, w, ori)) { printf("%s\nopen ori file failed\n", orifile); exit(1); }
        if (qv_yuv_read_luma(inputfile, frame, h, w, input)) { printf("%s\nopen input file failed\n", inputfile); exit(1); }
        return 0;
    }
    int read_frame(const char *orifile, const char *inputfile, int n)  // yuv_data.cpp:44-66
    {
        frame = 1;
        if (qv_yuv_read_frame(orifile, n, h, w, ori) || qv_yuv_read_frame(inputfile, n, h, w, input)) {
            printf("open file failed\n");
            return 1;
        }
        return 0;
    }
    double psnr(datatype *data) { return qv_psnr(data, ori, (size_t)nSize, nullptr); }   // yuv_data.cpp:87-97
    double psnr_pf(void)                                               // yuv_data.cpp:98-112
    {
        double p = 0;
        for (int n = 0; n < frame; ++n) {
            p = qv_psnr(recon + (size_t)n * h * w, ori + (size_t)n * h * w, (size_t)h * w, nullptr);
            printf("PSNR of Frame %d:%f\n", n, p);
        }
        return p;
    }
    int save_recon_as(const char *filename)                            // yuv_data.cpp:113-128
    {
        if (qv_yuv_write_recon(filename, recon, frame, h, w)) printf("write file failed\n");
        return 0;
    }
    ~vrcnn_data(void) { qv_host_free(ori); qv_host_free(input); qv_host_free(recon); }
    vrcnn_data(const vrcnn_data &) = delete;
    vrcnn_data &operator=(const vrcnn_data &) = delete;

    int frame, h, w, nSize, xSize;
    datatype *ori, *input, *recon;
};
