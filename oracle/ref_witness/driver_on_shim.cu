// TEST INFRASTRUCTURE: the reference's own driver -- testqvrcnn / run_all / main, inference/kernel.cu:74-138, the lines
// the Makefile extracts VERBATIM from /root/reference into oracle/_ref/ref_kernel_cu_74_138.inc at build time (never
// committed) -- compiled against THIS repository's drop-in headers (qcnn_gpu_b200/csrc/qvrcnn.cuh, yuv_data.h) and linked
// with libqvrcnn_b200.so instead of the reference's cnn.cu / mat.cu / qvrcnn.cu / yuv_data.cpp + cuDNN.  If the shim were
// not a drop-in for the path (names, signatures, public members such as I1.x_rec, datatype), this would not compile.
// Everything below the includes that is not the reference's text replaces what the reference takes from <windows.h>
// and from its compile-time configuration macros (inference/mat.cuh:23-37, kernel.cu:7-15).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>

#include <cuda_runtime.h>

#include "qvrcnn.cuh"          // -I qcnn_gpu_b200/csrc: the shim, not the reference's header

typedef union { long long QuadPart; } LARGE_INTEGER;
static inline void QueryPerformanceFrequency(LARGE_INTEGER *f) { f->QuadPart = 1000000000ll; }
static inline void QueryPerformanceCounter(LARGE_INTEGER *c)
{
    c->QuadPart = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define FRAME 1                              /* inference/mat.cuh: FRAME / CHANNEL are build-time macros there */
#define CHANNEL 1
#define MODEL_NAME "qvrcnn_nchw_vect_c_8bit_qfp_%d.data"   /* kernel.cu:7 without the author's D:\ directory */

#include "ref_kernel_cu_74_138.inc"
