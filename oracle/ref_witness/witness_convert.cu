// Reference converter witness (TEST INFRASTRUCTURE): calls the reference's own, unmodified
// model_qfp_HWCN2NCHW_VECT_C (inference/qvrcnn.cu:558-585 -> layer_qfp_HWCN2NCHW_VECT_C :535-557 ->
// HWCN2NCHW_VECT_C_CPU inference/mat.cu:97-119), linked from the objects the Makefile beside this file compiles
// out of /root/reference/inference.  Host-only code: needs no GPU.
// usage: qcnn_ref_convert <in_template_with_%d> <out_template_with_%d> <qp>
#include <cstdio>
#include <cstdlib>

#include "qvrcnn.cuh"

int main(int argc, char **argv)
{
    if (argc != 4) { fprintf(stderr, "usage: %s in_%%d.data out_%%d.data qp\n", argv[0]); return 2; }
    return model_qfp_HWCN2NCHW_VECT_C(argv[1], argv[2], atoi(argv[3]));
}
