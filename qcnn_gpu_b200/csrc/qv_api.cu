// C ABI of the network object (include/qvrcnn_b200.h): handle lifetime, model loading,
// the reference's load_data / forward_blu / x_rec read-back, and the batched / pipelined
// variants the multi-GPU driver uses.  Mirrors class qvrcnn (inference/qvrcnn.cuh:25-59,
// inference/qvrcnn.cu:4-68,168-242) and the hot loop of testqvrcnn (inference/kernel.cu:86-97).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "qv_fused.h"
#include "qv_internal.h"
#include "qv_layered.h"

using namespace qv;

#define QV_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return QV_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

// ---- spatial strips over several GPUs ---------------------------------------------------------
// One cudaMalloc'ed block per handle: a 256-byte header (published / completed sequence numbers, CTA counter) and two
// input slots holding this GPU's rows of the frame.  The block is what neighbours map (same process: peer access;
// other process: CUDA IPC) and read their 6 halo rows from, straight out of this GPU's HBM over NVLink.
namespace {
constexpr int STRIP_HALO = 6;                   // receptive-field radius of the net: 2 + 2 + 1 + 1 rows
constexpr size_t STRIP_HDR = 256, STRIP_PUB = 0, STRIP_DONE = 64, STRIP_CTR = 128;
constexpr uint32_t STRIP_MAGIC = 0x51565354u;   // "QVST"
struct StripDescImpl {                          // the content of qv_strip_desc
    uint32_t magic;
    int32_t dev, row0, row1, width, img_h;
    uint64_t nonce;                             // identifies the exporting process
    uint64_t base, slot_off, slot_stride;
    uint64_t net;                               // the exporting handle (meaningful inside the exporting process only)
    cudaIpcMemHandle_t ipc;
    char pci[16];                               // PCI bus id of the device: device ordinals differ between processes
};
static_assert(sizeof(StripDescImpl) <= sizeof(qv_strip_desc), "qv_strip_desc too small");
struct StripPeer {
    bool attached = false, ipc = false, same_device = false;
    uint8_t *base = nullptr;
    qv_net *net = nullptr;                      // same process only
    StripDescImpl d{};
};
struct StripState {
    bool on = false;
    int img_h = 0, row0 = 0, row1 = 0;
    uint8_t *block = nullptr;
    size_t slot_stride = 0;
    uint32_t seq = 0, slot_seq[2] = {0, 0};
    uint32_t fills[2] = {0, 0}, forwards[2] = {0, 0};   // per slot: how often acquired for filling / consumed (ordering checks between strips that share a GPU)
    cudaStream_t last_stream = nullptr;
    bool used = false;
    StripPeer peer[2];                          // QV_STRIP_ABOVE, QV_STRIP_BELOW
};
// Strip-mode handles of this process, per device.  Kernels that wait for one another must never share a GPU (nothing
// guarantees that two launches on one GPU run at the same time), so strips that live on the SAME device are ordered by
// the stream they are driven on instead: one stream for all of them, all loads of a frame before its forwards (checked).
std::mutex g_strip_mu;
std::vector<qv_net *> g_strip_nets[64];
uint64_t process_nonce()
{
    static const uint64_t n = [] {
        std::random_device rd;
        return ((uint64_t)rd() << 32) ^ (uint64_t)rd() ^ ((uint64_t)getpid() << 17);
    }();
    return n;
}
}  // namespace

struct qv_net {
    StripState strip;
    int dev = 0, batch = 0, H = 0, W = 0;
    cudaStream_t st = nullptr;
    cudaStream_t pst[2] = {nullptr, nullptr};      // pipeline slots of qv_forward_frames_host
    cudaEvent_t ev_compute = nullptr;              // orders the compute stages of the two slots (shared scratch)
    uint8_t *d_x = nullptr, *d_rec = nullptr;      // InputLayer::x / x_rec   (inference/cnn.cu:433-434)
    uint8_t *d_slot_in[2] = {nullptr, nullptr}, *d_slot_out[2] = {nullptr, nullptr};
    uint8_t *h_pin_in[2] = {nullptr, nullptr}, *h_pin_out[2] = {nullptr, nullptr};
    int8_t *d_a1 = nullptr, *d_a2 = nullptr, *d_a3 = nullptr;   // C1.v / Conc1.conc / Conc2.conc (layered path)
    int act_frames = 0;
    uint8_t *d_rows_tmp = nullptr;
    size_t rows_tmp_bytes = 0;
    ModelHost model;
    LayeredModel lm;
    FusedModel *fm = nullptr;
    bool uploaded = false;
    int impl = QV_IMPL_AUTO;
    long long launches = 0;
};

static int set_device(const qv_net *net) { QV_CUDA(cudaSetDevice(net->dev)); return QV_OK; }

static void free_layered(qv_net *net)
{
    for (int l = 0; l < QV_NLAYER; ++l) {
        cudaFree(net->lm.L[l].d_wpk);
        cudaFree(net->lm.L[l].d_bias);
        net->lm.L[l].d_wpk = nullptr;
        net->lm.L[l].d_bias = nullptr;
    }
}

static inline int pack4(const int8_t *p, int stride, int n)
{
    int v = 0;
    for (int j = 0; j < n; ++j) v |= ((int)(uint8_t)p[j * stride]) << (8 * j);
    return v;
}

// Re-pack the plain [K][C][R][S] host weights into the word layouts the layered kernels read.
static int upload_layered(qv_net *net)
{
    free_layered(net);
    for (int l = 0; l < QV_NLAYER; ++l) {
        const LayerShape &s = kLayers[l];
        const LayerHost &L = net->model.L[l];
        std::vector<int32_t> wpk;
        const int R = s.k;
        if (l == QV_C1) {                        // [r][k][2]
            wpk.assign(5 * 64 * 2, 0);
            for (int r = 0; r < 5; ++r)
                for (int k = 0; k < 64; ++k) {
                    const int8_t *row = &L.w[((size_t)k * 5 + r) * 5];
                    wpk[(r * 64 + k) * 2 + 0] = pack4(row, 1, 4);
                    wpk[(r * 64 + k) * 2 + 1] = pack4(row + 4, 1, 1);
                }
        } else if (l == QV_C4) {                 // [tap][c4]
            wpk.assign(9 * 12, 0);
            for (int t = 0; t < 9; ++t)
                for (int c4 = 0; c4 < 12; ++c4)
                    wpk[t * 12 + c4] = pack4(&L.w[((size_t)(4 * c4) * 3 + t / 3) * 3 + t % 3], 9, 4);
        } else {                                 // [group][tap][c4][16]
            const int C4 = s.cin / 4, G = s.cout / 16;
            wpk.assign((size_t)G * R * R * C4 * 16, 0);
            for (int g = 0; g < G; ++g)
                for (int t = 0; t < R * R; ++t)
                    for (int c4 = 0; c4 < C4; ++c4)
                        for (int j = 0; j < 16; ++j) {
                            const int k = g * 16 + j;
                            wpk[(((size_t)g * R * R + t) * C4 + c4) * 16 + j] =
                                pack4(&L.w[(((size_t)k * s.cin + 4 * c4) * R + t / R) * R + t % R], R * R, 4);
                        }
        }
        LayeredLayer &D = net->lm.L[l];
        D.cout = s.cout;
        D.q.blu = L.blu; D.q.mul = L.mul; D.q.shift = L.shift;
        D.q.rbias = (1 << (L.shift - 1)) / L.mul;                       // inference/mat.cu:268
        QV_CUDA(cudaMalloc(&D.d_wpk, wpk.size() * sizeof(int32_t)));
        QV_CUDA(cudaMalloc(&D.d_bias, L.b.size() * sizeof(int32_t)));
        QV_CUDA(cudaMemcpyAsync(D.d_wpk, wpk.data(), wpk.size() * sizeof(int32_t), cudaMemcpyHostToDevice, net->st));
        QV_CUDA(cudaMemcpyAsync(D.d_bias, L.b.data(), L.b.size() * sizeof(int32_t), cudaMemcpyHostToDevice, net->st));
        QV_CUDA(cudaStreamSynchronize(net->st));   // wpk is a local
    }
    net->lm.c4_bias = net->model.L[QV_C4].b[0];
    return QV_OK;
}

static int validate_qparams(const ModelHost &m)
{
    for (int l = 0; l < QV_NLAYER; ++l) {
        const LayerHost &L = m.L[l];
        if (L.mul <= 0 || L.shift < 1 || L.shift > 30) {
            set_error("layer %d: mul=%d shift=%d outside the supported range (mul>0, 1<=shift<=30)", l, L.mul, L.shift);
            return QV_ERR_RANGE;
        }
        if (l != QV_C4 && (L.blu < 0 || L.blu >= (1 << 24))) {
            set_error("layer %d: blu=%d outside [0, 2^24)", l, L.blu);
            return QV_ERR_RANGE;
        }
    }
    return QV_OK;
}

// Called whenever the host model changed; (re)builds device images once the model is complete.
static int sync_model(qv_net *net)
{
    net->uploaded = false;
    if (!net->model.complete()) return QV_OK;
    int rc = validate_qparams(net->model);
    if (rc) return rc;
    rc = check_fp32_exact_envelope(net->model);
    if (rc) return rc;
    if ((rc = set_device(net))) return rc;
    if ((rc = upload_layered(net))) return rc;
    if (net->fm) { fused_free(net->fm); net->fm = nullptr; }
    net->fm = fused_upload(net->model, net->st);     // may be null when the fused path is not built
    net->uploaded = true;
    return QV_OK;
}

static int ensure_acts(qv_net *net)
{
    if (net->d_a1) return QV_OK;
    // C1.v, Conc1.conc, Conc2.conc of the reference (inference/cnn.cu:64,281), NHWC here; the
    // reference also holds 644 B/px of fp32 `u` tensors which this design never materialises.
    int frames = std::min(net->batch, 8);
    const size_t px = (size_t)frames * net->H * net->W;
    QV_CUDA(cudaMalloc(&net->d_a1, px * 64));
    QV_CUDA(cudaMalloc(&net->d_a2, px * 48));
    QV_CUDA(cudaMalloc(&net->d_a3, px * 48));
    net->act_frames = frames;
    return QV_OK;
}

static int run_forward(qv_net *net, const uint8_t *d_in, uint8_t *d_out, int n, int H, int W, cudaStream_t st)
{
    if (!net->uploaded) {
        set_error("forward called before a complete model (weights + quant params) was loaded");
        return QV_ERR_STATE;
    }
    int impl = net->impl;
    if (impl == QV_IMPL_AUTO) impl = net->fm ? QV_IMPL_FUSED : QV_IMPL_LAYERED;
    if (impl == QV_IMPL_FUSED) {
        if (!net->fm) { set_error("fused tcgen05 path is not available in this build"); return QV_ERR_STATE; }
        QV_CUDA(fused_forward(net->fm, d_in, d_out, n, H, W, st, &net->launches));
        return QV_OK;
    }
    int rc = ensure_acts(net);
    if (rc) return rc;
    // activation scratch was sized for act_frames frames of net->H x net->W
    const size_t cap_px = (size_t)net->act_frames * net->H * net->W;
    const size_t fpx = (size_t)H * W;
    if (fpx > cap_px) { set_error("frame of %dx%d exceeds the handle's activation scratch", W, H); return QV_ERR_ARG; }
    const int chunk = (int)std::max<size_t>(1, cap_px / fpx);
    for (int f0 = 0; f0 < n; f0 += chunk) {
        const int c = std::min(chunk, n - f0);
        QV_CUDA(layered_forward(net->lm, d_in + (size_t)f0 * fpx, d_out + (size_t)f0 * fpx, c, H, W, net->d_a1,
                                net->d_a2, net->d_a3, st, &net->launches));
    }
    return QV_OK;
}

extern "C" {

int qv_create(int gpu_num, int batch, int channel, int height, int width, qv_net **out)
{
    if (!out) { set_error("qv_create: null output pointer"); return QV_ERR_ARG; }
    *out = nullptr;
    if (channel != 1) { set_error("qv_create: channel must be 1 (luma), got %d", channel); return QV_ERR_ARG; }
    if (batch < 1 || height < 1 || width < 1) { set_error("qv_create: bad geometry %dx%dx%d", batch, height, width); return QV_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("qv_create: no CUDA device (this library has no CPU fallback)");
        return QV_ERR_CUDA;
    }
    if (gpu_num < 0 || gpu_num >= ndev) { set_error("qv_create: gpu_num %d out of range (%d devices)", gpu_num, ndev); return QV_ERR_ARG; }
    QV_CUDA(cudaSetDevice(gpu_num));                                    // inference/qvrcnn.cu:6
    qv_net *net = new qv_net();
    net->dev = gpu_num; net->batch = batch; net->H = height; net->W = width;
    const size_t bytes = (size_t)batch * height * width;
    cudaError_t e = cudaStreamCreateWithFlags(&net->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&net->d_x, bytes);             // inference/cnn.cu:433
    if (e == cudaSuccess) e = cudaMalloc(&net->d_rec, bytes);           // inference/cnn.cu:434
    if (e != cudaSuccess) {
        set_error("qv_create: %s", cudaGetErrorString(e));
        qv_destroy(net);
        return QV_ERR_CUDA;
    }
    *out = net;
    return QV_OK;
}

int qv_destroy(qv_net *net)
{
    if (!net) return QV_OK;
    cudaSetDevice(net->dev);
    if (net->st) cudaStreamSynchronize(net->st);
    qv_strip_release(net);
    free_layered(net);
    if (net->fm) fused_free(net->fm);
    cudaFree(net->d_x); cudaFree(net->d_rec);
    cudaFree(net->d_a1); cudaFree(net->d_a2); cudaFree(net->d_a3);
    cudaFree(net->d_rows_tmp);
    for (int i = 0; i < 2; ++i) {
        cudaFree(net->d_slot_in[i]); cudaFree(net->d_slot_out[i]);
        cudaFreeHost(net->h_pin_in[i]); cudaFreeHost(net->h_pin_out[i]);
        if (net->pst[i]) cudaStreamDestroy(net->pst[i]);
    }
    if (net->ev_compute) cudaEventDestroy(net->ev_compute);
    if (net->st) cudaStreamDestroy(net->st);
    delete net;
    return QV_OK;
}

int qv_load_static_para_mem(qv_net *net, const void *image, size_t len)
{
    if (!net || !image) { set_error("qv_load_static_para_mem: null argument"); return QV_ERR_ARG; }
    ModelHost m;
    int rc = parse_model_vect_c((const uint8_t *)image, len, m);
    if (rc) return rc;
    net->model = m;
    return sync_model(net);
}

int qv_debug_fused_tables(const void *image, size_t len, uint8_t *wimg, uint32_t *ops, int32_t *consts, size_t sizes[3])
{
    if (!image || !sizes) { set_error("qv_debug_fused_tables: null argument"); return QV_ERR_ARG; }
    ModelHost m;
    int rc = parse_model_vect_c((const uint8_t *)image, len, m);
    if (rc) return rc;
    std::vector<uint8_t> w;
    std::vector<uint32_t> o;
    std::vector<int32_t> c;
    fused_debug_tables(m, w, o, c);
    if (wimg && sizes[0] >= w.size()) memcpy(wimg, w.data(), w.size());
    if (ops && sizes[1] >= o.size()) memcpy(ops, o.data(), o.size() * sizeof(uint32_t));
    if (consts && sizes[2] >= c.size()) memcpy(consts, c.data(), c.size() * sizeof(int32_t));
    sizes[0] = w.size(); sizes[1] = o.size(); sizes[2] = c.size();
    return QV_OK;
}

int qv_debug_fused_units(int sm_count, int n_frames, int height, int width, int row0, int row1, int row_window, int allow_line,
                         int32_t *units, size_t *n_units, int *grid)
{
    if (!n_units || !grid) { set_error("qv_debug_fused_units: null argument"); return QV_ERR_ARG; }
    std::vector<int> u;
    int g = 0;
    if (fused_debug_units(sm_count, n_frames, height, width, row0, row1, row_window != 0, allow_line != 0, u, g)) {
        set_error("qv_debug_fused_units: bad geometry");
        return QV_ERR_ARG;
    }
    if (units && *n_units >= u.size() / 5) memcpy(units, u.data(), u.size() * sizeof(int));
    *n_units = u.size() / 5;
    *grid = g;
    return QV_OK;
}

int qv_load_static_para(qv_net *net, const char *filename)
{
    if (!net) { set_error("qv_load_static_para: null handle"); return QV_ERR_ARG; }
    std::vector<uint8_t> buf;
    int rc = read_file(filename, buf);
    if (rc) { set_error("cannot open model file. (%s)", filename ? filename : "(null)"); return rc; }   // qvrcnn.cu:52
    return qv_load_static_para_mem(net, buf.data(), buf.size());
}

int qv_load_static_para_hwcn(qv_net *net, const char *filename)
{
    if (!net) { set_error("qv_load_static_para_hwcn: null handle"); return QV_ERR_ARG; }
    std::vector<uint8_t> buf;
    int rc = read_file(filename, buf);
    if (rc) return rc;
    ModelHost m;
    rc = parse_model_hwcn(buf.data(), buf.size(), m);
    if (rc) return rc;
    net->model = m;
    return sync_model(net);
}

int qv_load_quant_params(qv_net *net, const char *filename)
{
    if (!net) { set_error("qv_load_quant_params: null handle"); return QV_ERR_ARG; }
    int32_t q[18];
    int rc = qv_read_quant_params(filename, q);
    if (rc) return rc;
    for (int l = 0; l < QV_NLAYER; ++l) {
        LayerHost &L = net->model.L[l];
        L.blu = q[3 * l]; L.mul = q[3 * l + 1]; L.shift = q[3 * l + 2];
        L.have_q = true;
    }
    return sync_model(net);
}

int qv_set_weights(qv_net *net, int layer, const int8_t *w_kcrs, const int32_t *bias)
{
    if (!net || !w_kcrs || !bias || layer < 0 || layer >= QV_NLAYER) { set_error("qv_set_weights: bad argument"); return QV_ERR_ARG; }
    const LayerShape &s = kLayers[layer];
    LayerHost &L = net->model.L[layer];
    L.w.assign(w_kcrs, w_kcrs + (size_t)s.cout * s.cin * s.k * s.k);
    L.b.assign(bias, bias + s.cout);
    L.have_w = true;
    return sync_model(net);
}

int qv_get_quant_params(const qv_net *net, int32_t *out18)
{
    if (!net || !out18) { set_error("qv_get_quant_params: null argument"); return QV_ERR_ARG; }
    for (int l = 0; l < QV_NLAYER; ++l) {
        out18[3 * l] = net->model.L[l].blu; out18[3 * l + 1] = net->model.L[l].mul; out18[3 * l + 2] = net->model.L[l].shift;
    }
    return QV_OK;
}

int qv_load_data(qv_net *net, const uint8_t *host_luma)
{
    if (!net || !host_luma) { set_error("qv_load_data: null argument"); return QV_ERR_ARG; }
    int rc = set_device(net);
    if (rc) return rc;
    const size_t bytes = (size_t)net->batch * net->H * net->W;
    QV_CUDA(cudaMemcpyAsync(net->d_x, host_luma, bytes, cudaMemcpyHostToDevice, net->st));   // cnn.cu:441
    QV_CUDA(cudaStreamSynchronize(net->st));
    return QV_OK;
}

// After a synchronisation: did a CTA of the fused kernel report a failure?  (Never a silent wrong answer.)
static int kernel_report(qv_net *net)
{
    const int f = net->fm ? fused_take_failure(net->fm) : 0;
    if (!f) return QV_OK;
    set_error(f == 2 ? "fused kernel: a CTA saw other shared-memory / TMEM bases than the operand table was built for"
              : f == 3 ? "strip mode: a neighbour GPU's rows were never published (the frame data of this call is not valid)"
                       : "fused kernel: an mbarrier wait timed out (the frame data of this call is not valid)");
    return QV_ERR_CUDA;
}

int qv_forward_blu(qv_net *net)
{
    if (!net) { set_error("qv_forward_blu: null handle"); return QV_ERR_ARG; }
    int rc = set_device(net);
    if (rc) return rc;
    rc = run_forward(net, net->d_x, net->d_rec, net->batch, net->H, net->W, net->st);
    if (rc) return rc;
    QV_CUDA(cudaStreamSynchronize(net->st));        // the reference is synchronous (kernel.cu:95)
    return kernel_report(net);
}

int qv_get_recon(qv_net *net, uint8_t *host_out)
{
    if (!net || !host_out) { set_error("qv_get_recon: null argument"); return QV_ERR_ARG; }
    int rc = set_device(net);
    if (rc) return rc;
    const size_t bytes = (size_t)net->batch * net->H * net->W;
    QV_CUDA(cudaMemcpyAsync(host_out, net->d_rec, bytes, cudaMemcpyDeviceToHost, net->st));   // kernel.cu:96
    QV_CUDA(cudaStreamSynchronize(net->st));
    return kernel_report(net);
}

int qv_forward_frames_device(qv_net *net, const uint8_t *d_in, uint8_t *d_out, int n_frames, void *cuda_stream)
{
    if (!net || !d_in || !d_out || n_frames < 0) { set_error("qv_forward_frames_device: bad argument"); return QV_ERR_ARG; }
    int rc = set_device(net);
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : net->st;
    if (n_frames > 0) {
        rc = run_forward(net, d_in, d_out, n_frames, net->H, net->W, st);
        if (rc) return rc;
    }
    if (!cuda_stream) { QV_CUDA(cudaStreamSynchronize(st)); return kernel_report(net); }
    return QV_OK;
}

int qv_forward_rows_device(qv_net *net, const uint8_t *d_in, int img_height, int in_row0, int in_rows, uint8_t *d_out,
                           int out_row0, int out_row1, void *cuda_stream)
{
    if (!net || !d_in || !d_out) { set_error("qv_forward_rows_device: null argument"); return QV_ERR_ARG; }
    // The net's receptive-field radius is 2+2+1+1 = 6 rows: every output row needs the input rows within 6 of it that
    // lie inside the image; outside the image every layer pads its own input with zeros (inference/cnn.cu:44-49).
    const int in_row1 = in_row0 + in_rows;
    if (img_height < 1 || in_row0 < 0 || in_rows < 1 || in_row1 > img_height || out_row0 < in_row0 || out_row1 > in_row1 ||
        out_row0 > out_row1) { set_error("qv_forward_rows_device: inconsistent row ranges"); return QV_ERR_ARG; }
    if ((in_row0 != 0 && out_row0 - in_row0 < STRIP_HALO) || (in_row1 != img_height && in_row1 - out_row1 < STRIP_HALO)) {
        set_error("qv_forward_rows_device: need 6 halo rows on every interior strip edge");
        return QV_ERR_ARG;
    }
    int rc = set_device(net);
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : net->st;
    const size_t W = net->W;
    const bool fused_path = net->impl == QV_IMPL_FUSED || (net->impl == QV_IMPL_AUTO && net->fm);
    if (fused_path) {
        // the fused kernel takes the window as what it is: rows [in_row0, in_row1) of an img_height-row image, of which it
        // computes and writes exactly [out_row0, out_row1) -- no scratch frame, no copy
        if (!net->uploaded || !net->fm) { set_error("forward called before a complete model (weights + quant params) was loaded"); return QV_ERR_STATE; }
        FusedRows fr;
        fr.own0 = in_row0; fr.own1 = in_row1; fr.out0 = out_row0; fr.out1 = out_row1;
        QV_CUDA(fused_forward(net->fm, d_in, d_out, 1, img_height, net->W, st, &net->launches, &fr));
    } else {
        // layered path: the window is run as an image of its own (exact for rows >= 6 away from an interior window edge),
        // then the asked-for rows are copied out
        const size_t need = (size_t)in_rows * W;
        if (need > (size_t)net->batch * net->H * net->W) {
            set_error("qv_forward_rows_device: window larger than the handle's geometry");
            return QV_ERR_ARG;
        }
        if (net->rows_tmp_bytes < need) {
            QV_CUDA(cudaStreamSynchronize(st));
            cudaFree(net->d_rows_tmp);
            net->d_rows_tmp = nullptr; net->rows_tmp_bytes = 0;
            QV_CUDA(cudaMalloc(&net->d_rows_tmp, need));
            net->rows_tmp_bytes = need;
        }
        rc = run_forward(net, d_in, net->d_rows_tmp, 1, in_rows, net->W, st);
        if (rc) return rc;
        QV_CUDA(cudaMemcpyAsync(d_out, net->d_rows_tmp + (size_t)(out_row0 - in_row0) * W, (size_t)(out_row1 - out_row0) * W,
                                cudaMemcpyDeviceToDevice, st));
    }
    if (!cuda_stream) { QV_CUDA(cudaStreamSynchronize(st)); return kernel_report(net); }
    return QV_OK;
}

int qv_synchronize(qv_net *net, void *cuda_stream)
{
    if (!net) { set_error("qv_synchronize: null handle"); return QV_ERR_ARG; }
    int rc = set_device(net);
    if (rc) return rc;
    QV_CUDA(cudaStreamSynchronize(cuda_stream ? (cudaStream_t)cuda_stream : net->st));
    return kernel_report(net);
}

// ---- strips: setup / export / attach / load / forward -------------------------------------------------------------
static void strip_detach(qv_net *net)
{
    for (StripPeer &p : net->strip.peer) {
        if (p.attached && p.ipc && p.base) cudaIpcCloseMemHandle(p.base);
        p = StripPeer{};
    }
}

int qv_strip_release(qv_net *net)
{
    if (!net) { set_error("qv_strip_release: null handle"); return QV_ERR_ARG; }
    if (!net->strip.on) return QV_OK;
    cudaSetDevice(net->dev);
    cudaStreamSynchronize(net->st);
    strip_detach(net);
    cudaFree(net->strip.block);
    net->strip = StripState{};
    if (net->dev >= 0 && net->dev < 64) {
        std::lock_guard<std::mutex> l(g_strip_mu);
        auto &v = g_strip_nets[net->dev];
        v.erase(std::remove(v.begin(), v.end(), net), v.end());
    }
    return QV_OK;
}

int qv_strip_setup(qv_net *net, int img_height, int row0, int row1)
{
    if (!net) { set_error("qv_strip_setup: null handle"); return QV_ERR_ARG; }
    if (img_height < 1 || row0 < 0 || row1 <= row0 || row1 > img_height) { set_error("qv_strip_setup: rows [%d, %d) of %d", row0, row1, img_height); return QV_ERR_ARG; }
    int rc = set_device(net);
    if (rc) return rc;
    qv_strip_release(net);
    StripState &S = net->strip;
    S.img_h = img_height; S.row0 = row0; S.row1 = row1;
    S.slot_stride = ((size_t)(row1 - row0) * net->W + 255) / 256 * 256;
    QV_CUDA(cudaMalloc(&S.block, STRIP_HDR + 2 * S.slot_stride));
    QV_CUDA(cudaMemset(S.block, 0, STRIP_HDR));
    QV_CUDA(cudaDeviceSynchronize());
    S.on = true;
    if (net->dev >= 0 && net->dev < 64) {
        std::lock_guard<std::mutex> l(g_strip_mu);
        g_strip_nets[net->dev].push_back(net);
    }
    return QV_OK;
}

int qv_strip_export(qv_net *net, qv_strip_desc *out)
{
    if (!net || !out) { set_error("qv_strip_export: null argument"); return QV_ERR_ARG; }
    if (!net->strip.on) { set_error("qv_strip_export: qv_strip_setup has not been called"); return QV_ERR_STATE; }
    int rc = set_device(net);
    if (rc) return rc;
    StripDescImpl d{};
    d.magic = STRIP_MAGIC; d.dev = net->dev; d.row0 = net->strip.row0; d.row1 = net->strip.row1; d.width = net->W; d.img_h = net->strip.img_h;
    d.nonce = process_nonce();
    d.base = (uint64_t)(uintptr_t)net->strip.block; d.slot_off = STRIP_HDR; d.slot_stride = net->strip.slot_stride;
    d.net = (uint64_t)(uintptr_t)net;
    QV_CUDA(cudaIpcGetMemHandle(&d.ipc, net->strip.block));
    QV_CUDA(cudaDeviceGetPCIBusId(d.pci, (int)sizeof(d.pci), net->dev));
    memset(out, 0, sizeof(*out));
    memcpy(out, &d, sizeof(d));
    return QV_OK;
}

int qv_strip_attach(qv_net *net, int side, const qv_strip_desc *neighbour)
{
    if (!net || !neighbour || (side != QV_STRIP_ABOVE && side != QV_STRIP_BELOW)) { set_error("qv_strip_attach: bad argument"); return QV_ERR_ARG; }
    StripState &S = net->strip;
    if (!S.on) { set_error("qv_strip_attach: qv_strip_setup has not been called"); return QV_ERR_STATE; }
    StripDescImpl d;
    memcpy(&d, neighbour, sizeof(d));
    if (d.magic != STRIP_MAGIC) { set_error("qv_strip_attach: not a strip descriptor"); return QV_ERR_ARG; }
    if (d.width != net->W || d.img_h != S.img_h) { set_error("qv_strip_attach: the neighbour's frame is %dx%d, this one %dx%d", d.width, d.img_h, net->W, S.img_h); return QV_ERR_ARG; }
    if (side == QV_STRIP_ABOVE ? d.row1 != S.row0 : d.row0 != S.row1) {
        set_error("qv_strip_attach: rows [%d, %d) are not adjacent %s rows [%d, %d)", d.row0, d.row1, side == QV_STRIP_ABOVE ? "above" : "below", S.row0, S.row1);
        return QV_ERR_ARG;
    }
    if (d.row1 - d.row0 < STRIP_HALO) {
        // the halo would reach into the strip after next, which this protocol does not exchange
        set_error("qv_strip_attach: the neighbour holds %d rows, fewer than the %d-row halo (use fewer strips)", d.row1 - d.row0, STRIP_HALO);
        return QV_ERR_ARG;
    }
    int rc = set_device(net);
    if (rc) return rc;
    StripPeer &P = S.peer[side];
    if (P.attached && P.ipc && P.base) cudaIpcCloseMemHandle(P.base);
    P = StripPeer{};
    if (d.nonce == process_nonce()) {
        // same process: plain peer access to the neighbour's allocation
        if (d.dev != net->dev) {
            int can = 0;
            QV_CUDA(cudaDeviceCanAccessPeer(&can, net->dev, d.dev));
            if (!can) { set_error("qv_strip_attach: device %d cannot map memory of device %d (no peer access)", net->dev, d.dev); return QV_ERR_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(d.dev, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            QV_CUDA(e);
        }
        P.base = (uint8_t *)(uintptr_t)d.base;
        P.net = (qv_net *)(uintptr_t)d.net;
    } else {
        char my_pci[16] = {0};
        QV_CUDA(cudaDeviceGetPCIBusId(my_pci, (int)sizeof(my_pci), net->dev));
        if (strncmp(my_pci, d.pci, sizeof(my_pci)) == 0) {
            // two processes time-slicing one GPU: a kernel of one that waits for a kernel of the other may never see it run
            set_error("qv_strip_attach: the neighbour strip belongs to another process on the SAME GPU; strips of different processes must be on different GPUs");
            return QV_ERR_ARG;
        }
        void *p = nullptr;
        QV_CUDA(cudaIpcOpenMemHandle(&p, d.ipc, cudaIpcMemLazyEnablePeerAccess));
        P.base = (uint8_t *)p;
        P.ipc = true;
    }
    P.d = d;
    char my_pci[16] = {0};
    QV_CUDA(cudaDeviceGetPCIBusId(my_pci, (int)sizeof(my_pci), net->dev));
    P.same_device = strncmp(my_pci, d.pci, sizeof(my_pci)) == 0;
    P.attached = true;
    return QV_OK;
}

int qv_strip_input(qv_net *net, int slot, void **d_rows)
{
    if (!net || !d_rows || slot < 0 || slot > 1) { set_error("qv_strip_input: bad argument"); return QV_ERR_ARG; }
    if (!net->strip.on) { set_error("qv_strip_input: qv_strip_setup has not been called"); return QV_ERR_STATE; }
    *d_rows = net->strip.block + STRIP_HDR + (size_t)slot * net->strip.slot_stride;
    return QV_OK;
}

// Strips that share this handle's GPU (same process) are ordered by the one stream they must all be driven on.
static int check_shared_device_stream(qv_net *net, cudaStream_t st, const char *who)
{
    std::lock_guard<std::mutex> l(g_strip_mu);
    for (qv_net *o : g_strip_nets[net->dev])
        if (o != net && o->strip.used && o->strip.last_stream != st) {
            set_error("%s: several strips live on GPU %d: drive all of them on ONE caller-provided stream (kernels of different launches "
                      "must not wait for each other on one GPU, so stream order is what orders them)", who, net->dev);
            return QV_ERR_STATE;
        }
    return QV_OK;
}

int qv_strip_acquire(qv_net *net, int slot, void *cuda_stream)
{
    if (!net || slot < 0 || slot > 1) { set_error("qv_strip_acquire: bad argument"); return QV_ERR_ARG; }
    StripState &S = net->strip;
    if (!S.on) { set_error("qv_strip_acquire: qv_strip_setup has not been called"); return QV_ERR_STATE; }
    int rc = set_device(net);
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : net->st;
    if ((rc = check_shared_device_stream(net, st, "qv_strip_acquire"))) return rc;
    S.last_stream = st; S.used = true;
    const uint32_t s = S.slot_seq[slot];
    const uint32_t *w[2] = {nullptr, nullptr};
    for (int side = 0; side < 2; ++side) {
        const StripPeer &P = S.peer[side];
        if (!P.attached || s == 0) continue;
        if (P.same_device) {
            // the neighbour's forward that read this slot must already be on the (shared) stream
            if (P.net->strip.forwards[slot] < S.forwards[slot]) {
                set_error("qv_strip_acquire: the neighbour strip on the same GPU has not yet consumed slot %d (issue the strips' calls frame by frame)", slot);
                return QV_ERR_STATE;
            }
        } else {
            w[side] = reinterpret_cast<const uint32_t *>(P.base + STRIP_DONE);      // another GPU: wait for its "completed step s"
        }
    }
    if (net->fm && (w[0] || w[1])) {
        QV_CUDA(fused_wait_words(net->fm, w[0], s, w[1], s, st));
        net->launches += 1;
    }
    S.fills[slot] += 1;
    return QV_OK;
}

int qv_strip_load(qv_net *net, int slot, const uint8_t *host_rows, void *cuda_stream)
{
    if (!host_rows) { set_error("qv_strip_load: null argument"); return QV_ERR_ARG; }
    int rc = qv_strip_acquire(net, slot, cuda_stream);
    if (rc) return rc;
    StripState &S = net->strip;
    QV_CUDA(cudaMemcpyAsync(S.block + STRIP_HDR + (size_t)slot * S.slot_stride, host_rows, (size_t)(S.row1 - S.row0) * net->W,
                            cudaMemcpyHostToDevice, cuda_stream ? (cudaStream_t)cuda_stream : net->st));
    return QV_OK;
}

int qv_strip_forward(qv_net *net, int slot, uint8_t *d_out, void *cuda_stream)
{
    if (!net || !d_out || slot < 0 || slot > 1) { set_error("qv_strip_forward: bad argument"); return QV_ERR_ARG; }
    StripState &S = net->strip;
    if (!S.on) { set_error("qv_strip_forward: qv_strip_setup has not been called"); return QV_ERR_STATE; }
    if (!net->uploaded || !net->fm || net->impl == QV_IMPL_LAYERED) {
        set_error("qv_strip_forward: needs a loaded model and the fused path (peer-mapped halo rows are read by the fused kernel)");
        return QV_ERR_STATE;
    }
    if ((S.row0 > 0 && !S.peer[QV_STRIP_ABOVE].attached) || (S.row1 < S.img_h && !S.peer[QV_STRIP_BELOW].attached)) {
        set_error("qv_strip_forward: rows [%d, %d) of %d have an interior edge with no neighbour attached", S.row0, S.row1, S.img_h);
        return QV_ERR_STATE;
    }
    int rc = set_device(net);
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : net->st;
    if ((rc = check_shared_device_stream(net, st, "qv_strip_forward"))) return rc;
    const size_t W = net->W;
    FusedRows fr;
    fr.own0 = fr.out0 = S.row0; fr.own1 = fr.out1 = S.row1;
    fr.pub = reinterpret_cast<uint32_t *>(S.block + STRIP_PUB);
    fr.done = reinterpret_cast<uint32_t *>(S.block + STRIP_DONE);
    fr.done_ctr = reinterpret_cast<uint32_t *>(S.block + STRIP_CTR);
    for (int side = 0; side < 2; ++side) {
        if (side == QV_STRIP_ABOVE ? S.row0 == 0 : S.row1 == S.img_h) continue;
        const StripPeer &P = S.peer[side];
        const uint8_t *rows = P.base + P.d.slot_off + (size_t)slot * P.d.slot_stride;
        const uint32_t *flag = reinterpret_cast<const uint32_t *>(P.base + STRIP_PUB);
        if (P.same_device) {
            // a neighbour on this very GPU: no kernel may wait for it -- its rows are ordered before this launch by the shared
            // stream, provided its slot was filled for this frame already (checked here, on the host)
            if (P.net->strip.fills[slot] == 0) {
                set_error("qv_strip_forward: the neighbour strip on the same GPU has not filled slot %d yet (strips that share a GPU: all loads of a frame, then its forwards, on one stream)", slot);
                return QV_ERR_STATE;
            }
            flag = nullptr;
        }
        if (side == QV_STRIP_ABOVE) { fr.top_rows = STRIP_HALO; fr.d_top = rows + (size_t)(P.d.row1 - P.d.row0 - STRIP_HALO) * W; fr.flag_top = flag; }
        else { fr.bot_rows = STRIP_HALO; fr.d_bot = rows; fr.flag_bot = flag; }
    }
    const uint32_t seq = ++S.seq;
    S.slot_seq[slot] = seq;
    S.forwards[slot] += 1;
    S.last_stream = st; S.used = true;
    fr.seq = seq;
    const uint8_t *own = S.block + STRIP_HDR + (size_t)slot * S.slot_stride;
    QV_CUDA(fused_forward(net->fm, own, d_out, 1, S.img_h, net->W, st, &net->launches, &fr));
    if (!cuda_stream) { QV_CUDA(cudaStreamSynchronize(st)); return kernel_report(net); }
    return QV_OK;
}

int qv_forward_frames_host(qv_net *net, const uint8_t *h_in, uint8_t *h_out, int n_frames)
{
    if (!net || !h_in || !h_out || n_frames < 0) { set_error("qv_forward_frames_host: bad argument"); return QV_ERR_ARG; }
    int rc = set_device(net);
    if (rc) return rc;
    const size_t fpx = (size_t)net->H * net->W;
    // Pipeline granularity: the H2D of chunk k+1 and the D2H of chunk k-1 hide behind the compute of chunk k;
    // only the first upload and the last download are exposed, so the schedule is tapered -- small chunks at
    // both ends (1, 3, 4, 8, 16, 16, 8, 4, 3, 1 sixty-fourths of the frames; measured best of seven schedules at
    // 64 x 1080p, profiles/r1_e2e_chunk_schedules.log), each at most `batch` frames.
    std::vector<int> chunks;
    {
        static const int num[10] = {1, 3, 4, 8, 16, 16, 8, 4, 3, 1};
        int left = n_frames;
        if (n_frames >= 16) {
            for (int k = 0; k < 10 && left > 0; ++k) {
                int want = k == 9 ? left : std::max(1, n_frames * num[k] / 64);
                while (want > 0 && left > 0) {
                    const int c = std::min(std::min(want, left), net->batch);
                    chunks.push_back(c); want -= c; left -= c;
                }
            }
        }
        const char *ov = getenv("QV_E2E_CHUNKS");                  // tuning only: explicit chunk sizes, e.g. "4,4,8,16,16,8,4,4"
        if (ov && *ov) {
            chunks.clear(); left = n_frames;
            for (const char *p = ov; *p && left > 0;) {
                const int c = std::min(std::min(std::max(1, atoi(p)), left), net->batch);
                chunks.push_back(c); left -= c;
                while (*p && *p != ',') ++p;
                if (*p == ',') ++p;
            }
        }
        const int uni = std::max(1, std::min(net->batch, (n_frames + 3) / 4));
        while (left > 0) { const int c = std::min(uni, left); chunks.push_back(c); left -= c; }
    }
    const size_t cbytes = (size_t)net->batch * fpx;
    // Is the caller's memory page-locked already?  Then DMA straight from/to it.
    cudaPointerAttributes ai{}, ao{};
    bool pinned_in = cudaPointerGetAttributes(&ai, h_in) == cudaSuccess && ai.type == cudaMemoryTypeHost;
    bool pinned_out = cudaPointerGetAttributes(&ao, h_out) == cudaSuccess && ao.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!net->ev_compute) QV_CUDA(cudaEventCreateWithFlags(&net->ev_compute, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        if (!net->pst[i]) QV_CUDA(cudaStreamCreateWithFlags(&net->pst[i], cudaStreamNonBlocking));
        if (!net->d_slot_in[i]) QV_CUDA(cudaMalloc(&net->d_slot_in[i], cbytes));
        if (!net->d_slot_out[i]) QV_CUDA(cudaMalloc(&net->d_slot_out[i], cbytes));
        if (!pinned_in && !net->h_pin_in[i]) QV_CUDA(cudaMallocHost(&net->h_pin_in[i], cbytes));
        if (!pinned_out && !net->h_pin_out[i]) QV_CUDA(cudaMallocHost(&net->h_pin_out[i], cbytes));
    }
    // Two slots, each an in-order H2D -> forward -> D2H chain on its own stream, so the copies of
    // one chunk overlap the compute of the other (the reference's loop, inference/kernel.cu:91-97,
    // is fully serial: blocking memcpy, forward, sync, blocking memcpy).
    struct Pending { int f0 = -1, c = 0; } pend[2];
    auto drain = [&](int s) -> int {
        if (pend[s].f0 < 0) return QV_OK;
        QV_CUDA(cudaStreamSynchronize(net->pst[s]));
        if (!pinned_out) memcpy(h_out + (size_t)pend[s].f0 * fpx, net->h_pin_out[s], (size_t)pend[s].c * fpx);
        pend[s].f0 = -1;
        return QV_OK;
    };
    int slot = 0;
    int f0 = 0;
    for (size_t ci = 0; ci < chunks.size(); f0 += chunks[ci], ++ci, slot ^= 1) {
        const int c = chunks[ci];
        const size_t bytes = (size_t)c * fpx;
        if ((rc = drain(slot))) return rc;
        const uint8_t *src = h_in + (size_t)f0 * fpx;
        if (!pinned_in) { memcpy(net->h_pin_in[slot], src, bytes); src = net->h_pin_in[slot]; }
        QV_CUDA(cudaMemcpyAsync(net->d_slot_in[slot], src, bytes, cudaMemcpyHostToDevice, net->pst[slot]));
        // The layered path's compute stages must run one after the other (they share the handle's activation scratch).
        // The fused kernel has no scratch: its launches of the two slots are left unordered, so the persistent CTAs of the
        // next chunk move onto the SMs the previous chunk's last wave has already left.
        const bool fused_path = net->impl == QV_IMPL_FUSED || (net->impl == QV_IMPL_AUTO && net->fm);
        if (f0 > 0 && !fused_path) QV_CUDA(cudaStreamWaitEvent(net->pst[slot], net->ev_compute, 0));
        rc = run_forward(net, net->d_slot_in[slot], net->d_slot_out[slot], c, net->H, net->W, net->pst[slot]);
        if (rc) return rc;
        QV_CUDA(cudaEventRecord(net->ev_compute, net->pst[slot]));
        uint8_t *dst = pinned_out ? h_out + (size_t)f0 * fpx : net->h_pin_out[slot];
        QV_CUDA(cudaMemcpyAsync(dst, net->d_slot_out[slot], bytes, cudaMemcpyDeviceToHost, net->pst[slot]));
        pend[slot].f0 = f0; pend[slot].c = c;
    }
    if ((rc = drain(0))) return rc;
    if ((rc = drain(1))) return rc;
    return kernel_report(net);
}

// ---- streaming file pipeline -------------------------------------------------------------
namespace {
struct StreamSlot {
    uint8_t *h_in = nullptr, *h_ori = nullptr, *h_out = nullptr;      // pinned
    uint8_t *d_in = nullptr, *d_ori = nullptr, *d_out = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    int state = 0;                       // 0 free -> 1 filled by the reader -> 2 launched -> 0 (written)
    int f0 = 0, c = 0;
};
struct StreamShared {
    std::mutex mu;
    std::condition_variable cv;
    bool failed = false;
    char msg[256] = {0};
    void fail(const char *m) { std::lock_guard<std::mutex> l(mu); if (!failed) { failed = true; snprintf(msg, sizeof(msg), "%s", m); } cv.notify_all(); }
};
bool read_luma(FILE *f, long long frame, size_t fpx, uint8_t *dst)
{
    if (fseeko(f, (off_t)(frame * (long long)(fpx + fpx / 2)), SEEK_SET) != 0) return false;     // yuv_data.cpp:32-38
    return fread(dst, 1, fpx, f) == fpx;
}
}  // namespace

int qv_stream_yuv(qv_net *net, const char *anchor_yuv, const char *ori_yuv, const char *recon_yuv, int first_frame, int n_frames,
                  int64_t *sse_before, int64_t *sse_after)
{
    if (!net || !anchor_yuv || first_frame < 0 || n_frames < 0) { set_error("qv_stream_yuv: bad argument"); return QV_ERR_ARG; }
    if ((sse_before || sse_after) && !ori_yuv) { set_error("qv_stream_yuv: an SSE was asked for but no original file given"); return QV_ERR_ARG; }
    int rc = set_device(net);
    if (rc) return rc;
    if (sse_before) *sse_before = 0;
    if (sse_after) *sse_after = 0;
    if (n_frames == 0) return QV_OK;
    const size_t fpx = (size_t)net->H * net->W;
    const int C = net->batch, NS = 3;
    FILE *fa = fopen(anchor_yuv, "rb"), *fo = ori_yuv ? fopen(ori_yuv, "rb") : nullptr, *fr = nullptr;
    if (!fa) { set_error("open file failed. (%s)", anchor_yuv); return QV_ERR_IO; }                       // yuv_data.cpp:19-31
    if (ori_yuv && !fo) { fclose(fa); set_error("open file failed. (%s)", ori_yuv); return QV_ERR_IO; }
    if (recon_yuv) {
        // several ranks / threads may fill one file, each its own frame range: open without truncating, create if missing
        // (a second caller that lost the race to create it must not wipe what the first has written)
        const int fd = open(recon_yuv, O_CREAT | O_RDWR, 0644);
        fr = fd >= 0 ? fdopen(fd, "r+b") : nullptr;
        if (!fr && fd >= 0) close(fd);
        if (!fr) { fclose(fa); if (fo) fclose(fo); set_error("open file failed. (%s)", recon_yuv); return QV_ERR_IO; }
    }
    StreamSlot slot[3];
    StreamShared sh;
    int64_t *d_sse = nullptr;
    cudaError_t e = cudaMalloc(&d_sse, 2 * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemset(d_sse, 0, 2 * sizeof(int64_t));
    for (int s = 0; s < NS && e == cudaSuccess; ++s) {
        StreamSlot &S = slot[s];
        e = cudaMallocHost(&S.h_in, C * fpx);
        if (e == cudaSuccess) e = cudaMallocHost(&S.h_out, C * fpx);
        if (e == cudaSuccess && fo) e = cudaMallocHost(&S.h_ori, C * fpx);
        if (e == cudaSuccess) e = cudaMalloc(&S.d_in, C * fpx);
        if (e == cudaSuccess) e = cudaMalloc(&S.d_out, C * fpx);
        if (e == cudaSuccess && fo) e = cudaMalloc(&S.d_ori, C * fpx);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming);
    }
    const int nchunk = (n_frames + C - 1) / C;
    auto cleanup = [&]() {
        for (int s = 0; s < NS; ++s) {
            StreamSlot &S = slot[s];
            if (S.st) cudaStreamSynchronize(S.st);
            cudaFreeHost(S.h_in); cudaFreeHost(S.h_out); cudaFreeHost(S.h_ori);
            cudaFree(S.d_in); cudaFree(S.d_out); cudaFree(S.d_ori);
            if (S.done) cudaEventDestroy(S.done);
            if (S.st) cudaStreamDestroy(S.st);
        }
        cudaFree(d_sse);
        fclose(fa); if (fo) fclose(fo); if (fr) fclose(fr);
    };
    if (e != cudaSuccess) { cleanup(); set_error("qv_stream_yuv: %s", cudaGetErrorString(e)); return QV_ERR_CUDA; }

    // reader: files -> pinned memory, one chunk ahead of the GPU
    std::thread reader([&] {
        for (int k = 0; k < nchunk; ++k) {
            StreamSlot &S = slot[k % NS];
            {
                std::unique_lock<std::mutex> l(sh.mu);
                sh.cv.wait(l, [&] { return S.state == 0 || sh.failed; });
                if (sh.failed) return;
            }
            const int f0 = k * C, c = std::min(C, n_frames - f0);
            for (int j = 0; j < c; ++j) {
                if (!read_luma(fa, (long long)first_frame + f0 + j, fpx, S.h_in + (size_t)j * fpx) ||
                    (fo && !read_luma(fo, (long long)first_frame + f0 + j, fpx, S.h_ori + (size_t)j * fpx))) {
                    sh.fail("qv_stream_yuv: short read (file smaller than the requested frame range)");
                    return;
                }
            }
            { std::lock_guard<std::mutex> l(sh.mu); S.f0 = f0; S.c = c; S.state = 1; }
            sh.cv.notify_all();
        }
    });
    // writer: pinned memory -> recon file, one chunk behind the GPU
    const int dev = net->dev;
    std::thread writer([&] {
        cudaSetDevice(dev);
        std::vector<uint8_t> zeros(fpx / 2, 0);
        for (int k = 0; k < nchunk; ++k) {
            StreamSlot &S = slot[k % NS];
            {
                std::unique_lock<std::mutex> l(sh.mu);
                sh.cv.wait(l, [&] { return S.state == 2 || sh.failed; });
                if (sh.failed) return;
            }
            if (cudaEventSynchronize(S.done) != cudaSuccess) { sh.fail("qv_stream_yuv: the GPU stage failed"); return; }
            if (fr) {
                bool ok = fseeko(fr, (off_t)(((long long)first_frame + S.f0) * (long long)(fpx + fpx / 2)), SEEK_SET) == 0;
                for (int j = 0; j < S.c && ok; ++j)                                          // yuv_data.cpp:119-125
                    ok = fwrite(S.h_out + (size_t)j * fpx, 1, fpx, fr) == fpx && fwrite(zeros.data(), 1, fpx / 2, fr) == fpx / 2;
                if (!ok) { sh.fail("qv_stream_yuv: write to the reconstruction file failed"); return; }
            }
            { std::lock_guard<std::mutex> l(sh.mu); S.state = 0; }
            sh.cv.notify_all();
        }
    });
    // GPU stage, in chunk order
    const bool fused_path = net->impl == QV_IMPL_FUSED || (net->impl == QV_IMPL_AUTO && net->fm);
    cudaEvent_t ev_compute = nullptr;
    if (!fused_path && cudaEventCreateWithFlags(&ev_compute, cudaEventDisableTiming) != cudaSuccess) {
        sh.fail("qv_stream_yuv: the GPU stage failed");
        rc = QV_ERR_CUDA;
    }
    for (int k = 0; k < nchunk && rc == QV_OK; ++k) {
        StreamSlot &S = slot[k % NS];
        {
            std::unique_lock<std::mutex> l(sh.mu);
            sh.cv.wait(l, [&] { return S.state == 1 || sh.failed; });
            if (sh.failed) break;
        }
        const size_t bytes = (size_t)S.c * fpx;
        e = cudaMemcpyAsync(S.d_in, S.h_in, bytes, cudaMemcpyHostToDevice, S.st);
        if (e == cudaSuccess && fo) e = cudaMemcpyAsync(S.d_ori, S.h_ori, bytes, cudaMemcpyHostToDevice, S.st);
        // the layered path's launches share the handle's activation scratch: keep them in order across the slots' streams
        if (e == cudaSuccess && !fused_path && k > 0) e = cudaStreamWaitEvent(S.st, ev_compute, 0);
        if (e == cudaSuccess) {
            rc = run_forward(net, S.d_in, S.d_out, S.c, net->H, net->W, S.st);
            if (rc == QV_OK && !fused_path) e = cudaEventRecord(ev_compute, S.st);
            if (rc == QV_OK && fo) {
                e = sse_accumulate(S.d_in, S.d_ori, bytes, d_sse + 0, S.st);
                if (e == cudaSuccess) e = sse_accumulate(S.d_out, S.d_ori, bytes, d_sse + 1, S.st);
            }
        }
        if (e == cudaSuccess && rc == QV_OK) e = cudaMemcpyAsync(S.h_out, S.d_out, bytes, cudaMemcpyDeviceToHost, S.st);
        if (e == cudaSuccess && rc == QV_OK) e = cudaEventRecord(S.done, S.st);
        if (e != cudaSuccess) { set_error("qv_stream_yuv: %s", cudaGetErrorString(e)); rc = QV_ERR_CUDA; }
        if (rc != QV_OK) { sh.fail("qv_stream_yuv: the GPU stage failed"); break; }
        { std::lock_guard<std::mutex> l(sh.mu); S.state = 2; }
        sh.cv.notify_all();
    }
    reader.join();
    writer.join();
    if (ev_compute) { for (int s = 0; s < NS; ++s) cudaStreamSynchronize(slot[s].st); cudaEventDestroy(ev_compute); }
    int64_t h_sse[2] = {0, 0};
    if (rc == QV_OK && !sh.failed) {
        for (int s = 0; s < NS; ++s) cudaStreamSynchronize(slot[s].st);
        if (cudaMemcpy(h_sse, d_sse, sizeof(h_sse), cudaMemcpyDeviceToHost) != cudaSuccess) rc = QV_ERR_CUDA;
    }
    cleanup();
    if (rc == QV_OK && sh.failed) { set_error("%s", sh.msg); rc = strstr(sh.msg, "GPU") ? QV_ERR_CUDA : QV_ERR_IO; }
    if (rc != QV_OK) return rc;
    if (sse_before) *sse_before = h_sse[0];
    if (sse_after) *sse_after = h_sse[1];
    return kernel_report(net);
}

void *qv_host_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    void *p = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) return p;
    cudaGetLastError();
    return malloc(bytes);
}

void qv_host_free(void *p)
{
    if (!p) return;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost) { cudaFreeHost(p); return; }
    cudaGetLastError();
    free(p);
}

int qv_device_buffers(qv_net *net, void **d_x, void **d_x_rec)
{
    if (!net) { set_error("qv_device_buffers: null handle"); return QV_ERR_ARG; }
    if (d_x) *d_x = net->d_x;
    if (d_x_rec) *d_x_rec = net->d_rec;
    return QV_OK;
}

int qv_sse_device(const uint8_t *d_a, const uint8_t *d_b, size_t n, int64_t *d_sse_accum, void *cuda_stream)
{
    if (!d_a || !d_b || !d_sse_accum) { set_error("qv_sse_device: null argument"); return QV_ERR_ARG; }
    if (n == 0) return QV_OK;
    QV_CUDA(sse_accumulate(d_a, d_b, n, d_sse_accum, (cudaStream_t)cuda_stream));
    return QV_OK;
}

int qv_set_impl(qv_net *net, int impl)
{
    if (!net || impl < QV_IMPL_AUTO || impl > QV_IMPL_FUSED) { set_error("qv_set_impl: bad argument"); return QV_ERR_ARG; }
    net->impl = impl;
    return QV_OK;
}

int qv_get_impl(const qv_net *net)
{
    if (!net) return QV_ERR_ARG;
    if (net->impl != QV_IMPL_AUTO) return net->impl;
    return net->fm ? QV_IMPL_FUSED : QV_IMPL_LAYERED;
}

long long qv_launch_count(const qv_net *net) { return net ? net->launches : 0; }

int qv_get_activation(qv_net *net, int which, int8_t *host_out)
{
    if (!net || !host_out || which < 1 || which > 3) { set_error("qv_get_activation: bad argument"); return QV_ERR_ARG; }
    if (!net->d_a1) { set_error("qv_get_activation: the layered path has not run on this handle"); return QV_ERR_STATE; }
    int rc = set_device(net);
    if (rc) return rc;
    const int C = which == 1 ? 64 : 48;
    const int8_t *src = which == 1 ? net->d_a1 : (which == 2 ? net->d_a2 : net->d_a3);
    const size_t HW = (size_t)net->H * net->W;
    int8_t *tmp = nullptr;
    QV_CUDA(cudaMalloc(&tmp, HW * C));
    cudaError_t e = nhwc_to_planar(src, tmp, C, HW, net->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, tmp, HW * C, cudaMemcpyDeviceToHost, net->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(net->st);
    cudaFree(tmp);
    QV_CUDA(e);
    return QV_OK;
}

}  // extern "C"
