// File formats either side of the QVRCNN hot path: static model files (NCHW_VECT_C and HWCN
// flavours), quant-param files (pickle protocol <= 4 subset and raw 6x6 doubles), YUV 4:2:0
// luma I/O and the PSNR report.  Host-only C++; no CUDA here.
//
// Reference behaviour mirrored (paths relative to the reference repository):
//   CovLayer::load_static_para        inference/cnn.cu:90-112
//   qvrcnn::load_static_para          inference/qvrcnn.cu:47-63
//   HWCN2NCHW_VECT_C_CPU              inference/mat.cu:97-119
//   model_qfp_HWCN2NCHW_VECT_C        inference/qvrcnn.cu:535-585
//   quantNsave                        training/quantization.py:66-98
//   vrcnn_data::{read_data,read_frame,psnr,save_recon_as}   inference/yuv_data.cpp:15-66,87-128
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "qv_internal.h"

namespace qv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

static size_t vect_c_wsize(int l)
{
    const LayerShape &s = kLayers[l];
    return (size_t)s.k * s.k * ((s.cin + 3) / 4) * 4 * s.cout;   // inference/cnn.cu:24
}
size_t model_file_size_vect_c()
{
    size_t n = 0;
    for (int l = 0; l < QV_NLAYER; ++l) n += vect_c_wsize(l) + 4u * kLayers[l].cout + 12u;
    return n;
}
size_t model_file_size_hwcn()
{
    size_t n = 0;
    for (int l = 0; l < QV_NLAYER; ++l) {
        const LayerShape &s = kLayers[l];
        n += (size_t)s.k * s.k * s.cin * s.cout + 4u * s.cout + 12u;
    }
    return n;
}

static void read_tail(const uint8_t *&p, LayerHost &L, int cout)
{
    L.b.resize(cout);
    memcpy(L.b.data(), p, 4u * cout); p += 4u * cout;       // int32 bias[K]   cnn.cu:100
    memcpy(&L.blu, p, 4);   p += 4;                          // cnn.cu:101
    memcpy(&L.mul, p, 4);   p += 4;                          // cnn.cu:102
    memcpy(&L.shift, p, 4); p += 4;                          // cnn.cu:103
    L.have_w = L.have_q = true;
}

int parse_model_vect_c(const uint8_t *buf, size_t len, ModelHost &m)
{
    if (len != model_file_size_vect_c()) {
        set_error("static model image is %zu bytes, expected %zu", len, model_file_size_vect_c());
        return QV_ERR_IO;
    }
    const uint8_t *p = buf;
    for (int l = 0; l < QV_NLAYER; ++l) {
        const LayerShape &s = kLayers[l];
        const int c4 = (s.cin + 3) / 4, R = s.k;
        LayerHost &L = m.L[l];
        L.w.assign((size_t)s.cout * s.cin * R * R, 0);
        // w[K][ceil(C/4)][R][S][4]: lane = c & 3, group = c >> 2   (inference/mat.cu:109-117)
        for (int k = 0; k < s.cout; ++k)
            for (int c = 0; c < s.cin; ++c)
                for (int r = 0; r < R; ++r)
                    for (int q = 0; q < R; ++q)
                        L.w[(((size_t)k * s.cin + c) * R + r) * R + q] =
                            (int8_t)p[(size_t)k * (R * R * c4 * 4) + (size_t)(c >> 2) * (R * R * 4) +
                                      (size_t)r * (R * 4) + (size_t)q * 4 + (c & 3)];
        p += vect_c_wsize(l);
        read_tail(p, L, s.cout);
    }
    return QV_OK;
}

int parse_model_hwcn(const uint8_t *buf, size_t len, ModelHost &m)
{
    if (len != model_file_size_hwcn()) {
        set_error("HWCN model image is %zu bytes, expected %zu", len, model_file_size_hwcn());
        return QV_ERR_IO;
    }
    const uint8_t *p = buf;
    for (int l = 0; l < QV_NLAYER; ++l) {
        const LayerShape &s = kLayers[l];
        const int R = s.k;
        LayerHost &L = m.L[l];
        L.w.assign((size_t)s.cout * s.cin * R * R, 0);
        // TF order w[R][S][C][K]   (inference/qvrcnn.cu:542-545, mat.cu:116)
        for (int r = 0; r < R; ++r)
            for (int q = 0; q < R; ++q)
                for (int c = 0; c < s.cin; ++c)
                    for (int k = 0; k < s.cout; ++k)
                        L.w[(((size_t)k * s.cin + c) * R + r) * R + q] =
                            (int8_t)p[(((size_t)r * R + q) * s.cin + c) * s.cout + k];
        p += (size_t)R * R * s.cin * s.cout;
        read_tail(p, L, s.cout);
    }
    return QV_OK;
}

std::vector<uint8_t> serialize_model_vect_c(const ModelHost &m)
{
    std::vector<uint8_t> out(model_file_size_vect_c(), 0);
    uint8_t *p = out.data();
    for (int l = 0; l < QV_NLAYER; ++l) {
        const LayerShape &s = kLayers[l];
        const int c4 = (s.cin + 3) / 4, R = s.k;
        const LayerHost &L = m.L[l];
        for (int k = 0; k < s.cout; ++k)
            for (int c = 0; c < s.cin; ++c)
                for (int r = 0; r < R; ++r)
                    for (int q = 0; q < R; ++q)
                        p[(size_t)k * (R * R * c4 * 4) + (size_t)(c >> 2) * (R * R * 4) + (size_t)r * (R * 4) +
                          (size_t)q * 4 + (c & 3)] = (uint8_t)L.w[(((size_t)k * s.cin + c) * R + r) * R + q];
        p += vect_c_wsize(l);
        memcpy(p, L.b.data(), 4u * s.cout); p += 4u * s.cout;
        memcpy(p, &L.blu, 4);   p += 4;
        memcpy(p, &L.mul, 4);   p += 4;
        memcpy(p, &L.shift, 4); p += 4;
    }
    return out;
}

int read_file(const char *path, std::vector<uint8_t> &out)
{
    FILE *fp = path ? fopen(path, "rb") : nullptr;
    if (!fp) {
        set_error("cannot open file '%s'", path ? path : "(null)");
        return QV_ERR_IO;
    }
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), fp);
    fclose(fp);
    if (got != out.size()) {
        set_error("short read on '%s'", path);
        return QV_ERR_IO;
    }
    return QV_OK;
}

int check_fp32_exact_envelope(const ModelHost &m)
{
    for (int l = 0; l < QV_NLAYER; ++l) {
        const LayerShape &s = kLayers[l];
        const LayerHost &L = m.L[l];
        if (!L.have_w) continue;
        const size_t per_k = (size_t)s.cin * s.k * s.k;
        for (int k = 0; k < s.cout; ++k) {
            long long sum = 0;
            for (size_t i = 0; i < per_k; ++i) sum += std::abs((int)L.w[(size_t)k * per_k + i]);
            long long bound = 128 * sum + std::llabs((long long)L.b[k]);
            if (bound >= (1ll << 24)) {
                set_error("layer %d channel %d: 128*sum|w|+|b| = %lld >= 2^24; outside the exact-integer "
                          "envelope of the reference's fp32 accumulators", l, k, bound);
                return QV_ERR_RANGE;
            }
        }
    }
    return QV_OK;
}

// ---------------------------------------------------------------------------------------
// Minimal unpickler: just enough of protocols 2-4 to evaluate what pickle.dump() emits for a
// list of lists of {python int, python float, numpy float64/int64 scalar}
// (training/quantization.py:90-91).  numpy scalars arrive as
//   REDUCE(GLOBAL numpy.core.multiarray.scalar, (REDUCE(GLOBAL numpy.dtype, ('f8',..)) + BUILD, bytes))
// ---------------------------------------------------------------------------------------
namespace {
struct PV;
using PVp = std::shared_ptr<PV>;
struct PV {
    enum Kind { NONE, INT, FLOAT, BYTES, STR, LIST, TUPLE, GLOBAL, DTYPE, MARK, BOOL } kind = NONE;
    long long i = 0;
    double f = 0;
    std::string s;              // BYTES / STR / GLOBAL ("module name") / DTYPE (type string e.g. "f8")
    std::vector<PVp> items;     // LIST / TUPLE
    char byteorder = '<';       // DTYPE
};
PVp mk(PV::Kind k) { auto p = std::make_shared<PV>(); p->kind = k; return p; }

struct Unpickler {
    const uint8_t *p, *end;
    std::vector<PVp> st;
    std::map<long long, PVp> memo;
    long long memo_next = 0;
    bool fail = false;

    bool need(size_t n) { if ((size_t)(end - p) < n) { fail = true; return false; } return true; }
    uint32_t u32() { uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
    PVp pop() { if (st.empty()) { fail = true; return mk(PV::NONE); } PVp v = st.back(); st.pop_back(); return v; }
    std::vector<PVp> pop_mark()
    {
        std::vector<PVp> out;
        while (!st.empty() && st.back()->kind != PV::MARK) { out.insert(out.begin(), st.back()); st.pop_back(); }
        if (st.empty()) fail = true; else st.pop_back();
        return out;
    }
    std::string line()
    {
        std::string s;
        while (p < end && *p != '\n') s.push_back((char)*p++);
        if (p < end) ++p; else fail = true;
        return s;
    }
    PVp reduce(const PVp &callable, const PVp &args)
    {
        if (callable->kind != PV::GLOBAL || args->kind != PV::TUPLE) { fail = true; return mk(PV::NONE); }
        const std::string &g = callable->s;
        if (g == "numpy dtype") {
            PVp d = mk(PV::DTYPE);
            if (!args->items.empty() && args->items[0]->kind == PV::STR) d->s = args->items[0]->s;
            return d;
        }
        if (g == "_codecs encode") {
            // protocol <= 2 spells bytes as _codecs.encode(<utf-8 text>, 'latin1')
            if (args->items.empty() || args->items[0]->kind != PV::STR) { fail = true; return mk(PV::NONE); }
            PVp b = mk(PV::BYTES);
            const std::string &u8 = args->items[0]->s;
            for (size_t i = 0; i < u8.size(); ++i) {
                unsigned char c = (unsigned char)u8[i];
                if (c < 0x80) b->s.push_back((char)c);
                else if ((c & 0xE0) == 0xC0 && i + 1 < u8.size()) {
                    b->s.push_back((char)(((c & 0x1F) << 6) | ((unsigned char)u8[i + 1] & 0x3F)));
                    ++i;
                } else { fail = true; return mk(PV::NONE); }
            }
            return b;
        }
        if (g == "numpy.core.multiarray scalar" || g == "numpy._core.multiarray scalar") {
            if (args->items.size() != 2 || args->items[0]->kind != PV::DTYPE || args->items[1]->kind != PV::BYTES) {
                fail = true; return mk(PV::NONE);
            }
            const PV &d = *args->items[0];
            const std::string &raw = args->items[1]->s;
            uint8_t b[8] = {0};
            if (raw.size() > 8) { fail = true; return mk(PV::NONE); }
            for (size_t i = 0; i < raw.size(); ++i)
                b[d.byteorder == '>' ? raw.size() - 1 - i : i] = (uint8_t)raw[i];
            if (d.s == "f8") { PVp v = mk(PV::FLOAT); memcpy(&v->f, b, 8); return v; }
            if (d.s == "f4") { float t; memcpy(&t, b, 4); PVp v = mk(PV::FLOAT); v->f = t; return v; }
            if (d.s == "i8") { PVp v = mk(PV::INT); memcpy(&v->i, b, 8); return v; }
            if (d.s == "i4") { int32_t t; memcpy(&t, b, 4); PVp v = mk(PV::INT); v->i = t; return v; }
            fail = true;
            return mk(PV::NONE);
        }
        fail = true;
        return mk(PV::NONE);
    }
    PVp run()
    {
        while (p < end && !fail) {
            uint8_t op = *p++;
            switch (op) {
            case 0x80: if (need(1)) ++p; break;                                   // PROTO
            case 0x95: if (need(8)) p += 8; break;                                // FRAME
            case ']': st.push_back(mk(PV::LIST)); break;                          // EMPTY_LIST
            case ')': st.push_back(mk(PV::TUPLE)); break;                         // EMPTY_TUPLE
            case '(': st.push_back(mk(PV::MARK)); break;                          // MARK
            case 'N': st.push_back(mk(PV::NONE)); break;                          // NONE
            case 0x88: case 0x89: { PVp v = mk(PV::BOOL); v->i = op == 0x88; st.push_back(v); break; }
            case 'q': if (need(1)) { if (st.empty()) fail = true; else memo[*p] = st.back(); ++p; } break;  // BINPUT
            case 'r': if (need(4)) { uint32_t k = u32(); if (st.empty()) fail = true; else memo[k] = st.back(); } break;
            case 0x94: if (st.empty()) fail = true; else memo[memo_next++] = st.back(); break;               // MEMOIZE
            case 'h': if (need(1)) { auto it = memo.find(*p++); if (it == memo.end()) fail = true; else st.push_back(it->second); } break;
            case 'j': if (need(4)) { auto it = memo.find(u32()); if (it == memo.end()) fail = true; else st.push_back(it->second); } break;
            case 'c': { PVp g = mk(PV::GLOBAL); std::string m = line(); std::string n = line(); g->s = m + " " + n; st.push_back(g); break; }
            case 0x93: { PVp n = pop(), m = pop(); PVp g = mk(PV::GLOBAL); g->s = m->s + " " + n->s; st.push_back(g); break; }  // STACK_GLOBAL
            case 'X': if (need(4)) { uint32_t n = u32(); if (need(n)) { PVp v = mk(PV::STR); v->s.assign((const char *)p, n); p += n; st.push_back(v); } } break;
            case 0x8c: if (need(1)) { uint32_t n = *p++; if (need(n)) { PVp v = mk(PV::STR); v->s.assign((const char *)p, n); p += n; st.push_back(v); } } break;
            case 'C': if (need(1)) { uint32_t n = *p++; if (need(n)) { PVp v = mk(PV::BYTES); v->s.assign((const char *)p, n); p += n; st.push_back(v); } } break;
            case 'B': if (need(4)) { uint32_t n = u32(); if (need(n)) { PVp v = mk(PV::BYTES); v->s.assign((const char *)p, n); p += n; st.push_back(v); } } break;
            case 'K': if (need(1)) { PVp v = mk(PV::INT); v->i = *p++; st.push_back(v); } break;                     // BININT1
            case 'M': if (need(2)) { PVp v = mk(PV::INT); v->i = p[0] | (p[1] << 8); p += 2; st.push_back(v); } break;  // BININT2
            case 'J': if (need(4)) { PVp v = mk(PV::INT); v->i = (int32_t)u32(); st.push_back(v); } break;           // BININT
            case 0x8a:                                                            // LONG1
                if (need(1)) {
                    uint32_t n = *p++;
                    if (need(n) && n <= 8) {
                        long long v = 0;
                        for (uint32_t i = 0; i < n; ++i) v |= (long long)p[i] << (8 * i);
                        if (n && n < 8 && (p[n - 1] & 0x80)) v -= 1ll << (8 * n);
                        p += n;
                        PVp q = mk(PV::INT);
                        q->i = v;
                        st.push_back(q);
                    } else {
                        fail = true;
                    }
                }
                break;
            case 'G': if (need(8)) { uint8_t b[8]; for (int i = 0; i < 8; ++i) b[i] = p[7 - i]; p += 8; PVp v = mk(PV::FLOAT); memcpy(&v->f, b, 8); st.push_back(v); } break;  // BINFLOAT (big endian)
            case 0x85: { PVp t = mk(PV::TUPLE); t->items = {pop()}; st.push_back(t); break; }
            case 0x86: { PVp b = pop(), a = pop(); PVp t = mk(PV::TUPLE); t->items = {a, b}; st.push_back(t); break; }
            case 0x87: { PVp c = pop(), b = pop(), a = pop(); PVp t = mk(PV::TUPLE); t->items = {a, b, c}; st.push_back(t); break; }
            case 't': { PVp t = mk(PV::TUPLE); t->items = pop_mark(); st.push_back(t); break; }
            case 'R': { PVp args = pop(), callable = pop(); st.push_back(reduce(callable, args)); break; }
            case 'b': {                                                           // BUILD: dtype.__setstate__
                PVp state = pop();
                if (st.empty()) { fail = true; break; }
                PVp obj = st.back();
                if (obj->kind == PV::DTYPE && state->kind == PV::TUPLE && state->items.size() > 1 &&
                    state->items[1]->kind == PV::STR && !state->items[1]->s.empty())
                    obj->byteorder = state->items[1]->s[0];
                break;
            }
            case 'a': { PVp v = pop(); if (st.empty() || st.back()->kind != PV::LIST) fail = true; else st.back()->items.push_back(v); break; }
            case 'e': { auto items = pop_mark(); if (st.empty() || st.back()->kind != PV::LIST) fail = true; else for (auto &v : items) st.back()->items.push_back(v); break; }
            case '.': return pop();
            default: fail = true; break;
            }
        }
        fail = true;
        return mk(PV::NONE);
    }
};

bool pv_number(const PVp &v, double &out)
{
    if (v->kind == PV::INT || v->kind == PV::BOOL) { out = (double)v->i; return true; }
    if (v->kind == PV::FLOAT) { out = v->f; return true; }
    return false;
}
}  // namespace

// rows = 6 x [stepw, ratio, blu_adj, blu_q, mul, shift]; inference consumes columns 3..5.
static int rows_to_q(const double rows[6][6], int32_t *out18)
{
    for (int l = 0; l < 6; ++l)
        for (int j = 0; j < 3; ++j) {
            double v = rows[l][3 + j];
            if (!(std::fabs(v) < 2147483647.0) || v != std::floor(v)) {
                set_error("quant params: row %d column %d = %g is not an int32", l, 3 + j, v);
                return QV_ERR_IO;
            }
            out18[3 * l + j] = (int32_t)v;
        }
    for (int l = 0; l < 6; ++l) {
        int32_t mul = out18[3 * l + 1], sh = out18[3 * l + 2];
        if (mul <= 0 || sh < 1 || sh > 30) {
            set_error("quant params: layer %d has mul=%d shift=%d (need mul>0, 1<=shift<=30)", l, mul, sh);
            return QV_ERR_IO;
        }
    }
    return QV_OK;
}

int parse_quant_params(const uint8_t *buf, size_t len, int32_t *out18)
{
    double rows[6][6];
    if (len == 6 * 6 * sizeof(double)) {            // quant_params_cpp_<QP>.data  (quantization.py:93-96)
        memcpy(rows, buf, sizeof(rows));
        return rows_to_q(rows, out18);
    }
    Unpickler u{buf, buf + len, {}, {}, 0, false};
    PVp root = u.run();
    if (u.fail || root->kind != PV::LIST || root->items.size() != 6) {
        set_error("quant params: not a 288-byte raw table nor a pickle of 6 rows");
        return QV_ERR_IO;
    }
    for (int l = 0; l < 6; ++l) {
        const PVp &r = root->items[l];
        if ((r->kind != PV::LIST && r->kind != PV::TUPLE) || r->items.size() != 6) {
            set_error("quant params: row %d does not have 6 entries", l);
            return QV_ERR_IO;
        }
        for (int j = 0; j < 6; ++j)
            if (!pv_number(r->items[j], rows[l][j])) {
                set_error("quant params: row %d entry %d is not a number", l, j);
                return QV_ERR_IO;
            }
    }
    return rows_to_q(rows, out18);
}

}  // namespace qv

// ---------------------------------------------------------------------------------------
// C ABI: pure-host entry points
// ---------------------------------------------------------------------------------------
using namespace qv;

extern "C" {

const char *qv_last_error(void) { return get_error(); }
const char *qv_version(void) { return "qvrcnn-b200 0.1 (sm_100a)"; }

int qv_read_quant_params(const char *filename, int32_t *out18)
{
    if (!out18) { set_error("qv_read_quant_params: null output"); return QV_ERR_ARG; }
    std::vector<uint8_t> buf;
    int rc = read_file(filename, buf);
    if (rc) return rc;
    return parse_quant_params(buf.data(), buf.size(), out18);
}

int qv_convert_model_hwcn_to_vect_c(const char *file_in, const char *file_out)
{
    std::vector<uint8_t> buf;
    int rc = read_file(file_in, buf);
    if (rc) return rc;
    ModelHost m;
    rc = parse_model_hwcn(buf.data(), buf.size(), m);
    if (rc) return rc;
    std::vector<uint8_t> out = serialize_model_vect_c(m);
    FILE *fp = file_out ? fopen(file_out, "wb") : nullptr;
    if (!fp) { set_error("failed to open file %s", file_out ? file_out : "(null)"); return QV_ERR_IO; }
    size_t put = fwrite(out.data(), 1, out.size(), fp);
    fclose(fp);
    if (put != out.size()) { set_error("short write on %s", file_out); return QV_ERR_IO; }
    return QV_OK;
}

// ---------------------------------------------------------------------------------------
// Quant-parameter solver (training/quantization.py:5-64), re-derived in C++ with the same IEEE-754
// double operation order so that the 36 numbers match the Python module bit for bit.
// ---------------------------------------------------------------------------------------
namespace {
// python round(): half to even.  nearbyint under the default rounding mode does exactly that.
double py_round(double v) { return std::nearbyint(v); }

// mul_shift(max_u)   quantization.py:5-14: smallest shift i in [1,27] with 127 < max_u*round(127.5*2^i/max_u)/2^i < 127.5
void solve_mul_shift(double max_u, double &mul, int &shift)
{
    mul = 0;
    int i = 1;
    for (; i < 28; ++i) {
        const double max_int = 127.5 * std::ldexp(1.0, i);
        if (max_int > max_u) {
            mul = py_round(max_int / max_u);
            const double temp = max_u * mul / std::ldexp(1.0, i);
            if (temp > 127 && temp < 127.5) { shift = i; return; }
        }
    }
    shift = 27;                       // python: `return mul,i` after the loop ends with i == 27
}
// mul_shift_f(ratio)   quantization.py:15-24
void solve_mul_shift_f(double ratio, double &mul, int &shift)
{
    mul = 0;
    for (int i = 10; i < 28; ++i) {
        const double max_int = std::ldexp(1.0, i);
        if (max_int > ratio) {
            const double temp = max_int / ratio;
            mul = py_round(temp);
            if (std::fabs(max_int / mul - ratio) < 0.02 * ratio) { shift = i; return; }
        }
    }
    shift = 27;
}
void set_row(double *row, double stepw, double ratio, double blu_adj, double blu_q, double mul, int shift)
{
    row[0] = stepw; row[1] = ratio; row[2] = blu_adj; row[3] = blu_q; row[4] = mul; row[5] = shift;
}
// quant_qfp_layer   quantization.py:25-31
void solve_layer(double ratio, double stepw, double blu, double *row)
{
    double blu_q = py_round(blu * ratio / stepw), mul;
    int sh;
    solve_mul_shift(blu_q, mul, sh);
    const double blu_adj = 127 * std::ldexp(1.0, sh) / mul * stepw / ratio;
    blu_q = py_round(blu_adj * ratio / stepw);
    set_row(row, stepw, ratio, blu_adj, blu_q, mul, sh);
}
// quant_qfp_concat   quantization.py:32-49
void solve_concat(double ratio, double stepw1, double blu1, double stepw2, double blu2, double *row1, double *row2)
{
    if (blu1 < blu2) blu1 = blu2; else blu2 = blu1;
    const double blu_q1 = py_round(blu1 * ratio / stepw1), blu_q2 = py_round(blu2 * ratio / stepw2);
    double mul1, mul2;
    int sh1, sh2;
    solve_mul_shift(blu_q1, mul1, sh1);
    solve_mul_shift(blu_q2, mul2, sh2);
    if (mul1 / stepw1 / std::ldexp(1.0, sh1) > mul2 / stepw2 / std::ldexp(1.0, sh2))
        stepw1 = stepw2 * std::ldexp(1.0, sh2) / mul2 * mul1 / std::ldexp(1.0, sh1);
    else
        stepw2 = stepw1 * std::ldexp(1.0, sh1) / mul1 * mul2 / std::ldexp(1.0, sh2);
    const double blu1_adj = 127 * std::ldexp(1.0, sh1) / mul1 * stepw1 / ratio;
    const double blu2_adj = 127 * std::ldexp(1.0, sh2) / mul2 * stepw2 / ratio;
    set_row(row1, stepw1, ratio, blu1_adj, blu_q1, mul1, sh1);
    set_row(row2, stepw2, ratio, blu2_adj, blu_q2, mul2, sh2);
}
}  // namespace

int qv_solve_quant_params(const double *stepw_in, const double *blu_in, double *rows)
{
    if (!stepw_in || !blu_in || !rows) { set_error("qv_solve_quant_params: null argument"); return QV_ERR_ARG; }
    for (int l = 0; l < 6; ++l)
        if (!(stepw_in[l] > 0) || (l < 5 && !(blu_in[l] > 0))) { set_error("qv_solve_quant_params: stepw and blu must be positive (layer %d)", l); return QV_ERR_ARG; }
    double ratio = 255;                                                       // quantization.py:56
    solve_layer(ratio, stepw_in[0], blu_in[0], rows);
    ratio = ratio / rows[0] * rows[4] / std::ldexp(1.0, (int)rows[5]);        // :58
    solve_concat(ratio, stepw_in[1], blu_in[1], stepw_in[2], blu_in[2], rows + 6, rows + 12);
    ratio = ratio / rows[6] * rows[6 + 4] / std::ldexp(1.0, (int)rows[6 + 5]); // :60
    solve_concat(ratio, stepw_in[3], blu_in[3], stepw_in[4], blu_in[4], rows + 18, rows + 24);
    ratio = ratio / rows[18] * rows[18 + 4] / std::ldexp(1.0, (int)rows[18 + 5]);   // :62
    {                                                                         // quant_qfp_last :50-53
        double mul;
        int sh;
        solve_mul_shift_f(ratio / 255 / stepw_in[5], mul, sh);
        const double stepw_adj = ratio * mul / std::ldexp(1.0, sh) / 255;
        set_row(rows + 30, stepw_adj, ratio, 0, 0, mul, sh);
    }
    return QV_OK;
}

int qv_write_quant_params_cpp(const char *filename, const double *rows)
{
    if (!rows) { set_error("qv_write_quant_params_cpp: null argument"); return QV_ERR_ARG; }
    FILE *fp = filename ? fopen(filename, "wb") : nullptr;
    if (!fp) { set_error("cannot open %s for writing", filename ? filename : "(null)"); return QV_ERR_IO; }
    const size_t put = fwrite(rows, sizeof(double), 36, fp);
    fclose(fp);
    if (put != 36) { set_error("short write on %s", filename); return QV_ERR_IO; }
    return QV_OK;
}

int qv_quantize_layer(const float *w, size_t n_w, const float *b, size_t n_b, double stepw, double ratio, int8_t *w_q, int32_t *b_q)
{
    if (!w || !b || !w_q || !b_q || !(stepw > 0)) { set_error("qv_quantize_layer: bad argument"); return QV_ERR_ARG; }
    for (size_t i = 0; i < n_w; ++i) {
        // numpy: np.clip(np.around(wf / stepw), -128, 127); wf is float32, stepw a python float -> float64 division
        double q = py_round((double)w[i] / stepw);
        q = q < -128 ? -128 : (q > 127 ? 127 : q);
        w_q[i] = (int8_t)q;
    }
    for (size_t i = 0; i < n_b; ++i) {
        const double q = py_round((double)b[i] * ratio / stepw);
        if (!(std::fabs(q) < 2147483647.0)) { set_error("qv_quantize_layer: bias %zu does not fit int32", i); return QV_ERR_RANGE; }
        b_q[i] = (int32_t)q;
    }
    return QV_OK;
}

int qv_yuv_read_luma(const char *filename, int frames, int height, int width, uint8_t *out)
{
    if (!out || frames < 0 || height <= 0 || width <= 0) { set_error("qv_yuv_read_luma: bad argument"); return QV_ERR_ARG; }
    FILE *fp = filename ? fopen(filename, "rb") : nullptr;
    if (!fp) { set_error("open file failed: %s", filename ? filename : "(null)"); return QV_ERR_IO; }
    const size_t hw = (size_t)height * width;
    for (int i = 0; i < frames; ++i) {                        // inference/yuv_data.cpp:32-38
        if (fread(out + (size_t)i * hw, 1, hw, fp) != hw) {
            fclose(fp);
            set_error("%s: short read at frame %d", filename, i);
            return QV_ERR_IO;
        }
        fseek(fp, (long)(hw / 2), SEEK_CUR);                  // skip U and V
    }
    fclose(fp);
    return QV_OK;
}

int qv_yuv_read_frame(const char *filename, int n, int height, int width, uint8_t *out)
{
    if (!out || n < 0 || height <= 0 || width <= 0) { set_error("qv_yuv_read_frame: bad argument"); return QV_ERR_ARG; }
    FILE *fp = filename ? fopen(filename, "rb") : nullptr;
    if (!fp) { set_error("open file failed: %s", filename ? filename : "(null)"); return QV_ERR_IO; }
    const size_t hw = (size_t)height * width;
    fseeko(fp, (off_t)(hw * (size_t)n * 3 / 2), SEEK_CUR);    // inference/yuv_data.cpp:59
    size_t got = fread(out, 1, hw, fp);
    fclose(fp);
    if (got != hw) { set_error("%s: short read at frame %d", filename, n); return QV_ERR_IO; }
    return QV_OK;
}

int qv_yuv_write_recon(const char *filename, const uint8_t *luma, int frames, int height, int width)
{
    if (!luma || frames < 0 || height <= 0 || width <= 0) { set_error("qv_yuv_write_recon: bad argument"); return QV_ERR_ARG; }
    FILE *fp = filename ? fopen(filename, "wb") : nullptr;
    if (!fp) { set_error("write file failed: %s", filename ? filename : "(null)"); return QV_ERR_IO; }
    const size_t hw = (size_t)height * width;
    std::vector<uint8_t> uv(hw / 2, 0);                       // inference/yuv_data.cpp:119-120
    for (int i = 0; i < frames; ++i) {
        fwrite(luma + (size_t)i * hw, 1, hw, fp);
        fwrite(uv.data(), 1, uv.size(), fp);
    }
    fclose(fp);
    return QV_OK;
}

double qv_psnr(const uint8_t *data, const uint8_t *ori, size_t n, int64_t *sse_out)
{
    double mse = 0;                                            // inference/yuv_data.cpp:90-96
    int64_t sse = 0;
    for (size_t i = 0; i < n; ++i) {
        int d = (int)data[i] - (int)ori[i];
        mse += d * d;
        sse += (int64_t)d * d;
    }
    if (sse_out) *sse_out = sse;
    mse /= (double)n;
    return 10 * std::log10(65025.0 / mse);
}

double qv_psnr_from_sse(int64_t sse, size_t n)
{
    double mse = (double)sse;     // exact: every partial sum of the reference's loop is an integer < 2^53
    mse /= (double)n;
    return 10 * std::log10(65025.0 / mse);
}

}  // extern "C"
