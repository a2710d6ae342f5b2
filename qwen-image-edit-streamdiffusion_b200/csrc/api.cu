// libqie host side: error plumbing, TMA descriptor builder, scheduler/rope host helpers,
// the model handle and the forward orchestration (one C call enqueues the whole 60-block step).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace qie {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return QIE_ECUDA;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    });
    return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_rows, uint32_t box_cols, int elt_bytes) {
    PFN_encodeTiled enc = get_encode();
    QIE_REQUIRE(enc, QIE_ECUDA, "cuTensorMapEncodeTiled driver entry point unavailable");
    QIE_REQUIRE(((uintptr_t)base & 15) == 0 && row_stride_bytes % 16 == 0, QIE_EINVAL,
                "TMA operand must be 16-byte aligned (base %p, row stride %llu)", base,
                (unsigned long long)row_stride_bytes);
    QIE_REQUIRE(box_cols * elt_bytes == 128 && box_rows <= 256, QIE_EINVAL, "TMA box must be 128 B wide, <=256 rows");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    QIE_REQUIRE(r == CUDA_SUCCESS, QIE_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return QIE_OK;
}

int unpack_rows(const void* x, void* out, int batch, int n, int n_pad, int C, cudaStream_t st);
int attn_fwd_peers(const void* qkv_gathered, void* const* peer_out_dev, const qie_peers* pr, const int* tile_valid_dev,
                   int heads_local, int out_ld, int variant, void* stream);
int peer_bcast_rows(const void* src, void* const* peer_vel_dev, const qie_peers* pr, int img_rows, int img_offset, int C,
                    cudaStream_t st);
int peer_barrier_launch(const qie_peers* pr, cudaStream_t st);
int peer_bcast_span(const void* mine, void* const* peer_tab_dev, const qie_peers* pr, long long bstride_bytes, long long off_bytes,
                    long long len_bytes, cudaStream_t st);

// ---- per-(device, stream) kernel scratch (common.cuh) ----
namespace {
struct ScratchPool {
    bool ready = false;
    StreamScratch set[QIE_SCRATCH_SETS];
    cudaStream_t owner[QIE_SCRATCH_SETS];
    int used = 0;
};
ScratchPool g_pool[64];
std::mutex g_pool_mu;
}  // namespace

int stream_scratch_reserve() {
    int dev = 0;
    QIE_CUDA_OK(cudaGetDevice(&dev));
    QIE_REQUIRE(dev >= 0 && dev < 64, QIE_EINVAL, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    ScratchPool& p = g_pool[dev];
    if (p.ready) return QIE_OK;
    for (int i = 0; i < QIE_SCRATCH_SETS; ++i) {
        QIE_CUDA_OK(cudaMalloc(&p.set[i].ln_counters, 2 * sizeof(int)));
        QIE_CUDA_OK(cudaMemset(p.set[i].ln_counters, 0, 2 * sizeof(int)));
        QIE_CUDA_OK(cudaMalloc(&p.set[i].split_scratch, (size_t)SPLIT_SCRATCH_TILES * 256 * 256 * sizeof(float)));
        QIE_CUDA_OK(cudaMalloc(&p.set[i].split_tickets, SPLIT_SCRATCH_TILES * 8 * sizeof(int)));
        QIE_CUDA_OK(cudaMemset(p.set[i].split_tickets, 0, SPLIT_SCRATCH_TILES * 8 * sizeof(int)));
    }
    p.ready = true;
    return QIE_OK;
}

int stream_scratch(cudaStream_t st, StreamScratch* out) {
    int dev = 0;
    QIE_CUDA_OK(cudaGetDevice(&dev));
    QIE_REQUIRE(dev >= 0 && dev < 64, QIE_EINVAL, "device index %d out of range", dev);
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        ScratchPool& p = g_pool[dev];
        if (p.ready) {
            for (int i = 0; i < p.used; ++i)
                if (p.owner[i] == st) { *out = p.set[i]; return QIE_OK; }
            QIE_REQUIRE(p.used < QIE_SCRATCH_SETS, QIE_ESTATE,
                        "libqie kernels were launched on more than %d distinct streams of device %d (one scratch set per stream)",
                        QIE_SCRATCH_SETS, dev);
            p.owner[p.used] = st;
            *out = p.set[p.used++];
            return QIE_OK;
        }
    }
    // per-kernel entry points used without a handle on this device (unit tests): allocate the pool now (not capturable)
    int rc = stream_scratch_reserve();
    if (rc) return rc;
    return stream_scratch(st, out);
}

// -------------------------------------------------------------------------------------------
// model handle
// -------------------------------------------------------------------------------------------
}  // namespace qie

struct qie_handle {
    qie_model_cfg cfg;
    int device;
    int D;
    bool has_weights;
    int precision;   // 0 bf16, 1 fp8
    int fuse_qk;     // 1: RMSNorm+RoPE in the QKV GEMM epilogue, 0: standalone kernel
    int attn_variant;
    int attn_bounded = 1;               // option 3: bounded-score attention in the blocks whose bound allows it
    std::vector<float> score_bound;     // per block: bound on |q.k| * softmax_scale * log2(e), from the QK-RMSNorm weights
    qie_weights w;
    std::vector<qie_block_weights> blocks;
    // library-owned small device buffers
    float* d_rope;          // [rope_rows, 64, 2]: the table selected by the last BEGIN / QKV phase (owned by rope_cache)
    struct RopeEntry { std::vector<int> key; float* d; };
    std::vector<RopeEntry> rope_cache;   // one table per (img_shapes, rows, sequence-parallel shard); small, oldest evicted
    float* d_small;         // tproj [8,256] | t1 [8,D] | temb [8,D] | mod [L,2? ...] see offsets
    size_t small_bytes;
    // exact caches: per-timestep modulation tables and per-prompt text-stream embeddings
    float* d_sched_mod = nullptr;   // [n_sched][L*12*D]
    float* d_sched_fin = nullptr;   // [n_sched][2D]
    int n_sched = 0;
    float* d_prompt[4] = {nullptr, nullptr, nullptr, nullptr};   // fp32 [txt_pad, D] (pad rows zero)
    int prompt_rows[4] = {0, 0, 0, 0};
    int sel_sched[8] = {-1, -1, -1, -1, -1, -1, -1, -1};         // per batch row, -1 = compute
    int sel_prompt = -1;
    // fused Ulysses exchange (qie_set_peers): host copy + device table [0..7] qkv_gather, [8..15] attn_out, [16..23] vel, [24..31] mod
    bool has_peers = false;
    qie_peers peers{};
    void** d_peer_tab = nullptr;
    void* peer_tab_host[32] = {};            // what d_peer_tab holds (re-uploaded only when an address changes); [24..31] mod tables
    const int* d_tile_valid = nullptr;       // valid rows per 128-row tile of the gathered sequence of the installed geometry
    struct TileEntry { int img_total, txt_total, size; int* d; };
    std::vector<TileEntry> tile_cache;       // one list per geometry, kept until qie_destroy (captured graphs point at them)
    // optional per-kernel-class CUDA-event timing (bench.py roofline): class 0 gemm, 1 attention, 2 adaLN, 3 gemv, 4 other,
    // 5 peer barrier (sequence-parallel forward)
    int profile;
    struct Prof { int cls; cudaEvent_t a, b; double work; };
    std::vector<Prof> prof;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t get_event() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
};

using namespace qie;

extern "C" int qie_version(void) { return QIE_ABI_VERSION; }
extern "C" const char* qie_last_error(void) { return g_err; }
extern "C" int qie_device_sm_count(void) { return sm_count(); }

extern "C" int qie_make_seq(int batch, int img_rows, int txt_rows, qie_seq* out) {
    QIE_REQUIRE(out, QIE_EINVAL, "qie_make_seq: null out");
    QIE_REQUIRE(batch > 0 && img_rows > 0 && txt_rows > 0, QIE_ESHAPE, "qie_make_seq: B=%d img=%d txt=%d", batch,
                img_rows, txt_rows);
    out->batch = batch;
    out->img_rows = img_rows;
    out->txt_rows = txt_rows;
    out->img_pad = (img_rows + 127) / 128 * 128;
    out->txt_pad = (txt_rows + 127) / 128 * 128;
    for (int b = 0; b < 8; ++b) out->txt_rows_b[b] = b < batch ? txt_rows : 0;
    return QIE_OK;
}

extern "C" int qie_make_seq_ragged(int batch, int img_rows, const int* txt_rows_b, qie_seq* out) {
    QIE_REQUIRE(out && txt_rows_b && batch > 0 && batch <= 8, QIE_EINVAL, "qie_make_seq_ragged: bad argument");
    int mx = 0;
    for (int b = 0; b < batch; ++b) {
        QIE_REQUIRE(txt_rows_b[b] > 0, QIE_ESHAPE, "qie_make_seq_ragged: batch element %d has %d text tokens", b, txt_rows_b[b]);
        if (txt_rows_b[b] > mx) mx = txt_rows_b[b];
    }
    int rc = qie_make_seq(batch, img_rows, mx, out);
    if (rc) return rc;
    for (int b = 0; b < batch; ++b) out->txt_rows_b[b] = txt_rows_b[b];
    return QIE_OK;
}

// FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=linspace(1, 1/N, N), mu=calculate_shift(seq)) with the
// Qwen-Image scheduler config (dynamic exponential shifting, shift_terminal 0.02)  — SURVEY A.8
extern "C" int qie_flowmatch_sigmas(int num_steps, int image_seq_len, float* s) {
    QIE_REQUIRE(s && image_seq_len > 0, QIE_EINVAL, "qie_flowmatch_sigmas: bad argument");
    // one step: the stretch to the terminal sigma is (1 - s) / ((1 - s_last) / 0.98) with s = s_last = 1 -> 0 / 0 = NaN timesteps
    // in the reference scheduler (no error there); refuse instead of returning NaN
    QIE_REQUIRE(num_steps >= 2, QIE_EINVAL, "qie_flowmatch_sigmas: the dynamic-shift schedule needs at least 2 steps (got %d)",
                num_steps);
    const double m = (0.9 - 0.5) / (8192.0 - 256.0), b = 0.5 - m * 256.0;
    const double mu = image_seq_len * m + b, emu = exp(mu);
    std::vector<double> sig(num_steps);
    for (int i = 0; i < num_steps; ++i) {
        // np.linspace(1, 1/N, N) in float32
        const float lin = (float)(1.0 + (double)i * ((1.0 / num_steps - 1.0) / (num_steps - 1)));
        sig[i] = emu / (emu + (1.0 / (double)lin - 1.0));
    }
    const double scale = (1.0 - sig[num_steps - 1]) / (1.0 - 0.02);
    for (int i = 0; i < num_steps; ++i) s[i] = (float)(1.0 - (1.0 - sig[i]) / scale);
    s[num_steps] = 0.0f;
    return QIE_OK;
}

// QwenEmbedRope(theta=10000, axes, scale_rope=True) in the joint [img; txt] layout  — SURVEY A.5
extern "C" int qie_rope_table_host(const qie_model_cfg* cfg, const int* shp, int n_img, const qie_seq* seq,
                                   float* out) {
    QIE_REQUIRE(cfg && shp && seq && out && n_img > 0, QIE_EINVAL, "qie_rope_table_host: bad argument");
    const int a0 = cfg->rope_axes[0] / 2, a1 = cfg->rope_axes[1] / 2, a2 = cfg->rope_axes[2] / 2;
    QIE_REQUIRE(a0 + a1 + a2 == 64, QIE_ESHAPE, "qie_rope_table_host: rope axes must sum to 128");
    long long tok = 0;
    for (int i = 0; i < n_img; ++i) tok += (long long)shp[3 * i] * shp[3 * i + 1] * shp[3 * i + 2];
    QIE_REQUIRE(tok == seq->img_rows, QIE_ESHAPE, "qie_rope_table_host: img_shapes cover %lld tokens, img_rows=%d", tok,
                seq->img_rows);
    const int rpb = seq->img_pad + seq->txt_pad;
    for (long long i = 0; i < (long long)rpb * 128; i += 2) {
        out[i] = 1.f;
        out[i + 1] = 0.f;
    }
    auto put = [&](int row, int pair, int dim, int j, int index) {
        // fp32 like torch: outer(index.float(), 1 / theta^(arange(0,dim,2)/dim)) then polar()
        const float inv = 1.0f / powf(10000.0f, (float)(2 * j) / (float)dim);
        const float ang = (float)index * inv;
        out[((long long)row * 64 + pair) * 2] = cosf(ang);
        out[((long long)row * 64 + pair) * 2 + 1] = sinf(ang);
    };
    int row = 0, max_vid = 0;
    for (int i = 0; i < n_img; ++i) {
        const int f = shp[3 * i], h = shp[3 * i + 1], w = shp[3 * i + 2];
        for (int ff = 0; ff < f; ++ff)
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x, ++row) {
                    const int fi = i + ff;
                    const int hi = y - (h - h / 2);   // -ceil(h/2) .. floor(h/2)-1
                    const int wi = x - (w - w / 2);
                    for (int j = 0; j < a0; ++j) put(row, j, cfg->rope_axes[0], j, fi);
                    for (int j = 0; j < a1; ++j) put(row, a0 + j, cfg->rope_axes[1], j, hi);
                    for (int j = 0; j < a2; ++j) put(row, a0 + a1 + j, cfg->rope_axes[2], j, wi);
                }
        if (h / 2 > max_vid) max_vid = h / 2;
        if (w / 2 > max_vid) max_vid = w / 2;
    }
    for (int t = 0; t < seq->txt_rows; ++t) {
        const int r = seq->img_pad + t, idx = max_vid + t;
        for (int j = 0; j < a0; ++j) put(r, j, cfg->rope_axes[0], j, idx);
        for (int j = 0; j < a1; ++j) put(r, a0 + j, cfg->rope_axes[1], j, idx);
        for (int j = 0; j < a2; ++j) put(r, a0 + a1 + j, cfg->rope_axes[2], j, idx);
    }
    return QIE_OK;
}

extern "C" int qie_create(const qie_model_cfg* cfg, int device, qie_handle** out) {
    QIE_REQUIRE(cfg && out, QIE_EINVAL, "qie_create: null pointer");
    QIE_REQUIRE(cfg->head_dim == 128, QIE_ESHAPE, "qie_create: head_dim must be 128 (got %d)", cfg->head_dim);
    QIE_REQUIRE(cfg->num_layers > 0 && cfg->num_heads > 0 && cfg->in_channels % 8 == 0 && cfg->out_dim % 64 == 0 &&
                    cfg->joint_dim % 8 == 0,
                QIE_ESHAPE, "qie_create: unsupported model dims");
    QIE_REQUIRE(cfg->rope_axes[0] + cfg->rope_axes[1] + cfg->rope_axes[2] == 128, QIE_ESHAPE,
                "qie_create: rope axes must sum to head_dim");
    cudaDeviceProp prop;
    QIE_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    QIE_REQUIRE(prop.major == 10, QIE_EARCH, "qie_create: device %d is sm_%d%d, this library is sm_100a only", device,
                prop.major, prop.minor);
    QIE_CUDA_OK(cudaSetDevice(device));
    {   // kernel scratch pool of this device: allocate now, so that qie_forward stays allocation-free (CUDA-graph capturable)
        int rc0 = stream_scratch_reserve();
        if (rc0) return rc0;
    }
    qie_handle* h = new qie_handle();
    h->cfg = *cfg;
    h->device = device;
    h->D = cfg->num_heads * cfg->head_dim;
    h->has_weights = false;
    h->precision = 0;
    h->fuse_qk = 1;
    h->attn_variant = 0;
    h->d_rope = nullptr;
    h->profile = 0;
    // tproj [8,256] + t1 [8,D] + temb [8,D] + mod [8][L*2*6D] + final [8][2D]
    const size_t D = h->D, L = cfg->num_layers;
    h->small_bytes = sizeof(float) * 8 * (256 + 2 * D + L * 12 * D + 2 * D);
    cudaError_t e = cudaMalloc(&h->d_small, h->small_bytes);
    if (e != cudaSuccess) {
        delete h;
        return cuda_fail(e, "cudaMalloc(handle buffers)");
    }
    *out = h;
    return QIE_OK;
}

extern "C" int qie_destroy(qie_handle* h) {
    if (!h) return QIE_OK;
    cudaFree(h->d_small);
    for (auto& e : h->rope_cache) cudaFree(e.d);
    cudaFree(h->d_sched_mod);
    cudaFree(h->d_sched_fin);
    cudaFree(h->d_peer_tab);
    for (auto& e : h->tile_cache) cudaFree(e.d);
    for (float* p : h->d_prompt) cudaFree(p);
    for (auto& p : h->prof) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    delete h;
    return QIE_OK;
}

extern "C" int qie_set_weights(qie_handle* h, const qie_weights* w) {
    QIE_REQUIRE(h && w && w->blocks, QIE_EINVAL, "qie_set_weights: null pointer");
    const void* req[] = {w->img_in_w, w->img_in_b, w->txt_norm_w, w->txt_in_w, w->txt_in_b, w->t1_w,       w->t1_b,
                         w->t2_w,     w->t2_b,     w->mod_w,      w->mod_b,    w->norm_out_w, w->norm_out_b, w->proj_out_w,
                         w->proj_out_b};
    for (const void* p : req) QIE_REQUIRE(p, QIE_EINVAL, "qie_set_weights: a top-level weight pointer is null");
    h->blocks.assign(w->blocks, w->blocks + h->cfg.num_layers);
    for (int l = 0; l < h->cfg.num_layers; ++l)
        for (int s = 0; s < 2; ++s) {
            const qie_block_weights& b = h->blocks[l];
            QIE_REQUIRE(b.qkv_w[s] && b.qkv_b[s] && b.q_norm_w[s] && b.k_norm_w[s] && b.out_w[s] && b.out_b[s] &&
                            b.ff1_w[s] && b.ff1_b[s] && b.ff2_w[s] && b.ff2_b[s],
                        QIE_EINVAL, "qie_set_weights: block %d stream %d has a null pointer", l, s);
        }
    h->w = *w;
    h->w.blocks = h->blocks.data();
    // bound on the attention scores of every block.  After RMSNorm the unit row q^ has |q^|^2 = 128 and q = q^ * w; RoPE rotates the
    // channel pairs (2i, 2i+1), so with m_p = max(|w_2i|, |w_2i+1|):  q.k = sum_p rot(q_p).rot(k_p) <= sum_p |q_p| |k_p|
    // <= sum_p m_q,p m_k,p |q^_p| |k^_p| <= max_p(m_q,p m_k,p) * |q^| |k^| = 128 max_p(m_q,p m_k,p)   (Cauchy-Schwarz); image queries
    // meet text keys, so the maximum runs over both streams on either side.  2 % for the bf16 rounding of q and k.
    h->score_bound.assign(h->cfg.num_layers, INFINITY);
    for (int l = 0; l < h->cfg.num_layers; ++l) {
        float m[2][64];                  // [q / k][pair]: larger |weight| of the pair over both streams
        for (int k = 0; k < 2; ++k)
            for (int pr = 0; pr < 64; ++pr) m[k][pr] = 0.f;
        for (int s = 0; s < 2; ++s)
            for (int k = 0; k < 2; ++k) {
                float wv[128];
                QIE_CUDA_OK(cudaMemcpy(wv, k ? h->blocks[l].k_norm_w[s] : h->blocks[l].q_norm_w[s], sizeof(wv), cudaMemcpyDeviceToHost));
                for (int i = 0; i < 128; ++i) m[k][i >> 1] = fmaxf(m[k][i >> 1], isfinite(wv[i]) ? fabsf(wv[i]) : INFINITY);
            }
        float mx = 0.f;
        for (int pr = 0; pr < 64; ++pr) mx = fmaxf(mx, m[0][pr] * m[1][pr]);
        h->score_bound[l] = 128.f * mx * ATTN_SCALE_LOG2 * 1.02f;
    }
    h->has_weights = true;
    return QIE_OK;
}

// block `layer` runs the bounded-score attention: q leaves the fused QKV epilogue multiplied by softmax_scale * log2(e)
static bool layer_bounded(const qie_handle* h, int l) {
    return h->attn_bounded && h->fuse_qk && (h->attn_variant & 0x208) == 0 && h->score_bound[l] <= QIE_ATTN_SCORE_BOUND;
}

extern "C" int qie_attn_layer_variant(const qie_handle* h, int layer) {
    QIE_REQUIRE(h && h->has_weights && layer >= 0 && layer < h->cfg.num_layers, QIE_EINVAL, "qie_attn_layer_variant: bad argument");
    return layer_bounded(h, layer) ? (h->attn_variant | 0x200) : h->attn_variant;
}

extern "C" float qie_attn_score_bound(const qie_handle* h, int layer) {
    if (!h || !h->has_weights || layer < 0 || layer >= h->cfg.num_layers) return -1.f;
    return h->score_bound[layer];
}

extern "C" int qie_set_precision(qie_handle* h, int mode) {
    QIE_REQUIRE(h, QIE_EINVAL, "qie_set_precision: null handle");
    QIE_REQUIRE(mode >= 0 && mode <= 2, QIE_EINVAL, "qie_set_precision: mode must be 0 (bf16), 1 (fp8 e4m3) or 2 (int8)");
    if (mode >= 1) {
        QIE_REQUIRE(h->has_weights, QIE_ESTATE, "qie_set_precision: set weights first");
        for (auto& b : h->blocks)
            for (int s = 0; s < 2; ++s)
                QIE_REQUIRE(b.qkv_w8[s] && b.qkv_ws[s] && b.out_w8[s] && b.out_ws[s] && b.ff1_w8[s] && b.ff1_ws[s] &&
                                b.ff2_w8[s] && b.ff2_ws[s],
                            QIE_ESTATE, "qie_set_precision: fp8 weights missing");
    }
    h->precision = mode;
    return QIE_OK;
}

// debug / tuning knobs (not part of the reference surface): 0 = fuse_qk, 1 = attn_variant
extern "C" int qie_set_option(qie_handle* h, int key, int value) {
    QIE_REQUIRE(h, QIE_EINVAL, "qie_set_option: null handle");
    if (key == 0) h->fuse_qk = value;
    else if (key == 1) h->attn_variant = value;
    else if (key == 2) h->profile = value;
    else if (key == 3) h->attn_bounded = value;
    else QIE_REQUIRE(false, QIE_EINVAL, "qie_set_option: unknown key %d", key);
    return QIE_OK;
}

// workspace carve-up (all offsets 1 KB aligned)
namespace {
struct Ws {
    size_t resid, xm, xm8, xscale, qkv, attn, attn8, ffh, ffh8, xin, xtxt, outp, amax, total;
};
inline size_t al(size_t x) { return (x + 1023) & ~size_t(1023); }
Ws carve(const qie_handle* h, const qie_seq* s) {
    const size_t rows = (size_t)s->batch * (s->img_pad + s->txt_pad), D = h->D;
    Ws w{};
    size_t o = 0;
    w.resid = o; o += al(rows * D * 4);
    w.xm = o; o += al(rows * D * 2);
    w.qkv = o; o += al(rows * 3 * D * 2);
    w.attn = o; o += al(rows * D * 2);
    w.ffh = o; o += al(rows * 4 * D * 2);
    w.xin = o; o += al((size_t)s->batch * s->img_pad * h->cfg.in_channels * 2);
    w.xtxt = o; o += al((size_t)s->batch * s->txt_pad * h->cfg.joint_dim * 2);
    w.outp = o; o += al((size_t)s->batch * s->img_pad * h->cfg.out_dim * 2);
    w.xm8 = o; o += al(rows * D);
    w.attn8 = o; o += al(rows * D);
    w.ffh8 = o; o += al(rows * 4 * D);
    w.xscale = o; o += al(rows * 4 * 3);
    w.amax = o; o += al(rows * 4);
    w.total = o;
    return w;
}
}  // namespace

extern "C" unsigned long long qie_launch_count(void) { return g_launches; }

// per-launch timeline of the events recorded since the last read (does not consume them): start offset from the first recorded
// launch and duration in ms, kernel class; returns the number of entries written (<= max_n) or a negative status
extern "C" int qie_profile_timeline(qie_handle* h, float* start_ms, float* dur_ms, int* cls, int max_n) {
    QIE_REQUIRE(h && start_ms && dur_ms && cls && max_n >= 0, QIE_EINVAL, "qie_profile_timeline: bad argument");
    int n = 0;
    for (auto& p : h->prof) {
        if (n >= max_n) break;
        QIE_CUDA_OK(cudaEventSynchronize(p.b));
        QIE_CUDA_OK(cudaEventElapsedTime(&start_ms[n], h->prof.front().a, p.a));
        QIE_CUDA_OK(cudaEventElapsedTime(&dur_ms[n], p.a, p.b));
        cls[n++] = p.cls;
    }
    return n;
}

// sums the event-timed durations recorded since the last read; arrays of QIE_PROFILE_CLASSES classes (see qie_handle::profile)
extern "C" int qie_profile_read(qie_handle* h, double* ms, double* work, int* launches) {
    QIE_REQUIRE(h && ms && work && launches, QIE_EINVAL, "qie_profile_read: null pointer");
    for (int i = 0; i < QIE_PROFILE_CLASSES; ++i) { ms[i] = 0; work[i] = 0; launches[i] = 0; }
    for (auto& p : h->prof) {
        QIE_CUDA_OK(cudaEventSynchronize(p.b));
        float t = 0.f;
        QIE_CUDA_OK(cudaEventElapsedTime(&t, p.a, p.b));
        ms[p.cls] += t; work[p.cls] += p.work; launches[p.cls] += 1;
        h->ev_pool.push_back(p.a); h->ev_pool.push_back(p.b);
    }
    h->prof.clear();
    return QIE_OK;
}

namespace {
struct ProfScope {
    qie_handle* h; cudaStream_t st; bool on; qie_handle::Prof p;
    ProfScope(qie_handle* h_, cudaStream_t st_, int cls, double work) : h(h_), st(st_), on(h_->profile != 0) {
        if (on) { p.cls = cls; p.work = work; p.a = h->get_event(); p.b = h->get_event(); cudaEventRecord(p.a, st); }
    }
    ~ProfScope() { if (on) { cudaEventRecord(p.b, st); h->prof.push_back(p); } }
};
}  // namespace

extern "C" size_t qie_workspace_bytes(const qie_handle* h, const qie_seq* seq) {
    if (!h || !seq) return 0;
    return carve(h, seq).total;
}

// One implementation for the whole step and for its phases (sequence-parallel callers interleave NCCL all-to-alls):
//   BEGIN: RoPE table, temb + modulation vectors, stream embeddings        QKV : adaLN1 + QKV GEMM (+QK-norm+RoPE) of `layer`
//   ATTN : joint attention of `layer`                                      POST: out-proj, adaLN2, FF of `layer`
//   END  : norm_out + proj_out
static int forward_impl(qie_handle* h, int phases, int layer, const void* hidden, const void* enc, const float* timestep,
                        const int* img_shapes_host, int n_img, const qie_seq* seq, const qie_sp* sp, void* out,
                        void* workspace, size_t workspace_bytes, int n_blocks, void* stream) {
    QIE_REQUIRE(h && seq && workspace, QIE_EINVAL, "qie_forward: null pointer");
    if (int rc_sticky = peer_sticky_error()) return rc_sticky;   // a peer barrier of an earlier launch timed out
    if (phases & QIE_PHASE_BEGIN)
        QIE_REQUIRE(hidden && enc && timestep && img_shapes_host, QIE_EINVAL, "qie_forward: null input pointer");
    if (phases & QIE_PHASE_END) QIE_REQUIRE(out || (h->has_peers && sp), QIE_EINVAL, "qie_forward: null output pointer");
    QIE_REQUIRE(h->has_weights, QIE_ESTATE, "qie_forward: weights not set");
    QIE_REQUIRE(seq->batch >= 1 && seq->batch <= 8, QIE_ESHAPE, "qie_forward: batch must be 1..8");
    qie_seq chk;
    int rc = qie_make_seq(seq->batch, seq->img_rows, seq->txt_rows, &chk);
    if (rc) return rc;
    QIE_REQUIRE(chk.img_pad == seq->img_pad && chk.txt_pad == seq->txt_pad, QIE_ESHAPE,
                "qie_forward: seq padding must come from qie_make_seq");
    bool ragged = false;
    long long txt_valid_total = 0;
    for (int b = 0; b < seq->batch; ++b) {
        QIE_REQUIRE(seq->txt_rows_b[b] >= 1 && seq->txt_rows_b[b] <= seq->txt_rows, QIE_ESHAPE,
                    "qie_forward: batch element %d has %d text tokens (1..%d; build the layout with qie_make_seq / qie_make_seq_ragged)",
                    b, seq->txt_rows_b[b], seq->txt_rows);
        ragged = ragged || seq->txt_rows_b[b] != seq->txt_rows;
        txt_valid_total += seq->txt_rows_b[b];
    }
    QIE_REQUIRE(!ragged || !sp, QIE_ESHAPE, "qie_forward_phase: per-element text lengths are not supported in sequence-parallel shards");
    const Ws ws = carve(h, seq);
    QIE_REQUIRE(workspace_bytes >= ws.total, QIE_ENOMEM, "qie_forward: workspace %zu < required %zu", workspace_bytes,
                ws.total);
    QIE_REQUIRE(((uintptr_t)workspace & 1023) == 0, QIE_EINVAL, "qie_forward: workspace must be 1 KB aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // the installed peer tables apply to sequence-parallel calls only (sp given); a plain qie_forward on the same handle runs
    // the single-GPU path untouched
    const bool use_peers = h->has_peers && sp != nullptr;
    const int B = seq->batch, D = h->D, L = h->cfg.num_layers;
    const int nb = n_blocks < 0 || n_blocks > L ? L : n_blocks;
    const int rpb = seq->img_pad + seq->txt_pad;
    uint8_t* W = (uint8_t*)workspace;
    float* resid = (float*)(W + ws.resid);
    void* xm = W + ws.xm;
    void* qkv = W + ws.qkv;
    void* attn = W + ws.attn;
    void* ffh = W + ws.ffh;
    void* xin = W + ws.xin;
    void* xtxt = W + ws.xtxt;
    void* outp = W + ws.outp;
    void* xm8 = W + ws.xm8;
    void* attn8 = W + ws.attn8;
    void* ffh8 = W + ws.ffh8;
    float* xscale = (float*)(W + ws.xscale);   // [3][rows]: xm, attn, ffh scales
    float* ffh_amax = (float*)(W + ws.amax);   // [rows]: max|ffh| per token, folded in by the FF-up epilogue (8-bit modes)
    const size_t rows = (size_t)B * rpb;
    const int fp8 = h->precision;   // 0 bf16, 1 e4m3 W8A8, 2 int8 W8A8

    const double valid_rows = (double)B * seq->img_rows + (double)txt_valid_total;
    auto run_gemm = [&](qie_gemm_args& g) -> int {
        const double mv = ((g.streams & 1) ? (double)B * seq->img_rows : 0) + ((g.streams & 2) ? (double)txt_valid_total : 0);
        ProfScope ps(h, st, 0, 2.0 * mv * g.N * g.K);
        return qie_gemm(&g, seq, st);
    };
    // bounded-score attention of block l: q carries softmax_scale * log2(e) out of the fused QKV epilogue and the kernel skips the
    // running max (attn.cu); only where the norm weights bound the scores, never with the unfused debug path or a forced kernel
    auto bounded = [&](int l) -> bool { return layer_bounded(h, l); };
    auto run_attn = [&](int l) -> int {
        const int variant = bounded(l) ? (h->attn_variant | 0x200) : h->attn_variant;
        if (use_peers) {
            // my head group over the gathered sequence of every rank; the epilogue stores each token's output into the
            // attention buffer of the rank that owns the token
            const qie_peers& pr = h->peers;
            const int hl = h->cfg.num_heads / pr.size;
            const double S = (double)pr.img_total + pr.txt_total;
            ProfScope ps(h, st, 1, 4.0 * S * S * 128.0 * hl * pr.batch);
            return attn_fwd_peers(pr.qkv_gather[pr.rank], h->d_peer_tab + 8, &pr, h->d_tile_valid, hl, D, variant, st);
        }
        double ss = 0;
        for (int b = 0; b < B; ++b) ss += ((double)seq->img_rows + seq->txt_rows_b[b]) * ((double)seq->img_rows + seq->txt_rows_b[b]);
        ProfScope ps(h, st, 1, 4.0 * ss * 128.0 * h->cfg.num_heads);
        return qie_attn_fwd(qkv, attn, seq, h->cfg.num_heads, variant, st);
    };
    if (use_peers && (phases & (QIE_PHASE_QKV | QIE_PHASE_ATTN | QIE_PHASE_END)))
        QIE_REQUIRE(B == h->peers.batch && h->peers.img_pad == seq->img_pad && h->peers.txt_pad == seq->txt_pad && h->fuse_qk &&
                        sp->size == h->peers.size && sp->rank == h->peers.rank && sp->img_total == h->peers.img_total &&
                        sp->txt_total == h->peers.txt_total,
                    QIE_ESTATE, "qie_forward_phase: the installed peers describe another geometry (batch %d, shard %d+%d rows, "
                    "%d+%d tokens): call qie_set_peers for this one", h->peers.batch, h->peers.img_pad, h->peers.txt_pad,
                    h->peers.img_total, h->peers.txt_total);
    auto run_ln = [&](const float* mv, long long bs, long long ss, int sh, int sc, bool q8) -> int {
        ProfScope ps(h, st, 2, valid_rows * D * (q8 ? 5.0 : 6.0));   // algorithmic bytes: fp32 row in, bf16 (or 8-bit) row out
        // W8A8 modes: the GEMMs that follow read the 8-bit shadow only, the bf16 rows are not written
        return qie_ln_modulate(resid, mv, bs, ss, sh, sc, q8 ? nullptr : xm, q8 ? xm8 : nullptr, q8 ? xscale : nullptr, h->precision, D,
                               1e-6f, seq, st);
    };

    // ---- RoPE table (cached per shape key; host build + one H2D copy only when a new shape / shard appears).  Selected in
    // the BEGIN phase and again in every QKV phase, so one handle can serve several shards in turn (single-GPU emulation of
    // a sequence-parallel group) ----
    if ((phases & (QIE_PHASE_BEGIN | QIE_PHASE_QKV)) && img_shapes_host) {
        std::vector<int> key(img_shapes_host, img_shapes_host + 3 * n_img);
        key.push_back(seq->img_rows);
        key.push_back(seq->txt_rows);
        if (sp) {
            const int extra[6] = {sp->rank, sp->size, sp->img_total, sp->txt_total, sp->img_offset, sp->txt_offset};
            key.insert(key.end(), extra, extra + 6);
        }
        float* found = nullptr;
        for (auto& e : h->rope_cache)
            if (e.key == key) found = e.d;
        if (!found) {
            std::vector<float> tab((size_t)rpb * 128);
            if (!sp) {
                rc = qie_rope_table_host(&h->cfg, img_shapes_host, n_img, seq, tab.data());
                if (rc) return rc;
            } else {
                // sequence-parallel shard: build the table of the WHOLE sequence, keep the rows this rank owns
                QIE_REQUIRE(sp->img_offset + seq->img_rows <= sp->img_total &&
                                sp->txt_offset + seq->txt_rows <= sp->txt_total,
                            QIE_ESHAPE, "qie_forward: bad sequence-parallel shard");
                qie_seq full;
                if ((rc = qie_make_seq(1, sp->img_total, sp->txt_total, &full))) return rc;
                std::vector<float> all((size_t)(full.img_pad + full.txt_pad) * 128);
                if ((rc = qie_rope_table_host(&h->cfg, img_shapes_host, n_img, &full, all.data()))) return rc;
                for (size_t i = 0; i < tab.size(); i += 2) { tab[i] = 1.f; tab[i + 1] = 0.f; }
                memcpy(tab.data(), all.data() + (size_t)sp->img_offset * 128, (size_t)seq->img_rows * 128 * sizeof(float));
                memcpy(tab.data() + (size_t)seq->img_pad * 128, all.data() + (size_t)(full.img_pad + sp->txt_offset) * 128,
                       (size_t)seq->txt_rows * 128 * sizeof(float));
            }
            if (h->rope_cache.size() >= 16) {      // cudaFree synchronises the device: no kernel still reads the evicted table
                cudaFree(h->rope_cache.front().d);
                h->rope_cache.erase(h->rope_cache.begin());
            }
            QIE_CUDA_OK(cudaMalloc(&found, tab.size() * sizeof(float)));
            QIE_CUDA_OK(cudaMemcpyAsync(found, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, st));
            QIE_CUDA_OK(cudaStreamSynchronize(st));   // tab is a stack-owned host buffer; only when a new shape appears
            h->rope_cache.push_back({key, found});
        }
        h->d_rope = found;
    }
    QIE_REQUIRE(!(phases & QIE_PHASE_QKV) || h->d_rope, QIE_ESTATE, "qie_forward_phase: QKV phase before any BEGIN phase");

    // ---- small per-timestep vectors: temb, every block's modulation, final scale/shift ----
    float* tproj = h->d_small;
    float* t1 = tproj + 8 * 256;
    float* temb = t1 + 8 * (size_t)D;
    float* mod = temb + 8 * (size_t)D;                       // [B][L*2*6D]
    const long long modN = (long long)nb * 12 * D;   // batch stride of the modulation table
    float* fin = mod + 8 * (size_t)L * 12 * D;                     // [B][2D]
    // sequence parallel: the modulation table lives in peer-visible memory; every rank computes 1/P of its rows (the 13.6 GB
    // weight stream is the one part of the forward that does not shrink with the token shard) and stores them into the others
    if (use_peers) mod = (float*)h->peers.mod[h->peers.rank];
    if (phases & QIE_PHASE_BEGIN) {
    bool sched_cached = h->n_sched > 0;
    for (int b = 0; b < B; ++b) sched_cached = sched_cached && h->sel_sched[b] >= 0 && h->sel_sched[b] < h->n_sched;
    if (sched_cached) {
        // every vector below depends on the timestep only: reuse the tables built by qie_cache_schedule (bit-identical)
        for (int b = 0; b < B; ++b) {
            const size_t i = (size_t)h->sel_sched[b];
            if (nb > 0)
                QIE_CUDA_OK(cudaMemcpyAsync(mod + (size_t)b * modN, h->d_sched_mod + i * (size_t)L * 12 * D,
                                            (size_t)modN * sizeof(float), cudaMemcpyDeviceToDevice, st));
            QIE_CUDA_OK(cudaMemcpyAsync(fin + (size_t)b * 2 * D, h->d_sched_fin + i * 2 * D, (size_t)2 * D * sizeof(float),
                                        cudaMemcpyDeviceToDevice, st));
        }
    } else {
    if ((rc = qie_timestep_proj(timestep, tproj, B, 0, st))) return rc;
    if ((rc = qie_gemv(tproj, h->w.t1_w, h->w.t1_b, t1, B, D, 256, 0, st))) return rc;
    if ((rc = qie_gemv(t1, h->w.t2_w, h->w.t2_b, temb, B, D, D, 1, st))) return rc;
    if (nb > 0 && use_peers) {
        const qie_peers& pr = h->peers;
        const long long q = modN / 4, n0 = q * pr.rank / pr.size * 4, n1 = q * (pr.rank + 1) / pr.size * 4;   // 16-byte granules
        {
            ProfScope ps(h, st, 3, (double)(n1 - n0) * D * 2);
            if ((rc = gemv_strided(temb, (const __nv_bfloat16*)h->w.mod_w + n0 * D, h->w.mod_b + n0, mod + n0, B, n1 - n0, D, 1, modN, st)))
                return rc;
        }
        if ((rc = peer_bcast_span(mod, h->d_peer_tab + 24, &pr, modN * 4, n0 * 4, (n1 - n0) * 4, st))) return rc;
    } else if (nb > 0) {
        ProfScope ps(h, st, 3, (double)modN * D * 2);
        if ((rc = qie_gemv(temb, h->w.mod_w, h->w.mod_b, mod, B, modN, D, 1, st))) return rc;
    }
    if ((rc = qie_gemv(temb, h->w.norm_out_w, h->w.norm_out_b, fin, B, 2 * D, D, 1, st))) return rc;
    }
    // mod row for (b, layer l, stream s): mod + b*modN + (l*2+s)*6D, chunks [shift1|scale1|gate1|shift2|scale2|gate2]

    // ---- stream embeddings -> fp32 residual in the joint layout ----
    if ((rc = qie_pack_rows(hidden, xin, B, seq->img_rows, seq->img_pad, h->cfg.in_channels, st))) return rc;
    const bool prompt_cached = h->sel_prompt >= 0 && h->d_prompt[h->sel_prompt] && !sp && !ragged &&
                               h->prompt_rows[h->sel_prompt] == seq->txt_rows;
    if (!prompt_cached &&
        (rc = rmsnorm_pack_ragged(enc, h->w.txt_norm_w, xtxt, B, seq->txt_rows_b, seq->txt_rows, seq->txt_pad, h->cfg.joint_dim, 1e-6f,
                                  st)))
        return rc;
    {
        qie_gemm_args g{};
        g.a = xin; g.a_compact = 1; g.w[0] = h->w.img_in_w; g.bias[0] = h->w.img_in_b;
        g.out = resid; g.ldo = D; g.N = D; g.K = h->cfg.in_channels; g.streams = 1; g.epilogue = QIE_EPI_F32;
        if ((rc = run_gemm(g))) return rc;
        if (prompt_cached) {
            // txt_in(txt_norm(prompt_embeds)) depends on the prompt only: copy the cached fp32 rows into the residual
            for (int b = 0; b < B; ++b)
                QIE_CUDA_OK(cudaMemcpyAsync(resid + ((size_t)b * rpb + seq->img_pad) * D, h->d_prompt[h->sel_prompt],
                                            (size_t)seq->txt_pad * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        } else {
        qie_gemm_args t{};
        t.a = xtxt; t.a_compact = 1; t.w[1] = h->w.txt_in_w; t.bias[1] = h->w.txt_in_b;
        t.out = resid; t.ldo = D; t.N = D; t.K = h->cfg.joint_dim; t.streams = 2; t.epilogue = QIE_EPI_F32;
        if ((rc = run_gemm(t))) return rc;
        }
    }
    }   // QIE_PHASE_BEGIN

    // ---- transformer blocks ----
    const int l_begin = layer < 0 ? 0 : layer, l_end = layer < 0 ? nb : (layer < nb ? layer + 1 : nb);
    for (int l = l_begin; l < l_end; ++l) {
        const qie_block_weights& bw = h->blocks[l];
        const float* m = mod + (size_t)l * 12 * D;   // + b*modN + s*6D
        if (phases & QIE_PHASE_QKV) {
        // adaLN 1 -> xm
        if ((rc = run_ln(m, modN, 6LL * D, 0, D, fp8))) return rc;
        {   // QKV (+ QK-RMSNorm + RoPE)
            qie_gemm_args g{};
            g.N = 3 * D; g.K = D; g.streams = 3; g.out = qkv; g.ldo = 3 * D;
            g.epilogue = h->fuse_qk ? QIE_EPI_QKV_NORM_ROPE : QIE_EPI_BF16;
            g.rope = h->d_rope;
            g.q_scale = bounded(l) ? ATTN_SCALE_LOG2 : 1.f;
            for (int s = 0; s < 2; ++s) {
                g.bias[s] = bw.qkv_b[s];
                g.qk_norm_w[s][0] = bw.q_norm_w[s];
                g.qk_norm_w[s][1] = bw.k_norm_w[s];
                g.w[s] = fp8 ? bw.qkv_w8[s] : bw.qkv_w[s];
                g.w_scale[s] = bw.qkv_ws[s];
            }
            g.a = fp8 ? xm8 : xm; g.fp8 = fp8; g.a_scale = xscale;
            if (use_peers) {   // epilogue scatters q|k|v of head group g into rank g's gathered buffer (peer stores)
                const qie_peers& pr = h->peers;
                g.peer_out = h->d_peer_tab; g.sp_rank = pr.rank; g.sp_size = pr.size;
                g.sp_gathered_rows = pr.size * pr.img_pad + (pr.txt_total + 127) / 128 * 128;
                g.sp_txt_row0 = pr.size * pr.img_pad + sp->txt_offset;
            }
            if ((rc = run_gemm(g))) return rc;
            if (!h->fuse_qk) {
                const float* nw[4] = {bw.q_norm_w[0], bw.k_norm_w[0], bw.q_norm_w[1], bw.k_norm_w[1]};
                if ((rc = qie_qk_norm_rope(qkv, h->d_rope, nw, h->cfg.num_heads, 1e-6f, seq, st))) return rc;
            }
        }
        }   // QIE_PHASE_QKV
        if ((phases & QIE_PHASE_ATTN) && (rc = run_attn(l))) return rc;
        if (phases & QIE_PHASE_POST) {
        {   // out-proj + gate1 * y + residual
            qie_gemm_args g{};
            g.N = D; g.K = D; g.streams = 3; g.out = resid; g.ldo = D; g.epilogue = QIE_EPI_GATE_RESID_F32;
            g.gate = m + 2 * D; g.gate_bstride = modN; g.gate_sstride = 6LL * D;
            for (int s = 0; s < 2; ++s) {
                g.bias[s] = bw.out_b[s];
                g.w[s] = fp8 ? bw.out_w8[s] : bw.out_w[s];
                g.w_scale[s] = bw.out_ws[s];
            }
            g.a = attn;
            if (fp8) {
                ProfScope ps(h, st, 4, (double)rows * D * 3);
                if ((rc = qie_quant_rows(attn, attn8, xscale + rows, (long long)rows, D, h->precision, st))) return rc;
                g.a = attn8; g.fp8 = fp8; g.a_scale = xscale + rows;
            }
            if ((rc = run_gemm(g))) return rc;
        }
        // adaLN 2 -> xm
        if ((rc = run_ln(m, modN, 6LL * D, 3 * D, 4 * D, fp8))) return rc;
        {   // FF up + GELU(tanh)
            qie_gemm_args g{};
            g.N = 4 * D; g.K = D; g.streams = 3; g.out = ffh; g.ldo = 4 * D; g.epilogue = QIE_EPI_GELU_BF16;
            for (int s = 0; s < 2; ++s) {
                g.bias[s] = bw.ff1_b[s];
                g.w[s] = fp8 ? bw.ff1_w8[s] : bw.ff1_w[s];
                g.w_scale[s] = bw.ff1_ws[s];
            }
            g.a = fp8 ? xm8 : xm; g.fp8 = fp8; g.a_scale = xscale;
            if (fp8) {      // the per-token quantiser of the FF hidden needs the row max: the GELU epilogue folds it in
                QIE_CUDA_OK(cudaMemsetAsync(ffh_amax, 0, rows * sizeof(float), st));
                g.q8_amax = ffh_amax;
            }
            if ((rc = run_gemm(g))) return rc;
        }
        {   // FF down + gate2 * y + residual
            qie_gemm_args g{};
            g.N = D; g.K = 4 * D; g.streams = 3; g.out = resid; g.ldo = D; g.epilogue = QIE_EPI_GATE_RESID_F32;
            g.gate = m + 5 * D; g.gate_bstride = modN; g.gate_sstride = 6LL * D;
            for (int s = 0; s < 2; ++s) {
                g.bias[s] = bw.ff2_b[s];
                g.w[s] = fp8 ? bw.ff2_w8[s] : bw.ff2_w[s];
                g.w_scale[s] = bw.ff2_ws[s];
            }
            g.a = ffh;
            if (fp8) {
                ProfScope ps(h, st, 4, (double)rows * 4 * D * 3);      // one streaming pass: bf16 in, 8 bit out
                if ((rc = quant_rows_amax(ffh, ffh_amax, ffh8, xscale + 2 * rows, (long long)rows, 4 * D, h->precision, st))) return rc;
                g.a = ffh8; g.fp8 = fp8; g.a_scale = xscale + 2 * rows;
            }
            if ((rc = run_gemm(g))) return rc;
        }
        }   // QIE_PHASE_POST
    }

    // ---- norm_out (AdaLayerNormContinuous: scale first, then shift) + proj_out on the image stream ----
    if (!(phases & QIE_PHASE_END)) return QIE_OK;
    if ((rc = run_ln(fin, 2LL * D, 0, D, 0, false))) return rc;
    {
        qie_gemm_args g{};
        g.a = xm; g.w[0] = h->w.proj_out_w; g.bias[0] = h->w.proj_out_b;
        g.out = outp; g.out_compact = 1; g.ldo = h->cfg.out_dim; g.N = h->cfg.out_dim; g.K = D; g.streams = 1;
        g.epilogue = QIE_EPI_BF16;
        if ((rc = run_gemm(g))) return rc;
    }
    if (use_peers) {
        // sequence parallel: every rank of the group needs the whole velocity for the (replicated) Euler update — my rows go
        // straight into every rank's velocity buffer (NVLink peer stores); the caller's barrier makes them visible
        if ((rc = peer_bcast_rows(outp, h->d_peer_tab + 16, &h->peers, seq->img_rows, sp->img_offset, h->cfg.out_dim, st)))
            return rc;
        if (!out) return QIE_OK;
    }
    if (seq->img_pad == seq->img_rows) {
        QIE_CUDA_OK(cudaMemcpyAsync(out, outp, (size_t)B * seq->img_rows * h->cfg.out_dim * 2, cudaMemcpyDeviceToDevice,
                                    st));
    } else {
        if ((rc = unpack_rows(outp, out, B, seq->img_rows, seq->img_pad, h->cfg.out_dim, st))) return rc;
    }
    return QIE_OK;
}

extern "C" int qie_forward(qie_handle* h, const void* hidden, const void* enc, const float* timestep,
                           const int* img_shapes_host, int n_img, const qie_seq* seq, void* out, void* workspace,
                           size_t workspace_bytes, int n_blocks, void* stream) {
    return forward_impl(h, QIE_PHASE_ALL, -1, hidden, enc, timestep, img_shapes_host, n_img, seq, nullptr, out, workspace,
                        workspace_bytes, n_blocks, stream);
}

extern "C" int qie_forward_phase(qie_handle* h, int phases, int layer, const void* hidden, const void* enc,
                                 const float* timestep, const int* img_shapes_host, int n_img, const qie_seq* seq,
                                 const qie_sp* sp, void* out, void* workspace, size_t workspace_bytes, int n_blocks,
                                 void* stream) {
    QIE_REQUIRE(phases > 0 && (phases & ~QIE_PHASE_ALL) == 0, QIE_EINVAL, "qie_forward_phase: bad phase mask %d", phases);
    return forward_impl(h, phases, layer, hidden, enc, timestep, img_shapes_host, n_img, seq, sp, out, workspace,
                        workspace_bytes, n_blocks, stream);
}

// Shard of `rank` in a sequence-parallel group of `size`: image tokens and text tokens are each split contiguously, the first
// (total % size) ranks own one token more.  Fills the local layout (`seq_out`, per-rank padding identical on every rank) and the
// placement inside the whole sequence (`sp_out`).  Mirrors parallel.make_shard_plan.
extern "C" int qie_sp_shard(int batch, int img_total, int txt_total, int size, int rank, qie_seq* seq_out, qie_sp* sp_out) {
    QIE_REQUIRE(seq_out && sp_out && batch >= 1 && size >= 1 && size <= 8 && rank >= 0 && rank < size, QIE_EINVAL,
                "qie_sp_shard: bad argument");
    QIE_REQUIRE(img_total >= size && txt_total >= size, QIE_ESHAPE,
                "qie_sp_shard: every rank needs at least one image and one text token (img=%d, txt=%d, P=%d)", img_total,
                txt_total, size);
    auto share = [&](int total, int r) { return total / size + (r < total % size ? 1 : 0); };
    auto offset = [&](int total, int r) { return r * (total / size) + (r < total % size ? r : total % size); };
    auto pad = [](int n) { return (n + 127) / 128 * 128; };
    QIE_REQUIRE(pad(share(img_total, 0)) == pad(share(img_total, size - 1)) &&
                    pad(share(txt_total, 0)) == pad(share(txt_total, size - 1)),
                QIE_ESHAPE, "qie_sp_shard: shards straddle a 128-row boundary (a rank would own an all-padding tile)");
    seq_out->batch = batch;
    seq_out->img_rows = share(img_total, rank);
    seq_out->txt_rows = share(txt_total, rank);
    seq_out->img_pad = pad(share(img_total, 0));
    seq_out->txt_pad = pad(share(txt_total, 0));
    sp_out->rank = rank;
    sp_out->size = size;
    sp_out->img_total = img_total;
    sp_out->txt_total = txt_total;
    sp_out->img_offset = offset(img_total, rank);
    sp_out->txt_offset = offset(txt_total, rank);
    return QIE_OK;
}

// Valid rows of every 128-row tile of the gathered sequence [rank 0 image shard | ... | rank P-1 image shard | all text tokens]
extern "C" int qie_sp_tile_valid_host(int img_total, int txt_total, int size, int* out, int max_tiles) {
    QIE_REQUIRE(out && size >= 1 && size <= 8, QIE_EINVAL, "qie_sp_tile_valid_host: bad argument");
    qie_seq s;
    qie_sp p;
    int rc = qie_sp_shard(1, img_total, txt_total, size, 0, &s, &p);
    if (rc) return rc;
    const int n = (size * s.img_pad + (txt_total + 127) / 128 * 128) / 128;
    QIE_REQUIRE(n <= max_tiles, QIE_ENOMEM, "qie_sp_tile_valid_host: %d tiles, room for %d", n, max_tiles);
    int k = 0;
    for (int r = 0; r < size; ++r) {
        const int rows = img_total / size + (r < img_total % size ? 1 : 0);
        for (int t = 0; t < s.img_pad / 128; ++t) out[k++] = rows - t * 128 < 128 ? rows - t * 128 : 128;
    }
    for (int t = 0; t * 128 < txt_total; ++t) out[k++] = txt_total - t * 128 < 128 ? txt_total - t * 128 : 128;
    return k;
}

extern "C" int qie_set_peers(qie_handle* h, const qie_peers* peers, void* stream) {
    QIE_REQUIRE(h, QIE_EINVAL, "qie_set_peers: null handle");
    if (!peers) {
        h->has_peers = false;
        return peer_sticky_error(/*clear=*/true);     // reports (once) a barrier timeout of the group that is being dissolved
    }
    QIE_REQUIRE(peers->size >= 2 && peers->size <= 8 && peers->rank >= 0 && peers->rank < peers->size &&
                    h->cfg.num_heads % peers->size == 0 && peers->batch >= 1 && peers->batch <= 8,
                QIE_EINVAL, "qie_set_peers: bad group (size %d rank %d batch %d)", peers->size, peers->rank, peers->batch);
    qie_seq s;
    qie_sp p;
    int rc = qie_sp_shard(peers->batch, peers->img_total, peers->txt_total, peers->size, peers->rank, &s, &p);
    if (rc) return rc;
    QIE_REQUIRE(s.img_pad == peers->img_pad && s.txt_pad == peers->txt_pad, QIE_ESHAPE,
                "qie_set_peers: shard padding %d+%d does not match qie_sp_shard (%d+%d)", peers->img_pad, peers->txt_pad,
                s.img_pad, s.txt_pad);
    void* tab[32] = {};
    for (int i = 0; i < peers->size; ++i) {
        QIE_REQUIRE(peers->qkv_gather[i] && peers->attn_out[i] && peers->vel[i] && peers->flags[i] && peers->mod[i], QIE_EINVAL,
                    "qie_set_peers: a buffer of rank %d is null", i);
        tab[i] = peers->qkv_gather[i];
        tab[8 + i] = peers->attn_out[i];
        tab[16 + i] = peers->vel[i];
        tab[24 + i] = peers->mod[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->d_peer_tab) {
        QIE_CUDA_OK(cudaMalloc(&h->d_peer_tab, sizeof(tab)));
        memset(h->peer_tab_host, 0xff, sizeof(h->peer_tab_host));
    }
    if (memcmp(tab, h->peer_tab_host, sizeof(tab)) != 0) {
        // pageable source: the runtime stages it before returning, so `tab` may live on this stack; ordered on `stream`
        QIE_CUDA_OK(cudaMemcpyAsync(h->d_peer_tab, tab, sizeof(tab), cudaMemcpyHostToDevice, st));
        memcpy(h->peer_tab_host, tab, sizeof(tab));
    }
    // tile list of this geometry (kept for the lifetime of the handle: captured graphs keep pointing at it)
    int* d_tiles = nullptr;
    for (auto& e : h->tile_cache)
        if (e.img_total == peers->img_total && e.txt_total == peers->txt_total && e.size == peers->size) d_tiles = e.d;
    if (!d_tiles) {
        int tiles[1024];
        const int n = qie_sp_tile_valid_host(peers->img_total, peers->txt_total, peers->size, tiles, 1024);
        if (n < 0) return n;
        QIE_CUDA_OK(cudaMalloc(&d_tiles, n * sizeof(int)));
        QIE_CUDA_OK(cudaMemcpyAsync(d_tiles, tiles, n * sizeof(int), cudaMemcpyHostToDevice, st));
        QIE_CUDA_OK(cudaStreamSynchronize(st));      // `tiles` lives on this stack; only when a new geometry appears
        h->tile_cache.push_back({peers->img_total, peers->txt_total, peers->size, d_tiles});
    }
    h->d_tile_valid = d_tiles;
    h->peers = *peers;
    h->has_peers = true;
    return QIE_OK;
}

// all-ranks barrier of the installed group on `stream` (device-side epochs: replayable inside a CUDA graph)
extern "C" int qie_peer_barrier(qie_handle* h, void* stream) {
    QIE_REQUIRE(h && h->has_peers, QIE_ESTATE, "qie_peer_barrier: no peers installed");
    ProfScope ps(h, (cudaStream_t)stream, 5, 0.0);
    return peer_barrier_launch(&h->peers, (cudaStream_t)stream);
}

// One call = the whole sequence-parallel forward of this rank (fused peer-memory exchange): BEGIN, then per block
// QKV -> barrier -> ATTN -> barrier -> POST, then END (velocity rows stored into every rank's buffer) -> barrier -> copy of the
// whole velocity into `out_full` [B, img_total, out_dim].  No host synchronisation, no allocation: capturable in a CUDA graph.
extern "C" int qie_forward_sp(qie_handle* h, const void* hidden_local, const void* enc_local, const float* timestep,
                              const int* img_shapes_host, int n_img, const qie_seq* seq, const qie_sp* sp, void* out_full,
                              void* workspace, size_t workspace_bytes, void* stream) {
    QIE_REQUIRE(h && h->has_peers, QIE_ESTATE, "qie_forward_sp: no peers installed (qie_set_peers)");
    QIE_REQUIRE(sp && out_full, QIE_EINVAL, "qie_forward_sp: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = forward_impl(h, QIE_PHASE_BEGIN, -1, hidden_local, enc_local, timestep, img_shapes_host, n_img, seq, sp, nullptr,
                          workspace, workspace_bytes, -1, stream);
    if (rc) return rc;
    if ((rc = qie_peer_barrier(h, stream))) return rc;          // every rank's share of the modulation table has landed in mine
    for (int l = 0; l < h->cfg.num_layers; ++l) {
        // adaLN1 + QKV GEMM: the epilogue stores q|k|v of head group g into rank g's gather buffer
        if ((rc = forward_impl(h, QIE_PHASE_QKV, l, nullptr, nullptr, nullptr, img_shapes_host, n_img, seq, sp, nullptr, workspace,
                               workspace_bytes, -1, stream))) return rc;
        if ((rc = qie_peer_barrier(h, stream))) return rc;      // everybody's q|k|v has landed in my gather buffer
        // attention of my heads over all tokens: the epilogue stores into the token owners' attention buffers
        if ((rc = forward_impl(h, QIE_PHASE_ATTN, l, nullptr, nullptr, nullptr, nullptr, 0, seq, sp, nullptr, workspace,
                               workspace_bytes, -1, stream))) return rc;
        if ((rc = qie_peer_barrier(h, stream))) return rc;      // everybody's heads have landed in my attention buffer
        if ((rc = forward_impl(h, QIE_PHASE_POST, l, nullptr, nullptr, nullptr, nullptr, 0, seq, sp, nullptr, workspace,
                               workspace_bytes, -1, stream))) return rc;
    }
    if ((rc = forward_impl(h, QIE_PHASE_END, -1, nullptr, nullptr, nullptr, nullptr, 0, seq, sp, nullptr, workspace,
                           workspace_bytes, -1, stream))) return rc;
    if ((rc = qie_peer_barrier(h, stream))) return rc;          // the whole velocity is in my buffer
    const qie_peers& pr = h->peers;
    QIE_CUDA_OK(cudaMemcpyAsync(out_full, pr.vel[pr.rank], (size_t)pr.batch * pr.img_total * h->cfg.out_dim * 2,
                                cudaMemcpyDeviceToDevice, st));
    return QIE_OK;
}

// byte offset of an activation buffer inside the workspace: 0 = qkv [rows, 3D] bf16, 1 = attention output [rows, D] bf16
extern "C" long long qie_workspace_offset(const qie_handle* h, const qie_seq* seq, int which) {
    if (!h || !seq) return -1;
    const Ws ws = carve(h, seq);
    return which == 0 ? (long long)ws.qkv : which == 1 ? (long long)ws.attn : -1;
}

// ---- exact caches (replaces cached_pipeline_v2.py, README.md:125, and the precompute_conditions stub at
// qwen_realtime.py:140-165).  Setup-time calls: they allocate and synchronise. ----
extern "C" int qie_cache_schedule(qie_handle* h, const float* timesteps_host, int n, void* stream) {
    QIE_REQUIRE(h && timesteps_host && n > 0 && n <= 64, QIE_EINVAL, "qie_cache_schedule: bad argument");
    QIE_REQUIRE(h->has_weights, QIE_ESTATE, "qie_cache_schedule: weights not set");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t D = h->D, L = h->cfg.num_layers, modN = L * 12 * D;
    cudaFree(h->d_sched_mod);
    cudaFree(h->d_sched_fin);
    h->d_sched_mod = h->d_sched_fin = nullptr;
    h->n_sched = 0;
    QIE_CUDA_OK(cudaMalloc(&h->d_sched_mod, (size_t)n * modN * sizeof(float)));
    QIE_CUDA_OK(cudaMalloc(&h->d_sched_fin, (size_t)n * 2 * D * sizeof(float)));
    float* tproj = h->d_small;
    float* t1 = tproj + 8 * 256;
    float* temb = t1 + 8 * D;
    float* d_t = temb + 8 * D;     // borrow the head of the modulation table for the timestep values
    for (int i = 0; i < n; ++i) {  // one entry at a time: the same kernels, batch 1 -> bit-identical to the uncached forward
        QIE_CUDA_OK(cudaMemcpyAsync(d_t, timesteps_host + i, sizeof(float), cudaMemcpyHostToDevice, st));
        int rc;
        if ((rc = qie_timestep_proj(d_t, tproj, 1, 0, st))) return rc;
        if ((rc = qie_gemv(tproj, h->w.t1_w, h->w.t1_b, t1, 1, D, 256, 0, st))) return rc;
        if ((rc = qie_gemv(t1, h->w.t2_w, h->w.t2_b, temb, 1, D, D, 1, st))) return rc;
        if ((rc = qie_gemv(temb, h->w.mod_w, h->w.mod_b, h->d_sched_mod + (size_t)i * modN, 1, (long long)modN, D, 1, st))) return rc;
        if ((rc = qie_gemv(temb, h->w.norm_out_w, h->w.norm_out_b, h->d_sched_fin + (size_t)i * 2 * D, 1, 2 * D, D, 1, st))) return rc;
        QIE_CUDA_OK(cudaStreamSynchronize(st));
    }
    h->n_sched = n;
    return QIE_OK;
}

extern "C" int qie_cache_prompt(qie_handle* h, int slot, const void* enc, int txt_rows, void* stream) {
    QIE_REQUIRE(h && enc && slot >= 0 && slot < 4 && txt_rows > 0, QIE_EINVAL, "qie_cache_prompt: bad argument");
    QIE_REQUIRE(h->has_weights, QIE_ESTATE, "qie_cache_prompt: weights not set");
    cudaStream_t st = (cudaStream_t)stream;
    qie_seq s;
    int rc = qie_make_seq(1, 128, txt_rows, &s);
    if (rc) return rc;
    const size_t D = h->D;
    cudaFree(h->d_prompt[slot]);
    h->d_prompt[slot] = nullptr;
    h->prompt_rows[slot] = 0;
    void* xtxt = nullptr;
    QIE_CUDA_OK(cudaMalloc(&h->d_prompt[slot], (size_t)s.txt_pad * D * sizeof(float)));
    QIE_CUDA_OK(cudaMalloc(&xtxt, (size_t)s.txt_pad * h->cfg.joint_dim * 2));
    rc = qie_rmsnorm_pack(enc, h->w.txt_norm_w, xtxt, 1, txt_rows, s.txt_pad, h->cfg.joint_dim, 1e-6f, st);
    if (!rc) {
        qie_gemm_args t{};
        t.a = xtxt; t.a_compact = 1; t.w[1] = h->w.txt_in_w; t.bias[1] = h->w.txt_in_b;
        t.out = h->d_prompt[slot]; t.out_compact = 1; t.ldo = (int)D; t.N = (int)D; t.K = h->cfg.joint_dim; t.streams = 2;
        t.epilogue = QIE_EPI_F32;
        rc = qie_gemm(&t, &s, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(xtxt);
    if (rc) return rc;
    QIE_CUDA_OK(e);
    h->prompt_rows[slot] = txt_rows;
    return QIE_OK;
}

// selection used by the following qie_forward calls: sched_idx_host[b] (or NULL / -1 = compute), prompt_slot (-1 = compute)
extern "C" int qie_cache_select(qie_handle* h, const int* sched_idx_host, int batch, int prompt_slot) {
    QIE_REQUIRE(h && batch >= 0 && batch <= 8 && prompt_slot >= -1 && prompt_slot < 4, QIE_EINVAL, "qie_cache_select: bad argument");
    for (int b = 0; b < 8; ++b) h->sel_sched[b] = (sched_idx_host && b < batch) ? sched_idx_host[b] : -1;
    h->sel_prompt = prompt_slot;
    return QIE_OK;
}
