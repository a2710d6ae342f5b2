// K-gemm: persistent, warp-specialised tcgen05 GEMM for the MMDiT linears (sm_100a).
//
//   out[row, n] = epilogue( sum_k A[row, k] * W_stream(row)[n, k] + bias_stream(row)[n] )
//
// One launch covers BOTH streams of the dual-stream block (a two-problem grouped GEMM): image rows
// use W[0], text rows W[1]; every stream is padded to 128 rows so an M tile never straddles the two.
// Replaces, per QwenImageTransformerBlock: to_q/to_k/to_v + add_{q,k,v}_proj (N fused
// to 3D), to_out.0/to_add_out (+ gate*y + residual epilogue), img_mlp/txt_mlp net.0.proj (+GELU-tanh)
// and net.2 (+ gate*y + residual), plus img_in/txt_in/proj_out.
//
// Structure (192 threads, 1 CTA / SM, grid = min(tiles, #SM), static round-robin tile schedule):
//   warp 0 / lane 0 : TMA producer  — A tile [128 x 64] and W tile [BN x 64] (128B-swizzled rows) per stage
//   warp 1 / lane 0 : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16 128 x BN x 16, fp32 accum in TMEM
//   warps 2..5      : epilogue      — tcgen05.ld 32x32b -> registers -> fused math -> global
// Three pipelines: smem full/empty ring (TMA <-> MMA), 2 TMEM accumulator stages (MMA <-> epilogue, so the
// epilogue of tile i overlaps the main loop of tile i+1), and the tile loop.
// Roofline: tensor pipe; algorithmic FLOPs = 2*M*N*K.
#include "common.cuh"

namespace qie {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK_BYTES = 128;   // one 128B swizzle row: 64 bf16 or 128 e4m3
constexpr int GEMM_THREADS = 192;

struct GemmDev {
    qie_seq seq;
    int N, K;                // K in elements
    int n_blocks;            // N / BN
    int streams;             // bit mask
    int a_compact, out_compact;
    int ldo;
    void* out;
    const float* bias[2];
    const float* gate;
    long long gate_bstride, gate_sstride;
    const float* rope;
    const float* qk_norm_w[2][2];
    const float* a_scale;
    const float* w_scale[2];
    int model_dim;           // D (QKV epilogue: column block -> q/k/v)
};

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK_BYTES;
    static constexpr int B_BYTES = BN * GEMM_BK_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // +1024: manual 1 KB alignment
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;             // two accumulator stages
};

// decode the m-block index into (batch, stream, tile-in-stream)
struct MBlock {
    int b, s, ti;
};
__device__ __forceinline__ MBlock decode_mblock(const GemmDev& p, int mb) {
    const int t0 = (p.streams & 1) ? p.seq.img_pad / GEMM_BM : 0;
    const int t1 = (p.streams & 2) ? p.seq.txt_pad / GEMM_BM : 0;
    MBlock r;
    r.b = mb / (t0 + t1);
    const int rem = mb % (t0 + t1);
    r.s = rem >= t0 ? 1 : 0;
    r.ti = r.s ? rem - t0 : rem;
    return r;
}

template <int EPI>
__device__ __forceinline__ void store_chunk(const GemmDev& p, float (&v)[32], long long orow, int n0, bool valid) {
    if constexpr (EPI == QIE_EPI_BF16 || EPI == QIE_EPI_GELU_BF16 || EPI == QIE_EPI_QKV_NORM_ROPE) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + n0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 u;
            if (valid) {
                u.x = pack_bf16(v[i * 8 + 0], v[i * 8 + 1]);
                u.y = pack_bf16(v[i * 8 + 2], v[i * 8 + 3]);
                u.z = pack_bf16(v[i * 8 + 4], v[i * 8 + 5]);
                u.w = pack_bf16(v[i * 8 + 6], v[i * 8 + 7]);
            } else {
                u = make_uint4(0, 0, 0, 0);
            }
            *reinterpret_cast<uint4*>(o + i * 8) = u;
        }
    } else {
        float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 f = valid ? make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3])
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(o + i * 4) = f;
        }
    }
}

template <int BN, int EPI, bool FP8>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
            const __grid_constant__ CUtensorMap tmB1, const GemmDev p) {
    using S = GemmSmem<BN>;
    constexpr int STAGES = S::STAGES;
    constexpr int BK = FP8 ? 128 : 64;           // elements per k-block
    constexpr uint32_t IDESC = FP8 ? umma_idesc_e4m3(GEMM_BM, BN) : umma_idesc_bf16(GEMM_BM, BN);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* full_bar = bars;                   // [STAGES]
    uint64_t* empty_bar = bars + STAGES;         // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;     // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;// [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int t0 = (p.streams & 1) ? p.seq.img_pad / GEMM_BM : 0;
    const int t1 = (p.streams & 2) ? p.seq.txt_pad / GEMM_BM : 0;
    const int m_blocks = p.seq.batch * (t0 + t1);
    const int num_tiles = m_blocks * p.n_blocks;
    const int k_blocks = (p.K + BK - 1) / BK;
    const int rpb = p.seq.img_pad + p.seq.txt_pad;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB0);
        tma_prefetch_desc(&tmB1);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<S::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mb = tile / p.n_blocks, nb = tile % p.n_blocks;
                const MBlock m = decode_mblock(p, mb);
                const int seg_pad = m.s ? p.seq.txt_pad : p.seq.img_pad;
                const int a_row = p.a_compact ? m.b * seg_pad + m.ti * GEMM_BM
                                              : m.b * rpb + (m.s ? p.seq.img_pad : 0) + m.ti * GEMM_BM;
                const CUtensorMap* tmB = m.s ? &tmB1 : &tmB0;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::STAGE_BYTES;
                    uint8_t* sb = sa + S::A_BYTES;
                    mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                    tma_load_2d(sa, &tmA, kb * BK, a_row, &full_bar[stage]);
                    tma_load_2d(sb, tmB, kb * BK, nb * BN, &full_bar[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint64_t da = umma_desc_kmajor_sw128(sa);
                    const uint64_t db = umma_desc_kmajor_sw128(sa + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // 4 x 32 B along the swizzled 128 B row
                        if constexpr (FP8)
                            umma_ss_f8(d_tmem, da + 2 * k, db + 2 * k, IDESC, (kb | k) ? 1u : 0u);
                        else
                            umma_ss_f16(d_tmem, da + 2 * k, db + 2 * k, IDESC, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[acc]);         // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ================= epilogue (warps 2..5) =================
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int mb = tile / p.n_blocks, nb = tile % p.n_blocks;
            const MBlock m = decode_mblock(p, mb);
            const int seg_pad = m.s ? p.seq.txt_pad : p.seq.img_pad;
            const int seg_rows = m.s ? p.seq.txt_rows : p.seq.img_rows;
            const int local = m.ti * GEMM_BM + quad * 32 + lane;
            const bool valid = local < seg_rows;
            const int jrow_in_batch = (m.s ? p.seq.img_pad : 0) + local;     // row inside the joint layout
            const long long jrow = (long long)m.b * rpb + jrow_in_batch;
            const long long orow = p.out_compact ? (long long)m.b * seg_pad + local : jrow;
            const long long arow = p.a_compact ? (long long)m.b * seg_pad + local : jrow;
            const float* bias = p.bias[m.s];
            float a_sc = 1.f;
            if constexpr (FP8) a_sc = p.a_scale[arow];

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;

            auto load_chunk = [&](int c, float (&v)[32]) {
                uint32_t r[32];
                tmem_ld32(t_addr + c * 32, r);
                tmem_ld_wait();
                const int n0 = nb * BN + c * 32;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 bv = bias ? *reinterpret_cast<const float4*>(bias + n0 + i) : make_float4(0, 0, 0, 0);
                    if constexpr (FP8) {
                        const float4 ws = *reinterpret_cast<const float4*>(p.w_scale[m.s] + n0 + i);
                        v[i + 0] = __uint_as_float(r[i + 0]) * (a_sc * ws.x) + bv.x;
                        v[i + 1] = __uint_as_float(r[i + 1]) * (a_sc * ws.y) + bv.y;
                        v[i + 2] = __uint_as_float(r[i + 2]) * (a_sc * ws.z) + bv.z;
                        v[i + 3] = __uint_as_float(r[i + 3]) * (a_sc * ws.w) + bv.w;
                    } else {
                        v[i + 0] = __uint_as_float(r[i + 0]) + bv.x;
                        v[i + 1] = __uint_as_float(r[i + 1]) + bv.y;
                        v[i + 2] = __uint_as_float(r[i + 2]) + bv.z;
                        v[i + 3] = __uint_as_float(r[i + 3]) + bv.w;
                    }
                }
            };

            if constexpr (EPI == QIE_EPI_QKV_NORM_ROPE) {
                // a BN-wide tile holds BN/128 whole heads of exactly one of q / k / v (D % BN == 0)
                const int which = (nb * BN) / p.model_dim;    // 0 q, 1 k, 2 v
#pragma unroll 1
                for (int hh = 0; hh < BN / 128; ++hh) {
                    float v[32];
                    if (which == 2) {
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            load_chunk(hh * 4 + c, v);
                            store_chunk<EPI>(p, v, orow, nb * BN + hh * 128 + c * 32, valid);
                        }
                    } else {
                        float ss = 0.f;
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            load_chunk(hh * 4 + c, v);
#pragma unroll
                            for (int i = 0; i < 32; ++i) ss += v[i] * v[i];
                        }
                        const float rinv = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
                        const float* nw = p.qk_norm_w[m.s][which];
                        const float* rp = p.rope + (long long)jrow_in_batch * 128;
#pragma unroll 1
                        for (int c = 0; c < 4; ++c) {
                            load_chunk(hh * 4 + c, v);
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 w = *reinterpret_cast<const float4*>(nw + c * 32 + i);
                                const float4 cs = valid ? *reinterpret_cast<const float4*>(rp + c * 32 + i)
                                                        : make_float4(1.f, 0.f, 1.f, 0.f);
                                const float x0 = v[i] * rinv * w.x, x1 = v[i + 1] * rinv * w.y;
                                const float x2 = v[i + 2] * rinv * w.z, x3 = v[i + 3] * rinv * w.w;
                                v[i] = x0 * cs.x - x1 * cs.y;
                                v[i + 1] = x0 * cs.y + x1 * cs.x;
                                v[i + 2] = x2 * cs.z - x3 * cs.w;
                                v[i + 3] = x2 * cs.w + x3 * cs.z;
                            }
                            store_chunk<EPI>(p, v, orow, nb * BN + hh * 128 + c * 32, valid);
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    float v[32];
                    load_chunk(c, v);
                    const int n0 = nb * BN + c * 32;
                    if constexpr (EPI == QIE_EPI_GELU_BF16) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
                        store_chunk<EPI>(p, v, orow, n0, valid);
                    } else if constexpr (EPI == QIE_EPI_GATE_RESID_F32) {
                        if (valid) {
                            float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n0;
                            const float* g = p.gate + m.b * p.gate_bstride + m.s * p.gate_sstride + n0;
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                float4 r4 = *reinterpret_cast<float4*>(o + i);
                                const float4 g4 = *reinterpret_cast<const float4*>(g + i);
                                r4.x += g4.x * v[i];
                                r4.y += g4.y * v[i + 1];
                                r4.z += g4.z * v[i + 2];
                                r4.w += g4.w * v[i + 3];
                                *reinterpret_cast<float4*>(o + i) = r4;
                            }
                        }
                    } else {
                        store_chunk<EPI>(p, v, orow, n0, valid);
                    }
                }
            }
            // release this accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<S::TMEM_COLS>(tmem_base);
    }
}

template <int BN, int EPI, bool FP8>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB0, const CUtensorMap& tmB1, const GemmDev& p,
                       int num_tiles, cudaStream_t st) {
    using S = GemmSmem<BN>;
    static bool configured = false;
    if (!configured) {
        QIE_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, EPI, FP8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::TOTAL));
        configured = true;
    }
    int grid = sm_count();
    if (grid > num_tiles) grid = num_tiles;
    gemm_kernel<BN, EPI, FP8><<<grid, GEMM_THREADS, S::TOTAL, st>>>(tmA, tmB0, tmB1, p);
    QIE_LAUNCH_OK("gemm_kernel");
    return QIE_OK;
}

template <int BN, bool FP8>
static int dispatch_epi(int epi, const CUtensorMap& a, const CUtensorMap& b0, const CUtensorMap& b1, const GemmDev& p,
                        int tiles, cudaStream_t st) {
    switch (epi) {
        case QIE_EPI_BF16: return launch_gemm<BN, QIE_EPI_BF16, FP8>(a, b0, b1, p, tiles, st);
        case QIE_EPI_GELU_BF16: return launch_gemm<BN, QIE_EPI_GELU_BF16, FP8>(a, b0, b1, p, tiles, st);
        case QIE_EPI_F32: return launch_gemm<BN, QIE_EPI_F32, FP8>(a, b0, b1, p, tiles, st);
        case QIE_EPI_GATE_RESID_F32: return launch_gemm<BN, QIE_EPI_GATE_RESID_F32, FP8>(a, b0, b1, p, tiles, st);
        case QIE_EPI_QKV_NORM_ROPE:
            if constexpr (BN >= 128) return launch_gemm<BN, QIE_EPI_QKV_NORM_ROPE, FP8>(a, b0, b1, p, tiles, st);
    }
    set_error("qie_gemm: unsupported epilogue %d for block_n %d", epi, BN);
    return QIE_EINVAL;
}

}  // namespace qie

using namespace qie;

extern "C" int qie_gemm(const qie_gemm_args* g, const qie_seq* seq, void* stream) {
    QIE_REQUIRE(g && seq && g->a && g->out, QIE_EINVAL, "qie_gemm: null pointer");
    QIE_REQUIRE(g->streams >= 1 && g->streams <= 3, QIE_EINVAL, "qie_gemm: streams mask must be 1..3");
    QIE_REQUIRE(seq->img_pad % 128 == 0 && seq->txt_pad % 128 == 0 && seq->img_pad >= seq->img_rows &&
                    seq->txt_pad >= seq->txt_rows && seq->batch > 0,
                QIE_ESHAPE, "qie_gemm: bad sequence layout");
    const bool compact = g->a_compact || g->out_compact;
    QIE_REQUIRE(!compact || g->streams == 1 || g->streams == 2, QIE_EINVAL,
                "qie_gemm: compact A/out needs exactly one stream enabled");
    const int eb = g->fp8 ? 1 : 2;
    const int bk = g->fp8 ? 128 : 64;
    QIE_REQUIRE(g->K > 0 && (g->K * eb) % 16 == 0, QIE_ESHAPE, "qie_gemm: K=%d row stride must be 16 B aligned", g->K);
    int bn = g->block_n;
    if (bn == 0) bn = g->N % 256 == 0 ? 256 : (g->N % 128 == 0 ? 128 : 64);
    QIE_REQUIRE((bn == 64 || bn == 128 || bn == 256) && g->N % bn == 0, QIE_ESHAPE,
                "qie_gemm: N=%d not a multiple of block_n=%d", g->N, bn);
    for (int s = 0; s < 2; ++s)
        if (g->streams & (1 << s)) {
            QIE_REQUIRE(g->w[s], QIE_EINVAL, "qie_gemm: weight of stream %d is null", s);
            QIE_REQUIRE((s ? seq->txt_pad : seq->img_pad) > 0, QIE_ESHAPE, "qie_gemm: stream %d enabled but empty", s);
            if (g->fp8) QIE_REQUIRE(g->w_scale[s] && g->a_scale, QIE_EINVAL, "qie_gemm: fp8 needs scales");
        }
    if (g->epilogue == QIE_EPI_GATE_RESID_F32) QIE_REQUIRE(g->gate, QIE_EINVAL, "qie_gemm: gate is null");

    GemmDev p{};
    p.seq = *seq;
    p.N = g->N;
    p.K = g->K;
    p.n_blocks = g->N / bn;
    p.streams = g->streams;
    p.a_compact = g->a_compact;
    p.out_compact = g->out_compact;
    p.ldo = g->ldo;
    p.out = g->out;
    p.gate = g->gate;
    p.gate_bstride = g->gate_bstride;
    p.gate_sstride = g->gate_sstride;
    p.rope = g->rope;
    p.a_scale = g->a_scale;
    for (int s = 0; s < 2; ++s) {
        p.bias[s] = g->bias[s];
        p.w_scale[s] = g->w_scale[s];
        for (int k = 0; k < 2; ++k) p.qk_norm_w[s][k] = g->qk_norm_w[s][k];
    }
    p.model_dim = g->N / 3;
    if (g->epilogue == QIE_EPI_QKV_NORM_ROPE) {
        QIE_REQUIRE(g->N % 3 == 0 && p.model_dim % bn == 0 && bn >= 128 && g->rope, QIE_ESHAPE,
                    "qie_gemm: QKV epilogue needs N=3D, D %% block_n == 0, block_n>=128, rope table");
        for (int s = 0; s < 2; ++s)
            if (g->streams & (1 << s))
                QIE_REQUIRE(p.qk_norm_w[s][0] && p.qk_norm_w[s][1], QIE_EINVAL, "qie_gemm: qk norm weights null");
    }

    const int rpb = seq->img_pad + seq->txt_pad;
    const int only = g->streams == 2 ? 1 : 0;
    const long long a_rows = g->a_compact ? (long long)seq->batch * (only ? seq->txt_pad : seq->img_pad)
                                          : (long long)seq->batch * rpb;
    CUtensorMap tmA, tmB[2];
    int rc = make_tmap_2d(&tmA, g->a, (uint64_t)a_rows, (uint64_t)g->K, (uint64_t)g->K * eb, GEMM_BM, bk, eb);
    if (rc) return rc;
    for (int s = 0; s < 2; ++s) {
        const void* w = g->w[s] ? g->w[s] : g->w[1 - s];
        rc = make_tmap_2d(&tmB[s], w, (uint64_t)g->N, (uint64_t)g->K, (uint64_t)g->K * eb, bn, bk, eb);
        if (rc) return rc;
    }
    const int t0 = (g->streams & 1) ? seq->img_pad / 128 : 0, t1 = (g->streams & 2) ? seq->txt_pad / 128 : 0;
    const int tiles = seq->batch * (t0 + t1) * p.n_blocks;
    cudaStream_t st = (cudaStream_t)stream;
    if (g->fp8) {
        switch (bn) {
            case 64: return dispatch_epi<64, true>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st);
            case 128: return dispatch_epi<128, true>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st);
            default: return dispatch_epi<256, true>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st);
        }
    }
    switch (bn) {
        case 64: return dispatch_epi<64, false>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st);
        case 128: return dispatch_epi<128, false>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st);
        default: return dispatch_epi<256, false>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st);
    }
}
