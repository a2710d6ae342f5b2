// K-attn: joint image+text flash-style attention forward on tcgen05 (sm_100a), head_dim 128, non-causal.
// Replaces F.scaled_dot_product_attention(q, k, v) over the concatenated sequence inside
// QwenDoubleStreamAttnProcessor2_0; q,k arrive RMS-normed + roped (QKV GEMM epilogue).
//
// One CTA = one (batch, head) x 256 query rows = two 128-row Q tiles that ping-pong on the tensor pipe:
//   warp 0 / lane 0 : TMA producer (Q tiles once; K_j, V_j tiles through a 3-slot ring, 128 rows x 128 dims each)
//   warp 1 / lane 0 : MMA issuer   S_t = Q_t K_j^T (SS, K-major x K-major)  -> TMEM S_t   (128 x 128 fp32)
//                                  O_t += P_t V_j  (SS, P from smem, V MN-major) -> TMEM O_t (128 x 128 fp32)
//   warps 4..7      : softmax warpgroup of tile 0 (thread r owns query row r == TMEM lane r)
//   warps 8..11     : softmax warpgroup of tile 1
// TMEM: S0 | S1 | O0 | O1 = 4 x 128 columns.  Online softmax with lazy rescaling: the running reference max
// only moves when the row max grows by > 8 (log2 domain), so the O read-modify-write in TMEM is rare.
// KV rows in the padding of either stream are masked to -inf (each 128-row KV tile belongs to one stream).
// Roofline: tensor pipe (co-limited by MUFU ex2); algorithmic FLOPs = 4 * S^2 * 128 per (batch, head).
#include "common.cuh"

namespace qie {

constexpr int ATT_THREADS = 384;
constexpr int ATT_TILE = 128;                       // q rows per tile, kv rows per tile, head dim
constexpr int ATT_HALF_BYTES = ATT_TILE * 128;      // 128 rows x 64 bf16 (one swizzled half tile) = 16 KB
constexpr int ATT_TILE_BYTES = 2 * ATT_HALF_BYTES;  // 32 KB
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_SMEM = (2 + 2 + ATT_KV_STAGES) * ATT_TILE_BYTES + 256 + 1024;

struct AttnDev {
    qie_seq seq;
    int H;
    __nv_bfloat16* out;
    float scale_log2;        // softmax scale * log2(e)
    uint32_t v_lbo, v_sbo;   // V (MN-major) descriptor strides, bytes
    uint32_t v_kstep;        // byte advance of the V descriptor per 16 kv rows
};

__device__ __forceinline__ int kv_valid_rows(const qie_seq& s, int j) {
    const int r0 = j * ATT_TILE;
    if (r0 < s.img_pad) return min(ATT_TILE, s.img_rows - r0);
    return min(ATT_TILE, s.txt_rows - (r0 - s.img_pad));
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnDev p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                   // [2 tiles][2 halves][128 x 128 B]
    uint8_t* sP = smem + 2 * ATT_TILE_BYTES;              // [2 tiles][2 halves]
    uint8_t* sKV = smem + 4 * ATT_TILE_BYTES;             // [stages][2 halves]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (4 + ATT_KV_STAGES) * ATT_TILE_BYTES);
    uint64_t* q_full = bars;            // [1]
    uint64_t* kv_full = bars + 1;       // [3]
    uint64_t* kv_empty = bars + 4;      // [3]
    uint64_t* s_full = bars + 7;        // [2]  MMA -> softmax: S_t ready
    uint64_t* p_full = bars + 9;        // [2]  softmax -> MMA: P_t in smem, S_t consumed, O_t rescaled
    uint64_t* pv_done = bars + 11;      // [2]  MMA -> softmax: O_t += P_t V_j retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int rpb = p.seq.img_pad + p.seq.txt_pad;
    const int n_kv = rpb / ATT_TILE;
    const int head = blockIdx.y, b = blockIdx.z;
    const int q_row0 = blockIdx.x * 2 * ATT_TILE;          // row inside the batch element
    const bool tile1_on = q_row0 + ATT_TILE < rpb;
    const int D = p.H * ATT_TILE;
    const int colQ = head * ATT_TILE, colK = D + head * ATT_TILE, colV = 2 * D + head * ATT_TILE;
    const int row_base = b * rpb;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < ATT_KV_STAGES; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[t], 128);
            mbar_init(&pv_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            const int nt = tile1_on ? 2 : 1;
            mbar_expect_tx(q_full, nt * ATT_TILE_BYTES);
            for (int t = 0; t < nt; ++t)
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_2d(sQ + t * ATT_TILE_BYTES + hf * ATT_HALF_BYTES, &tmQKV, colQ + hf * 64,
                                row_base + q_row0 + t * ATT_TILE, q_full);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_kv; ++j) {
                for (int kv = 0; kv < 2; ++kv) {   // K_j then V_j
                    mbar_wait(&kv_empty[stage], phase ^ 1);
                    mbar_expect_tx(&kv_full[stage], ATT_TILE_BYTES);
                    const int col = kv ? colV : colK;
                    for (int hf = 0; hf < 2; ++hf)
                        tma_load_2d(sKV + stage * ATT_TILE_BYTES + hf * ATT_HALF_BYTES, &tmQKV, col + hf * 64,
                                    row_base + j * ATT_TILE, &kv_full[stage]);
                    if (++stage == ATT_KV_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, false);
            constexpr uint32_t IDESC_O = umma_idesc_bf16(128, 128, true);   // B = V is MN-major
            const int nt = tile1_on ? 2 : 1;
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&]() {
                if (++stage == ATT_KV_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            auto issue_S = [&](int t, int kstage) {   // S_t = Q_t K^T
                const uint32_t q = smem_u32(sQ + t * ATT_TILE_BYTES), k = smem_u32(sKV + kstage * ATT_TILE_BYTES);
#pragma unroll
                for (int s = 0; s < 8; ++s) {         // 8 x 16 head dims; half = s/4, 32 B steps inside the swizzled row
                    const uint32_t off = (s >> 2) * ATT_HALF_BYTES + (s & 3) * 32;
                    umma_ss_f16(tmem_base + t * 128, umma_desc_kmajor_sw128(q + off), umma_desc_kmajor_sw128(k + off),
                                IDESC_S, s ? 1u : 0u);
                }
                umma_commit(&s_full[t]);
            };
            auto issue_PV = [&](int t, int vstage, bool first) {   // O_t (+)= P_t V
                const uint32_t pp = smem_u32(sP + t * ATT_TILE_BYTES), v = smem_u32(sKV + vstage * ATT_TILE_BYTES);
#pragma unroll
                for (int s = 0; s < 8; ++s) {         // 8 x 16 kv rows
                    const uint32_t aoff = (s >> 2) * ATT_HALF_BYTES + (s & 3) * 32;
                    umma_ss_f16(tmem_base + 256 + t * 128, umma_desc_kmajor_sw128(pp + aoff),
                                umma_desc_mnmajor_sw128(v + s * p.v_kstep, p.v_lbo, p.v_sbo), IDESC_O,
                                (first && s == 0) ? 0u : 1u);
                }
                umma_commit(&pv_done[t]);
            };
            mbar_wait(q_full, 0);
            tc_fence_after();
            // prologue: S_t(0)
            mbar_wait(&kv_full[stage], phase);   // K_0
            tc_fence_after();
            for (int t = 0; t < nt; ++t) issue_S(t, stage);
            umma_commit(&kv_empty[stage]);
            advance();
            for (int j = 0; j < n_kv; ++j) {
                // V_j is the current ring slot, K_{j+1} the next
                const int vstage = stage;
                const uint32_t vphase = phase;
                advance();
                const int kstage = stage;
                const uint32_t kphase = phase;
                const bool more = j + 1 < n_kv;
                mbar_wait(&kv_full[vstage], vphase);
                tc_fence_after();
                for (int t = 0; t < nt; ++t) {
                    mbar_wait(&p_full[t], j & 1);          // P_t(j) written, S_t free, O_t rescaled
                    tc_fence_after();
                    issue_PV(t, vstage, j == 0);
                    if (t == nt - 1) umma_commit(&kv_empty[vstage]);
                    if (more) {
                        if (t == 0) {
                            mbar_wait(&kv_full[kstage], kphase);
                            tc_fence_after();
                        }
                        issue_S(t, kstage);
                        if (t == nt - 1) umma_commit(&kv_empty[kstage]);
                    }
                }
                if (more) advance();
            }
        }
    } else if (warp >= 4) {
        // ================= softmax warpgroups =================
        const int t = (warp - 4) >> 2;                 // tile handled by this warpgroup
        if (t == 0 || tile1_on) {
            const int quad = warp & 3;
            const int r = quad * 32 + lane;            // query row inside the tile == TMEM lane
            const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
            const uint32_t tS = tmem_base + lane_addr + t * 128;
            const uint32_t tO = tmem_base + lane_addr + 256 + t * 128;
            uint8_t* prow = sP + t * ATT_TILE_BYTES + r * 128;   // row r of each 64-column half
            const float c = p.scale_log2;
            float m_ref = -INFINITY, l = 0.f;
            for (int j = 0; j < n_kv; ++j) {
                const int nv = kv_valid_rows(p.seq, j);
                mbar_wait(&s_full[t], j & 1);
                tc_fence_after();
                // ---- pass 1: row max ----
                float mx = -INFINITY;
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t s[32];
                    tmem_ld32(tS + ch * 32, s);
                    tmem_ld_wait();
                    if (ch * 32 + 32 <= nv) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (ch * 32 + i < nv) mx = fmaxf(mx, __uint_as_float(s[i]));
                    }
                }
                mx *= c;
                float alpha = 1.f;
                const bool grow = mx > m_ref + 8.0f;   // also true on the first tile (m_ref = -inf)
                if (grow) {
                    alpha = exp2f(m_ref - mx);         // 0 on the first tile
                    m_ref = mx;
                    l *= alpha;
                }
                if (j > 0) {
                    // O_t and the P_t buffer are owned by the tensor pipe until PV(j-1) retires
                    mbar_wait(&pv_done[t], (j - 1) & 1);
                    tc_fence_after();
                    if (__any_sync(0xffffffffu, grow)) {
#pragma unroll 1
                        for (int ch = 0; ch < 4; ++ch) {
                            uint32_t o[32];
                            tmem_ld32(tO + ch * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st32(tO + ch * 32, o);
                        }
                        tmem_st_wait();
                    }
                }
                // ---- pass 2: p = exp2(s*c - m_ref), row sum, bf16 P into swizzled smem ----
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t s[32];
                    tmem_ld32(tS + ch * 32, s);
                    tmem_ld_wait();
                    float e[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float x = fast_exp2(fmaf(__uint_as_float(s[i]), c, -m_ref));
                        e[i] = (ch * 32 + i < nv) ? x : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) l += e[i];
                    // columns ch*32..+31 -> half (ch>>1), 16-byte chunks ((ch&1)*4 .. +3), XOR-swizzled by row&7
                    uint8_t* hrow = prow + (ch >> 1) * ATT_HALF_BYTES;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint4 u;
                        u.x = pack_bf16(e[q4 * 8 + 0], e[q4 * 8 + 1]);
                        u.y = pack_bf16(e[q4 * 8 + 2], e[q4 * 8 + 3]);
                        u.z = pack_bf16(e[q4 * 8 + 4], e[q4 * 8 + 5]);
                        u.w = pack_bf16(e[q4 * 8 + 6], e[q4 * 8 + 7]);
                        const int chunk = ((ch & 1) * 4 + q4) ^ (r & 7);
                        *reinterpret_cast<uint4*>(hrow + chunk * 16) = u;
                    }
                }
                fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor (async) proxy
                tc_fence_before();          // orders this thread's TMEM loads/stores before the arrive
                mbar_arrive(&p_full[t]);
            }
            // ---- epilogue: O / l -> bf16 -> global ----
            mbar_wait(&pv_done[t], (n_kv - 1) & 1);
            tc_fence_after();
            const int qrow = q_row0 + t * ATT_TILE + r;
            const float inv = 1.f / l;
            __nv_bfloat16* orow = p.out + (long long)(row_base + qrow) * D + head * ATT_TILE;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t o[32];
                tmem_ld32(tO + ch * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 u;
                    u.x = pack_bf16(__uint_as_float(o[q4 * 8 + 0]) * inv, __uint_as_float(o[q4 * 8 + 1]) * inv);
                    u.y = pack_bf16(__uint_as_float(o[q4 * 8 + 2]) * inv, __uint_as_float(o[q4 * 8 + 3]) * inv);
                    u.z = pack_bf16(__uint_as_float(o[q4 * 8 + 4]) * inv, __uint_as_float(o[q4 * 8 + 5]) * inv);
                    u.w = pack_bf16(__uint_as_float(o[q4 * 8 + 6]) * inv, __uint_as_float(o[q4 * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + ch * 32 + q4 * 8) = u;
                }
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace qie

using namespace qie;

extern "C" int qie_attn_fwd(const void* qkv, void* out, const qie_seq* seq, int num_heads, int variant, void* stream) {
    QIE_REQUIRE(qkv && out && seq, QIE_EINVAL, "qie_attn_fwd: null pointer");
    QIE_REQUIRE(seq->img_pad % 128 == 0 && seq->txt_pad % 128 == 0 && seq->batch > 0 && num_heads > 0 &&
                    seq->img_rows > 0 && seq->txt_rows > 0 && seq->img_rows > seq->img_pad - 128 &&
                    seq->txt_rows > seq->txt_pad - 128,
                QIE_ESHAPE, "qie_attn_fwd: bad sequence layout (every 128-row KV tile needs >= 1 valid row)");
    const int rpb = seq->img_pad + seq->txt_pad;
    const int D = num_heads * 128;
    CUtensorMap tm;
    int rc = make_tmap_2d(&tm, qkv, (uint64_t)seq->batch * rpb, (uint64_t)3 * D, (uint64_t)3 * D * 2, 128, 64, 2);
    if (rc) return rc;
    AttnDev p{};
    p.seq = *seq;
    p.H = num_heads;
    p.out = (__nv_bfloat16*)out;
    p.scale_log2 = 0.08838834764831845f * 1.4426950408889634f;   // 1/sqrt(128) * log2(e)
    // V tile in smem: two halves (64 dims each, 16 KB apart) of 128 kv rows x 128 B, 128B-swizzled by TMA.
    // MN-major canonical layout: 8 kv rows x 128 B = one 1024 B atom (SBO), next 64 dims LBO away.
    p.v_lbo = ATT_HALF_BYTES;
    p.v_sbo = 1024;
    p.v_kstep = 2048;
    if (variant == 1) {   // probe: swapped LBO/SBO roles
        p.v_lbo = 1024;
        p.v_sbo = ATT_HALF_BYTES;
    }
    static bool configured = false;
    if (!configured) {
        QIE_CUDA_OK(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured = true;
    }
    dim3 grid((rpb + 255) / 256, num_heads, seq->batch);
    attn_kernel<<<grid, ATT_THREADS, ATT_SMEM, (cudaStream_t)stream>>>(tm, p);
    QIE_LAUNCH_OK("attn_kernel");
    return QIE_OK;
}
