// K-attn: joint image+text flash-style attention forward on tcgen05 (sm_100a), head_dim 128, non-causal.
// Replaces F.scaled_dot_product_attention(q, k, v) over the concatenated sequence inside
// QwenDoubleStreamAttnProcessor2_0; q,k arrive RMS-normed + roped (QKV GEMM epilogue).
//
// Two kernels: attn_pair_kernel (default: a CTA pair per 256 query rows, cta_group::2 MMAs, 256-wide KV tiles, P in TMEM) and
// attn_kernel (single-CTA fallback, variant bit 0x8; also the A/B yard-stick).  The kernels that lost the round-1 A/B runs
// (first pair kernels, persistent / speculative / traced builds; the round-2 ping-pong kernel with two Q super-tiles per pair; the
// bounded-score kernel with S delivered in two staggered column halves, with and without a MUFU hand-over between partner warps)
// live in the git history (commits 550b982 and before; "Ping-pong CTA-pair attention kernel"; "staggered bounded-score kernel"),
// measurements in profiles/.
//
// attn_kernel: one CTA = one (batch, head) x 256 query rows = two 128-row Q tiles that ping-pong on the tensor pipe:
//   warp 0 / lane 0 : TMA producer (Q tiles once; K_j, V_j tiles through a ring, 128 rows x 128 dims each)
//   warp 1 / lane 0 : MMA issuer   S_t = Q_t K_j^T (SS, K-major x K-major)          -> TMEM S_t (128 x 128 fp32)
//                                  O_t += P_t V_j  (P_t from TMEM [TS] or smem [SS], V MN-major) -> TMEM O_t
//   warps 4..7      : softmax warpgroup of tile 0 (thread r owns query row r == TMEM lane r)
//   warps 8..11     : softmax warpgroup of tile 1
// TMEM: S0 | S1 | O0 | O1 = 4 x 128 columns; the bf16 P_t tile aliases the first 64 columns of S_t (the MMA pipe runs
// PV_t(j) before S_t(j+1) in issue order, so the alias is safe).  Online softmax with lazy rescaling: the reference max
// only moves when the row max grows by > 8 (log2 domain), so the O read-modify-write in TMEM is rare.  The softmax
// inner loop uses packed fp32x2 FFMA2/FADD2 and 3-input FMNMX3; exp2 runs on the MUFU.
// KV rows in the padding of either stream are masked to -inf (each 128-row KV tile belongs to one stream).
// Roofline: tensor pipe (co-limited by MUFU ex2); algorithmic FLOPs = 4 * S^2 * 128 per (batch, head).
#include "common.cuh"

namespace qie {

constexpr int ATT_THREADS = 384;
constexpr int ATT_TILE = 128;                       // q rows per tile, kv rows per tile, head dim
constexpr int ATT_HALF_BYTES = ATT_TILE * 128;      // 128 rows x 64 bf16 (one swizzled half tile) = 16 KB
constexpr int ATT_TILE_BYTES = 2 * ATT_HALF_BYTES;  // 32 KB

template <bool P_TMEM>
struct AttCfg {
    static constexpr int KV_STAGES = P_TMEM ? 5 : 3;                  // P in smem costs two 32 KB tiles
    static constexpr int P_TILES = P_TMEM ? 0 : 2;
    static constexpr int SMEM = (2 + P_TILES + KV_STAGES) * ATT_TILE_BYTES + 256 + 1024;
};

struct AttnDev {
    qie_seq seq;
    int H;
    __nv_bfloat16* out;
    float scale_log2;        // softmax scale * log2(e)
    uint32_t v_lbo, v_sbo;   // V (MN-major) descriptor strides, bytes
    uint32_t v_kstep;        // byte advance of the V descriptor per 16 kv rows
    const int* tile_valid;   // optional: valid rows per 128-row KV tile (gathered sequence-parallel layout); NULL = from seq
    // sequence-parallel scatter of the output (pair kernel only; qie_peers).  The sequence is the gathered layout
    // [rank 0 image shard (sp_img_pad rows) | ... | rank P-1 image shard | all text tokens]; query row qg of batch b belongs to
    //   image: rank s = qg / sp_img_pad, local row qg % sp_img_pad
    //   text : token t = qg - sp_img_region, owned by rank s (the first sp_txt_rem ranks own sp_txt_base + 1 tokens), local row
    //          sp_img_pad + (t - first token of s)
    // and goes to peer_out[s] [batch][sp_rows_pad][out_ld] at head column (head_off + head) * 128
    void* const* peer_out;
    int sp_img_pad, sp_img_region, sp_rows_pad, sp_txt_total, sp_txt_base, sp_txt_rem;
    int out_ld, head_off;
};

// valid rows of 128-row KV tile j of batch element b (0 for a text tile behind the element's own text length)
__device__ __forceinline__ int kv_valid_rows(const qie_seq& s, int j, int b) {
    const int r0 = j * ATT_TILE;
    if (r0 < s.img_pad) return min(ATT_TILE, s.img_rows - r0);
    return max(0, min(ATT_TILE, s.txt_rows_b[b] - (r0 - s.img_pad)));
}

// packed fp32x2 helpers (Blackwell FFMA2 / FADD2)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// row max over the 128 scores of this thread's row; FULL = every kv column valid (no per-element predicate)
template <bool FULL>
__device__ __forceinline__ float row_max(uint32_t tS, int nv) {
    uint32_t sa[32], sb[32];
    float m0 = -INFINITY, m1 = -INFINITY;
    tmem_ld32(tS, sa);
    tmem_ld_wait();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        uint32_t(&cur)[32] = (ch & 1) ? sb : sa;
        uint32_t(&nxt)[32] = (ch & 1) ? sa : sb;
        if (ch < 3) tmem_ld32(tS + (ch + 1) * 32, nxt);
        if constexpr (FULL) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                m0 = max3(m0, __uint_as_float(cur[i]), __uint_as_float(cur[i + 1]));
                m1 = max3(m1, __uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (ch * 32 + i < nv) m0 = fmaxf(m0, __uint_as_float(cur[i]));
        }
        if (ch < 3) tmem_ld_wait();
    }
    return fmaxf(m0, m1);
}

// p = exp2(s*c - m), accumulate the row sum (packed fp32x2), emit bf16 P either as packed words (TMEM) or into the
// 128B-swizzled smem tile (row r: 16-byte chunk index XOR (r & 7))
template <bool FULL, bool P_TMEM, int POLY>
__device__ __forceinline__ void softmax_pass2(uint32_t tS, int nv, uint64_t c2, uint64_t nm2, uint64_t& l2,
                                              uint32_t (&pw)[64], uint8_t* prow, int r) {
    uint32_t sa[32], sb[32];
    tmem_ld32(tS, sa);
    tmem_ld_wait();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        uint32_t(&cur)[32] = (ch & 1) ? sb : sa;
        uint32_t(&nxt)[32] = (ch & 1) ? sa : sb;
        if (ch < 3) tmem_ld32(tS + (ch + 1) * 32, nxt);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const uint64_t X = fma2(pk2u(cur[i], cur[i + 1]), c2, nm2);
            float e0, e1;
            if (FULL && ((i >> 1) & 7) < POLY) {
                // FMA-pipe exp2 (the MUFU is the co-bottleneck of this kernel): x = n + f, n = round(x), |f| <= 1/2,
                // 2^f by a degree-3 minimax polynomial (rel err 7.5e-5 << bf16 rounding of P), 2^n via the exponent bits
                float x0, x1;
                upk2(X, x0, x1);
                const uint64_t Xc = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
                const uint64_t T = add2(Xc, pk2(12582912.f, 12582912.f));          // 1.5 * 2^23: low mantissa bits = n
                const uint64_t N = add2(T, pk2(-12582912.f, -12582912.f));
                const uint64_t Fr = fma2(N, pk2(-1.f, -1.f), Xc);
                uint64_t P = fma2(Fr, pk2(0.0551716574f, 0.0551716574f), pk2(0.2426111400f, 0.2426111400f));
                P = fma2(P, Fr, pk2(0.6932609677f, 0.6932609677f));
                P = fma2(P, Fr, pk2(0.9999280572f, 0.9999280572f));
                float t0, t1, p0, p1;
                upk2(T, t0, t1);
                upk2(P, p0, p1);
                e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
                e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
            } else {
                float x0, x1;
                upk2(X, x0, x1);
                if constexpr (!FULL) {
                    if (ch * 32 + i >= nv) x0 = -INFINITY;
                    if (ch * 32 + i + 1 >= nv) x1 = -INFINITY;
                }
                e0 = fast_exp2(x0);
                e1 = fast_exp2(x1);
            }
            l2 = add2(l2, pk2(e0, e1));
            w[i >> 1] = pack_bf16(e0, e1);
        }
        if constexpr (P_TMEM) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pw[ch * 16 + i] = w[i];
        } else {
            uint8_t* hrow = prow + (ch >> 1) * ATT_HALF_BYTES;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int chunk = ((ch & 1) * 4 + q4) ^ (r & 7);
                *reinterpret_cast<uint4*>(hrow + chunk * 16) =
                    make_uint4(w[q4 * 4], w[q4 * 4 + 1], w[q4 * 4 + 2], w[q4 * 4 + 3]);
            }
        }
        if (ch < 3) tmem_ld_wait();
    }
}

template <bool P_TMEM, int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnDev p) {
    using C = AttCfg<P_TMEM>;
    constexpr int KV_STAGES = C::KV_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                           // [2 tiles][2 halves][128 x 128 B]
    uint8_t* sP = smem + 2 * ATT_TILE_BYTES;                      // [2 tiles][2 halves]   (SS variant only)
    uint8_t* sKV = smem + (2 + C::P_TILES) * ATT_TILE_BYTES;      // [stages][2 halves]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + C::P_TILES + KV_STAGES) * ATT_TILE_BYTES);
    uint64_t* q_full = bars;                        // [1]
    uint64_t* kv_full = bars + 1;                   // [KV_STAGES]
    uint64_t* kv_empty = bars + 1 + KV_STAGES;      // [KV_STAGES]
    uint64_t* s_full = bars + 1 + 2 * KV_STAGES;    // [2]  MMA -> softmax: S_t ready (implies PV_t of the previous tile retired)
    uint64_t* p_full = s_full + 2;                  // [2]  softmax -> MMA: P_t written, S_t consumed, O_t rescaled
    uint64_t* pv_done = s_full + 4;                 // [2]  MMA -> softmax: O_t += P_t V_j retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);

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
        for (int i = 0; i < KV_STAGES; ++i) {
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
                    if (++stage == KV_STAGES) {
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
                if (++stage == KV_STAGES) {
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
                const uint32_t v = smem_u32(sKV + vstage * ATT_TILE_BYTES);
#pragma unroll
                for (int s = 0; s < 8; ++s) {         // 8 x 16 kv rows
                    const uint64_t vdesc = umma_desc_mnmajor_sw128(v + s * p.v_kstep, p.v_lbo, p.v_sbo);
                    const uint32_t accum = (first && s == 0) ? 0u : 1u;
                    if constexpr (P_TMEM) {
                        // A = P_t from TMEM: 16 bf16 of K per step = 8 packed 32-bit columns
                        umma_ts_f16(tmem_base + 256 + t * 128, tmem_base + t * 128 + s * 8, vdesc, IDESC_O, accum);
                    } else {
                        const uint32_t pp = smem_u32(sP + t * ATT_TILE_BYTES);
                        const uint32_t aoff = (s >> 2) * ATT_HALF_BYTES + (s & 3) * 32;
                        umma_ss_f16(tmem_base + 256 + t * 128, umma_desc_kmajor_sw128(pp + aoff), vdesc, IDESC_O, accum);
                    }
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
                        issue_S(t, kstage);                // in-order after PV_t(j): may overwrite the aliased P_t
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
            [[maybe_unused]] uint8_t* prow = sP + t * ATT_TILE_BYTES + r * 128;   // row r of each 64-column half
            const float c = p.scale_log2;
            const uint64_t c2 = pk2(c, c);
            float m_ref = -INFINITY;
            uint64_t l2 = pk2(0.f, 0.f);
            for (int j = 0; j < n_kv; ++j) {
                const int nv = p.tile_valid ? __ldg(p.tile_valid + j) : kv_valid_rows(p.seq, j, b);
                const bool full = nv == ATT_TILE;
                mbar_wait(&s_full[t], j & 1);
                tc_fence_after();
                // ---- pass 1: row max (TMEM loads software-pipelined: chunk ch+1 is in flight while ch is reduced) ----
                float mx = full ? row_max<true>(tS, nv) : row_max<false>(tS, nv);
                mx *= c;
                float alpha = 1.f;
                const bool grow = mx > m_ref + 8.0f;   // also true on the first tile (m_ref = -inf)
                if (grow) {
                    alpha = fast_exp2(m_ref - mx);     // 0 on the first tile
                    m_ref = mx;
                    l2 = fma2(l2, pk2(alpha, alpha), pk2(0.f, 0.f));
                }
                if (j > 0 && __any_sync(0xffffffffu, grow)) {
                    // s_full(j) was committed after PV_t(j-1) in issue order, so O_t is quiescent here
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
                // ---- pass 2: p = exp2(s*c - m_ref), row sum, bf16 P ----
                const uint64_t nm2 = pk2(-m_ref, -m_ref);
                [[maybe_unused]] uint32_t pw[64];      // P row as 64 packed bf16x2 words (TMEM variant)
                if (full) softmax_pass2<true, P_TMEM, POLY>(tS, nv, c2, nm2, l2, pw, prow, r);
                else softmax_pass2<false, P_TMEM, POLY>(tS, nv, c2, nm2, l2, pw, prow, r);
                if constexpr (P_TMEM) {
                    // all S columns of this row are consumed: overwrite S_t[0,64) with the packed P row
                    uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&pw[0]);
                    uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&pw[32]);
                    tmem_st32(tS, lo);
                    tmem_st32(tS + 32, hi);
                    tmem_st_wait();
                } else {
                    if (j > 0) {   // the P_t smem tile is read by PV_t(j-1); s_full(j) implies it retired, nothing to wait
                    }
                    fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor (async) proxy
                }
                tc_fence_before();          // orders this thread's TMEM loads/stores before the arrive
                mbar_arrive(&p_full[t]);
            }
            // ---- epilogue: O / l -> bf16 -> global ----
            mbar_wait(&pv_done[t], (n_kv - 1) & 1);
            tc_fence_after();
            const int qrow = q_row0 + t * ATT_TILE + r;
            float l_lo, l_hi;
            upk2(l2, l_lo, l_hi);
            const float inv = 1.f / (l_lo + l_hi);
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

template <bool P_TMEM, int POLY>
static int launch_attn(const CUtensorMap& tm, const AttnDev& p, dim3 grid, cudaStream_t st) {
    QIE_CONFIGURE_ONCE(cudaFuncSetAttribute(attn_kernel<P_TMEM, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         AttCfg<P_TMEM>::SMEM));
    attn_kernel<P_TMEM, POLY><<<grid, ATT_THREADS, AttCfg<P_TMEM>::SMEM, st>>>(tm, p);
    QIE_LAUNCH_OK("attn_kernel");
    return QIE_OK;
}

// =====================================================================================================================
// CTA-pair attention with 256-wide KV tiles (default kernel).  The round-1 traces showed the tensor pipe's operand ingest to be
// the limiter of 128-wide tiles: a 256 x 128 x 16 SS MMA needs 4 KB of A and 4 KB of B per SM every 64 cycles = 128 B/clk.
// This kernel raises the arithmetic intensity of both MMAs:
//   S = Q K_j^T   : one 256 x 256 x 128 SS MMA group per 256 KV rows (96 B/clk of operands, the GEMM main-loop shape)
//   O += P_j V_j  : P (bf16) lives in TMEM and is the A operand (TS MMA), so only V comes from shared memory (64 B/clk)
// TMEM: O [0,128) | S [128,384) fp32 | P [384,512) bf16x2.  One softmax stream: both warpgroups work on the SAME KV tile,
// warpgroup w on score columns [128 w, 128 w + 128); they agree on the row max through shared memory once per tile (so
// both halves of P use one reference) and keep separate partial row sums.  S is pulled into registers at once and handed
// back (s_free), so S(j+1) runs under the exponentials of tile j; PV(j) runs under the softmax of tile j+1.
// Issue order of the leader: S(0) | S(1) PV(0) | S(2) PV(1) | ...
// FIXED = bounded-score form (variant 0x200; the default wherever the QK-RMSNorm weights of the block bound the scores, see
// qie_attn_score_bound): q arrives multiplied by softmax_scale * log2(e) and p = 2^s is taken without any reference — no row
// max, no exchange between the warpgroups, no O rescale.  FIXED = false is the online softmax with lazy rescaling.
// =====================================================================================================================
constexpr int AT5_KSTAGES = 3, AT5_VSTAGES = 2;
constexpr int AT5_THREADS = 384;              // warps 0-7 softmax (two warpgroups), warp 8 TMA, warp 9 S issuer, warp 10 PV issuer, 11 idle
constexpr int AT5_STAGE_BYTES = 32 * 1024;     // K: my 128 kv rows x 128 dims; V: 256 kv rows x my 64 dims
constexpr int AT5_SMEM = ATT_TILE_BYTES + (AT5_KSTAGES + AT5_VSTAGES) * AT5_STAGE_BYTES + 2048 + 1024 + 512 + 1024;

// Epilogue shared by the CTA-pair kernels: O / (l_wg0 + l_wg1) -> bf16 -> global; warpgroup wg stores head dims [64 wg, 64 wg + 64)
// of its 32 rows.  With qie_peers installed every query row goes to the attention buffer of the rank that owns its token.
__device__ __forceinline__ void pair_epilogue(const AttnDev& p, uint64_t l2, float* xl, uint64_t* pv_done, uint32_t pv_parity,
                                              uint32_t tO, int wg, int quad, int r, bool q_valid, int q_row0, int row_base, int D,
                                              int head, int b) {
        float l_lo, l_hi;
        upk2(l2, l_lo, l_hi);
        xl[wg * 128 + r] = l_lo + l_hi;
        mbar_wait(pv_done, pv_parity);
        tc_fence_after();
        named_bar_sync(1 + quad, 64);
        const float inv = 1.f / (xl[r] + xl[128 + r]);
        bool store = q_valid;
        __nv_bfloat16* orow = p.out + (long long)(row_base + q_row0 + r) * D + head * ATT_TILE + wg * 64;
        if (p.peer_out && q_valid) {
            // gathered sequence-parallel layout: this row's token is owned by rank `srank` (see AttnDev); NVLink peer store
            const int qg = q_row0 + r;
            int srank, local;
            if (qg < p.sp_img_region) {
                srank = qg / p.sp_img_pad;
                local = qg - srank * p.sp_img_pad;
            } else {
                const int t = qg - p.sp_img_region, cut = p.sp_txt_rem * (p.sp_txt_base + 1);
                store = t < p.sp_txt_total;               // pad rows of the text region belong to nobody
                int first;
                if (t < cut) { srank = t / (p.sp_txt_base + 1); first = srank * (p.sp_txt_base + 1); }
                else { srank = p.sp_txt_rem + (t - cut) / p.sp_txt_base; first = cut + (srank - p.sp_txt_rem) * p.sp_txt_base; }
                local = p.sp_img_pad + (t - first);
            }
            if (store)
                orow = reinterpret_cast<__nv_bfloat16*>(__ldg(reinterpret_cast<const unsigned long long*>(p.peer_out) + srank)) +
                       ((long long)b * p.sp_rows_pad + local) * p.out_ld + (p.head_off + head) * ATT_TILE + wg * 64;
        }
        if (q_valid) {      // warp-uniform: tcgen05.ld is a warp-collective; only the global stores are per-row predicated
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                uint32_t o[32];
                tmem_ld32(tO + ch * 32, o);
                tmem_ld_wait();
                if (store) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(o[q4 * 8 + i]) * inv;
                        *reinterpret_cast<uint4*>(orow + ch * 32 + q4 * 8) =
                            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    }
                }
            }
        }
}

template <int POLY, bool FIXED>
__global__ void __launch_bounds__(AT5_THREADS, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap tm128, const AttnDev p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                       // [2 d-halves][128 rows x 128 B]
    uint8_t* sK = smem + ATT_TILE_BYTES;                      // [K stages][2 d-halves][128 kv rows x 128 B]
    uint8_t* sV = sK + AT5_KSTAGES * AT5_STAGE_BYTES;         // [V stages][256 kv rows x 128 B (my 64 dims)]
    float* xm = reinterpret_cast<float*>(sV + AT5_VSTAGES * AT5_STAGE_BYTES);   // [2 parities][2 WGs][128 rows] tile max
    float* xl = xm + 512;                                                        // [2 WGs][128 rows] partial row sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(xl + 256);
    uint64_t* q_full = bars;                       // leader
    uint64_t* k_full = bars + 1;                   // [K stages] leader
    uint64_t* k_empty = k_full + AT5_KSTAGES;      // both
    uint64_t* v_full = k_empty + AT5_KSTAGES;      // [V stages] leader
    uint64_t* v_empty = v_full + AT5_VSTAGES;      // both
    uint64_t* s_full = v_empty + AT5_VSTAGES;      // both
    uint64_t* s_free = s_full + 1;                 // leader, 16 warp arrivals: S is in registers everywhere
    uint64_t* p_full = s_free + 1;                 // leader, 16 warp arrivals: P is in TMEM, O rescaled
    uint64_t* pv_done = p_full + 1;                // both
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int cta_rank = (int)cluster_ctarank();
    const int rpb = p.seq.img_pad + p.seq.txt_pad;
    const int n128 = rpb / ATT_TILE;
    const int n_kv = (n128 + 1) / 2;               // 256-row KV tiles (the last one may be half empty)
    const int head = blockIdx.y, b = blockIdx.z;
    const int q_row0 = (blockIdx.x >> 1) * 2 * ATT_TILE + cta_rank * ATT_TILE;
    const bool q_valid = q_row0 < rpb;
    const int D = p.H * ATT_TILE;
    const int colQ = head * ATT_TILE, colK = D + head * ATT_TILE, colV = 2 * D + head * ATT_TILE;
    const int row_base = b * rpb;
    griddep_launch_dependents();

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm128);
        mbar_init(q_full, 1);
        for (int i = 0; i < AT5_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < AT5_VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(s_free, 16);
        mbar_init(p_full, 16);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc_cg2<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t COL_S = 128, COL_P = 384;
    griddep_wait();      // the prologue above overlapped the QKV GEMM's last wave; q|k|v are visible from here on
    // Register redistribution (the kernel is compiled for 384 threads x 168 registers): the control warpgroup (warps 8-11)
    // hands registers to the two softmax warpgroups, whose threads hold a whole 128-value score row.

    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      if (warp == 8) {
        {
            // ================= TMA producer (whole warp in the loop, one elected lane issues) =================
            const bool issuer = elect_one_sync();
            const int qr = q_valid ? q_row0 : 0;
            if (issuer) {
                if (cta_rank == 0) mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_2d_cg2(sQ + hf * ATT_HALF_BYTES, &tm128, colQ + hf * 64, row_base + qr, leader_smem_u32(q_full));
            }
            __syncwarp();
            int ks = 0, vs = 0;
            uint32_t kph = 0, vph = 0;
            auto load_k = [&](int j) {       // my 128 kv rows (half of the 256-row tile) x 128 head dims, two 64-dim halves
                mbar_wait(&k_empty[ks], kph ^ 1);
                if (issuer) {
                    if (cta_rank == 0) mbar_expect_tx(&k_full[ks], 2 * AT5_STAGE_BYTES);
                    const uint32_t bar = leader_smem_u32(&k_full[ks]);
                    for (int hf = 0; hf < 2; ++hf)
                        tma_load_2d_cg2(sK + ks * AT5_STAGE_BYTES + hf * ATT_HALF_BYTES, &tm128, colK + hf * 64,
                                        row_base + j * 256 + cta_rank * ATT_TILE, bar);
                }
                __syncwarp();
                if (++ks == AT5_KSTAGES) { ks = 0; kph ^= 1; }
            };
            auto load_v = [&](int j) {       // 256 kv rows x my 64 head dims
                mbar_wait(&v_empty[vs], vph ^ 1);
                if (issuer) {
                    if (cta_rank == 0) mbar_expect_tx(&v_full[vs], 2 * AT5_STAGE_BYTES);
                    const uint32_t bar = leader_smem_u32(&v_full[vs]);
                    for (int hf = 0; hf < 2; ++hf)
                        tma_load_2d_cg2(sV + vs * AT5_STAGE_BYTES + hf * ATT_HALF_BYTES, &tm128, colV + cta_rank * 64,
                                        row_base + j * 256 + hf * ATT_TILE, bar);
                }
                __syncwarp();
                if (++vs == AT5_VSTAGES) { vs = 0; vph ^= 1; }
            };
            load_k(0);
            for (int j = 0; j < n_kv; ++j) {
                if (j + 1 < n_kv) load_k(j + 1);
                load_v(j);
            }
        }
      } else if (warp == 9) {
        // Two issuer threads on two schedulers (warps 9 and 10, each the highest warp id of its scheduler, which the arbiter
        // serves first): issuing a tcgen05.mma costs the thread ~75-100 cycles here (descriptor arithmetic + the instruction
        // itself, competing for issue slots with two busy softmax warps), so ONE thread needs ~2000 cycles per KV tile for the 8 S
        // and 16 PV instructions — as long as the whole softmax chain (round-2 timing experiments, profiles/r02_attention.md).
        // S and PV touch disjoint TMEM columns and are ordered against the softmax by s_free / p_full / pv_done, not against
        // each other, so the two streams may interleave freely on the (in-order) tensor pipe.
        if (cta_rank == 0) {
            // ================= S issuer (leader CTA) =================
            // The whole warp runs the loop (barrier waits included) and one elected lane issues: loop state and operands stay
            // provably warp-uniform, so ptxas feeds UTCHMMA from uniform registers directly.  With the loop inside an
            // `if (lane == 0)` it wrapped EVERY tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (75-100 cycles per issue).
            constexpr uint32_t IDESC_S = umma_idesc_bf16(256, 256, false);
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool issuer = elect_one_sync();
            int ks = 0;
            uint32_t kph = 0;
            const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ));
            mbar_wait(q_full, 0);
            for (int j = 0; j < n_kv; ++j) {
                if (j > 0) mbar_wait(s_free, (j - 1) & 1);   // S(j-1) is in registers in both CTAs
                mbar_wait(&k_full[ks], kph);
                tc_fence_after();
                const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(sK + ks * AT5_STAGE_BYTES));
                if (issuer) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) {   // 8 x 16 head dims: + 32 B inside a swizzled row, + 16 KB for the second d-half
                        const uint64_t off = (uint64_t)(((s >> 2) * ATT_HALF_BYTES + (s & 3) * 32) >> 4);
                        umma_ss_f16_cg2(tb + COL_S, dq + off, dk + off, IDESC_S, s ? 1u : 0u);
                    }
                    umma_commit_cg2(s_full, 3);
                    umma_commit_cg2(&k_empty[ks], 3);
                }
                __syncwarp();
                if (++ks == AT5_KSTAGES) { ks = 0; kph ^= 1; }
            }
        }
      } else if (warp == 10) {
        if (cta_rank == 0) {
            // ================= PV issuer (leader CTA; whole warp in the loop, one elected lane issues) =================
            constexpr uint32_t IDESC_O = umma_idesc_bf16(256, 128, true);   // B = V is MN-major
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool issuer = elect_one_sync();
            int vs = 0;
            uint32_t vph = 0;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&v_full[vs], vph);
                mbar_wait(p_full, j & 1);                    // P(j) is in TMEM in both CTAs, O rescaled
                tc_fence_after();
                const uint64_t dv = umma_desc_mnmajor_sw128(smem_u32(sV + vs * AT5_STAGE_BYTES), ATT_HALF_BYTES, 1024);
                if (issuer) {
#pragma unroll
                    for (int s = 0; s < 16; ++s)    // 16 x 16 kv rows (2 KB of V each); A = P (8 packed columns per step)
                        umma_ts_f16_cg2(tb, tb + COL_P + s * 8, dv + (uint64_t)(s * 128), IDESC_O, (j == 0 && s == 0) ? 0u : 1u);
                    umma_commit_cg2(pv_done, 3);
                    umma_commit_cg2(&v_empty[vs], 3);
                }
                __syncwarp();
                if (++vs == AT5_VSTAGES) { vs = 0; vph ^= 1; }
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ================= softmax (warps 0-7): warpgroup wg owns score columns [128 wg, 128 wg + 128) of every KV tile =================
        const int wg = warp >> 2;
        const int quad = warp & 3;                   // TMEM lane quadrant a warp may touch = warp id % 4
        const int r = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + COL_S + wg * 128;
        const uint32_t tP = tmem_base + lane_addr + COL_P + wg * 64;
        const uint32_t tO = tmem_base + lane_addr + wg * 64;          // the half of O this warpgroup rescales / stores
        const float c = p.scale_log2;
        const uint64_t c2 = pk2(c, c);
        float m_ref = -INFINITY;
        uint64_t l2 = pk2(0.f, 0.f);
        auto valid_rows = [&](int t128) -> int {     // valid kv rows of my half of a 256-row tile
            return t128 < n128 ? (p.tile_valid ? __ldg(p.tile_valid + t128) : kv_valid_rows(p.seq, t128, b)) : 0;
        };
        int nv_next = valid_rows(wg);
        for (int j = 0; j < n_kv; ++j) {
            const int nv = nv_next;
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t s[128];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t(&dst)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[ch * 32]);
                tmem_ld32(tS + ch * 32, dst);
            }
            nv_next = valid_rows(2 * (j + 1) + wg);          // its constant-bank / global latency hides under the TMEM load
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_smem_u32(s_free));   // the tensor pipe may refill S now
            [[maybe_unused]] bool grow = false;
            [[maybe_unused]] float alpha = 1.f;
            if constexpr (FIXED) {
                // ---- bounded scores (QK-RMSNorm: |q.k| <= 128 max|w_q| max|w_k|), q pre-multiplied by scale*log2(e) in the QKV
                // epilogue: p = 2^s needs no reference, hence no row max, no exchange between the warpgroups, no O rescale and no
                // scale/subtract FFMA2; p in [2^-B, 2^B] with B <= 80 neither underflows nor overflows fp32 / bf16 / the fp32
                // accumulators (8448 * 2^80 * |v|).  The clamp of the polynomial path is not needed either. ----
                if (nv == ATT_TILE) {
#pragma unroll
                    for (int i = 0; i < 128; i += 2) {
                        const float x0 = __uint_as_float(s[i]), x1 = __uint_as_float(s[i + 1]);
                        float e0, e1;
                        if (((i >> 1) & 7) < POLY) {
                            const uint64_t X = pk2(x0, x1);
                            const uint64_t T = add2(X, pk2(12582912.f, 12582912.f));
                            const uint64_t N = add2(T, pk2(-12582912.f, -12582912.f));
                            const uint64_t Fr = fma2(N, pk2(-1.f, -1.f), X);
                            uint64_t P = fma2(Fr, pk2(0.0551716574f, 0.0551716574f), pk2(0.2426111400f, 0.2426111400f));
                            P = fma2(P, Fr, pk2(0.6932609677f, 0.6932609677f));
                            P = fma2(P, Fr, pk2(0.9999280572f, 0.9999280572f));
                            float t0, t1, p0, p1;
                            upk2(T, t0, t1);
                            upk2(P, p0, p1);
                            e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
                            e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
                        } else {
                            e0 = fast_exp2(x0);
                            e1 = fast_exp2(x1);
                        }
                        l2 = add2(l2, pk2(e0, e1));
                        s[i >> 1] = pack_bf16(e0, e1);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 128; i += 2) {
                        const float e0 = i < nv ? fast_exp2(__uint_as_float(s[i])) : 0.f;
                        const float e1 = i + 1 < nv ? fast_exp2(__uint_as_float(s[i + 1])) : 0.f;
                        l2 = add2(l2, pk2(e0, e1));
                        s[i >> 1] = pack_bf16(e0, e1);
                    }
                }
            } else {
            // ---- row max of my 128 columns, then the row max of the whole 256-wide tile through shared memory ----
            float m0 = -INFINITY, m1 = -INFINITY;
            if (nv == ATT_TILE) {
                float m2 = -INFINITY, m3 = -INFINITY;     // four chains of 16 instead of two of 32
#pragma unroll
                for (int i = 0; i < 128; i += 8) {
                    m0 = max3(m0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
                    m1 = max3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
                    m2 = max3(m2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
                    m3 = max3(m3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
                }
                m0 = fmaxf(m0, m2);
                m1 = fmaxf(m1, m3);
            } else {
#pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i < nv) m0 = fmaxf(m0, __uint_as_float(s[i]));
            }
            float* xmj = xm + (j & 1) * 256;
            xmj[wg * 128 + r] = fmaxf(m0, m1);
            named_bar_sync(1 + quad, 64);                 // only my partner warp (same rows, other warpgroup)
            const float mx = fmaxf(fmaxf(m0, m1), xmj[(wg ^ 1) * 128 + r]) * c;
            grow = mx > m_ref + 8.0f;                 // identical decision in both warpgroups (same row, same inputs)
            if (grow) {
                alpha = fast_exp2(m_ref - mx);
                m_ref = mx;
                l2 = fma2(l2, pk2(alpha, alpha), pk2(0.f, 0.f));
            }
            const uint64_t nm2 = pk2(-m_ref, -m_ref);
            // ---- exponentials in place: s[i/2] <- bf16x2(p_i, p_i+1) ----
            if (nv == ATT_TILE) {
#pragma unroll
                for (int i = 0; i < 128; i += 2) {
                    const uint64_t X = fma2(pk2u(s[i], s[i + 1]), c2, nm2);
                    float x0, x1, e0, e1;
                    upk2(X, x0, x1);
                    if (((i >> 1) & 7) < POLY) {
                        // FMA-pipe exp2 (the MUFU is the co-bottleneck): x = n + f, 2^f by a degree-3 polynomial, 2^n via the exponent
                        const uint64_t Xc = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
                        const uint64_t T = add2(Xc, pk2(12582912.f, 12582912.f));
                        const uint64_t N = add2(T, pk2(-12582912.f, -12582912.f));
                        const uint64_t Fr = fma2(N, pk2(-1.f, -1.f), Xc);
                        uint64_t P = fma2(Fr, pk2(0.0551716574f, 0.0551716574f), pk2(0.2426111400f, 0.2426111400f));
                        P = fma2(P, Fr, pk2(0.6932609677f, 0.6932609677f));
                        P = fma2(P, Fr, pk2(0.9999280572f, 0.9999280572f));
                        float t0, t1, p0, p1;
                        upk2(T, t0, t1);
                        upk2(P, p0, p1);
                        e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
                        e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
                    } else {
                        e0 = fast_exp2(x0);
                        e1 = fast_exp2(x1);
                    }
                    l2 = add2(l2, pk2(e0, e1));
                    s[i >> 1] = pack_bf16(e0, e1);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 128; i += 2) {
                    float x0, x1;
                    upk2(fma2(pk2u(s[i], s[i + 1]), c2, nm2), x0, x1);
                    const float e0 = i < nv ? fast_exp2(x0) : 0.f, e1 = i + 1 < nv ? fast_exp2(x1) : 0.f;
                    l2 = add2(l2, pk2(e0, e1));
                    s[i >> 1] = pack_bf16(e0, e1);
                }
            }
            }   // !FIXED
            if (j > 0) {
                // PV(j-1) reads the P buffer and owns O: it must have retired before either is touched
                mbar_wait(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (!FIXED && __any_sync(0xffffffffu, grow)) {
#pragma unroll 1
                    for (int ch = 0; ch < 2; ++ch) {
                        uint32_t o[32];
                        tmem_ld32(tO + ch * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tO + ch * 32, o);
                    }
                }
            }
            {
                uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
                uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
                tmem_st32(tP, lo);
                tmem_st32(tP + 32, hi);
                tmem_st_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_smem_u32(p_full));
        }
        // ---- epilogue: O / (l_wg0 + l_wg1) -> bf16 -> global; warpgroup wg stores head dims [64 wg, 64 wg + 64) ----
        pair_epilogue(p, l2, xl, pv_done, (n_kv - 1) & 1, tO, wg, quad, r, q_valid, q_row0, row_base, D, head, b);
    }

    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc_cg2<512>(tmem_base);
    }
}

template <int POLY, bool FIXED>
static int launch_attn_pair(const CUtensorMap& tm128, const AttnDev& p, dim3 grid, cudaStream_t st) {
    QIE_CONFIGURE_ONCE(cudaFuncSetAttribute(attn_pair_kernel<POLY, FIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT5_SMEM));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(AT5_THREADS);
    cfg.dynamicSmemBytes = AT5_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = launch_attrs(attr, 2);
    QIE_CUDA_OK(cudaLaunchKernelEx(&cfg, attn_pair_kernel<POLY, FIXED>, tm128, p));
    QIE_LAUNCH_OK("attn_pair_kernel");
    return QIE_OK;
}



}  // namespace qie

using namespace qie;

struct AttnScatter {
    void* const* peer_out;
    const qie_peers* pr;
    int out_ld, head_off;
};
static int attn_launch(const void* qkv, void* out, const qie_seq* seq, const int* tile_valid, int num_heads, int variant,
                       void* stream, const AttnScatter* sc = nullptr);

extern "C" int qie_attn_fwd_tiles(const void* qkv, void* out, int n_tiles, const int* tile_valid_dev, int num_heads,
                                  int variant, void* stream) {
    QIE_REQUIRE(qkv && out && tile_valid_dev && n_tiles > 0 && num_heads > 0, QIE_EINVAL, "qie_attn_fwd_tiles: bad argument");
    qie_seq s{};
    s.batch = 1;
    s.img_rows = s.img_pad = n_tiles * 128;
    return attn_launch(qkv, out, &s, tile_valid_dev, num_heads, variant, stream);
}

// variant 0 = the tuned default (CTA-pair kernel, 2 of every 8 score pairs on the FMA-pipe polynomial exp2).  Explicit choices:
// bit 3 (0x8) = single-CTA fallback kernel; bits 4..7 = how many of every 8 score pairs take the polynomial (0 with bit 8
// (0x100) set = all MUFU, 2, 3, 4); bit 9 (0x200) = bounded-score form: the caller promises that q is already multiplied by
// softmax_scale * log2(e) and that |q.k| <= 80 for every pair (true after QK-RMSNorm with bounded norm weights: qie_forward
// derives the bound of every block from its norm weights and picks this form only where it holds), so p = 2^s needs no
// running max, no exchange between the two softmax warpgroups and no O rescale.
extern "C" int qie_attn_fwd(const void* qkv, void* out, const qie_seq* seq, int num_heads, int variant, void* stream) {
    QIE_REQUIRE(qkv && out && seq, QIE_EINVAL, "qie_attn_fwd: null pointer");
    QIE_REQUIRE(seq->img_pad % 128 == 0 && seq->txt_pad % 128 == 0 && seq->batch > 0 && num_heads > 0 &&
                    seq->img_rows > 0 && seq->txt_rows > 0 && seq->img_rows > seq->img_pad - 128 &&
                    seq->txt_rows > seq->txt_pad - 128,
                QIE_ESHAPE, "qie_attn_fwd: bad sequence layout (every 128-row KV tile needs >= 1 valid row)");
    return attn_launch(qkv, out, seq, nullptr, num_heads, variant, stream);
}

// attention of one rank of a sequence-parallel group over its gathered q|k|v, output scattered to the token owners
// (called by the ATTN phase of qie_forward_phase when peers are installed)
namespace qie {
int attn_fwd_peers(const void* qkv_gathered, void* const* peer_out_dev, const qie_peers* pr, const int* tile_valid_dev,
                   int heads_local, int out_ld, int variant, void* stream) {
    QIE_REQUIRE(qkv_gathered && peer_out_dev && pr && tile_valid_dev && heads_local > 0, QIE_EINVAL, "attn_fwd_peers: bad argument");
    qie_seq s{};
    s.batch = pr->batch;
    s.img_rows = s.img_pad = pr->size * pr->img_pad + (pr->txt_total + 127) / 128 * 128;
    AttnScatter sc{peer_out_dev, pr, out_ld, pr->rank * heads_local};
    return attn_launch(qkv_gathered, const_cast<void*>(qkv_gathered) /* unused */, &s, tile_valid_dev, heads_local, variant, stream, &sc);
}
}  // namespace qie

static int attn_launch(const void* qkv, void* out, const qie_seq* seq, const int* tile_valid, int num_heads, int variant,
                       void* stream, const AttnScatter* sc) {
    if (variant == 0) variant = 0x20;
    if (variant == 0x200) variant = 0x230;    // bounded-score default: 3 of every 8 score pairs on the polynomial
    const int poly = (variant >> 4) & 15, single = (variant >> 3) & 1, fixed = (variant >> 9) & 1;
    QIE_REQUIRE((variant & ~0x3F8) == 0 && (poly == 0 || poly == 2 || poly == 3 || poly == 4), QIE_EINVAL,
                "qie_attn_fwd: bad variant 0x%x", variant);
    QIE_REQUIRE(!(fixed && single), QIE_EINVAL, "qie_attn_fwd: the bounded-score form (0x200) is a CTA-pair kernel");
    const int rpb = seq->img_pad + seq->txt_pad;
    const int D = num_heads * 128;
    CUtensorMap tm;
    int rc = make_tmap_2d(&tm, qkv, (uint64_t)seq->batch * rpb, (uint64_t)3 * D, (uint64_t)3 * D * 2, 128, 64, 2);
    if (rc) return rc;
    AttnDev p{};
    p.seq = *seq;
    p.H = num_heads;
    p.tile_valid = tile_valid;
    p.out = (__nv_bfloat16*)out;
    if (sc) {
        QIE_REQUIRE(!single, QIE_EINVAL, "the attention output scatter needs a CTA-pair kernel");
        const qie_peers* pr = sc->pr;
        p.peer_out = sc->peer_out;
        p.sp_img_pad = pr->img_pad;
        p.sp_img_region = pr->size * pr->img_pad;
        p.sp_rows_pad = pr->img_pad + pr->txt_pad;
        p.sp_txt_total = pr->txt_total;
        p.sp_txt_base = pr->txt_total / pr->size;
        p.sp_txt_rem = pr->txt_total % pr->size;
        p.out_ld = sc->out_ld;
        p.head_off = sc->head_off;
    }
    p.scale_log2 = ATTN_SCALE_LOG2;
    // V tile in smem: two halves (64 dims each, 16 KB apart) of 128 kv rows x 128 B, 128B-swizzled by TMA.
    // MN-major canonical layout: 8 kv rows x 128 B = one 1024 B atom (SBO), next 64 dims LBO away.
    p.v_lbo = ATT_HALF_BYTES;
    p.v_sbo = 1024;
    p.v_kstep = 2048;
    dim3 grid((rpb + 255) / 256, num_heads, seq->batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (single) {
        switch (poly) {
            case 0: return launch_attn<true, 0>(tm, p, grid, st);
            case 2: return launch_attn<true, 2>(tm, p, grid, st);
            case 3: return launch_attn<true, 3>(tm, p, grid, st);
            case 4: return launch_attn<true, 4>(tm, p, grid, st);
        }
    }
    grid.x *= 2;     // a cluster of two CTAs per 256 query rows
    if (fixed) {
        switch (poly) {
            case 0: return launch_attn_pair<0, true>(tm, p, grid, st);
            case 2: return launch_attn_pair<2, true>(tm, p, grid, st);
            case 3: return launch_attn_pair<3, true>(tm, p, grid, st);
            case 4: return launch_attn_pair<4, true>(tm, p, grid, st);
        }
    }
    switch (poly) {
        case 0: return launch_attn_pair<0, false>(tm, p, grid, st);
        case 2: return launch_attn_pair<2, false>(tm, p, grid, st);
        case 3: return launch_attn_pair<3, false>(tm, p, grid, st);
        case 4: return launch_attn_pair<4, false>(tm, p, grid, st);
    }
    return QIE_EINVAL;
}
