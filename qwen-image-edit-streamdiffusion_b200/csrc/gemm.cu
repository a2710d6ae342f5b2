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
// Structure (192 threads, 1 CTA / SM, persistent, static round-robin tile schedule; 320 threads = eight epilogue warps for the
// short-K 8-bit shapes, see gemm_threads):
//   warps 0..EW-1   : epilogue      — tcgen05.ld 32x32b -> registers -> per-warp smem transpose -> coalesced global
//   warp EW         : TMA producer  — A tile [128 x 128 B] and W tile (128B-swizzled rows) per stage
//   warp EW+1       : MMA issuer    — tcgen05.mma kind::f16 / kind::f8f6f4 / kind::i8, accumulators in TMEM.  Highest warp id of
//                                     its scheduler on purpose: the arbiter serves the highest id first, so the issuer is
//                                     never queued behind the epilogue warp it shares the scheduler with.
//                     Both walk the schedule and wait on the barriers as whole warps; ONE lane (elect.sync) issues, so that the
//                     operands stay provably warp-uniform (no R2UR waterfall around UTCHMMA / UTMALDG, see the issuer loop).
// Three pipelines: smem full/empty ring (TMA <-> MMA), 2 TMEM accumulator stages (MMA <-> epilogue, so the
// epilogue of tile i overlaps the main loop of tile i+1), and the tile loop.
//
// CG = 1: one CTA computes a 128 x BN tile (cta_group::1).
// CG = 2: a cluster of two CTAs on one TPC computes a 256 x BN tile with ONE tcgen05.mma.cta_group::2 issued by the
//         leader CTA: each CTA loads its own 128 A rows (any two 128-row blocks of the same stream) and HALF of the W
//         tile (BN/2 rows), so L2->SM operand traffic per FLOP drops by a third and W smem reads are halved.  The peer's
//         TMA credits the leader's full barrier; the leader's tcgen05.commit multicasts to both CTAs' barriers.
// Roofline: tensor pipe; algorithmic FLOPs = 2*M*N*K.
#include "common.cuh"

namespace qie {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK_BYTES = 128;   // one 128B swizzle row: 64 bf16 or 128 e4m3
// Epilogue warps (template parameter EW): four (one per TMEM lane quadrant) for bf16 operands, where a 256-wide tile's main loop
// (~24.6 k cycles at K = 3072) covers the epilogue; EIGHT for the 8-bit operand types at BN = 256 and K < 8192, whose main loop is
// half as long and was waiting for the latency-bound epilogue warps (QKV + RMSNorm + RoPE at 0.54 of cuBLASLt e4m3, out-proj at 0.69:
// tools/q8_gemm_bench.py).  Two warps then share a lane quadrant and split the tile's columns (one 128-column head each in the
// QKV epilogue).  The long-K shape (FF-down) keeps four: its main loop covers the epilogue and it prefers the sixth ring stage.
__host__ __device__ constexpr int gemm_threads(int EW) { return (EW + 2) * 32; }

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
    float q_scale;            // QKV epilogue: factor on the q columns (softmax scale * log2 e for the bounded-score attention)
    const float* a_scale;
    const float* w_scale[2];
    int model_dim;           // D (QKV epilogue: column block -> q/k/v)
    int l2_hints;            // bit 0: weights evict-last, bit 1: activations evict-first (CTA-pair kernels)
    // split-K tail: the last `tail_tiles` tiles (the partial wave of the persistent schedule) are cut into `tail_split` K
    // ranges, one work item each, so the tail wave costs 1/tail_split of a tile; partial accumulators meet in `scratch`
    int tail_tiles, tail_split, tail_dbg;
    int tail_nsplit;         // 1: the tail tiles are cut along N instead (two BN/2-column halves, no reduction; tail_split == 2)
    int group_m;             // m-units per raster band (qie_tune key 5; 0 = GEMM_GROUP_M)
    float* scratch;          // [tail_tiles][tail_split][CG*128 rows][BN] fp32
    int* tickets;            // [tail_tiles][8 row slices]
    void* const* peer_out;   // QKV epilogue, sequence parallel: device table of the ranks' gathered q|k|v buffers (or NULL)
    int sp_rank, sp_hl;            // my rank, heads per rank
    int sp_gathered_rows;          // rows of one batch element in the gathered layout [P image shards | all text tokens]
    int sp_txt_row0;               // gathered row of my first text token
    float* q8_amax;                // optional: per-row max|out| folded in by the bf16 epilogues (feeds the 8-bit quantiser)
};

template <int BN, int CG, int EW = 4>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK_BYTES;
    static constexpr int B_BYTES = (BN / CG) * GEMM_BK_BYTES;               // each CTA of a pair holds half of the W tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_ROW_BYTES = 144;                               // 32 fp32 + 16 B pad: conflict-free both ways
    static constexpr int EPI_WARP_BYTES = 32 * EPI_ROW_BYTES;               // per-warp transpose staging
    static constexpr int COLV_BYTES = 2 * BN * 4;                           // per-warp copy of the tile's bias | weight scales (QKV epilogue)
    // four epilogue warps: 196 KB of operand ring; eight: what 227 KB leave beside the doubled staging (5 stages at BN = 256, pairs)
    static constexpr int RING_BYTES = EW == 4 ? 196 * 1024 : 227 * 1024 - 1536 - EW * (EPI_WARP_BYTES + COLV_BYTES);
    static constexpr int STAGES = RING_BYTES / STAGE_BYTES > 8 ? 8 : RING_BYTES / STAGE_BYTES;
    static constexpr int BAR_BYTES = 512;                                   // ring + TMEM barriers, tmem slot, 16 LN landing barriers
    static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + EW * EPI_WARP_BYTES + EW * COLV_BYTES + 1024;   // +1024: manual alignment
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;             // two accumulator stages
};

// Work decomposition.  An "m-unit" is what one CTA (CG=1) or one CTA pair (CG=2) covers along M:
// CG consecutive 128-row blocks of ONE stream of ONE batch element (the last unit of a stream may be half empty).
struct MUnit {
    int b, s, ti;      // batch, stream, first 128-row block inside the stream
};
template <int CG>
__device__ __forceinline__ int units_in_stream(int pad_rows) {
    return (pad_rows / GEMM_BM + CG - 1) / CG;
}
template <int CG>
__device__ __forceinline__ MUnit decode_munit(const GemmDev& p, int mu) {
    const int u0 = (p.streams & 1) ? units_in_stream<CG>(p.seq.img_pad) : 0;
    const int u1 = (p.streams & 2) ? units_in_stream<CG>(p.seq.txt_pad) : 0;
    MUnit r;
    r.b = mu / (u0 + u1);
    const int rem = mu % (u0 + u1);
    r.s = rem >= u0 ? 1 : 0;
    r.ti = (r.s ? rem - u0 : rem) * CG;
    return r;
}

// Tile rasterisation: m-units are walked in bands of GEMM_GROUP_M; inside a band m varies fastest, so the ~74 tiles that
// are resident at once cover ~8 m-units x ~9 n-blocks (A and W panels of a few MB each, L2-resident) instead of one m-unit
// x all n-blocks (which streams the whole W matrix through L2 for every 256 rows: measured 5x the algorithmic DRAM reads).
constexpr int GEMM_GROUP_M = 8;
__device__ __forceinline__ void tile_to_mn(int tile, int m_units, int n_blocks, int& mu, int& nb, int group_m = GEMM_GROUP_M) {
    const int band = tile / (group_m * n_blocks);
    const int first = band * group_m;
    const int gm = min(group_m, m_units - first);
    const int local = tile - band * group_m * n_blocks;
    mu = first + local % gm;
    nb = local / gm;
}

// Work items of the persistent schedule: the first (num_tiles - tail_tiles) items are whole tiles; every remaining tile
// is `tail_split` items, each covering one K range.
struct WorkItem {
    int tile, part, kb0, kb1;
    bool split;      // K range of a tail tile: partial sums meet in scratch
    bool nhalf;      // column half `part` of a tail tile: an independent BN/2-wide tile
};
__device__ __forceinline__ WorkItem decode_item(int item, int num_tiles, int tail_tiles, int tail_split, int k_blocks,
                                                int tail_nsplit) {
    WorkItem w;
    const int whole = num_tiles - tail_tiles;
    w.nhalf = false;
    if (item < whole) {
        w.tile = item; w.part = 0; w.kb0 = 0; w.kb1 = k_blocks; w.split = false;
    } else if (tail_nsplit) {
        const int i = item - whole;
        w.tile = whole + (i >> 1); w.part = i & 1; w.kb0 = 0; w.kb1 = k_blocks; w.split = false; w.nhalf = true;
    } else {
        const int i = item - whole;
        w.tile = whole + i / tail_split;
        w.part = i % tail_split;
        const int per = (k_blocks + tail_split - 1) / tail_split;
        w.kb0 = w.part * per;
        w.kb1 = min(k_blocks, w.kb0 + per);
        w.split = true;
    }
    return w;
}

template <int BN, int EPI, int QT, int CG, int EW = 4>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
            const __grid_constant__ CUtensorMap tmB1, const GemmDev p) {
    // epilogue warps 0 .. EW-1, TMA producer = warp EW, MMA issuer = warp EW + 1
    using S = GemmSmem<BN, CG, EW>;
    constexpr int STAGES = S::STAGES;
    constexpr bool FP8 = QT != 0;                // 8-bit operands (e4m3 or int8): 128-element k-blocks, dequant epilogue
    constexpr int BK = FP8 ? 128 : 64;           // elements per k-block
    constexpr uint32_t IDESC = QT == 2 ? umma_idesc_s8(GEMM_BM * CG, BN)
                               : QT == 1 ? umma_idesc_e4m3(GEMM_BM * CG, BN) : umma_idesc_bf16(GEMM_BM * CG, BN);
    constexpr int BNH = BN / 2;                  // column half of a tail tile (N-split tail)
    constexpr uint32_t IDESC_H = QT == 2 ? umma_idesc_s8(GEMM_BM * CG, BNH)
                                 : QT == 1 ? umma_idesc_e4m3(GEMM_BM * CG, BNH) : umma_idesc_bf16(GEMM_BM * CG, BNH);
    auto accf = [](uint32_t r) -> float { return QT == 2 ? (float)(int)r : __uint_as_float(r); };   // int32 accumulators for kind::i8

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* full_bar = bars;                   // [STAGES]   (CG=2: only the leader's copy is used)
    uint64_t* empty_bar = bars + STAGES;         // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;     // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;// [2]        (CG=2: only the leader's copy is used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    griddep_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
    const int u0 = (p.streams & 1) ? units_in_stream<CG>(p.seq.img_pad) : 0;
    const int u1 = (p.streams & 2) ? units_in_stream<CG>(p.seq.txt_pad) : 0;
    const int m_units = p.seq.batch * (u0 + u1);
    const int num_tiles = m_units * p.n_blocks;                     // tiles of one CTA (CG=1) / one pair (CG=2)
    const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;
    const int k_blocks = (p.K + BK - 1) / BK;
    const int num_items = num_tiles - p.tail_tiles + p.tail_tiles * p.tail_split;
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
            mbar_init(&tempty_bar[i], EW * CG);
        }
        fence_barrier_init();
    }
    if (warp == EW + 1) {
        if constexpr (CG == 2) tmem_alloc_cg2<S::TMEM_COLS>(tmem_slot);
        else tmem_alloc<S::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();      // everything above overlapped the previous kernel's tail; its output is visible from here on

    if (warp == EW) {
        {
            // ================= TMA producer (every CTA loads its own A rows and its share of W) =================
            // whole warp in the loop, one elected lane issues (uniform operands: no R2UR waterfall around UTMALDG, see the MMA issuer)
            const bool issuer = elect_one_sync();
            int stage = 0;
            uint32_t phase = 0;
            // L2 residency: the weight matrix is re-read by every m-unit (33 times at config 2) and fits the 126 MB L2, the
            // activation panel of a band is needed by the ~n_blocks tiles that run together: weights evict last, activations
            // first, so the weights are fetched from HBM once per launch instead of once per band (p.l2_hints, qie_tune key 2)
            const uint64_t pol_w = l2_policy_evict_last(), pol_a = l2_policy_evict_first();
            const int hints = p.l2_hints;
            for (int item = tile0; item < num_items; item += tile_step) {
                const WorkItem wi = decode_item(item, num_tiles, p.tail_tiles, p.tail_split, k_blocks, p.tail_nsplit);
                int mu, nb;
                tile_to_mn(wi.tile, m_units, p.n_blocks, mu, nb, p.group_m);
                const MUnit m = decode_munit<CG>(p, mu);
                const int seg_pad = m.s ? p.seq.txt_pad : p.seq.img_pad;
                int ti = m.ti + cta_rank;
                if (ti * GEMM_BM >= seg_pad) ti = m.ti;          // odd block count: the peer re-reads the leader's rows
                const int a_row = p.a_compact ? m.b * seg_pad + ti * GEMM_BM
                                              : m.b * rpb + (m.s ? p.seq.img_pad : 0) + ti * GEMM_BM;
                const CUtensorMap* tmB = m.s ? &tmB1 : &tmB0;
                // column half of a tail tile: the pair's BN/2-wide B operand is BN/2/CG rows per CTA; the box stays BN/CG rows (the
                // rows behind the ones the MMA reads are loaded and ignored: tail items only)
                const int b_row = wi.nhalf ? nb * BN + wi.part * BNH + cta_rank * (BNH / CG) : nb * BN + cta_rank * (BN / CG);
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::STAGE_BYTES;
                    uint8_t* sb = sa + S::A_BYTES;
                    if (issuer) {
                    if constexpr (CG == 2) {
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
                        const uint32_t bar = leader_smem_u32(&full_bar[stage]);
                        if (hints & 2) tma_load_2d_cg2_hint(sa, &tmA, kb * BK, a_row, bar, pol_a);
                        else tma_load_2d_cg2(sa, &tmA, kb * BK, a_row, bar);
                        if (hints & 1) tma_load_2d_cg2_hint(sb, tmB, kb * BK, b_row, bar, pol_w);
                        else tma_load_2d_cg2(sb, tmB, kb * BK, b_row, bar);
                    } else {
                        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                        tma_load_2d(sa, &tmA, kb * BK, a_row, &full_bar[stage]);
                        tma_load_2d(sb, tmB, kb * BK, b_row, &full_bar[stage]);
                    }
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == EW + 1) {
        if (cta_rank == 0) {
            // ================= MMA issuer (leader CTA only) =================
            // The whole warp walks the schedule and waits on the barriers; one elected lane issues.  Loop state and operands are
            // then provably warp-uniform and UTCHMMA takes them from uniform registers; inside an `if (lane == 0)` loop ptxas
            // wrapped every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (5 broadcasts, ~75-100 cycles per issue).
            const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
            const bool issuer = elect_one_sync();
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = tile0; item < num_items; item += tile_step) {
                const WorkItem wi = decode_item(item, num_tiles, p.tail_tiles, p.tail_split, k_blocks, p.tail_nsplit);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tb + acc * BN;
                const uint32_t idesc = wi.nhalf ? IDESC_H : IDESC;
                for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint64_t da = umma_desc_kmajor_sw128(sa);
                    const uint64_t db = umma_desc_kmajor_sw128(sa + S::A_BYTES);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {   // 4 x 32 B along the swizzled 128 B row
                            const uint32_t accum = ((kb - wi.kb0) | k) ? 1u : 0u;
                            umma_ss<QT, CG>(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                        }
                        // frees the smem slot (in both CTAs of a pair) when these MMAs retire
                        if constexpr (CG == 2) umma_commit_cg2(&empty_bar[stage], 3);
                        else umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                // accumulator complete -> epilogue warps (of both CTAs)
                if (issuer) {
                    if constexpr (CG == 2) umma_commit_cg2(&tfull_bar[acc], 3);
                    else umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ================= epilogue (warps 0 .. EW-1) =================
        // Each warp owns the 32 accumulator rows of its TMEM lane quadrant.  Per 32-column chunk:
        //   phase 1 (lane == row)   : tcgen05.ld -> registers -> padded per-warp smem tile
        //   phase 2 (lanes == cols) : 8 lanes x float4 cover one 128 B row segment, 4 rows per instruction, so every
        //                             global load/store of bias, gate, residual, rope and output is a coalesced line.
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        [[maybe_unused]] const int half = warp >> 2;  // EW = 8: which half of the tile's columns this warp takes
        uint8_t* stg = smem + STAGES * S::STAGE_BYTES + S::BAR_BYTES + warp * S::EPI_WARP_BYTES;
        const int sub = lane >> 3, c4 = (lane & 7) * 4;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = tile0; item < num_items; item += tile_step) {
            const WorkItem wi = decode_item(item, num_tiles, p.tail_tiles, p.tail_split, k_blocks, p.tail_nsplit);
            int mu, nb;
            tile_to_mn(wi.tile, m_units, p.n_blocks, mu, nb, p.group_m);
            const MUnit m = decode_munit<CG>(p, mu);
            const int seg_pad = m.s ? p.seq.txt_pad : p.seq.img_pad;
            const int seg_rows = m.s ? p.seq.txt_rows_b[m.b] : p.seq.img_rows;
            const int ti = m.ti + cta_rank;
            const bool dummy = ti * GEMM_BM >= seg_pad;                       // peer half of an odd last unit: nothing to store
            const int local0 = ti * GEMM_BM + quad * 32;                      // first row of this warp inside the stream
            const int jrow0_in_batch = (m.s ? p.seq.img_pad : 0) + local0;   // ... inside the joint layout
            const long long jrow0 = (long long)m.b * rpb + jrow0_in_batch;
            const long long orow0 = p.out_compact ? (long long)m.b * seg_pad + local0 : jrow0;
            const long long arow0 = p.a_compact ? (long long)m.b * seg_pad + local0 : jrow0;
            const float* bias = p.bias[m.s];
            const int n_base = nb * BN + (wi.nhalf ? wi.part * BNH : 0);      // first output column of this item
            const int n_chunks = wi.nhalf ? BNH / 32 : BN / 32;               // 32-column chunks of this item
            // my chunks of the item.  EW = 8: the two warps of a lane quadrant split them — one 128-column head each in the QKV
            // epilogue (row statistics run over a whole head; the single head of a column-half tail item goes to the first warp)
            int c_begin = 0, c_end = n_chunks;
            if constexpr (EW == 8) {
                const int per = (EPI == QIE_EPI_QKV_NORM_ROPE && n_chunks == 4) ? 4 : n_chunks / 2;
                c_begin = half * per < n_chunks ? half * per : n_chunks;
                c_end = c_begin + per < n_chunks ? c_begin + per : n_chunks;
            }

            // QKV epilogue, q / k tiles: every thread of the row layout needs the bias (and weight scale) of every column of the
            // tile, twice (RMS pre-pass, staging).  As global loads their L2 latency was exposed once per 32-column chunk and pass
            // (the top stall sites of the ncu source view); a per-warp copy in shared memory is fetched before the accumulator is
            // waited for, so its latency hides behind the main loop.
            [[maybe_unused]] float* cbias = reinterpret_cast<float*>(smem + STAGES * S::STAGE_BYTES + S::BAR_BYTES + EW * S::EPI_WARP_BYTES +
                                                                      warp * S::COLV_BYTES);
            if constexpr (EPI == QIE_EPI_QKV_NORM_ROPE) {
                if (n_base / p.model_dim != 2 && !dummy) {
                    for (int i = lane * 4; i < n_chunks * 32; i += 128) {
                        *reinterpret_cast<float4*>(cbias + i) = *reinterpret_cast<const float4*>(bias + n_base + i);
                        if constexpr (FP8)
                            *reinterpret_cast<float4*>(cbias + BN + i) = *reinterpret_cast<const float4*>(p.w_scale[m.s] + n_base + i);
                    }
                    __syncwarp();
                }
            }

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;

            [[maybe_unused]] float as_row = 1.f;
            if constexpr (FP8) as_row = dummy ? 0.f : p.a_scale[arow0 + lane];
            // phase 1: chunk c of the accumulator -> staging, optionally (bias + per-row scale) applied in row layout
            auto stage_chunk = [&](int c, float row_scale, bool add_bias_first) {
                uint32_t r[32];
                tmem_ld32(t_addr + c * 32, r);
                tmem_ld_wait();
                float* srow = reinterpret_cast<float*>(stg + lane * S::EPI_ROW_BYTES);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 f = make_float4(accf(r[i]), accf(r[i + 1]), accf(r[i + 2]), accf(r[i + 3]));
                    if (add_bias_first) {
                        const float4 bv = *reinterpret_cast<const float4*>(cbias + c * 32 + i);
                        if constexpr (FP8) {
                            const float4 ws = *reinterpret_cast<const float4*>(cbias + BN + c * 32 + i);
                            f.x *= as_row * ws.x; f.y *= as_row * ws.y; f.z *= as_row * ws.z; f.w *= as_row * ws.w;
                        }
                        f.x = (f.x + bv.x) * row_scale; f.y = (f.y + bv.y) * row_scale;
                        f.z = (f.z + bv.z) * row_scale; f.w = (f.w + bv.w) * row_scale;
                    }
                    *reinterpret_cast<float4*>(srow + i) = f;
                }
                __syncwarp();
            };

            // ---- split-K tail item: park my raw partial accumulators in scratch; the warp whose slice arrives last sums the
            // partials of all K ranges in part order (deterministic) and runs the normal epilogue on the sums
            bool from_scratch = false, released = false;
            [[maybe_unused]] const float* part0 = nullptr;
            if constexpr (EPI != QIE_EPI_QKV_NORM_ROPE) {
                if (wi.split) {
                    const int tail_idx = wi.tile - (num_tiles - p.tail_tiles);
                    const size_t part_elems = (size_t)CG * GEMM_BM * BN;
                    float* mine = p.scratch + ((size_t)tail_idx * p.tail_split + wi.part) * part_elems +
                                  (size_t)(cta_rank * GEMM_BM + quad * 32) * BN;
                    if (!dummy && !(p.tail_dbg & 1)) {
#pragma unroll 1
                        for (int c = c_begin; c < c_end; ++c) {       // K ranges are whole BN-wide tiles: n_chunks == BN / 32
                            stage_chunk(c, 1.f, false);
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int rr = it * 4 + sub;
                                *reinterpret_cast<float4*>(mine + (size_t)rr * BN + c * 32 + c4) =
                                    *reinterpret_cast<const float4*>(stg + rr * S::EPI_ROW_BYTES + c4 * 4);
                            }
                            __syncwarp();
                        }
                    }
                    // the accumulator stage is free again as soon as it has been copied out
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(leader_smem_u32(&tempty_bar[acc]));
                        else mbar_arrive(&tempty_bar[acc]);
                    }
                    released = true;
                    bool last = false;
                    if (!dummy && !(p.tail_dbg & 2)) {
                        __threadfence();
                        __syncwarp();
                        int t = 0;
                        int* ticket = p.tickets + tail_idx * 8 + cta_rank * 4 + quad;
                        if (lane == 0) t = atomicAdd(ticket, 1);
                        t = __shfl_sync(0xffffffffu, t, 0);
                        last = t == p.tail_split * (EW / 4) - 1;    // every warp of this lane quadrant, of every K range, has parked
                        if (last) {
                            c_begin = 0;                            // the last one sums and finishes ALL columns of the quadrant's rows
                            c_end = n_chunks;
                            if (lane == 0) *ticket = 0;          // re-armed for the next launch
                            __threadfence();
                            from_scratch = true;
                            part0 = p.scratch + (size_t)tail_idx * p.tail_split * part_elems +
                                    (size_t)(cta_rank * GEMM_BM + quad * 32) * BN;
                        }
                    }
                    if (!last) {
                        if (++acc == 2) {
                            acc = 0;
                            acc_phase ^= 1;
                        }
                        continue;
                    }
                }
            }
            auto stage_from_scratch = [&](int c) {          // sum of the K-range partials, in part order -> staging tile
                const size_t part_elems = (size_t)CG * GEMM_BM * BN;
                // all loads of a range are issued before the first add, and range 1 before range 0 is consumed: one L2 round trip
                // per chunk for the common two-range split instead of one per (row group, range)
                float4 a[8], b4[8];
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    a[it] = __ldcg(reinterpret_cast<const float4*>(part0 + (size_t)(it * 4 + sub) * BN + c * 32 + c4));
                for (int pp = 1; pp < p.tail_split; ++pp) {
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        b4[it] = __ldcg(reinterpret_cast<const float4*>(part0 + pp * part_elems + (size_t)(it * 4 + sub) * BN + c * 32 + c4));
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        a[it].x += b4[it].x; a[it].y += b4[it].y; a[it].z += b4[it].z; a[it].w += b4[it].w;
                    }
                }
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    *reinterpret_cast<float4*>(stg + (it * 4 + sub) * S::EPI_ROW_BYTES + c4 * 4) = a[it];
                __syncwarp();
            };

            [[maybe_unused]] float rinv_head = 1.f;
            [[maybe_unused]] float rmax[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // q8_amax: max|out| of my 8 rows over this tile
            if (!dummy) {
                // `which` (0 q, 1 k, 2 v) is a property of the tile: D % BN == 0
                const int which = EPI == QIE_EPI_QKV_NORM_ROPE ? n_base / p.model_dim : 2;
                const bool normed = EPI == QIE_EPI_QKV_NORM_ROPE && which != 2;
                // Global operands of a chunk (bias, gate, norm weight, rope rows / residual rows): the loads of chunk c + 1 are
                // issued before the math and the stores of chunk c, so their L2 latency (~1 us under load, it was > 50 % of the
                // epilogue warps' stall samples in the ncu source view) is covered by work instead of being exposed once per chunk.
                struct ChunkOps { float4 bv, wsc, g4, nw4, aux[8]; };
                auto fetch_ops = [&](int c, ChunkOps& o) {
                    const int n0 = n_base + c * 32;
                    o.bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias && !normed) o.bv = *reinterpret_cast<const float4*>(bias + n0 + c4);
                    o.wsc = make_float4(1.f, 1.f, 1.f, 1.f);
                    if constexpr (FP8) o.wsc = *reinterpret_cast<const float4*>(p.w_scale[m.s] + n0 + c4);
                    o.g4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if constexpr (EPI == QIE_EPI_GATE_RESID_F32)
                        o.g4 = *reinterpret_cast<const float4*>(p.gate + m.b * p.gate_bstride + m.s * p.gate_sstride + n0 + c4);
                    o.nw4 = make_float4(1.f, 1.f, 1.f, 1.f);
                    if constexpr (EPI == QIE_EPI_QKV_NORM_ROPE) {
                        if (normed) {
                            o.nw4 = *reinterpret_cast<const float4*>(p.qk_norm_w[m.s][which] + (c & 3) * 32 + c4);
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int rr = it * 4 + sub;
                                o.aux[it] = local0 + rr < seg_rows
                                                ? __ldg(reinterpret_cast<const float4*>(
                                                      p.rope + (long long)(jrow0_in_batch + rr) * 128 + (c & 3) * 32 + c4))
                                                : make_float4(1.f, 0.f, 1.f, 0.f);
                            }
                        }
                    }
                    if constexpr (EPI == QIE_EPI_GATE_RESID_F32) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int rr = it * 4 + sub;
                            o.aux[it] = local0 + rr < seg_rows
                                            ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) +
                                                                               (orow0 + rr) * p.ldo + n0 + c4)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                };
                [[maybe_unused]] float asc[8];     // per-row activation scale (8-bit operands): the same for every chunk of the tile
                if constexpr (FP8) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) asc[it] = normed ? 1.f : p.a_scale[arow0 + it * 4 + sub];
                }
                // EW = 8: two epilogue warps per scheduler cover each other's load latency, and the register file is split ten
                // ways (168 registers per thread: the second operand set spilled, and spill reloads miss the few KB of L1 that the
                // operand ring leaves): the operands of chunk c are fetched at the top of chunk c, under its TMEM load and staging
                ChunkOps nxt;
                if constexpr (EW == 4)
                    if (c_begin < c_end) fetch_ops(c_begin, nxt);
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    const int n0 = n_base + c * 32;
                    if constexpr (EPI == QIE_EPI_QKV_NORM_ROPE) {
                        if (which != 2 && (c & 3) == 0) {                     // first chunk of a head: row RMS over 128 cols
                            // the four TMEM loads of the head are pipelined two deep (chunk cc + 1 in flight while cc is squared)
                            // (EW = 8: one buffer — the other warp of the scheduler covers the load, 32 registers fewer)
                            float ss = 0.f;
                            auto square_chunk = [&](const uint32_t(&r)[32], int cc) {      // ss += |acc * scales + bias|^2 of chunk c + cc
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
                                    const float4 bv = *reinterpret_cast<const float4*>(cbias + (c + cc) * 32 + i);
                                    float4 ws = make_float4(1.f, 1.f, 1.f, 1.f);
                                    if constexpr (FP8) {
                                        ws = *reinterpret_cast<const float4*>(cbias + BN + (c + cc) * 32 + i);
                                        ws.x *= as_row; ws.y *= as_row; ws.z *= as_row; ws.w *= as_row;
                                    }
                                    const float a0 = accf(r[i]) * ws.x + bv.x;
                                    const float a1 = accf(r[i + 1]) * ws.y + bv.y;
                                    const float a2 = accf(r[i + 2]) * ws.z + bv.z;
                                    const float a3 = accf(r[i + 3]) * ws.w + bv.w;
                                    ss += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
                                }
                            };
                            if constexpr (EW == 8) {
                                uint32_t ra[32];
#pragma unroll
                                for (int cc = 0; cc < 4; ++cc) {
                                    tmem_ld32(t_addr + (c + cc) * 32, ra);
                                    tmem_ld_wait();
                                    square_chunk(ra, cc);
                                }
                            } else {
                                uint32_t ra[32], rb[32];
                                tmem_ld32(t_addr + c * 32, ra);
#pragma unroll
                                for (int cc = 0; cc < 4; ++cc) {
                                    uint32_t(&r)[32] = (cc & 1) ? rb : ra;
                                    uint32_t(&rn)[32] = (cc & 1) ? ra : rb;
                                    tmem_ld_wait();
                                    if (cc < 3) tmem_ld32(t_addr + (c + cc + 1) * 32, rn);
                                    square_chunk(r, cc);
                                }
                            }
                            rinv_head = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
                            if (which == 0) rinv_head *= p.q_scale;   // RMSNorm, RoPE and this factor are all linear in the row
                        }
                        if constexpr (EW == 8) fetch_ops(c, nxt);
                        stage_chunk(c, which != 2 ? rinv_head : 1.f, which != 2);
                    } else {
                        if constexpr (EW == 8) fetch_ops(c, nxt);
                        if (from_scratch) stage_from_scratch(c);
                        else stage_chunk(c, 1.f, false);
                    }
                    // sequence-parallel scatter: this 32-column chunk belongs to one head, i.e. to one destination rank;
                    // image rows go to row (my_rank * img_pad + local row) of that rank's gathered [q|k|v] buffer, text rows
                    // behind the image shards of all ranks at their index in the whole text sequence
                    [[maybe_unused]] __nv_bfloat16* scat = nullptr;
                    [[maybe_unused]] long long scat_ld = 0;
                    if constexpr (EPI == QIE_EPI_QKV_NORM_ROPE) {
                        if (p.peer_out) {
                            const int head = (n0 - which * p.model_dim) >> 7, g = head / p.sp_hl;
                            scat_ld = 3LL * p.sp_hl * 128;
                            scat = reinterpret_cast<__nv_bfloat16*>(__ldg(reinterpret_cast<const unsigned long long*>(p.peer_out) + g)) +
                                   ((long long)m.b * p.sp_gathered_rows + (m.s ? p.sp_txt_row0 : p.sp_rank * p.seq.img_pad) + local0) * scat_ld +
                                   (which * p.sp_hl + (head - g * p.sp_hl)) * 128 + (n0 & 127) + c4;
                        }
                    }

                    // phase 2: this chunk's operands were fetched one chunk ago; issue the next chunk's loads, then math and stores
                    const ChunkOps cur = nxt;
                    if constexpr (EW == 4)
                        if (c + 1 < c_end) fetch_ops(c + 1, nxt);
                    const float4 bv = cur.bv, g4 = cur.g4, nw4 = cur.nw4;
                    [[maybe_unused]] const float4 wsc = cur.wsc;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + sub;
                        const bool valid = local0 + rr < seg_rows;
                        float4 v = *reinterpret_cast<const float4*>(stg + rr * S::EPI_ROW_BYTES + c4 * 4);
                        if constexpr (FP8) {
                            if (!normed) {
                                v.x *= asc[it] * wsc.x; v.y *= asc[it] * wsc.y; v.z *= asc[it] * wsc.z; v.w *= asc[it] * wsc.w;
                            }
                        }
                        if (normed) {
                            const float4 cs = cur.aux[it];
                            const float x0 = v.x * nw4.x, x1 = v.y * nw4.y, x2 = v.z * nw4.z, x3 = v.w * nw4.w;
                            v.x = x0 * cs.x - x1 * cs.y; v.y = x0 * cs.y + x1 * cs.x;
                            v.z = x2 * cs.z - x3 * cs.w; v.w = x2 * cs.w + x3 * cs.z;
                        } else {
                            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                        }
                        if constexpr (EPI == QIE_EPI_GELU_BF16) {
                            v.x = gelu_tanh(v.x); v.y = gelu_tanh(v.y); v.z = gelu_tanh(v.z); v.w = gelu_tanh(v.w);
                            if (p.q8_amax) rmax[it] = fmaxf(rmax[it], fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                        }
                        if constexpr (EPI == QIE_EPI_GATE_RESID_F32) {
                            if (valid) {
                                float4 r4 = cur.aux[it];
                                r4.x += g4.x * v.x; r4.y += g4.y * v.y; r4.z += g4.z * v.z; r4.w += g4.w * v.w;
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (orow0 + rr) * p.ldo + n0 + c4) = r4;
                            }
                        } else if constexpr (EPI == QIE_EPI_F32) {
                            if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (orow0 + rr) * p.ldo + n0 + c4) = v;
                        } else {
                            const uint2 u = valid ? make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w)) : make_uint2(0u, 0u);
                            if (EPI == QIE_EPI_QKV_NORM_ROPE && scat) {
                                // NVLink peer store (or local when g == my rank); pad rows are not stored: in the gathered text
                                // region they are another rank's tokens
                                if (valid) *reinterpret_cast<uint2*>(scat + rr * scat_ld) = u;
                            } else
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (orow0 + rr) * p.ldo + n0 + c4) = u;
                        }
                    }
                    __syncwarp();   // staging tile is reused by the next chunk
                }
            }
            if constexpr (EPI == QIE_EPI_GELU_BF16) {
                // the 8-bit GEMM that consumes this output quantises per token: fold max|out| of every row (as the bf16 value
                // the quantiser will see; rounding is monotonic) into q8_amax[row] so that the cast needs no max pass of its own
                if (p.q8_amax && !dummy) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        float m = rmax[it];
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                        const int rr = it * 4 + sub;
                        if ((lane & 7) == 0 && local0 + rr < seg_rows)
                            atomicMax(reinterpret_cast<int*>(p.q8_amax + orow0 + rr), __float_as_int(__bfloat162float(__float2bfloat16(m))));
                    }
                }
            }
            // release this accumulator stage back to the MMA warp (of the leader CTA)
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && !released) {
                if constexpr (CG == 2) mbar_arrive_cluster(leader_smem_u32(&tempty_bar[acc]));
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    __syncwarp();        // reconverge the single-lane producer / MMA roles before the aligned barriers below
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers / smem must outlive the leader's last commit
    else __syncthreads();
    if (warp == EW + 1) {
        tc_fence_after();
        if constexpr (CG == 2) tmem_dealloc_cg2<S::TMEM_COLS>(tmem_base);
        else tmem_dealloc<S::TMEM_COLS>(tmem_base);
    }
}

template <int BN, int EPI, int QT, int CG, int EW>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB0, const CUtensorMap& tmB1, const GemmDev& p,
                       int num_tiles, cudaStream_t st) {
    using S = GemmSmem<BN, CG, EW>;
    QIE_CONFIGURE_ONCE(cudaFuncSetAttribute(gemm_kernel<BN, EPI, QT, CG, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::TOTAL));
    int units = sm_count() / CG;                 // CTAs (CG=1) or CTA pairs (CG=2) resident at once
    if (units > num_tiles) units = num_tiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(units * CG);
    cfg.blockDim = dim3(gemm_threads(EW));
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = launch_attrs(attr, CG);
    QIE_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, EPI, QT, CG, EW>, tmA, tmB0, tmB1, p));
    QIE_LAUNCH_OK("gemm_kernel");
    return QIE_OK;
}

template <int BN, int QT, int CG>
static int dispatch_epi(int epi, const CUtensorMap& a, const CUtensorMap& b0, const CUtensorMap& b1, const GemmDev& p,
                        int tiles, cudaStream_t st) {
    if constexpr (QT != 0 && BN == 256) {        // eight epilogue warps (see gemm_threads) for the short-K 8-bit shapes
        if (p.K < 8192) {
            switch (epi) {
                case QIE_EPI_BF16: return launch_gemm<BN, QIE_EPI_BF16, QT, CG, 8>(a, b0, b1, p, tiles, st);
                case QIE_EPI_GELU_BF16: return launch_gemm<BN, QIE_EPI_GELU_BF16, QT, CG, 8>(a, b0, b1, p, tiles, st);
                case QIE_EPI_F32: return launch_gemm<BN, QIE_EPI_F32, QT, CG, 8>(a, b0, b1, p, tiles, st);
                case QIE_EPI_GATE_RESID_F32: return launch_gemm<BN, QIE_EPI_GATE_RESID_F32, QT, CG, 8>(a, b0, b1, p, tiles, st);
                case QIE_EPI_QKV_NORM_ROPE: return launch_gemm<BN, QIE_EPI_QKV_NORM_ROPE, QT, CG, 8>(a, b0, b1, p, tiles, st);
            }
        }
    }
    switch (epi) {
        case QIE_EPI_BF16: return launch_gemm<BN, QIE_EPI_BF16, QT, CG, 4>(a, b0, b1, p, tiles, st);
        case QIE_EPI_GELU_BF16: return launch_gemm<BN, QIE_EPI_GELU_BF16, QT, CG, 4>(a, b0, b1, p, tiles, st);
        case QIE_EPI_F32: return launch_gemm<BN, QIE_EPI_F32, QT, CG, 4>(a, b0, b1, p, tiles, st);
        case QIE_EPI_GATE_RESID_F32: return launch_gemm<BN, QIE_EPI_GATE_RESID_F32, QT, CG, 4>(a, b0, b1, p, tiles, st);
        case QIE_EPI_QKV_NORM_ROPE:
            if constexpr (BN >= 128) return launch_gemm<BN, QIE_EPI_QKV_NORM_ROPE, QT, CG, 4>(a, b0, b1, p, tiles, st);
    }
    set_error("qie_gemm: unsupported epilogue %d for block_n %d", epi, BN);
    return QIE_EINVAL;
}

}  // namespace qie


using namespace qie;

int qie::g_pdl = 1;          // qie_tune(7, v): programmatic dependent launch of the per-block kernels (default on: bit-identical,
                            // +0.5 % on one GPU at the power cap, more for the short kernels of the sequence-parallel shards)
int g_gemm_l2_hints = 0;    // set through qie_tune(2, v)
int g_gemm_group_m = 0;     // qie_tune(5, v): m-units per raster band, 0 = default
int g_gemm_split_tail = 17; // qie_tune(4, v): bit 0 K split of the tail (long-K tiles, two ranges), bit 3 K split wherever it fits, bit 4 N split
                            // of the tail where no K split applies (default 1 | 16), bit 5 N split takes precedence over the K split

extern "C" int qie_gemm(const qie_gemm_args* g, const qie_seq* seq, void* stream) {
    QIE_REQUIRE(g && seq && g->a && g->out, QIE_EINVAL, "qie_gemm: null pointer");
    QIE_REQUIRE(g->streams >= 1 && g->streams <= 3, QIE_EINVAL, "qie_gemm: streams mask must be 1..3");
    QIE_REQUIRE(g->fp8 >= 0 && g->fp8 <= 2, QIE_EINVAL, "qie_gemm: fp8 (operand type) must be 0 bf16, 1 e4m3, 2 int8");
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
    p.q_scale = g->q_scale == 0.f ? 1.f : g->q_scale;
    p.a_scale = g->a_scale;
    for (int s = 0; s < 2; ++s) {
        p.bias[s] = g->bias[s];
        p.w_scale[s] = g->w_scale[s];
        for (int k = 0; k < 2; ++k) p.qk_norm_w[s][k] = g->qk_norm_w[s][k];
    }
    p.model_dim = g->N / 3;
    QIE_REQUIRE(!g->q8_amax || g->epilogue == QIE_EPI_GELU_BF16, QIE_EINVAL, "qie_gemm: q8_amax is produced by the GELU epilogue only");
    p.q8_amax = g->q8_amax;
    p.l2_hints = g_gemm_l2_hints;
    // raster band height: wide outputs (QKV, FF-up: >= 24 n-blocks) re-read A less with 16 m-units per band (ncu DRAM reads
    // 339 -> 239 MB and 429 -> 290 MB per launch), the N = 3072 shapes are best at 8 (profiles/r01_gemm_traffic.json)
    p.group_m = g_gemm_group_m > 0 ? g_gemm_group_m : (p.n_blocks >= 24 ? 16 : GEMM_GROUP_M);
    if (g->peer_out) {
        QIE_REQUIRE(g->epilogue == QIE_EPI_QKV_NORM_ROPE && g->sp_size >= 1 && g->sp_size <= 8 &&
                        g->sp_rank >= 0 && g->sp_rank < g->sp_size && (g->N / 3 / 128) % g->sp_size == 0 &&
                        g->sp_gathered_rows >= g->sp_size * seq->img_pad && g->sp_txt_row0 >= g->sp_size * seq->img_pad &&
                        g->sp_txt_row0 + seq->txt_rows <= g->sp_gathered_rows && g->streams == 3,
                    QIE_EINVAL, "qie_gemm: peer scatter needs the QKV epilogue, both streams, heads %% sp_size == 0 and a gathered "
                    "layout that holds every rank's image shard and the text tokens");
        p.peer_out = g->peer_out;
        p.sp_rank = g->sp_rank;
        p.sp_hl = g->N / 3 / 128 / g->sp_size;
        p.sp_gathered_rows = g->sp_gathered_rows;
        p.sp_txt_row0 = g->sp_txt_row0;
    }
    if (g->epilogue == QIE_EPI_QKV_NORM_ROPE) {
        QIE_REQUIRE(g->N % 3 == 0 && p.model_dim % bn == 0 && bn >= 128 && g->rope, QIE_ESHAPE,
                    "qie_gemm: QKV epilogue needs N=3D, D %% block_n == 0, block_n>=128, rope table");
        for (int s = 0; s < 2; ++s)
            if (g->streams & (1 << s)) {
                QIE_REQUIRE(p.qk_norm_w[s][0] && p.qk_norm_w[s][1], QIE_EINVAL, "qie_gemm: qk norm weights null");
                QIE_REQUIRE(p.bias[s], QIE_EINVAL, "qie_gemm: QKV epilogue needs a bias");
            }
    }

    // CTA pairs pay off once there are enough 256-row tiles to fill all 74 TPCs
    const int t0 = (g->streams & 1) ? seq->img_pad / 128 : 0, t1 = (g->streams & 2) ? seq->txt_pad / 128 : 0;
    const int pair_tiles = seq->batch * ((t0 + 1) / 2 + (t1 + 1) / 2) * p.n_blocks;
    int cg = g->cta_group;
    if (cg == 0) {
        // a CTA pair and a single CTA finish a 128-row block of a tile in the same time: take the form with fewer waves (small
        // sequence-parallel shards: 9 blocks x 48 n-blocks are 3 waves of 148 CTAs but 4 waves of 74 pairs), pairs on a tie
        const int sms = sm_count(), single_tiles = seq->batch * (t0 + t1) * p.n_blocks;
        const int waves2 = (pair_tiles + sms / 2 - 1) / (sms / 2), waves1 = (single_tiles + sms - 1) / sms;
        cg = (bn >= 128 && pair_tiles >= sms / 2 && waves2 <= waves1) ? 2 : 1;
    }
    QIE_REQUIRE(cg == 1 || (cg == 2 && bn >= 128), QIE_EINVAL, "qie_gemm: cta_group must be 1 or 2 (2 needs block_n >= 128)");

    const int rpb = seq->img_pad + seq->txt_pad;
    const int only = g->streams == 2 ? 1 : 0;
    const long long a_rows = g->a_compact ? (long long)seq->batch * (only ? seq->txt_pad : seq->img_pad)
                                          : (long long)seq->batch * rpb;
    CUtensorMap tmA, tmB[2];
    int rc = make_tmap_2d(&tmA, g->a, (uint64_t)a_rows, (uint64_t)g->K, (uint64_t)g->K * eb, GEMM_BM, bk, eb);
    if (rc) return rc;
    for (int s = 0; s < 2; ++s) {
        const void* w = g->w[s] ? g->w[s] : g->w[1 - s];
        rc = make_tmap_2d(&tmB[s], w, (uint64_t)g->N, (uint64_t)g->K, (uint64_t)g->K * eb, bn / cg, bk, eb);
        if (rc) return rc;
    }
    const int tiles = cg == 2 ? pair_tiles : seq->batch * (t0 + t1) * p.n_blocks;
    cudaStream_t st = (cudaStream_t)stream;
    // split-K tail (qie_tune key 4): only when the persistent schedule ends in a partial wave that a K split can
    // shorten, never for the QKV epilogue (row statistics over whole heads) or int8 (int32 partials would not survive fp32)
    p.tail_tiles = 0;
    p.tail_split = 1;
    {
        const int units = sm_count() / cg;
        const int tail = tiles % units, kblocks = (g->K + bk - 1) / bk;
        // (the eight-epilogue-warp instantiations — 8-bit operands, K < 8192 — take the N split below: their K split is short-K
        // by construction and has no GPU test)
        if ((g_gemm_split_tail & 9) && cg == 2 && bn == 256 && tiles > units && tail > 0 && g->epilogue != QIE_EPI_QKV_NORM_ROPE &&
            g->fp8 != 2 && !(g->fp8 && g->K < 8192)) {
            // the gain is (1 - 1/split) of a tile, the reduction costs a fixed few microseconds: worth it for long-K tiles only
            // (mode 1), or everywhere (mode 8 | 1, experiments / tests)
            int split = units / tail;
            if (split > 8) split = 8;
            if (!(g_gemm_split_tail & 8)) {
                if (split > 2) split = 2;
                if (kblocks < 96) split = 1;
            }
            while (split > 1 && kblocks / split < 6) --split;
            if (split > 1) {
                qie::StreamScratch sc;
                int rc0 = qie::stream_scratch(st, &sc);
                if (rc0) return rc0;
                float* scratch = sc.split_scratch;
                int* tickets = sc.split_tickets;
                if (tail * split <= qie::SPLIT_SCRATCH_TILES) {
                    p.tail_tiles = tail;
                    p.tail_split = split;
                    p.tail_dbg = (g_gemm_split_tail >> 1) & 3;
                    p.scratch = scratch;
                    p.tickets = tickets;
                }
            }
        }
    }
    // N-split tail (qie_tune key 4, bit 4): when the partial last wave holds at most half as many tiles as there are CTA pairs,
    // every tail tile becomes two independent BN/2-column items — the tail wave then costs ~0.55 of a tile instead of 1 (the
    // N = 128 MMA runs at ~90 % of the N = 256 rate) and needs no reduction, so it also serves the QKV epilogue (a half is
    // exactly one 128-column head) and the short-K shapes the K split does not pay for
    if ((g_gemm_split_tail & 16) && bn == 256 && (p.tail_tiles == 0 || (g_gemm_split_tail & 32))) {
        const int units = sm_count() / cg, tail = tiles % units;
        if (tail > 0 && 2 * tail <= units && g->N % 128 == 0) {
            p.tail_tiles = tail;
            p.tail_split = 2;
            p.tail_nsplit = 1;
            p.tail_dbg = 0;
        }
    }
#define QIE_GEMM_DISPATCH(F8, CGV)                                                                         \
    switch (bn) {                                                                                          \
        case 64:                                                                                           \
            if constexpr (CGV == 1) return dispatch_epi<64, F8, 1>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st); \
            break;                                                                                         \
        case 128: return dispatch_epi<128, F8, CGV>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st);      \
        default: return dispatch_epi<256, F8, CGV>(g->epilogue, tmA, tmB[0], tmB[1], p, tiles, st);       \
    }
    if (g->fp8 == 1) {
        if (cg == 2) { QIE_GEMM_DISPATCH(1, 2) } else { QIE_GEMM_DISPATCH(1, 1) }
    } else if (g->fp8 == 2) {
        if (cg == 2) { QIE_GEMM_DISPATCH(2, 2) } else { QIE_GEMM_DISPATCH(2, 1) }
    } else {
        if (cg == 2) { QIE_GEMM_DISPATCH(0, 2) } else { QIE_GEMM_DISPATCH(0, 1) }
    }
#undef QIE_GEMM_DISPATCH
    set_error("qie_gemm: no kernel for block_n=%d cta_group=%d", bn, cg);
    return QIE_EINVAL;
}
