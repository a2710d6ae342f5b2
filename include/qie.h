/*
 * libqie — C ABI of the B200-native MMDiT denoise step (Qwen-Image-Edit-2509).
 *
 * This is the drop-in boundary for the ONE hot path of shi3z/Qwen-Image-Edit-StreamDiffusion:
 * the `pipeline.transformer(...)` call (the attribute the reference itself swaps at
 * benchmark_lightning_compile.py:89-93, test_compiled.py:39-43, benchmark_compile.py:102-106)
 * plus the per-step true-CFG combine / FlowMatch-Euler update the diffusers pipeline runs
 * around it (reached from server.py:137-153, qwen_realtime.py:247-255, webui_realtime.py:77-85).
 *
 * Conventions
 *   - plain C types only; every pointer is a CUDA device pointer unless the name says `host`.
 *   - every entry point returns 0 (QIE_OK) or a negative qie_status; the text of the last error
 *     of the calling thread is available from qie_last_error().  No exception crosses the ABI.
 *   - all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *     no entry point synchronises the device in the steady state, none allocates inside
 *     qie_forward (CUDA-graph capturable).
 *   - the library never frees caller memory; caller keeps weights/workspace alive.
 *   - threading: a handle (its timestep / modulation buffers) and a workspace serve ONE forward at a time; concurrent
 *     forwards (cond / uncond on two streams, README.md:127-128) take one handle + workspace per stream.  Kernel-internal
 *     scratch (row counters, split-K partials) is kept per (device, stream), so launches on different streams never share
 *     mutable state; at most 16 distinct streams per device are served (QIE_ESTATE beyond).
 *   - sm_100a only: on any other device qie_create returns QIE_EARCH.  There is no CPU path.
 *
 * Joint sequence layout used by all per-token entry points ("qie_seq"):
 *   each batch element owns rows_per_batch = img_pad + txt_pad rows; rows [0,img_rows) are image
 *   tokens (noise latents then reference-image latents, the order of `hidden_states`), rows
 *   [img_pad, img_pad+txt_rows) are text tokens; img_pad/txt_pad are img_rows/txt_rows rounded up
 *   to 128.  Attention is permutation invariant over keys, so [img; txt] is equivalent to the
 *   reference's cat([txt, img]) once each token carries its own RoPE row.
 */
#ifndef QIE_H_
#define QIE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QIE_ABI_VERSION 1

typedef enum qie_status {
    QIE_OK = 0,
    QIE_EINVAL = -1, /* bad argument */
    QIE_ESHAPE = -2, /* unsupported shape */
    QIE_ECUDA = -3,  /* CUDA runtime / driver error */
    QIE_ENOMEM = -4, /* workspace too small */
    QIE_EARCH = -5,  /* device is not sm_100 */
    QIE_ESTATE = -6  /* weights not set etc. */
} qie_status;

typedef struct qie_handle qie_handle;

/* replaces: QwenImageTransformer2DModel.config (diffusers transformer/config.json) */
typedef struct qie_model_cfg {
    int num_layers;  /* 60 */
    int num_heads;   /* 24 */
    int head_dim;    /* 128 (the only supported value) */
    int in_channels; /* 64 */
    int out_dim;     /* patch_size^2 * out_channels = 64 */
    int joint_dim;   /* 3584 */
    int rope_axes[3]; /* 16,56,56 ; sums to head_dim */
} qie_model_cfg;

typedef struct qie_seq {
    int batch;
    int img_rows; /* valid image tokens per batch element */
    int txt_rows; /* valid text tokens per batch element: the MAXIMUM over the batch when the lengths differ */
    int img_pad;  /* img_rows rounded up to 128 */
    int txt_pad;  /* txt_rows rounded up to 128 */
    int txt_rows_b[8]; /* valid text tokens of every batch element (1 .. txt_rows; qie_make_seq fills them with txt_rows).  Rows behind
                        * them are padding for every kernel: masked keys in attention, zero rows elsewhere.  This is what lets the cond and
                        * the uncond prompt of a true-CFG step (different lengths) share ONE forward — replaces the reference's
                        * batched_cfg_pipeline.py (README.md:126) without letting pad tokens take part in attention */
} qie_seq;

/* weights of one QwenImageTransformerBlock; index [0] = image stream, [1] = text stream.
 * All matrices are nn.Linear layout [out, in] row-major bf16 (or e4m3 when the *_w8 pointers are
 * set); biases, norm weights and dequant scales are fp32.  state_dict keys: SURVEY App. A.10. */
typedef struct qie_block_weights {
    const void* qkv_w[2];          /* [3D, D]  to_q|to_k|to_v  /  add_q_proj|add_k_proj|add_v_proj */
    const float* qkv_b[2];         /* [3D] */
    const float* q_norm_w[2];      /* [128] norm_q / norm_added_q */
    const float* k_norm_w[2];      /* [128] norm_k / norm_added_k */
    const void* out_w[2];          /* [D, D]   to_out.0 / to_add_out */
    const float* out_b[2];
    const void* ff1_w[2];          /* [4D, D]  {img,txt}_mlp.net.0.proj */
    const float* ff1_b[2];
    const void* ff2_w[2];          /* [D, 4D]  {img,txt}_mlp.net.2 */
    const float* ff2_b[2];
    /* optional FP8 (e4m3) copies with per-output-channel fp32 scales; NULL = bf16 path */
    const void* qkv_w8[2];  const float* qkv_ws[2];
    const void* out_w8[2];  const float* out_ws[2];
    const void* ff1_w8[2];  const float* ff1_ws[2];
    const void* ff2_w8[2];  const float* ff2_ws[2];
} qie_block_weights;

typedef struct qie_weights {
    const void* img_in_w;  const float* img_in_b;      /* [D, in_channels] */
    const float* txt_norm_w;                           /* [joint_dim] */
    const void* txt_in_w;  const float* txt_in_b;      /* [D, joint_dim] */
    const void* t1_w;      const float* t1_b;          /* time_text_embed.timestep_embedder.linear_1 [D,256] */
    const void* t2_w;      const float* t2_b;          /* ...linear_2 [D,D] */
    const void* mod_w;     const float* mod_b;         /* all img_mod.1/txt_mod.1 stacked: [L][2][6D, D] bf16, [L][2][6D] */
    const void* norm_out_w; const float* norm_out_b;   /* norm_out.linear [2D, D] */
    const void* proj_out_w; const float* proj_out_b;   /* [out_dim, D] */
    const qie_block_weights* blocks;                   /* HOST array of num_layers entries (copied) */
} qie_weights;

/* ---- library ---- */
int qie_version(void);
const char* qie_last_error(void);
int qie_device_sm_count(void);

/* ---- model handle: replaces constructing/holding `pipeline.transformer` (server.py:66-69) ---- */
int qie_create(const qie_model_cfg* cfg, int device, qie_handle** out);
int qie_destroy(qie_handle* h);
/* qie_set_weights keeps the pointers (the caller owns the memory) and reads the four QK-RMSNorm weight vectors of every block back
 * once (synchronous 512 B copies) to bound the attention scores (qie_attn_score_bound): call it again after changing them in place */
int qie_set_weights(qie_handle* h, const qie_weights* w);
/* 0 = bf16 GEMMs, 1 = FP8 e4m3 W8A8, 2 = INT8 W8A8 (both need the *_w8 / *_ws pointers); replaces int8_linear.py /
 * cublaslt_int8.py / triton_int8_gemm.py named at README.md:136-141 */
int qie_set_precision(qie_handle* h, int mode);
/* tuning/debug knobs outside the reference surface: key 0 = fuse QK-norm+RoPE into the QKV GEMM epilogue (default 1),
 * key 1 = attention kernel variant, key 2 = record CUDA events around every kernel class inside qie_forward,
 * key 3 = bounded-score attention where the norm weights allow it (see qie_attn_score_bound) */
int qie_set_option(qie_handle* h, int key, int value);
/* key 3 of qie_set_option: 1 (default) = blocks whose score bound (below) is <= QIE_ATTN_SCORE_BOUND run the bounded-score
 * attention (q pre-scaled in the QKV epilogue, qie_attn_fwd variant 0x200), 0 = every block runs the online-softmax kernel.
 * qie_attn_score_bound: the bound on |q.k| * softmax_scale * log2(e) of block `layer`, derived at qie_set_weights from the
 * four QK-RMSNorm weight vectors of the block (read back once, 2 % margin for the bf16 rounding of q and k); < 0 on error */
float qie_attn_score_bound(const qie_handle* h, int layer);
/* The attention variant that matches the q|k|v the QKV phase of block `layer` produces (>= 0; negative = status): callers of
 * the phase API that run the attention themselves (qie_forward_phase QKV -> their own exchange -> qie_attn_fwd / _tiles)
 * MUST pass this value, because q is pre-scaled in the blocks that run the bounded-score form. */
int qie_attn_layer_variant(const qie_handle* h, int layer);
/* measurement aids: kernels launched by the library so far; event-timed ms / algorithmic work / launches per kernel
 * class since the last read (class 0 GEMM [FLOP], 1 attention [FLOP], 2 adaLN [bytes], 3 modulation GEMV [bytes], 4 other,
 * 5 peer barrier); arrays of QIE_PROFILE_CLASSES entries.  qie_profile_timeline lists the launches recorded since the last
 * qie_profile_read (start offset from the first one, duration, class) without consuming them. */
#define QIE_PROFILE_CLASSES 6
unsigned long long qie_launch_count(void);
/* process-wide experiment / launch knobs: 0 adaLN threads/block, 1 adaLN smem reservation, 2 GEMM L2 hints (bit 0 weights
 * evict-last, bit 1 activations evict-first), 3 adaLN kernel form (2 = CTA rows where D % 1024 == 0 [default], 1 = warp-per-row
 * streaming ring, 0 = one warp per row), 4 GEMM split-K tail, 5 GEMM raster band,
 * 7 programmatic dependent launch of the GEMM / attention / adaLN / barrier kernels (1 on = default, 0 off).
 * qie_tune_get returns the current value (>= 0) or QIE_EINVAL for an unknown key. */
int qie_tune(int key, int value);
int qie_tune_get(int key);
int qie_profile_read(qie_handle* h, double* ms, double* work, int* launches);
int qie_profile_timeline(qie_handle* h, float* start_ms, float* dur_ms, int* cls, int max_n);
/* host helper: pad a (img_rows, txt_rows) pair into the joint layout */
int qie_make_seq(int batch, int img_rows, int txt_rows, qie_seq* out);
/* ... with one text length per batch element (encoder_hidden_states is then [batch, max length, joint_dim], rows behind an
 * element's length are ignored) */
int qie_make_seq_ragged(int batch, int img_rows, const int* txt_rows_b_host, qie_seq* out);
size_t qie_workspace_bytes(const qie_handle* h, const qie_seq* seq);

/* replaces: QwenImageTransformer2DModel.forward(hidden_states, encoder_hidden_states, timestep,
 *           img_shapes, txt_seq_lens), called by the pipeline loop.
 *   hidden   bf16 [B, img_rows, in_channels]        enc  bf16 [B, txt_rows, joint_dim]
 *   timestep fp32 [B] device (sigma in [0,1], i.e. already /1000 as the pipeline passes it)
 *   img_shapes_host  int[n_img*3] (f,h,w) of ONE batch element (QwenEmbedRope uses element 0)
 *   out      bf16 [B, img_rows, out_dim]
 *   n_blocks < 0 runs all layers (a positive value truncates the stack — baseline/bench aid). */
int qie_forward(qie_handle* h, const void* hidden, const void* enc, const float* timestep,
                const int* img_shapes_host, int n_img, const qie_seq* seq, void* out,
                void* workspace, size_t workspace_bytes, int n_blocks, void* stream);

/* ---- phase-wise forward for sequence-parallel (Ulysses) callers: the caller runs NCCL all-to-alls between the QKV
 * and ATTN phases and between ATTN and POST (replaces nothing in the reference — it has no multi-GPU attention; it is the
 * partition north_star asks for).  `seq` then describes the LOCAL token shard of this rank and `sp` places it inside the
 * whole sequence so the RoPE rows match.  qie_forward == qie_forward_phase(QIE_PHASE_ALL, layer -1, sp NULL). */
#define QIE_PHASE_BEGIN 1
#define QIE_PHASE_QKV 2
#define QIE_PHASE_ATTN 4
#define QIE_PHASE_POST 8
#define QIE_PHASE_END 16
#define QIE_PHASE_ALL 31
typedef struct qie_sp {
    int rank, size;
    int img_total, txt_total;   /* tokens of the whole (unsharded) sequence */
    int img_offset, txt_offset; /* first image / text token owned by this rank */
} qie_sp;
int qie_forward_phase(qie_handle* h, int phases, int layer /* -1 = all layers */, const void* hidden, const void* enc,
                      const float* timestep, const int* img_shapes_host, int n_img, const qie_seq* seq, const qie_sp* sp,
                      void* out, void* workspace, size_t workspace_bytes, int n_blocks, void* stream);
/* byte offset inside the workspace of: 0 = qkv [rows, 3D] bf16, 1 = attention output [rows, D] bf16 */
long long qie_workspace_offset(const qie_handle* h, const qie_seq* seq, int which);
/* joint attention over an explicit list of 128-row KV tiles: qkv bf16 [n_tiles*128, 3*heads*128], tile_valid (device
 * int[n_tiles]) = valid rows of each tile (>= 1); every row is also a query row. */
int qie_attn_fwd_tiles(const void* qkv, void* out, int n_tiles, const int* tile_valid_dev, int num_heads, int variant,
                       void* stream);

/* ---- fused Ulysses exchange over NVLink peer memory (SURVEY §8e "fusion target"): instead of pack kernel + NCCL
 * all-to-all + unpack kernel around the attention of every block, the producing kernels store straight into the
 * consumer rank's buffers through peer-mapped pointers:
 *   - the QKV GEMM epilogue (RMSNorm + RoPE applied) writes head group g of q|k|v into rank g's gathered buffer
 *     qkv_gather[g] [size*rows_pad, 3*(H/size)*128] at row (my_rank*rows_pad + local row);
 *   - the attention epilogue of rank g writes its heads' output for the tokens of rank s into attn_out[s] [rows_pad, H*128].
 *   - the END phase stores the rank's velocity rows into every rank's velocity buffer vel[s] [batch, img_total, out_dim].
 * The only synchronisation left is qie_peer_barrier between the phases (one flag store per peer + one spin per peer).
 * Buffers must be peer-accessible: allocate them with qie_peer_alloc and exchange the 64-byte handles between the
 * ranks (torch.distributed all_gather of bytes), map with qie_peer_open.  Single-GPU emulation (tests): point the
 * tables at local buffers of the emulated ranks and run the ranks' phases one after the other, without the barrier. */
typedef struct qie_peers {
    int rank, size;          /* my rank in the sequence-parallel group, group size (2..8, divides num_heads) */
    int batch;               /* frames per forward (all of them stay inside the group) */
    int img_pad, txt_pad;    /* padded rows of ONE rank's image / text shard as qie_sp_shard returns them (identical on every rank) */
    int img_total, txt_total;/* tokens of the whole (unsharded) sequence of one frame */
    void* qkv_gather[8];     /* rank g's gathered q|k|v buffer bf16 [batch][size*img_pad + pad128(txt_total)][3*(H/size)*128]:
                              * image shard of rank r at rows [r*img_pad, ...), ALL text tokens contiguous behind the image shards
                              * (the gathered sequence is as long as the single-GPU one: no per-rank text padding inside it) */
    void* attn_out[8];       /* rank s's attention-output buffer = its workspace + qie_workspace_offset(..., 1) */
    void* vel[8];            /* rank s's velocity buffer bf16 [batch][img_total][out_dim]: every rank receives all rows */
    void* flags[8];          /* rank s's barrier words: 16 zero-initialised uint32 ([0..7] arrivals, [8] its barrier count) */
    void* mod[8];            /* rank s's modulation table fp32 [batch][num_layers*12*D]: in the BEGIN phase every rank computes 1/size
                              * of the rows (the 13.6 GB modulation-weight stream does not shrink with the token shard) and stores them
                              * into the other ranks' tables; a barrier must separate BEGIN from the first QKV phase */
} qie_peers;
/* installs (or with NULL removes) the peer tables of ONE geometry (batch, img_total, txt_total), used by the sequence-parallel
 * entry points (qie_forward_phase with `sp`, qie_forward_sp); a plain qie_forward on the same handle ignores them.  Call it again whenever the
 * geometry of the next forward differs (cheap: host-side, the tile list of each geometry is built once and kept).  A forward whose
 * seq / sp do not match the installed geometry returns QIE_ESTATE.  qie_set_peers(h, NULL, ...) returns QIE_ECUDA (once) if a
 * barrier of the dissolved group had timed out. */
int qie_set_peers(qie_handle* h, const qie_peers* peers, void* stream);
/* host helper: token shard of `rank` (image and text tokens each split contiguously, the first total % size ranks own one
 * more) -> local layout + placement; QIE_ESHAPE when a rank would own an all-padding 128-row tile */
int qie_sp_shard(int batch, int img_total, int txt_total, int size, int rank, qie_seq* seq_out, qie_sp* sp_out);
/* host helper: valid rows of every 128-row tile of the gathered sequence; returns the tile count */
int qie_sp_tile_valid_host(int img_total, int txt_total, int size, int* out_host, int max_tiles);
/* cudaMalloc'ed, zero-filled, IPC-exportable device buffer; handle_out receives the 64-byte cudaIpcMemHandle_t */
int qie_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle_out64);
int qie_peer_free(void* dev_ptr);
/* stream-ordered device-to-device copy out of / into such a buffer (tests: the buffers have no torch tensor around them) */
int qie_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
/* maps a buffer exported by another process of this node (enables peer access to its device on first use) */
int qie_peer_open(const unsigned char* handle64, void** dev_ptr);
int qie_peer_close(void* dev_ptr);
/* all-ranks barrier of the installed group on `stream`, system-scope release/acquire over the ranks' flag words.  The epoch
 * lives in device memory (flags[rank][8]) and is advanced by the kernel itself, so the launch is replayable from a CUDA graph;
 * every rank must execute the same sequence of barriers.  A barrier that waits longer than ~2 s gives up and raises a sticky,
 * process-wide error: every later qie_forward / qie_forward_phase / qie_forward_sp returns QIE_ECUDA (see qie_last_error). */
int qie_peer_barrier(qie_handle* h, void* stream);
/* 0, or 1 once a barrier has timed out (read from mapped host memory: no stream is synchronised) */
int qie_peer_barrier_timeouts(void);
/* The whole sequence-parallel forward of this rank as ONE call (peers installed): BEGIN -> barrier, per block QKV -> barrier -> ATTN ->
 * barrier -> POST, END -> barrier; `out_full` bf16 [batch, img_total, out_dim] receives the velocity of ALL tokens (every rank's
 * END stores its rows into every rank's `vel` buffer).  No allocation, no host synchronisation: CUDA-graph capturable. */
int qie_forward_sp(qie_handle* h, const void* hidden_local, const void* enc_local, const float* timestep,
                   const int* img_shapes_host, int n_img, const qie_seq* seq, const qie_sp* sp, void* out_full,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- exact caches ("next" row N1): replaces cached_pipeline_v2.py (README.md:125) and the
 * precompute_conditions stub of qwen_realtime.py:140-165.  temb, every block's modulation vectors and the final
 * scale/shift depend on the timestep only; txt_in(txt_norm(prompt_embeds)) on the prompt only.  Cached forwards are
 * bit-identical to uncached ones.  qie_cache_schedule / qie_cache_prompt are setup-time calls (allocate + synchronise). */
int qie_cache_schedule(qie_handle* h, const float* timesteps_host, int n, void* stream);
int qie_cache_prompt(qie_handle* h, int slot /* 0..3 */, const void* enc /* bf16 [txt_rows, joint_dim] */, int txt_rows,
                     void* stream);
/* selection applied by the following qie_forward calls: sched_idx_host[b] per batch row (NULL or -1 = compute), prompt
 * slot (-1 = compute) */
int qie_cache_select(qie_handle* h, const int* sched_idx_host, int batch, int prompt_slot);

/* replaces: the true-CFG combine + norm rescale of QwenImageEditPlusPipeline.__call__ and
 * FlowMatchEulerDiscreteScheduler.step, fused; v_uncond may be NULL (cond-only).
 *   v_* bf16 rows of `v_row_stride` elements, first `channels` used; latents bf16 [rows, channels] in/out. */
int qie_cfg_euler_step(const void* v_cond, const void* v_uncond, void* latents, float true_cfg_scale,
                       float sigma, float sigma_next, int batch, int tokens, int channels,
                       int v_tokens_stride, void* stream);

/* ---- "next" row N3: the layout step either side of the denoise loop, fused with the VAE-latent
 * (de)normalisation.  replaces QwenImageEditPlusPipeline._pack_latents / _unpack_latents and the
 * `(z - latents_mean) / latents_std` (encode side) / `z * latents_std + latents_mean` (decode side) tensor ops.
 *   z       bf16 [B, C, h, w]   (C = in_channels / 4 = 16 latent channels, h and w even)
 *   tokens  bf16 [B, (h/2)*(w/2), 4*C]; channel index of a token = c*4 + dy*2 + dx
 *   mean/std fp32 [C] device, both NULL = pure re-layout (noise latents) */
int qie_pack_latents(const void* z, const float* mean, const float* std, void* tokens, int batch, int channels, int h,
                     int w, void* stream);
int qie_unpack_latents(const void* tokens, const float* mean, const float* std, void* z, int batch, int channels, int h,
                       int w, void* stream);

/* host-only: FlowMatchEulerDiscreteScheduler.set_timesteps (dynamic exponential shift + terminal stretch);
 * writes num_steps+1 sigmas.  replaces scheduler.set_timesteps(sigmas, mu) */
int qie_flowmatch_sigmas(int num_steps, int image_seq_len, float* sigmas_host);
/* host-only: QwenEmbedRope table in the joint layout: out_host float[rows_per_batch*64*2] (cos,sin) */
int qie_rope_table_host(const qie_model_cfg* cfg, const int* img_shapes_host, int n_img, const qie_seq* seq,
                        float* out_host);

/* ---- per-kernel entry points (unit tests, profiling) ---- */
enum qie_epilogue {
    QIE_EPI_BF16 = 0,          /* out_bf16 = acc + bias                                  */
    QIE_EPI_GELU_BF16 = 1,     /* out_bf16 = gelu_tanh(acc + bias)                       */
    QIE_EPI_F32 = 2,           /* out_f32  = acc + bias                                  */
    QIE_EPI_GATE_RESID_F32 = 3,/* out_f32 += gate[b, stream, n] * (acc + bias)           */
    QIE_EPI_QKV_NORM_ROPE = 4  /* bf16 qkv with per-head RMSNorm + RoPE on q,k columns   */
};
typedef struct qie_gemm_args {
    const void* a;            /* bf16 (or e4m3) [rows, K] */
    int a_compact;            /* 0: A rows are joint-layout rows; 1: A is [B*seg_pad, K] of the single enabled stream */
    const void* w[2];         /* per-stream weights [N, K] */
    const float* bias[2];     /* per-stream bias [N] (may be NULL) */
    void* out;                /* [rows, ldo] */
    int out_compact;          /* same meaning as a_compact for the output rows */
    int ldo;                  /* output leading dimension (elements) */
    int N, K;
    int streams;              /* bit0: image rows, bit1: text rows */
    int epilogue;             /* enum qie_epilogue */
    const float* gate;        /* GATE_RESID: gate vectors, element (b, stream, n) at gate[b*gate_bstride + stream*gate_sstride + n] */
    long long gate_bstride, gate_sstride;
    const float* rope;        /* QKV_NORM_ROPE: [rows_per_batch, 64, 2] */
    const float* qk_norm_w[2][2]; /* QKV_NORM_ROPE: [stream][q/k] -> [128] */
    int fp8;                  /* operand type: 0 bf16; 1 e4m3, 2 int8: a_scale [rows] and w_scale[stream][N] dequantise in the epilogue */
    const float* a_scale;
    const float* w_scale[2];
    int block_n;              /* 0 = auto */
    int cta_group;            /* 0 = auto, 1 = one CTA per 128-row tile, 2 = CTA pair per 256-row tile (tcgen05 cta_group::2) */
    /* QKV_NORM_ROPE only: scatter the output over the sequence-parallel group instead of writing `out` (see qie_peers):
     * peer_out = DEVICE array of sp_size pointers to the ranks' gathered buffers [batch][sp_gathered_rows][3*(H/sp_size)*128];
     * a valid image row i of batch b goes to row b*sp_gathered_rows + sp_rank*img_pad + i, a valid text row i to row
     * b*sp_gathered_rows + sp_txt_row0 + i (sp_txt_row0 = sp_size*img_pad + first text token of this rank) */
    void* const* peer_out;
    int sp_rank, sp_size, sp_gathered_rows, sp_txt_row0;
    /* GELU_BF16 / BF16 epilogues feeding an 8-bit GEMM (fp8 / int8 modes): when q8_amax is set the epilogue also folds
     * max|out| of every output row into q8_amax[row] (fp32, atomicMax on the bit pattern; the caller zeroes it) so that the
     * per-token quantiser that follows needs no separate max pass */
    float* q8_amax;
    /* QKV_NORM_ROPE only: factor applied to the q columns after RMSNorm and RoPE, before the bf16 rounding (0 = 1.0).
     * qie_forward passes softmax_scale * log2(e) in the blocks whose attention runs in the bounded-score form (qie_attn_fwd 0x200) */
    float q_scale;
} qie_gemm_args;
int qie_gemm(const qie_gemm_args* args, const qie_seq* seq, void* stream);

/* joint attention over all valid rows of each batch element; qkv bf16 [rows, 3*H*128] (q|k|v), q and k already
 * normed + roped; out bf16 [rows, H*128].  replaces F.scaled_dot_product_attention in
 * QwenDoubleStreamAttnProcessor2_0. variant: 0 = tuned default (CTA-pair kernel); 0x8 selects the single-CTA
 * fallback kernel; bits 4..7 = how many of every 8 score pairs use the FMA-pipe polynomial exp2 (0x100 = none, 2, 3, 4).
 * 0x200 = bounded-score form of the CTA-pair kernel: the CALLER promises that q already carries softmax_scale * log2(e)
 * (qie_gemm_args.q_scale) and that |q.k| <= QIE_ATTN_SCORE_BOUND for every (query, key) pair, so p = 2^(q.k) needs no running
 * max, no exchange between the softmax warpgroups and no rescale of O; the result is the same softmax.  After QK-RMSNorm the
 * bound follows from the norm weights: |q.k| * softmax_scale * log2(e) <= 128 * max_p(m_q,p * m_k,p) * softmax_scale * log2(e) with
 * m_p the larger |weight| of RoPE pair p over both streams (RoPE rotates inside a pair; Cauchy-Schwarz over the pairs). */
#define QIE_ATTN_SCORE_BOUND 80.0f
int qie_attn_fwd(const void* qkv, void* out, const qie_seq* seq, int num_heads, int variant, void* stream);
/* LayerNorm(no affine, eps) + x*(1+scale)+shift; x fp32 [rows, D] -> out bf16 [rows, D].
 * shift/scale for (b, stream) at mod[b*mod_bstride + stream*mod_sstride + {shift_off,scale_off} + c].
 * if out8/out_scale non-NULL also emits e4m3 / int8 rows with per-row scale (for the W8A8 GEMMs); `out` may then be NULL (the
 * W8A8 forward reads the 8-bit rows only: a third of the kernel's bytes are not written). */
int qie_ln_modulate(const float* x, const float* mod, long long mod_bstride, long long mod_sstride, int shift_off,
                    int scale_off, void* out, void* out8, float* out_scale, int qmode /* 1 e4m3, 2 int8 */, int D, float eps,
                    const qie_seq* seq, void* stream);
/* y[b, n] = bias[n] + sum_k act(x[b,k]) * W[n,k]; act: 0 none, 1 SiLU. W bf16 [N,K], x/y fp32. batch<=8 */
int qie_gemv(const float* x, const void* w, const float* bias, float* y, int batch, long long N, int K, int act,
             void* stream);
/* sinusoidal timestep projection (Timesteps(256, flip_sin_to_cos, scale=1000)): t fp32 [B] -> out fp32 [B,256] */
int qie_timestep_proj(const float* t, float* out, int batch, int round_bf16, void* stream);
/* per-head RMSNorm(weight) + interleaved-pair RoPE, in place on the q and k column blocks of qkv */
int qie_qk_norm_rope(void* qkv, const float* rope, const float* const* norm_w /* [2 streams][2 q,k] */,
                     int num_heads, float eps, const qie_seq* seq, void* stream);
/* RMSNorm(weight) over rows: x bf16 [B, n, D] -> out bf16 [B, n_pad, D] (pad rows zero) */
int qie_rmsnorm_pack(const void* x, const float* w, void* out, int batch, int n, int n_pad, int D, float eps,
                     void* stream);
/* copy [B, n, C] bf16 -> [B, n_pad, C] with zero pad rows */
int qie_pack_rows(const void* x, void* out, int batch, int n, int n_pad, int C, void* stream);
/* per-row dynamic symmetric quantisation: x bf16 [rows, K] -> q [rows,K] (qmode 1: e4m3, scale amax/448; 2: int8, scale
 * amax/127, round-to-nearest-even), scale fp32 [rows] */
int qie_quant_rows(const void* x, void* q, float* scale, long long rows, int K, int qmode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QIE_H_ */
