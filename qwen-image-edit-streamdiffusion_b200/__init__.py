"""B200-native MMDiT denoise step for Qwen-Image-Edit-2509 (hot path of shi3z/Qwen-Image-Edit-StreamDiffusion)."""
from ._lib import QieError, build_library, lib, LIB_PATH, make_seq  # noqa: F401
from .transformer import B200QwenImageTransformer2DModel, QwenImageDiTConfig, Transformer2DModelOutput, merge_lora  # noqa: F401
from .pipeline import (StreamingDenoiser, cfg_euler_step, flowmatch_sigmas, model_timestep, pack_latents, run_denoise,  # noqa: F401
                       unpack_latents)
from .parallel import (ParallelLayout, ShardPlan, UlyssesTransformer, exchange_velocities, make_layout, make_shard_plan,  # noqa: F401
                       pack_heads, run_denoise_parallel, split_sizes, unpack_heads, emulate_fused_ulysses, PeerRankBuffers, make_peers, scatter_qkv_reference,
                       scatter_attn_reference, exchange_bytes_per_forward)
