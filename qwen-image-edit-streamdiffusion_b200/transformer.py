"""Drop-in for `pipeline.transformer` of diffusers' QwenImageEditPlusPipeline.

The reference swaps that attribute itself (benchmark_lightning_compile.py:89-93,
test_compiled.py:39-43); this module is the same seam: an `nn.Module` whose `forward` keeps the
`QwenImageTransformer2DModel.forward` keyword surface and runs the whole 60-block
step through `qie_forward` in libqie.so (hand-written sm_100a kernels).  PyTorch is only the
allocator / stream owner here; there is no eager fallback.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib as L


@dataclass(frozen=True)
class QwenImageDiTConfig:
    """transformer/config.json of Qwen/Qwen-Image-Edit-2509."""
    patch_size: int = 2
    in_channels: int = 64
    out_channels: int = 16
    num_layers: int = 60
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 3584
    guidance_embeds: bool = False
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def out_dim(self) -> int:
        return self.patch_size * self.patch_size * self.out_channels


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor


def _block_keys(i: int) -> Dict[str, str]:
    p = f"transformer_blocks.{i}."
    return {
        "q": [p + "attn.to_q", p + "attn.add_q_proj"], "k": [p + "attn.to_k", p + "attn.add_k_proj"],
        "v": [p + "attn.to_v", p + "attn.add_v_proj"],
        "nq": [p + "attn.norm_q", p + "attn.norm_added_q"], "nk": [p + "attn.norm_k", p + "attn.norm_added_k"],
        "out": [p + "attn.to_out.0", p + "attn.to_add_out"],
        "ff1": [p + "img_mlp.net.0.proj", p + "txt_mlp.net.0.proj"],
        "ff2": [p + "img_mlp.net.2", p + "txt_mlp.net.2"],
        "mod": [p + "img_mod.1", p + "txt_mod.1"],
    }


def merge_lora(sd: Dict[str, torch.Tensor], lora_sd: Dict[str, torch.Tensor], scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """W' = W + scale * (alpha / r) * B @ A for every targeted Linear, folded before packing / quantisation, so the two extra
    skinny GEMMs PEFT runs per Linear (`pipeline.load_lora_weights`, server.py:76-79) disappear.
    Accepts PEFT / diffusers naming (`<module>.lora_A.weight`, `<module>.lora_B.weight`, optional `<module>.alpha`) and the
    kohya-style `lora_down` / `lora_up`, with an optional `transformer.` or `diffusion_model.` prefix."""
    out = dict(sd)
    pairs = {}
    for k, v in lora_sd.items():
        base = k
        for pre in ("transformer.", "diffusion_model.", "base_model.model."):
            if base.startswith(pre):
                base = base[len(pre):]
        for a_tag, b_tag in ((".lora_A.weight", ".lora_B.weight"), (".lora_down.weight", ".lora_up.weight"),
                             (".lora_A.default.weight", ".lora_B.default.weight")):
            if base.endswith(a_tag):
                pairs.setdefault(base[: -len(a_tag)], {})["A"] = v
            elif base.endswith(b_tag):
                pairs.setdefault(base[: -len(b_tag)], {})["B"] = v
        if base.endswith(".alpha"):
            pairs.setdefault(base[: -len(".alpha")], {})["alpha"] = float(v)
    for mod, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise L.QieError(f"LoRA entry for {mod} is incomplete")
        key = mod + ".weight"
        if key not in out:
            raise L.QieError(f"LoRA targets {mod}, which is not a Linear of this transformer")
        A, B = ab["A"].float(), ab["B"].float()
        r = A.shape[0]
        s_ = scale * (ab.get("alpha", float(r)) / r)
        out[key] = (out[key].float() + s_ * (B @ A)).to(out[key].dtype)
    return out


class B200QwenImageTransformer2DModel(nn.Module):
    """`pipe.transformer = B200QwenImageTransformer2DModel.from_state_dict(sd)`."""

    def __init__(self, config: QwenImageDiTConfig = QwenImageDiTConfig(), device: str | torch.device = "cuda:0"):
        super().__init__()
        if config.attention_head_dim != 128:
            raise L.QieError("the sm_100a kernels support attention_head_dim == 128 only")
        self.cfg = config
        self.config = SimpleNamespace(**{k: getattr(config, k) for k in config.__dataclass_fields__})
        self._device = torch.device(device)
        self._t: Dict[str, torch.Tensor] = {}      # packed device tensors kept alive for the library
        self._handle = C.c_void_p()
        self._ws: Optional[torch.Tensor] = None
        self._blocks_arr = None
        self._fp8_ready = False
        # nn.Module bookkeeping so `.parameters()`, `.dtype`, `.device` behave (benchmark_int8.py:22)
        self._anchor = nn.Parameter(torch.zeros(1, dtype=torch.bfloat16, device=self._device), requires_grad=False)
        lib = L.lib()
        c = L.ModelCfg(config.num_layers, config.num_attention_heads, config.attention_head_dim, config.in_channels,
                       config.out_dim, config.joint_attention_dim, (C.c_int * 3)(*config.axes_dims_rope))
        self._ccfg = c
        with torch.cuda.device(self._device):
            L.check(lib.qie_create(C.byref(c), self._device.index or 0, C.byref(self._handle)), "qie_create")

    # ------------------------------------------------------------------ module surface
    @property
    def dtype(self) -> torch.dtype:
        return torch.bfloat16

    @property
    def device(self) -> torch.device:
        return self._device

    def to(self, *args, **kwargs):
        """`pipeline.to("cuda")` reaches every component (server.py:70-71).  The packed weights live where the module was
        built and are bf16 by construction: the same device / bf16 is a no-op, anything else fails loudly instead of moving
        only the bookkeeping parameter."""
        device, dtype, _, _ = torch._C._nn._parse_to(*args, **kwargs)
        if device is not None:
            device = torch.device(device)
            if device.type != "cuda" or (device.index is not None and device.index != (self._device.index or 0)):
                raise L.QieError(f"the packed weights live on {self._device}; build the module on the target device "
                                 f"instead of .to({device}) (there is no CPU path)")
        if dtype is not None and dtype != torch.bfloat16:
            raise L.QieError(f"the kernels store weights and activations in bf16; .to({dtype}) is not supported")
        return self

    def cache_context(self, name: str):      # CacheMixin.cache_context("cond"/"uncond") — no-op here
        return contextlib.nullcontext()

    def __del__(self):
        try:
            if self._handle:
                L.lib().qie_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def _alloc(self, name: str, shape, dtype) -> torch.Tensor:
        t = torch.empty(shape, dtype=dtype, device=self._device)
        self._t[name] = t
        return t

    def _allocate_packed(self):
        c, D = self.cfg, self.cfg.inner_dim
        Lr = c.num_layers
        bf, f32 = torch.bfloat16, torch.float32
        a = self._alloc
        a("img_in_w", (D, c.in_channels), bf); a("img_in_b", (D,), f32)
        a("txt_norm_w", (c.joint_attention_dim,), f32)
        a("txt_in_w", (D, c.joint_attention_dim), bf); a("txt_in_b", (D,), f32)
        a("t1_w", (D, 256), bf); a("t1_b", (D,), f32)
        a("t2_w", (D, D), bf); a("t2_b", (D,), f32)
        a("mod_w", (Lr, 2, 6 * D, D), bf); a("mod_b", (Lr, 2, 6 * D), f32)
        a("norm_out_w", (2 * D, D), bf); a("norm_out_b", (2 * D,), f32)
        a("proj_out_w", (c.out_dim, D), bf); a("proj_out_b", (c.out_dim,), f32)
        a("qkv_w", (Lr, 2, 3 * D, D), bf); a("qkv_b", (Lr, 2, 3 * D), f32)
        a("qn_w", (Lr, 2, 128), f32); a("kn_w", (Lr, 2, 128), f32)
        a("out_w", (Lr, 2, D, D), bf); a("out_b", (Lr, 2, D), f32)
        a("ff1_w", (Lr, 2, 4 * D, D), bf); a("ff1_b", (Lr, 2, 4 * D), f32)
        a("ff2_w", (Lr, 2, D, 4 * D), bf); a("ff2_b", (Lr, 2, D), f32)

    def _register(self):
        t, c = self._t, self.cfg
        Lr = c.num_layers
        blocks = (L.BlockWeights * Lr)()
        names = [("qkv_w", "qkv_w"), ("qkv_b", "qkv_b"), ("q_norm_w", "qn_w"), ("k_norm_w", "kn_w"),
                 ("out_w", "out_w"), ("out_b", "out_b"), ("ff1_w", "ff1_w"), ("ff1_b", "ff1_b"),
                 ("ff2_w", "ff2_w"), ("ff2_b", "ff2_b")]
        fp8 = [("qkv_w8", "qkv_w8"), ("qkv_ws", "qkv_ws"), ("out_w8", "out_w8"), ("out_ws", "out_ws"),
               ("ff1_w8", "ff1_w8"), ("ff1_ws", "ff1_ws"), ("ff2_w8", "ff2_w8"), ("ff2_ws", "ff2_ws")]
        for l in range(Lr):
            for field, key in names + (fp8 if self._fp8_ready else []):
                arr = getattr(blocks[l], field)
                for s in range(2):
                    arr[s] = t[key][l, s].data_ptr()
        w = L.Weights()
        for k in ("img_in_w", "img_in_b", "txt_norm_w", "txt_in_w", "txt_in_b", "t1_w", "t1_b", "t2_w", "t2_b",
                  "mod_w", "mod_b", "norm_out_w", "norm_out_b", "proj_out_w", "proj_out_b"):
            setattr(w, k, t[k].data_ptr())
        w.blocks = C.cast(blocks, C.POINTER(L.BlockWeights))
        self._blocks_arr = blocks
        L.check(L.lib().qie_set_weights(self._handle, C.byref(w)), "qie_set_weights")

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], config: QwenImageDiTConfig = QwenImageDiTConfig(),
                        device="cuda:0") -> "B200QwenImageTransformer2DModel":
        """Pack a diffusers-named state_dict into the library layout (bf16 weights)."""
        m = cls(config, device)
        m._allocate_packed()
        t = m._t

        def put(dst: torch.Tensor, src: torch.Tensor):
            dst.copy_(src.to(device=dst.device, dtype=dst.dtype, non_blocking=True))

        for dst, key in (("img_in", "img_in"), ("txt_in", "txt_in"), ("proj_out", "proj_out"),
                         ("t1", "time_text_embed.timestep_embedder.linear_1"),
                         ("t2", "time_text_embed.timestep_embedder.linear_2"), ("norm_out", "norm_out.linear")):
            put(t[dst + "_w"], sd[key + ".weight"]); put(t[dst + "_b"], sd[key + ".bias"])
        put(t["txt_norm_w"], sd["txt_norm.weight"])
        D = config.inner_dim
        for l in range(config.num_layers):
            k = _block_keys(l)
            for s in range(2):
                for j, name in enumerate(("q", "k", "v")):
                    put(t["qkv_w"][l, s, j * D:(j + 1) * D], sd[k[name][s] + ".weight"])
                    put(t["qkv_b"][l, s, j * D:(j + 1) * D], sd[k[name][s] + ".bias"])
                put(t["qn_w"][l, s], sd[k["nq"][s] + ".weight"]); put(t["kn_w"][l, s], sd[k["nk"][s] + ".weight"])
                for dst, name in (("out", "out"), ("ff1", "ff1"), ("ff2", "ff2"), ("mod", "mod")):
                    put(t[dst + "_w"][l, s], sd[k[name][s] + ".weight"]); put(t[dst + "_b"][l, s], sd[k[name][s] + ".bias"])
        m._register()
        return m

    @classmethod
    def from_safetensors(cls, paths, config: QwenImageDiTConfig = QwenImageDiTConfig(), device="cuda:0", lora: Optional[str] = None,
                         lora_scale: float = 1.0) -> "B200QwenImageTransformer2DModel":
        """Load a diffusers `transformer/` checkpoint (one or several .safetensors shards, diffusers key names) and optionally
        merge a LoRA file offline (next-row N2: the reference loads the Lightning LoRA un-merged via PEFT, server.py:76-79)."""
        from safetensors import safe_open
        sd = {}
        for p in ([paths] if isinstance(paths, (str, bytes)) or hasattr(paths, "__fspath__") else list(paths)):
            with safe_open(str(p), framework="pt", device="cpu") as f:
                for k in f.keys():
                    sd[k] = f.get_tensor(k)
        if lora is not None:
            with safe_open(str(lora), framework="pt", device="cpu") as f:
                sd = merge_lora(sd, {k: f.get_tensor(k) for k in f.keys()}, lora_scale)
        return cls.from_state_dict(sd, config, device)

    @classmethod
    def from_random(cls, config: QwenImageDiTConfig = QwenImageDiTConfig(), seed: int = 0, device="cuda:0",
                    std: float = 0.02) -> "B200QwenImageTransformer2DModel":
        """Random-init weights of the architecture generated directly in HBM (no checkpoint, no network).
        Recipe of SURVEY §8c: matrices N(0, 1/fan_in), modulation matrices / biases N(0, std^2), norm weights 1+N(0,std^2)."""
        m = cls(config, device)
        m._allocate_packed()
        g = torch.Generator(device=m._device).manual_seed(seed)
        for name, t in m._t.items():
            if name in ("txt_norm_w", "qn_w", "kn_w"):
                t.copy_(1.0 + std * torch.randn(t.shape, generator=g, device=t.device, dtype=torch.float32))
            elif name.endswith("_b"):
                t.copy_(std * torch.randn(t.shape, generator=g, device=t.device, dtype=torch.float32))
            else:
                scale = std if name in ("mod_w", "norm_out_w") else 1.0 / math.sqrt(t.shape[-1])
                flat = t.view(-1, t.shape[-1])
                step = max(1, (1 << 28) // t.shape[-1])       # bounded temporaries for the 10+ GB tensors
                for r0 in range(0, flat.shape[0], step):
                    blk = flat[r0:r0 + step]
                    blk.copy_(scale * torch.randn(blk.shape, generator=g, device=t.device, dtype=torch.float32))
        m._register()
        return m

    def export_state_dict(self) -> Dict[str, torch.Tensor]:
        """diffusers-named views of the packed weights (what `state_dict()` of the reference module holds)."""
        t, c, D = self._t, self.cfg, self.cfg.inner_dim
        sd = {}
        for src, key in (("img_in", "img_in"), ("txt_in", "txt_in"), ("proj_out", "proj_out"),
                         ("t1", "time_text_embed.timestep_embedder.linear_1"),
                         ("t2", "time_text_embed.timestep_embedder.linear_2"), ("norm_out", "norm_out.linear")):
            sd[key + ".weight"], sd[key + ".bias"] = t[src + "_w"], t[src + "_b"]
        sd["txt_norm.weight"] = t["txt_norm_w"]
        for l in range(c.num_layers):
            k = _block_keys(l)
            for s in range(2):
                for j, name in enumerate(("q", "k", "v")):
                    sd[k[name][s] + ".weight"] = t["qkv_w"][l, s, j * D:(j + 1) * D]
                    sd[k[name][s] + ".bias"] = t["qkv_b"][l, s, j * D:(j + 1) * D]
                sd[k["nq"][s] + ".weight"], sd[k["nk"][s] + ".weight"] = t["qn_w"][l, s], t["kn_w"][l, s]
                for src, name in (("out", "out"), ("ff1", "ff1"), ("ff2", "ff2"), ("mod", "mod")):
                    sd[k[name][s] + ".weight"], sd[k[name][s] + ".bias"] = t[src + "_w"][l, s], t[src + "_b"][l, s]
        return sd

    # ------------------------------------------------------------------ FP8 W8A8 path
    def quantize_fp8(self):
        """Offline weight quantisation (replaces quantize_transformer.py / Int8Linear, README.md:136-138):
        per-output-channel symmetric e4m3 of the four big per-block linears; activations are quantised
        per token on the fly inside the adaLN / quant kernels."""
        for name in ("qkv", "out", "ff1", "ff2"):
            w = self._t[name + "_w"]
            w8 = self._alloc(name + "_w8", w.shape, torch.float8_e4m3fn)
            ws = self._alloc(name + "_ws", w.shape[:-1], torch.float32)
            for l in range(w.shape[0]):
                wf = w[l].float()
                s = wf.abs().amax(dim=-1).clamp_min(1e-12) / 448.0
                ws[l].copy_(s)
                w8[l].copy_((wf / s.unsqueeze(-1)).to(torch.float8_e4m3fn))
        self._fp8_ready = True
        self._q8_kind = "fp8"
        self._register()
        return self

    def quantize_int8(self):
        """Offline per-output-channel symmetric int8 weights (s_w = max|W_n|/127, round-to-nearest-even) — the W side of the
        README's Int8Linear / triton_int8_gemm ("quantize + matmul + dequantize", README.md:136-141)."""
        for name in ("qkv", "out", "ff1", "ff2"):
            w = self._t[name + "_w"]
            w8 = self._alloc(name + "_w8", w.shape, torch.int8)
            ws = self._alloc(name + "_ws", w.shape[:-1], torch.float32)
            for l in range(w.shape[0]):
                wf = w[l].float()
                s = wf.abs().amax(dim=-1).clamp_min(1e-12) / 127.0
                ws[l].copy_(s)
                w8[l].copy_(torch.round(wf / s.unsqueeze(-1)).clamp_(-127, 127).to(torch.int8))
        self._fp8_ready = True
        self._q8_kind = "int8"
        self._register()
        return self

    def set_precision(self, mode: str):
        if mode in ("fp8", "int8") and getattr(self, "_q8_kind", None) != mode:
            self.quantize_fp8() if mode == "fp8" else self.quantize_int8()
        L.check(L.lib().qie_set_precision(self._handle, {"bf16": 0, "fp8": 1, "int8": 2}[mode]), "qie_set_precision")
        return self

    def set_option(self, key: int, value: int):
        L.check(L.lib().qie_set_option(self._handle, key, value), "qie_set_option")
        return self

    # ------------------------------------------------------------------ exact caches
    def cache_schedule(self, timesteps: Sequence[float]):
        """Precompute temb, all modulation vectors and the final scale/shift for a fixed list of timestep values (the
        values `forward` will receive, i.e. pipeline.model_timestep(sigma)).  Replaces the fixed-schedule part of the
        reference's cached_pipeline_v2.py (README.md:125): the 13.6 GB modulation GEMV leaves the per-frame loop."""
        vals = [float(t) for t in timesteps]
        arr = (C.c_float * len(vals))(*vals)
        with torch.cuda.device(self._device):
            L.check(L.lib().qie_cache_schedule(self._handle, arr, len(vals), L.cur_stream()), "qie_cache_schedule")
        self._sched_vals = [C.c_float(v).value for v in vals]
        return self

    def cache_prompt(self, name: str, prompt_embeds: torch.Tensor):
        """Precompute txt_in(txt_norm(prompt_embeds)) once per prompt (reference: CachedConditions /
        precompute_conditions, qwen_realtime.py:69-89,140-165 — a stub there).  `name` e.g. "cond" / "uncond"."""
        names = self.__dict__.setdefault("_prompt_slots", {})
        if name not in names:
            if len(names) >= 4:
                raise L.QieError("at most 4 cached prompts")
            names[name] = len(names)
        e = prompt_embeds.to(device=self._device, dtype=torch.bfloat16)
        if e.dim() == 3:
            if e.shape[0] != 1:
                raise L.QieError("cache_prompt takes one prompt ([T, C] or [1, T, C])")
            e = e[0]
        e = e.contiguous()
        with torch.cuda.device(self._device):
            L.check(L.lib().qie_cache_prompt(self._handle, names[name], L.ptr(e), e.shape[0], L.cur_stream()), "qie_cache_prompt")
        self.__dict__.setdefault("_prompt_rows", {})[name] = e.shape[0]
        return self

    def _select_caches(self, B: int, T: int, timestep_values, cached_prompt):
        idx = None
        vals = getattr(self, "_sched_vals", None)
        if timestep_values is not None and vals:
            tv = [C.c_float(float(v)).value for v in timestep_values]
            tv = tv * B if len(tv) == 1 else tv
            if len(tv) == B and all(v in vals for v in tv):
                idx = (C.c_int * B)(*[vals.index(v) for v in tv])
        slot = -1
        if cached_prompt is not None:
            slots = getattr(self, "_prompt_slots", {})
            if cached_prompt not in slots or self._prompt_rows[cached_prompt] != T:
                raise L.QieError(f"no cached prompt {cached_prompt!r} with {T} tokens")
            slot = slots[cached_prompt]
        L.check(L.lib().qie_cache_select(self._handle, idx, B if idx is not None else 0, slot), "qie_cache_select")

    # ------------------------------------------------------------------ measurement aids
    PROFILE_CLASSES = ("gemm", "attention", "adaln", "mod_gemv", "other", "barrier")
    _profiling = False

    def profile(self, on: bool):
        """record CUDA events around every kernel class inside qie_forward (read with read_profile)."""
        self._profiling = bool(on)
        return self.set_option(2, 1 if on else 0)

    def read_profile(self) -> Dict[str, Dict[str, float]]:
        n_cls = L.PROFILE_CLASSES
        ms, work, n = (C.c_double * n_cls)(), (C.c_double * n_cls)(), (C.c_int * n_cls)()
        L.check(L.lib().qie_profile_read(self._handle, ms, work, n), "qie_profile_read")
        return {k: {"ms": ms[i], "work": work[i], "launches": n[i]} for i, k in enumerate(self.PROFILE_CLASSES)}

    def read_timeline(self, max_n: int = 4096):
        """[(start_ms, duration_ms, class name)] of every launch recorded since the last read_profile (call it BEFORE
        read_profile, which consumes the events): where the time of one forward goes, gaps between the kernels included."""
        st, du, cl = (C.c_float * max_n)(), (C.c_float * max_n)(), (C.c_int * max_n)()
        n = L.lib().qie_profile_timeline(self._handle, st, du, cl, max_n)
        if n < 0:
            L.check(n, "qie_profile_timeline")
        return [(st[i], du[i], self.PROFILE_CLASSES[cl[i]]) for i in range(n)]

    # ------------------------------------------------------------------ forward
    def _workspace(self, seq: L.Seq) -> torch.Tensor:
        need = L.lib().qie_workspace_bytes(self._handle, C.byref(seq))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.zeros(need + 1024, dtype=torch.uint8, device=self._device)
        return self._ws

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor = None,
                encoder_hidden_states_mask: torch.Tensor = None, timestep: torch.Tensor = None,
                img_shapes: Optional[List] = None, txt_seq_lens: Optional[List[int]] = None,
                guidance: torch.Tensor = None, attention_kwargs: Optional[dict] = None,
                controlnet_block_samples=None, return_dict: bool = True, num_blocks: int = -1,
                timestep_values: Optional[Sequence[float]] = None, cached_prompt: Optional[str] = None, **_ignored):
        if controlnet_block_samples is not None:
            raise L.QieError("controlnet residuals are not on the reference's path and are not supported")
        if hidden_states.device != self._device:
            raise L.QieError(f"hidden_states on {hidden_states.device}, model on {self._device}")
        B, S_i, Cin = hidden_states.shape
        T = encoder_hidden_states.shape[1]
        if txt_seq_lens is not None and max(txt_seq_lens) != T:
            # the 0.36.0 processor attends over all T rows (no mask); only the RoPE length comes from txt_seq_lens
            if max(txt_seq_lens) > T:
                raise L.QieError("txt_seq_lens exceeds encoder_hidden_states length")
            encoder_hidden_states = encoder_hidden_states[:, :max(txt_seq_lens)]
            T = encoder_hidden_states.shape[1]
        # one length per batch element (batched true-CFG: cond and uncond prompt in one forward; replaces the reference's
        # batched_cfg_pipeline.py, README.md:126): rows behind an element's own length are padding for every kernel — masked keys in
        # attention — so the result equals the separate forwards (upstream would let the pad tokens take part)
        ragged = txt_seq_lens is not None and len(txt_seq_lens) == B and len(set(int(v) for v in txt_seq_lens)) > 1
        shapes = img_shapes[0] if (isinstance(img_shapes, (list, tuple)) and isinstance(img_shapes[0], (list, tuple))
                                   and isinstance(img_shapes[0][0], (list, tuple))) else img_shapes
        flat = [int(v) for fhw in shapes for v in fhw]
        seq = L.make_seq_ragged(S_i, txt_seq_lens) if ragged else L.make_seq(B, S_i, T)
        ws = self._workspace(seq)
        base = (ws.data_ptr() + 1023) // 1024 * 1024
        hs = hidden_states.to(torch.bfloat16).contiguous()
        enc = encoder_hidden_states.to(device=self._device, dtype=torch.bfloat16).contiguous()
        ts = timestep.to(device=self._device, dtype=torch.float32).reshape(-1).expand(B).contiguous()
        out = torch.empty(B, S_i, self.cfg.out_dim, dtype=torch.bfloat16, device=self._device)
        shp = (C.c_int * len(flat))(*flat)
        self._select_caches(B, T, timestep_values, cached_prompt)
        with torch.cuda.device(self._device):
            L.check(L.lib().qie_forward(self._handle, L.ptr(hs), L.ptr(enc), L.ptr(ts), shp, len(flat) // 3,
                                        C.byref(seq), L.ptr(out), C.c_void_p(base), ws.numel() - (base - ws.data_ptr()),
                                        num_blocks, L.cur_stream()), "qie_forward")
        out = out.to(hidden_states.dtype) if hidden_states.dtype != torch.bfloat16 else out
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)
