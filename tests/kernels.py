"""Thin test-side wrappers that call the per-kernel C-ABI entry points of libqie.so on torch tensors."""
import ctypes as C

import torch

import qie_b200
from qie_b200 import _lib as L


def seq(batch, img_rows, txt_rows):
    return L.make_seq(batch, img_rows, txt_rows)


def rows(s):
    return s.batch * (s.img_pad + s.txt_pad)


def to_joint(s, img, txt):
    """[B,n_img,C], [B,n_txt,C] -> padded joint layout [B*rpb, C] (pad rows zero)."""
    B, C_ = img.shape[0], img.shape[-1]
    out = torch.zeros(B, s.img_pad + s.txt_pad, C_, dtype=img.dtype, device=img.device)
    out[:, :s.img_rows] = img
    out[:, s.img_pad:s.img_pad + s.txt_rows] = txt
    return out.reshape(-1, C_).contiguous()


def from_joint(s, x):
    x = x.reshape(s.batch, s.img_pad + s.txt_pad, -1)
    return x[:, :s.img_rows], x[:, s.img_pad:s.img_pad + s.txt_rows]


def gemm(s, a, w, bias, out, epilogue, streams=3, a_compact=0, out_compact=0, gate=None, gate_bstride=0,
         gate_sstride=0, rope=None, qk_norm_w=None, block_n=0, fp8=False, a_scale=None, w_scale=None, cta_group=0):
    g = L.GemmArgs()
    g.a = a.data_ptr(); g.a_compact = a_compact
    for i in range(2):
        g.w[i] = w[i].data_ptr() if w[i] is not None else None
        g.bias[i] = bias[i].data_ptr() if bias is not None and bias[i] is not None else None
        if w_scale is not None and w_scale[i] is not None:
            g.w_scale[i] = w_scale[i].data_ptr()
    g.out = out.data_ptr(); g.out_compact = out_compact; g.ldo = out.shape[-1]
    ws = [x for x in w if x is not None][0]
    g.N, g.K = ws.shape[0], ws.shape[1]
    g.streams = streams; g.epilogue = epilogue
    if gate is not None:
        g.gate = gate.data_ptr(); g.gate_bstride = gate_bstride; g.gate_sstride = gate_sstride
    if rope is not None:
        g.rope = rope.data_ptr()
    if qk_norm_w is not None:
        for st in range(2):
            for k in range(2):
                g.qk_norm_w[st][k] = qk_norm_w[st][k].data_ptr()
    g.fp8 = int(fp8)
    if a_scale is not None:
        g.a_scale = a_scale.data_ptr()
    g.block_n = block_n
    g.cta_group = cta_group
    L.check(L.lib().qie_gemm(C.byref(g), C.byref(s), L.cur_stream()), "qie_gemm")
    return out


def attn(s, qkv, heads, variant=0):
    out = torch.empty(qkv.shape[0], heads * 128, dtype=torch.bfloat16, device=qkv.device)
    L.check(L.lib().qie_attn_fwd(L.ptr(qkv), L.ptr(out), C.byref(s), heads, variant, L.cur_stream()), "qie_attn_fwd")
    return out


def ln_modulate(s, x, mod, bstride, sstride, shift_off, scale_off, D, fp8=False, qmode=1, want_bf16=True):
    out = torch.empty(x.shape[0], D, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    out8 = torch.empty(x.shape[0], D, dtype=torch.uint8, device=x.device) if fp8 else None
    sc = torch.empty(x.shape[0], dtype=torch.float32, device=x.device) if fp8 else None
    L.check(L.lib().qie_ln_modulate(L.ptr(x), L.ptr(mod), bstride, sstride, shift_off, scale_off, L.ptr(out),
                                    L.ptr(out8), L.ptr(sc), qmode, D, 1e-6, C.byref(s), L.cur_stream()), "qie_ln_modulate")
    return (out, out8, sc) if fp8 else out


def gemv(x, w, bias, act):
    B, K = x.shape
    N = w.shape[0]
    y = torch.empty(B, N, dtype=torch.float32, device=x.device)
    L.check(L.lib().qie_gemv(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(y), B, N, K, act, L.cur_stream()), "qie_gemv")
    return y


def timestep_proj(t, round_bf16=0):
    out = torch.empty(t.shape[0], 256, dtype=torch.float32, device=t.device)
    L.check(L.lib().qie_timestep_proj(L.ptr(t), L.ptr(out), t.shape[0], round_bf16, L.cur_stream()), "qie_timestep_proj")
    return out


def qk_norm_rope(s, qkv, rope, norm_w, heads):
    arr = (C.c_void_p * 4)(*[w.data_ptr() for w in norm_w])
    L.check(L.lib().qie_qk_norm_rope(L.ptr(qkv), L.ptr(rope), arr, heads, 1e-6, C.byref(s), L.cur_stream()),
            "qie_qk_norm_rope")
    return qkv


def rmsnorm_pack(x, w, n_pad):
    B, n, D = x.shape
    out = torch.empty(B, n_pad, D, dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().qie_rmsnorm_pack(L.ptr(x), L.ptr(w), L.ptr(out), B, n, n_pad, D, 1e-6, L.cur_stream()),
            "qie_rmsnorm_pack")
    return out


def quant_rows(x, qmode=1):
    q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    sc = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    L.check(L.lib().qie_quant_rows(L.ptr(x), L.ptr(q), L.ptr(sc), x.shape[0], x.shape[1], qmode, L.cur_stream()),
            "qie_quant_rows")
    return q, sc


def rope_table(cfg_c, img_shapes, s):
    flat = [int(v) for fhw in img_shapes for v in fhw]
    n = (s.img_pad + s.txt_pad) * 128
    buf = (C.c_float * n)()
    L.check(L.lib().qie_rope_table_host(C.byref(cfg_c), (C.c_int * len(flat))(*flat), len(flat) // 3, C.byref(s), buf),
            "qie_rope_table_host")
    return torch.frombuffer(buf, dtype=torch.float32).reshape(s.img_pad + s.txt_pad, 64, 2).clone()


def rel_err(a, b):
    """max |a-b| / max |b|  — the tolerance metric of BASELINE.json (per-step velocity max-rel-err)."""
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
