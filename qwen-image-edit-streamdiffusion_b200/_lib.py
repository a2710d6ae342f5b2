"""ctypes binding of libqie.so (the C ABI declared in include/qie.h).

The product path has NO fallback: if the shared library is missing, or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("QIE_LIB", _HERE / "libqie.so"))


class QieError(RuntimeError):
    pass


class ModelCfg(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("num_heads", C.c_int), ("head_dim", C.c_int),
                ("in_channels", C.c_int), ("out_dim", C.c_int), ("joint_dim", C.c_int),
                ("rope_axes", C.c_int * 3)]


class Seq(C.Structure):
    _fields_ = [("batch", C.c_int), ("img_rows", C.c_int), ("txt_rows", C.c_int),
                ("img_pad", C.c_int), ("txt_pad", C.c_int), ("txt_rows_b", C.c_int * 8)]

    def __init__(self, batch=0, img_rows=0, txt_rows=0, img_pad=0, txt_pad=0, txt_rows_b=None):
        """txt_rows_b: valid text rows of every batch element (default: txt_rows for all, as qie_make_seq fills them)"""
        super().__init__(batch, img_rows, txt_rows, img_pad, txt_pad)
        for b in range(min(batch, 8)):
            self.txt_rows_b[b] = txt_rows if txt_rows_b is None else int(txt_rows_b[b])

    @property
    def rows_per_batch(self) -> int:
        return self.img_pad + self.txt_pad


class Sp(C.Structure):
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("img_total", C.c_int), ("txt_total", C.c_int),
                ("img_offset", C.c_int), ("txt_offset", C.c_int)]


_P2 = C.c_void_p * 2


class BlockWeights(C.Structure):
    _fields_ = [("qkv_w", _P2), ("qkv_b", _P2), ("q_norm_w", _P2), ("k_norm_w", _P2),
                ("out_w", _P2), ("out_b", _P2), ("ff1_w", _P2), ("ff1_b", _P2),
                ("ff2_w", _P2), ("ff2_b", _P2),
                ("qkv_w8", _P2), ("qkv_ws", _P2), ("out_w8", _P2), ("out_ws", _P2),
                ("ff1_w8", _P2), ("ff1_ws", _P2), ("ff2_w8", _P2), ("ff2_ws", _P2)]


class Weights(C.Structure):
    _fields_ = [("img_in_w", C.c_void_p), ("img_in_b", C.c_void_p), ("txt_norm_w", C.c_void_p),
                ("txt_in_w", C.c_void_p), ("txt_in_b", C.c_void_p),
                ("t1_w", C.c_void_p), ("t1_b", C.c_void_p), ("t2_w", C.c_void_p), ("t2_b", C.c_void_p),
                ("mod_w", C.c_void_p), ("mod_b", C.c_void_p),
                ("norm_out_w", C.c_void_p), ("norm_out_b", C.c_void_p),
                ("proj_out_w", C.c_void_p), ("proj_out_b", C.c_void_p),
                ("blocks", C.POINTER(BlockWeights))]


class GemmArgs(C.Structure):
    _fields_ = [("a", C.c_void_p), ("a_compact", C.c_int), ("w", _P2), ("bias", _P2),
                ("out", C.c_void_p), ("out_compact", C.c_int), ("ldo", C.c_int),
                ("N", C.c_int), ("K", C.c_int), ("streams", C.c_int), ("epilogue", C.c_int),
                ("gate", C.c_void_p), ("gate_bstride", C.c_longlong), ("gate_sstride", C.c_longlong),
                ("rope", C.c_void_p), ("qk_norm_w", (C.c_void_p * 2) * 2),
                ("fp8", C.c_int), ("a_scale", C.c_void_p), ("w_scale", _P2), ("block_n", C.c_int), ("cta_group", C.c_int),
                ("peer_out", C.c_void_p), ("sp_rank", C.c_int), ("sp_size", C.c_int), ("sp_gathered_rows", C.c_int),
                ("sp_txt_row0", C.c_int), ("q8_amax", C.c_void_p), ("q_scale", C.c_float)]


class Peers(C.Structure):
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("batch", C.c_int), ("img_pad", C.c_int), ("txt_pad", C.c_int),
                ("img_total", C.c_int), ("txt_total", C.c_int), ("qkv_gather", C.c_void_p * 8), ("attn_out", C.c_void_p * 8),
                ("vel", C.c_void_p * 8), ("flags", C.c_void_p * 8), ("mod", C.c_void_p * 8)]


PROFILE_CLASSES = 6


EPI_BF16, EPI_GELU_BF16, EPI_F32, EPI_GATE_RESID_F32, EPI_QKV_NORM_ROPE = range(5)

# every symbol include/qie.h declares: name -> (restype, argtypes)
_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
SYMBOLS = {
    "qie_version": (_i, []),
    "qie_last_error": (C.c_char_p, []),
    "qie_device_sm_count": (_i, []),
    "qie_create": (_i, [C.POINTER(ModelCfg), _i, C.POINTER(_vp)]),
    "qie_destroy": (_i, [_vp]),
    "qie_set_weights": (_i, [_vp, C.POINTER(Weights)]),
    "qie_set_precision": (_i, [_vp, _i]),
    "qie_set_option": (_i, [_vp, _i, _i]),
    "qie_attn_score_bound": (_f, [_vp, _i]),
    "qie_attn_layer_variant": (_i, [_vp, _i]),
    "qie_launch_count": (C.c_ulonglong, []),
    "qie_tune": (_i, [_i, _i]),
    "qie_tune_get": (_i, [_i]),
    "qie_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i)]),
    "qie_profile_timeline": (_i, [_vp, C.POINTER(_f), C.POINTER(_f), C.POINTER(_i), _i]),
    "qie_make_seq": (_i, [_i, _i, _i, C.POINTER(Seq)]),
    "qie_make_seq_ragged": (_i, [_i, _i, C.POINTER(_i), C.POINTER(Seq)]),
    "qie_workspace_bytes": (C.c_size_t, [_vp, C.POINTER(Seq)]),
    "qie_forward": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_i), _i, C.POINTER(Seq), _vp, _vp, C.c_size_t, _i, _vp]),
    "qie_forward_phase": (_i, [_vp, _i, _i, _vp, _vp, _vp, C.POINTER(_i), _i, C.POINTER(Seq), C.POINTER(Sp), _vp, _vp,
                               C.c_size_t, _i, _vp]),
    "qie_workspace_offset": (_ll, [_vp, C.POINTER(Seq), _i]),
    "qie_attn_fwd_tiles": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    "qie_set_peers": (_i, [_vp, C.POINTER(Peers), _vp]),
    "qie_peer_alloc": (_i, [C.c_size_t, C.POINTER(_vp), C.c_char_p]),
    "qie_peer_free": (_i, [_vp]),
    "qie_peer_copy": (_i, [_vp, _vp, C.c_size_t, _vp]),
    "qie_peer_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "qie_peer_close": (_i, [_vp]),
    "qie_peer_barrier": (_i, [_vp, _vp]),
    "qie_peer_barrier_timeouts": (_i, []),
    "qie_sp_shard": (_i, [_i, _i, _i, _i, _i, C.POINTER(Seq), C.POINTER(Sp)]),
    "qie_sp_tile_valid_host": (_i, [_i, _i, _i, C.POINTER(_i), _i]),
    "qie_forward_sp": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_i), _i, C.POINTER(Seq), C.POINTER(Sp), _vp, _vp, C.c_size_t, _vp]),
    "qie_cache_schedule": (_i, [_vp, C.POINTER(_f), _i, _vp]),
    "qie_cache_prompt": (_i, [_vp, _i, _vp, _i, _vp]),
    "qie_cache_select": (_i, [_vp, C.POINTER(_i), _i, _i]),
    "qie_cfg_euler_step": (_i, [_vp, _vp, _vp, _f, _f, _f, _i, _i, _i, _i, _vp]),
    "qie_pack_latents": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "qie_unpack_latents": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "qie_flowmatch_sigmas": (_i, [_i, _i, C.POINTER(_f)]),
    "qie_rope_table_host": (_i, [C.POINTER(ModelCfg), C.POINTER(_i), _i, C.POINTER(Seq), C.POINTER(_f)]),
    "qie_gemm": (_i, [C.POINTER(GemmArgs), C.POINTER(Seq), _vp]),
    "qie_attn_fwd": (_i, [_vp, _vp, C.POINTER(Seq), _i, _i, _vp]),
    "qie_ln_modulate": (_i, [_vp, _vp, _ll, _ll, _i, _i, _vp, _vp, _vp, _i, _i, _f, C.POINTER(Seq), _vp]),
    "qie_gemv": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _i, _i, _vp]),
    "qie_timestep_proj": (_i, [_vp, _vp, _i, _i, _vp]),
    "qie_qk_norm_rope": (_i, [_vp, _vp, C.POINTER(_vp), _i, _f, C.POINTER(Seq), _vp]),
    "qie_rmsnorm_pack": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "qie_pack_rows": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "qie_quant_rows": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp]),
}

_lib = None


def build_library(verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into libqie.so (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", str(_HERE / "csrc"), "-j4"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise QieError("building libqie.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load libqie.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise QieError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU / PyTorch fallback for the hot path)")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)   # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        if l.qie_version() != 1:
            raise QieError("libqie ABI version mismatch")
        if os.environ.get("QIE_PDL"):       # A/B switch for whole test / bench runs: programmatic dependent launch (qie_tune key 7)
            l.qie_tune(7, int(os.environ["QIE_PDL"]))
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().qie_last_error().decode(errors="replace")
        raise QieError(f"{what or 'libqie'} failed with status {rc}: {msg}")


def make_seq(batch: int, img_rows: int, txt_rows: int) -> Seq:
    s = Seq()
    check(lib().qie_make_seq(batch, img_rows, txt_rows, C.byref(s)), "qie_make_seq")
    return s


def make_seq_ragged(img_rows: int, txt_rows_b) -> Seq:
    """one text length per batch element (the cond and the uncond prompt of a true-CFG step in ONE forward)"""
    s = Seq()
    arr = (C.c_int * len(txt_rows_b))(*[int(v) for v in txt_rows_b])
    check(lib().qie_make_seq_ragged(len(txt_rows_b), img_rows, arr, C.byref(s)), "qie_make_seq_ragged")
    return s


def ptr(t) -> C.c_void_p:
    """device/host pointer of a torch tensor (None -> NULL)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def cur_stream() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
