"""ctypes binding of libub200.so — the only way Python reaches the CUDA kernels.

There is deliberately no fallback: if the library is missing or the device is not sm_100 the
import of the compute path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libub200.so")

F16, BF16 = 0, 1
EPI_BIAS, EPI_DROPOUT, EPI_RESIDUAL, EPI_GELU = 1, 2, 4, 8
EPI_DGELU, EPI_ACCUM, EPI_OUT_F32, EPI_COLSUM = 16, 32, 64, 128
EPI_ATOMIC = 256
EPI_TANH = 512
EPI_LN = 1024


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64),
        ("a_major", C.c_int32), ("b_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("dtype", C.c_int32), ("epilogue", C.c_int32),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("aux", C.c_void_p),
        ("out", C.c_void_p), ("out2", C.c_void_p), ("colsum", C.c_void_p),
        ("ldr", C.c_int64), ("ldaux", C.c_int64), ("ldo", C.c_int64),
        ("dropout_p", C.c_float),
        ("rng_seed", C.c_uint64), ("rng_stream", C.c_uint64),
        ("tile_n", C.c_int32), ("max_ctas", C.c_int32), ("cluster", C.c_int32),
        ("k_splits", C.c_int32), ("n_valid", C.c_int32),
        ("rng_offset_dev", C.c_void_p),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_out", C.c_void_p), ("ldln", C.c_int64),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("ctx", C.c_void_p), ("lse", C.c_void_p), ("cu_seqlens", C.c_void_p),
        ("batch", C.c_int32), ("total_tokens", C.c_int32), ("max_seqlen", C.c_int32),
        ("hidden", C.c_int32), ("num_heads", C.c_int32), ("dtype", C.c_int32),
        ("dropout_p", C.c_float), ("rng_seed", C.c_uint64), ("rng_stream", C.c_uint64),
        ("dctx", C.c_void_p), ("dqkv", C.c_void_p), ("workspace", C.c_void_p), ("dbias", C.c_void_p),
        ("rng_offset_dev", C.c_void_p),
    ]


class LnBwdArgs(C.Structure):
    _fields_ = [
        ("dy", C.c_void_p), ("x", C.c_void_p), ("gamma", C.c_void_p), ("dx", C.c_void_p),
        ("dx_drop", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dbias", C.c_void_p),
        ("rows", C.c_int32), ("hidden", C.c_int32), ("dtype", C.c_int32),
        ("dropout_p", C.c_float), ("rng_seed", C.c_uint64), ("rng_stream", C.c_uint64),
        ("row_kind", C.c_void_p), ("kind", C.c_int32), ("dropout_on_dy", C.c_int32),
        ("rng_offset_dev", C.c_void_p), ("stats_ws", C.c_void_p),
    ]


class EmbedPrepArgs(C.Structure):
    _fields_ = [
        ("pack_idx", C.c_void_p), ("gather_index", C.c_void_p), ("input_ids", C.c_void_p),
        ("position_ids", C.c_void_p), ("txt_type_ids", C.c_void_p), ("img_type_ids", C.c_void_p),
        ("img_masks", C.c_void_p),
        ("T", C.c_int32), ("L", C.c_int32), ("Lt", C.c_int32), ("Li", C.c_int32),
        ("pos_rows", C.c_int32), ("mode", C.c_int32),
        ("kind", C.c_void_p), ("word_id", C.c_void_p), ("pos_id", C.c_void_p),
        ("type_id", C.c_void_p), ("img_src", C.c_void_p), ("mask_flag", C.c_void_p),
    ]


class EmbedRowsArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "kind", "word_id", "pos_id", "type_id", "img_src", "word_emb", "pos_emb", "type_emb",
        "ln_txt_g", "ln_txt_b", "img_linear_out", "pos_feat", "w_pos", "b_pos",
        "ln_img_g", "ln_img_b", "ln_pos_g", "ln_pos_b", "ln_out_g", "ln_out_b", "x", "u", "ppre")] + [
        ("T", C.c_int32), ("hidden", C.c_int32), ("dtype", C.c_int32),
        ("dropout_p", C.c_float), ("rng_seed", C.c_uint64), ("rng_stream", C.c_uint64),
        ("rng_offset_dev", C.c_void_p)]


class EmbedColsumArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("type_id", C.c_void_p), ("kind", C.c_void_p),
                ("img_src", C.c_void_p), ("pos_feat", C.c_void_p), ("out", C.c_void_p),
                ("T", C.c_int32), ("hidden", C.c_int32), ("mode", C.c_int32),
                ("type_vocab", C.c_int32), ("dtype", C.c_int32)]


class AdamSegment(C.Structure):
    _fields_ = [("grad", C.c_void_p), ("master", C.c_void_p), ("exp_avg", C.c_void_p),
                ("exp_avg_sq", C.c_void_p), ("model", C.c_void_p), ("n", C.c_int64),
                ("step_size", C.c_float), ("lr_wd", C.c_float),
                ("grad_dtype", C.c_int32), ("model_dtype", C.c_int32),
                ("weight_decay", C.c_float), ("group", C.c_int32), ("step_offset", C.c_int32),
                ("flags", C.c_int32)]


MAX_PEERS = 8


class PeerAllreduceArgs(C.Structure):
    _fields_ = [("buf", C.c_void_p * MAX_PEERS), ("stage", C.c_void_p * MAX_PEERS),
                ("flags", C.POINTER(C.c_uint32) * MAX_PEERS),
                ("rank", C.c_int32), ("world", C.c_int32),
                ("offset", C.c_int64), ("count", C.c_int64), ("stage_bytes", C.c_int64),
                ("dtype", C.c_int32), ("max_ctas", C.c_int32), ("scale", C.c_float),
                ("timeout_ms", C.c_int32)]


F32 = 2
_lib = None


def load():
    """Load libub200.so (building is the job of ``uniter_b200.build`` / ``__graft_entry__``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libub200.so not found at %s — run `python -m uniter_b200.build` "
            "(there is no non-CUDA fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.ub200_version.restype = C.c_int
    lib.ub200_last_error_string.restype = C.c_char_p
    lib.ub200_device_check.restype = C.c_int
    lib.ub200_set_sm_reserve.restype = C.c_int
    lib.ub200_set_sm_reserve.argtypes = [C.c_int]
    lib.ub200_gemm.restype = C.c_int
    lib.ub200_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    lib.ub200_gemm_grouped.restype = C.c_int
    lib.ub200_gemm_grouped.argtypes = [C.POINTER(GemmArgs), C.c_int32, C.c_void_p]
    for name in ("ub200_attn_fwd", "ub200_attn_bwd"):
        getattr(lib, name).restype = C.c_int
        getattr(lib, name).argtypes = [C.POINTER(AttnArgs), C.c_void_p]
    lib.ub200_attn_bwd_workspace_bytes.restype = C.c_int64
    lib.ub200_attn_bwd_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.ub200_layernorm_fwd.restype = C.c_int
    lib.ub200_layernorm_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_void_p]
    lib.ub200_layernorm_bwd.restype = C.c_int
    lib.ub200_layernorm_bwd.argtypes = [C.POINTER(LnBwdArgs), C.c_void_p]
    lib.ub200_colsum.restype = C.c_int
    lib.ub200_colsum.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                 C.c_void_p]
    lib.ub200_embed_prep.restype = C.c_int
    lib.ub200_embed_prep.argtypes = [C.POINTER(EmbedPrepArgs), C.c_void_p]
    lib.ub200_embed_gather_cast.restype = C.c_int
    lib.ub200_embed_gather_cast.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.ub200_embed_rows_fwd.restype = C.c_int
    lib.ub200_embed_rows_fwd.argtypes = [C.POINTER(EmbedRowsArgs), C.c_void_p]
    lib.ub200_embed_bwd_scatter.restype = C.c_int
    lib.ub200_embed_bwd_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.ub200_embed_bwd_colsums.restype = C.c_int
    lib.ub200_embed_bwd_colsums.argtypes = [C.POINTER(EmbedColsumArgs), C.c_void_p]
    lib.ub200_cvt_from_f32_strided.restype = C.c_int
    lib.ub200_cvt_from_f32_strided.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                               C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.ub200_ce_fwd.restype = C.c_int
    lib.ub200_ce_fwd.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_void_p]
    lib.ub200_ce_bwd.restype = C.c_int
    lib.ub200_ce_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.ub200_dgelu_mul.restype = C.c_int
    lib.ub200_dgelu_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.ub200_dtanh_mul.restype = C.c_int
    lib.ub200_dtanh_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.ub200_cvt_from_f32.restype = C.c_int
    lib.ub200_cvt_from_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.ub200_adam_chunk.restype = C.c_int32
    lib.ub200_grad_sumsq.restype = C.c_int
    lib.ub200_grad_sumsq.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.ub200_adamw_step.restype = C.c_int
    lib.ub200_adamw_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
    lib.ub200_adam_prep.restype = C.c_int
    lib.ub200_adam_prep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ub200_gather_rows.restype = C.c_int
    lib.ub200_gather_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.ub200_peer_flags_bytes.restype = C.c_int64
    lib.ub200_peer_stage_bytes.restype = C.c_int64
    lib.ub200_peer_stage_bytes.argtypes = [C.c_int64, C.c_int32]
    lib.ub200_peer_allreduce.restype = C.c_int
    lib.ub200_peer_allreduce.argtypes = [C.POINTER(PeerAllreduceArgs), C.c_void_p]
    lib.ub200_peer_ipc_export.restype = C.c_int
    lib.ub200_peer_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    lib.ub200_peer_ipc_open.restype = C.c_int
    lib.ub200_peer_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.ub200_peer_ipc_close.restype = C.c_int
    lib.ub200_peer_ipc_close.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libub200 error %d: %s" % (rc, load().ub200_last_error_string().decode()))


def dtype_code(t, allow_f32=False):
    import torch
    if t == torch.bfloat16:
        return BF16
    if t == torch.float16:
        return F16
    if t == torch.float32 and allow_f32:
        return F32
    raise TypeError("libub200 computes in fp16 or bf16, got %s" % t)


def ptr(t):
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


class PinnedRing(object):
    """Rotating pinned host staging buffers for small per-step H2D copies (packing metadata,
    optimizer segment tables).  A copy from PAGEABLE memory makes the host wait until the stream
    has drained, i.e. it costs a full synchronisation per step; from pinned memory it is just
    another asynchronous stream operation.  A slot is reused only after the copy that last read it
    has completed (event), which in steady state is always already true."""

    def __init__(self, slots=8):
        self.slots = [None] * slots
        self.events = [None] * slots
        self.i = 0

    def upload(self, host_tensor, device):
        """Async copy of a contiguous CPU tensor to `device` through a pinned slot."""
        import torch
        nbytes = host_tensor.numel() * host_tensor.element_size()
        k = self.i
        self.i = (self.i + 1) % len(self.slots)
        if self.events[k] is not None:
            self.events[k].synchronize()
        buf = self.slots[k]
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
            self.slots[k] = buf
        stage = buf[:nbytes].view(host_tensor.dtype)
        stage.copy_(host_tensor.reshape(-1))
        dev = torch.empty(host_tensor.numel(), dtype=host_tensor.dtype, device=device)
        dev.copy_(stage, non_blocking=True)
        ev = self.events[k]
        if ev is None:
            ev = torch.cuda.Event()
            self.events[k] = ev
        ev.record(torch.cuda.current_stream(device))
        return dev
