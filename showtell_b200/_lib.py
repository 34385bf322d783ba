"""ctypes binding of libshowtell_b200.so (the C ABI declared in include/showtell_b200.h).

There is deliberately no fallback: if the shared library is missing or a tensor is not a
contiguous CUDA tensor of the expected dtype, the call raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("SHOWTELL_B200_LIBNAME", "libshowtell_b200.so"))   # A/B builds only

ST_GRU, ST_LSTM = 0, 1
ST_MAX_STEPS = 128
_STATUS = {-1: "bad shape", -2: "lengths not sorted", -3: "unsupported", -4: "CUDA error",
           -5: "NULL pointer", -6: "workspace too small"}

_lib = None


class RnnWeights(C.Structure):
    """Mirror of st_rnn_weights."""
    _fields_ = [("kind", C.c_int), ("L", C.c_int), ("E", C.c_int), ("H", C.c_int), ("V", C.c_int),
                ("emb", C.c_void_p), ("Wih_host", C.POINTER(C.c_void_p)),
                ("Whh_host", C.POINTER(C.c_void_p)), ("bih_host", C.POINTER(C.c_void_p)),
                ("bhh_host", C.POINTER(C.c_void_p)), ("Wv", C.c_void_p), ("bv", C.c_void_p),
                ("gemm_mode", C.c_int)]


_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64
_IP = C.POINTER(C.c_int)
_SIGS = {
    "st_version": (C.c_int, []),
    "st_last_error": (C.c_char_p, []),
    "st_launch_count": (C.c_int64, []),
    "st_device_info": (_I, [_IP, _IP, _IP, C.POINTER(_L)]),
    "st_sgemm": (_I, [_I, _I, _I, _I, _I, _F, _P, _I, _P, _I, _F, _P, _I, _P, _P]),
    "st_gemm_bf16": (_I, [_I, _I, _I, _P, _I, _P, _I, _P, _I, _I, _P, _F, _F, _P]),
    "st_gemm_bf16_ex": (_I, [_I, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _F, _F, _P]),
    "st_cast_bf16": (_I, [_P, _I, _I, _I, _P, _I, _P, _I, _P]),
    "st_vocab_ce_parts": (_I, [_I]),
    "st_debug_gemm_variant": (_I, [_I]),
    "st_split_tf32": (_I, [_P, _I, _I, _I, _P, _P, _I, _P]),
    "st_gemm_tf32x3": (_I, [_I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _P, _F, _F, _P]),
    "st_topk_parts": (_I, [_I]),
    "st_collate_sort": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "st_gather_rows_bytes": (_I, [_P, _P, _P, _I, C.c_int64, _P]),
    "st_collate_captions": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "st_allreduce_set_timeout_ms": (_I, [C.c_int64]),
    "st_allreduce_error": (_I, []),
    "st_debug_decode_table": (_I, [_I]),
    "st_debug_decode_screen": (_I, [_I]),
    "st_debug_relayout_legacy": (_I, [_I]),
    "st_debug_allreduce_pair_p2p": (_I, [_I]),
    "st_row_norm_max": (_I, [_P, _I, _I, _P, _P]),
    "st_vocab_topk_screen": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P]),
    "st_gemm_bf16_screen": (_I, [_I, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P, _P]),
    "st_gemm_tf32x3_topk": (_I, [_I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "st_debug_set_pdl": (_I, [_I]),
    "st_debug_set_bwd_kp": (_I, [_I]),
    "st_debug_set_bwd_ks": (_I, [_I]),
    "st_debug_set_coop": (_I, [_I]),
    "st_rnn_seq_tc_bwd_ctas": (_I, [_I, _I, _I]),
    "st_gemm_set_sm_limit": (_I, [_I]),
    "st_gemm_set_c_zeroed": (_I, [_I]),
    "st_scale_multi": (_I, [_I, _P, _P, _P, _P, _P]),
    "st_bn1d_fwd": (_I, [_P, _I, _I, _I, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P, _I, _P]),
    "st_bn1d_bwd": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P]),
    "st_sgd_step": (_I, [_I, _P, _P, _P, _P, _P, _F, _F, _I, _P, _P]),
    "st_adam_step": (_I, [_I, _P, _P, _P, _P, _P, _P, _F, C.c_double, C.c_double, _F, _L, _P, _P]),
    "st_allreduce_flag_words": (_I, [_I]),
    "st_allreduce_sum_f32": (_I, [C.POINTER(C.c_void_p), _P, C.POINTER(C.c_void_p), _I, _I, _L, _I, _P]),
    "st_vocab_ce_fwd": (_I, [_I, _I, _I, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "st_vocab_ce_bwd": (_I, [_I, _I, _I, _P, _I, _P, _I, _P, _P, _P, _F, _P, _I, _P, _I, _P]),
    "st_pack_inputs": (_I, [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _IP, _P]),
    "st_pack_inputs_bf16": (_I, [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _IP, _P]),
    "st_pack_inputs_bwd": (_I, [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _IP, _P]),
    "st_pack_targets": (_I, [_P, _P, _I, _I, _I, _IP, _P]),
    "st_token_error": (_I, [C.POINTER(_L), _I]),
    "st_gather_rows": (_I, [_P, _I, _P, _I, _P, _I, _I, _P]),
    "st_rowsum_bf16": (_I, [_P, _P, _I, _I, _I, _P]),
    "st_colsum": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "st_rnn_seq_fwd": (_I, [_I, _I, _I, _IP, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "st_rnn_seq_bwd": (_I, [_I, _I, _I, _IP, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "st_rnn_seq_tc_supported": (_I, [_I, _I]),
    "st_rnn_cluster_supported": (_I, [_I, _I]),
    "st_rnn_cluster_fwd": (_I, [_I, _I, _I, _IP, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "st_rnn_seq_tc_fwd": (_I, [_I, _I, _I, _IP, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "st_rnn_seq_tc_bwd": (_I, [_I, _I, _I, _IP, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P,
                               _P, _P, _P]),
    "st_rnn_step_x_tc_supported": (_I, [_I, _I, _I]),
    "st_rnn_step_x_tc_fwd": (_I, [_I, _I, _I, _I, _IP, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "st_rnn_step_x_tc_bwd": (_I, [_I, _I, _I, _IP, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _I,
                                  _P, _P, _I, _P, _I, _I, _P]),
    "st_shift_states": (_I, [_P, _P, _P, _I, _I, _IP, _P]),
    "st_shift_states_bf16": (_I, [_P, _P, _P, _I, _I, _IP, _P]),
    "st_attn_relayout": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "st_grid_mean_bpc": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "st_attn_relayout_bf16in": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "st_attn_step_fwd": (_I, [_I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _I, _P, _P, _I, _P, _I, _I, _P]),
    "st_attn_step_bwd": (_I, [_I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P]),
    "st_attn_hoist_bwd": (_I, [_I, _IP, _I, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P]),
    "st_attn_ctx_all": (_I, [_I, _IP, _I, _I, _I, _P, _I, _P, _P, _P, _I, _I, _P]),
    "st_attn_embed_q": (_I, [_I, _IP, _I, _I, _I, _P, _P, _I, _P, _P]),
    "st_attn_penalty": (_I, [_I, _P, _F, _P, _P, _P]),
    "st_add_rows": (_I, [_P, _P, _I, _I, _P]),
    "st_ce_fwd_bwd": (_I, [_P, _I, _P, _I, _I, _P, _P, _P, _F, _P]),
    "st_argmax_rows": (_I, [_P, _I, _I, _I, _P, _I, _P]),
    "st_topk_rows": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P]),
    "st_decode_workspace_bytes": (_L, [C.POINTER(RnnWeights), _I, _I, _I]),
    "st_decode_greedy": (_I, [C.POINTER(RnnWeights), _P, _I, _I, _P, _P, _L, _P]),
    "st_decode_beam_chain": (_I, [C.POINTER(RnnWeights), _P, _I, _I, _I, _P, _P, _P, _P, _L, _P]),
    "st_decode_beam_tree": (_I, [C.POINTER(RnnWeights), _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _L, _P]),
}
# symbols added by later source files are registered here as they appear
_OPTIONAL_SIGS = {}


def load():
    """Load (once) and return the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m showtell_b200.build` "
            "(showtell_b200 has no CPU / PyTorch fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(_SIGS.items()) + list(_OPTIONAL_SIGS.items()):
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("SHOWTELL_PDL", "1") == "0":       # A/B switch for programmatic dependent launch
        lib.st_debug_set_pdl(0)
    if os.environ.get("SHOWTELL_GEMM_VARIANT"):            # A/B switch: st_debug_gemm_variant flags (e.g. 0x10000: no pair kernel)
        lib.st_debug_gemm_variant(int(os.environ["SHOWTELL_GEMM_VARIANT"], 0))
    if os.environ.get("SHOWTELL_COOP", "0") == "1":        # cooperative launches of the persistent BPTT kernel
        lib.st_debug_set_coop(1)
    if os.environ.get("SHOWTELL_BWD_KS"):                  # A/B switch for the BPTT kernel's K split / tile height
        lib.st_debug_set_bwd_ks(int(os.environ["SHOWTELL_BWD_KS"]))
    _lib = lib
    return lib


def exported_symbols():
    return sorted(list(_SIGS) + list(_OPTIONAL_SIGS))


def last_error():
    return load().st_last_error().decode("utf-8", "replace")


def check(status, what=""):
    """Translate an st_status into the exception the reference API would have raised:
    RuntimeError for unsorted lengths (pack_padded_sequence, rnn.py:31) and CUDA failures,
    ValueError for bad options / shapes (main.py:102,113)."""
    if status == 0:
        return
    msg = f"{what}: {_STATUS.get(status, status)}: {last_error()}"
    if status in (-2, -4):
        raise RuntimeError(msg)
    raise ValueError(msg)


def raise_token_error():
    """IndexError if a packing kernel has met a token id outside [0, vocab_size) (checked at call boundaries:
    the step that contained it may have been issued one call earlier; the access itself was clamped)."""
    bad = _L(0)
    n = load().st_token_error(C.byref(bad), 1)
    if n:
        raise IndexError(f"showtell_b200: {n} caption token id(s) outside [0, vocab_size) in a previous step "
                         f"(first: {bad.value}); nn.Embedding / CrossEntropyLoss of the reference assert on this")


def ptr(t, dtype=None):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"expected a tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError("showtell_b200 runs on CUDA tensors only (no CPU fallback); "
                           "move the module and its inputs to the GPU")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected dtype {dtype}, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def int_array(values):
    arr = (C.c_int * len(values))(*[int(v) for v in values])
    return arr


def batch_sizes(lengths):
    """lengths (sorted descending, as utils.py:66 produces them) -> batch_sizes per step.
    Raises RuntimeError like pack_padded_sequence(enforce_sorted=True) does (rnn.py:31)."""
    lengths = [int(l) for l in lengths]
    if len(lengths) == 0:
        raise RuntimeError("empty batch")
    if any(lengths[i] < lengths[i + 1] for i in range(len(lengths) - 1)):
        raise RuntimeError("`lengths` array must be sorted in decreasing order")
    if lengths[-1] <= 0:
        raise RuntimeError("Length of all samples has to be greater than 0")
    T = lengths[0]
    if T > ST_MAX_STEPS:
        raise ValueError(f"caption length {T} exceeds ST_MAX_STEPS={ST_MAX_STEPS}")
    out, j = [], len(lengths)
    for t in range(T):
        while j > 0 and lengths[j - 1] <= t:
            j -= 1
        out.append(j)
    return out
