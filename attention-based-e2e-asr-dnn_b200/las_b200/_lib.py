"""ctypes binding of liblas_b200.so (the C ABI declared in include/las_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liblas_b200.so')

c_f32p = C.c_void_p     # all device pointers travel as raw addresses
c_ll = C.c_longlong


class LasGemmF32(C.Structure):
    _fields_ = [
        ('A', C.c_void_p), ('B', C.c_void_p), ('C', C.c_void_p), ('bias1', C.c_void_p), ('bias2', C.c_void_p),
        ('M', C.c_int), ('N', C.c_int), ('K', C.c_int), ('batch', C.c_int),
        ('a_m_so', c_ll), ('a_m_si', c_ll), ('a_m_inner', C.c_int),
        ('a_k_so', c_ll), ('a_k_si', c_ll), ('a_k_inner', C.c_int),
        ('b_k_so', c_ll), ('b_k_si', c_ll), ('b_k_inner', C.c_int),
        ('b_n_s', c_ll),
        ('c_m_so', c_ll), ('c_m_si', c_ll), ('c_m_inner', C.c_int),
        ('bsA', c_ll), ('bsB', c_ll), ('bsC', c_ll),
        ('alpha', C.c_float), ('beta', C.c_float),
        ('prof_tag', C.c_int),
    ]


class LasGemmTc(C.Structure):
    _fields_ = [
        ('A', C.c_void_p), ('B', C.c_void_p), ('C', C.c_void_p), ('bias1', C.c_void_p), ('bias2', C.c_void_p),
        ('M', C.c_int), ('N', C.c_int), ('K', C.c_int),
        ('a_batches', C.c_int), ('k_batches', C.c_int),
        ('a_s1', c_ll), ('a_s2', c_ll), ('b_s1', c_ll), ('b_s2', c_ll),
        ('c_bs', c_ll), ('ldc', c_ll),
        ('a_mn_major', C.c_int), ('b_mn_major', C.c_int), ('accumulate', C.c_int),
        ('lens', C.c_void_p),
        ('prof_tag', C.c_int),
        ('prof_flops', C.c_double),
        ('splitk', C.c_int),
        ('workspace', C.c_void_p),
        ('max_ctas', C.c_int),
        ('a_f16', C.c_int), ('b_f16', C.c_int),
        ('k_chunk', C.c_int), ('k_chunk_stride', c_ll),
    ]


class LasAttnStep(C.Structure):
    _fields_ = [
        ('q', C.c_void_p), ('ld_q', c_ll),
        ('K', C.c_void_p), ('V', C.c_void_p), ('lens', C.c_void_p),
        ('w', C.c_void_p), ('ld_w', c_ll),
        ('w_b0', C.c_void_p),
        ('ctx', C.c_void_p), ('ld_ctx', c_ll),
        ('ctx2', C.c_void_p), ('ld_ctx2', c_ll),
        ('dctx', C.c_void_p), ('ld_dctx', c_ll),
        ('dctx2', C.c_void_p), ('ld_dctx2', c_ll),
        ('dq', C.c_void_p), ('ld_dq', c_ll), ('dq_accumulate', C.c_int),
        ('de', C.c_void_p),
        ('B', C.c_int), ('T', C.c_int), ('P', C.c_int), ('heads', C.c_int),
        ('scale', C.c_float),
        ('ctx2_bf16', C.c_void_p), ('ld_ctx2_bf16', c_ll),
        ('dq_bf16', C.c_void_p), ('ld_dq_bf16', c_ll),
        ('kv_bf16', C.c_int),
        ('fmask', C.c_void_p), ('ld_fmask', c_ll),
        ('w2', C.c_void_p),
        ('dctx2_nsplit', C.c_int), ('dctx2_split_stride', c_ll),
    ]


class LasSpeller(C.Structure):
    _fields_ = [
        ('B', C.c_int), ('T', C.c_int), ('P', C.c_int), ('E', C.c_int), ('DH', C.c_int), ('DO', C.c_int), ('V', C.c_int),
        ('heads', C.c_int), ('steps', C.c_int),
        ('sos_idx', C.c_int), ('pad_idx', C.c_int),
        ('training', C.c_int),
        ('use_tc', C.c_int),
        ('kv_bf16', C.c_int),
        ('init_force', C.c_int),
        ('emb', C.c_void_p), ('cls_b', C.c_void_p),
        ('w_ih0', C.c_void_p), ('w_hh0', C.c_void_p), ('b_ih0', C.c_void_p), ('b_hh0', C.c_void_p),
        ('w_ih1', C.c_void_p), ('w_hh1', C.c_void_p), ('b_ih1', C.c_void_p), ('b_hh1', C.c_void_p),
        ('wq', C.c_void_p), ('bq', C.c_void_p),
        ('init_query', C.c_void_p),
        ('K', C.c_void_p), ('V_', C.c_void_p), ('enc_lens', C.c_void_p),
        ('dec_y', C.c_void_p), ('ld_y', c_ll),
        ('use_gold_host', C.c_void_p),
        ('drop0', C.c_void_p), ('drop1', C.c_void_p),
        ('logits', C.c_void_p), ('att0', C.c_void_p), ('chars', C.c_void_p),
        ('fws', C.c_void_p), ('fws_floats', C.c_size_t),
        ('iws', C.c_void_p), ('iws_ints', C.c_size_t),
        ('K_f16', C.c_void_p), ('V_f16', C.c_void_p),
    ]


class LasSpellerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'dlogits', 'd_emb', 'd_cls_b', 'd_w_ih0', 'd_w_hh0', 'd_b_ih0', 'd_b_hh0', 'd_w_ih1', 'd_w_hh1', 'd_b_ih1', 'd_b_hh1',
        'd_wq', 'd_bq', 'd_init_query', 'dK', 'dV')]


class LasAdamTensor(C.Structure):
    _fields_ = [('p', C.c_void_p), ('g', C.c_void_p), ('m', C.c_void_p), ('v', C.c_void_p), ('vmax', C.c_void_p),
                ('numel', c_ll), ('step_size', C.c_float), ('bias_c2_sqrt', C.c_float)]


class LasAdamChunk(C.Structure):
    _fields_ = [('tensor', C.c_int), ('pad_', C.c_int), ('offset', c_ll)]


ADAM_CHUNK = 65536

# name -> (restype, argtypes); must list every symbol include/las_b200.h declares
SIGNATURES = {
    'las_abi_version': (C.c_int, []),
    'las_init': (C.c_int, [C.c_int]),
    'las_last_error': (C.c_char_p, []),
    'las_launch_count': (c_ll, []),
    'las_launch_count_reset': (None, []),
    'las_prof_enable': (None, [C.c_uint]),
    'las_prof_reset': (None, []),
    'las_prof_collect': (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(c_ll), C.POINTER(C.c_double)]),
    'las_gemm_f32': (C.c_int, [C.POINTER(LasGemmF32), C.c_void_p]),
    'las_gemm_bf16_tc': (C.c_int, [C.POINTER(LasGemmTc), C.c_void_p]),
    'las_cast_f32_to_bf16': (C.c_int, [C.c_void_p, c_ll, c_ll, c_ll, C.c_void_p, c_ll, c_ll, C.c_int, C.c_int, C.c_void_p]),
    'las_cast_f32_to_f16': (C.c_int, [C.c_void_p, c_ll, c_ll, c_ll, C.c_void_p, c_ll, c_ll, C.c_int, C.c_int, C.c_void_p]),
    'las_colsum_scratch_floats': (C.c_size_t, [C.c_int]),
    'las_colsum_f32': (C.c_int, [C.c_void_p, c_ll, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'las_lstm_rec_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'las_lstm_rec_fwd_f32': (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'las_lstm_rec_bwd_f32': (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'las_lstm_rec_tc_supported': (C.c_int, [C.c_int, C.c_int, C.c_int]),
    'las_lstm_rec_tc_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'las_lstm_rec_fwd_tc': (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'las_set_launch_start_stream': (None, [C.c_void_p]),
    'las_launch_start_mode': (C.c_int, []),
    'las_lstm_rec_fwd_arm_progress': (None, [C.c_void_p, C.c_int]),
    'las_lstm_rec_bwd_arm_progress': (None, [C.c_void_p, C.c_int]),
    'las_lstm_rec_fwd_progress_info': (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'las_stream_wait_value_geq': (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint]),
    'las_lstm_rec_fwd_tc_ex': (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    'las_lstm_rec_tc_set_debug': (None, [C.c_void_p]),
    'las_lstm_rec_bwd_tc': (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'las_lstm_rec_bwd_tc_dbias_slices': (C.c_int, [C.c_int, C.c_int, C.c_int]),
    'las_lstm_rec_bwd_tc_db': (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    'las_transpose_cast_bf16': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'las_attn_step_fwd_f32': (C.c_int, [C.POINTER(LasAttnStep), C.c_void_p]),
    'las_attn_step_bwd_f32': (C.c_int, [C.POINTER(LasAttnStep), C.c_void_p]),
    'las_lstm_cell_fwd_f32': (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p]),
    'las_lstm_cell_bwd_f32': (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p]),
    'las_transcript_cut_i32': (C.c_int, [C.c_void_p, c_ll, c_ll, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'las_speller_persistent': (C.c_int, [C.c_int] * 9),
    'las_speller_workspace_floats': (C.c_size_t, [C.POINTER(LasSpeller)]),
    'las_speller_workspace_ints': (C.c_size_t, [C.POINTER(LasSpeller)]),
    'las_speller_fwd_f32': (C.c_int, [C.POINTER(LasSpeller), C.c_void_p]),
    'las_collate_specaug_f32': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    'las_masked_ce_scratch_floats': (C.c_size_t, [C.c_int, C.c_int]),
    'las_masked_ce_f32': (C.c_int, [C.c_void_p, c_ll, C.c_void_p, c_ll, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    'las_speller_graph_stats': (None, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    'las_speller_bwd_f32': (C.c_int, [C.POINTER(LasSpeller), C.POINTER(LasSpellerGrads), C.c_void_p]),
    'las_speller_bwd_phases_f32': (C.c_int, [C.POINTER(LasSpeller), C.POINTER(LasSpellerGrads), C.c_int, C.c_void_p]),
    'las_adamw_amsgrad_fused': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_float,
                                          C.c_double, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: build it with `python __graft_entry__.py` (or `make -C '
                               f'attention-based-e2e-asr-dnn_b200/csrc`). las_b200 has no CPU / PyTorch fallback.')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.las_abi_version() != 1:
            raise RuntimeError('liblas_b200.so ABI version mismatch')
        _lib = lib
    return _lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = load().las_last_error()
        raise RuntimeError(f'las_b200 {what} failed (code {rc}): {msg.decode() if msg else "?"}')


def ptr(t):
    """Raw device address of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
