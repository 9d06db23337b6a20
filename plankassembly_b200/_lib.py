"""ctypes binding of include/plank_b200.h.  Fails loudly: no CPU or eager fallback exists."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libplank_b200.so')

vp, i32, i64, u64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t


class AttnFwdArgs(C.Structure):
    _fields_ = [('q', vp), ('k', vp), ('v', vp), ('ldq', i64), ('ldk', i64), ('ldv', i64),
                ('o', vp), ('ldo', i64), ('lse', vp), ('kpm', vp),
                ('B', i32), ('H', i32), ('Lq', i32), ('Lk', i32), ('dh', i32), ('causal', i32),
                ('scale', f32), ('p_drop', f32), ('seed', u64), ('offset', u64), ('impl', i32), ('round_out', i32),
                ('drop_rows', vp), ('drop_cols', vp), ('kv_len', vp)]


class AttnBwdArgs(C.Structure):
    _fields_ = [('q', vp), ('k', vp), ('v', vp), ('ldq', i64), ('ldk', i64), ('ldv', i64),
                ('o', vp), ('d_o', vp), ('ldo', i64), ('lse', vp), ('delta', vp),
                ('dq', vp), ('dk', vp), ('dv', vp), ('lddq', i64), ('lddk', i64), ('lddv', i64),
                ('kpm', vp),
                ('B', i32), ('H', i32), ('Lq', i32), ('Lk', i32), ('dh', i32), ('causal', i32),
                ('scale', f32), ('p_drop', f32), ('seed', u64), ('offset', u64), ('impl', i32), ('round_out', i32),
                ('drop_rows', vp), ('drop_cols', vp), ('dbias', vp), ('kv_len', vp)]


class GemmArgs(C.Structure):
    _fields_ = [('a', vp), ('lda', i64), ('a_mn', i32), ('b', vp), ('ldb', i64), ('b_mn', i32),
                ('c', vp), ('ldc', i64), ('bias', vp), ('relu', i32), ('p_drop', f32), ('seed', u64), ('offset', u64),
                ('alpha', f32), ('M', i32), ('N', i32), ('K', i32), ('batch', i32),
                ('a_batch_rows', i64), ('b_batch_rows', i64), ('c_batch_stride', i64), ('split_k', i32), ('accumulate', i32), ('round_out', i32),
                ('mask_out', vp), ('mask_in', vp), ('colsum', vp), ('mask_scale', f32)]


fp = C.c_void_p
PA_MAX_DEC_LAYERS = 8


class DecodeLayer(C.Structure):
    _fields_ = [(n, fp) for n in ('w_sqkv', 'b_sqkv', 'w_so', 'b_so', 'g1', 'be1', 'w_cq', 'b_cq', 'w_co', 'b_co', 'g2', 'be2',
                                  'w_f1', 'b_f1', 'w_f2', 'b_f2', 'g3', 'be3', 'self_k', 'self_v', 'cross_kv')]


class DecodeFusedArgs(C.Structure):
    _fields_ = [('B', i32), ('S', i32), ('T', i32), ('d', i32), ('H', i32), ('ff', i32), ('V', i32), ('L', i32), ('dof', i32),
                ('end_token', i32), ('layer_eps', f32), ('final_eps', f32), ('layers', DecodeLayer * PA_MAX_DEC_LAYERS),
                ('gf', fp), ('bf', fp), ('w_heads', fp), ('b_heads', fp), ('e_val', fp), ('e_coord', fp), ('e_pos', fp),
                ('kpm', fp), ('y', fp), ('o', fp), ('part', fp), ('part_bytes', i64), ('hfin', fp), ('samples', fp),
                ('attach', fp), ('first_end', fp), ('state', fp), ('chains', i32), ('profile', i32)]


# name -> (restype, argtypes); mirrors include/plank_b200.h one to one
SIGNATURES = {
    'pa_abi_version': (i32, []),
    'pa_last_error': (C.c_char_p, []),
    'pa_device_ok': (i32, []),
    'pa_embed_input_fwd': (i32, [C.POINTER(vp), C.POINTER(vp), i32, i64, i32, vp, vp, vp]),
    'pa_embed_input_bwd': (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), i32, i64, i32, vp]),
    'pa_embed_output_fwd': (i32, [vp, i64, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp]),
    'pa_embed_output_bwd': (i32, [vp, vp, i64, i32, i32, i32, vp, vp, vp, i32, vp]),
    'pa_add_ln_fwd': (i32, [vp, vp, vp, vp, vp, f32, f32, u64, u64, i64, i32, vp, vp, vp, vp, vp]),
    'pa_add_ln_bwd_workspace': (sz, [i64, i32]),
    'pa_add_ln_bwd': (i32, [vp, vp, vp, vp, vp, vp, f32, u64, u64, i64, i32, vp, vp, i32, vp, vp, vp, vp, vp]),
    'pa_relu_dropout_fwd': (i32, [vp, i64, f32, u64, u64, vp]),
    'pa_relu_dropout_bwd': (i32, [vp, vp, i64, f32, i32, vp]),
    'pa_relu_dropout_bwd_colsum': (i32, [vp, vp, i64, i32, f32, i32, vp, vp]),
    'pa_round_tf32': (i32, [vp, vp, i64, vp]),
    'pa_split3_tf32': (i32, [vp, i64, vp, i64, i32, i32, vp]),
    'pa_dropout_mask_words': (sz, [i32, i32, i32, i32]),
    'pa_dropout_mask': (i32, [vp, vp, i32, i32, i32, f32, u64, u64, vp]),
    'pa_attn_fwd': (i32, [C.POINTER(AttnFwdArgs), vp]),
    'pa_debug_attn_prof': (i32, [vp]),
    'pa_attn_bwd': (i32, [C.POINTER(AttnBwdArgs), vp]),
    'pa_gemm_tf32': (i32, [C.POINTER(GemmArgs), vp]),
    'pa_dist_loss_fwd': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, vp, vp, vp]),
    'pa_dist_loss_bwd': (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, vp, vp, i32, vp]),
    'pa_dist_train_full': (i32, [vp, vp, vp, i32, i32, i32, f32, vp, vp]),
    'pa_tokenize_lines': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'pa_tokenize_planks': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    'pa_parse_sequences': (i32, [vp, i64, i32, i32, i32, i32, vp, i32, vp, vp, vp]),
    'pa_plank_iou': (i32, [vp, vp, vp, i32, vp, vp, i32, i32, vp, vp, vp, vp]),
    'pa_adam_chunk_elems': (i32, []),
    'pa_adam_flat': (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, vp]),
    'pa_gemm_skinny_f32': (i32, [vp, i64, vp, i64, vp, vp, i64, i32, i32, i32, i32, vp]),
    'pa_decode_advance': (i32, [vp, vp]),
    'pa_decode_embed': (i32, [vp, i64, i32, i32, vp, i32, vp, vp, vp, i32, vp, vp]),
    'pa_decode_attn': (i32, [vp, i64, vp, vp, i64, vp, vp, i64, i64, i32, i32, vp, vp, vp, i32, i32, i32, f32, vp, vp]),
    'pa_decode_head': (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp, i32, vp, vp, i64, vp, vp]),
    'pa_decode_fused_workspace': (sz, [i32, i32, i32, i32]),
    'pa_decode_fused': (i32, [C.POINTER(DecodeFusedArgs), vp]),
    'pa_debug_decode_prof': (i32, [vp]),
}

_lib = None


class PlankB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlankB200Error(
            f'{LIB_PATH} is missing: build it with `python -m plankassembly_b200.build` '
            '(or __graft_entry__.build()).  plankassembly_b200 has no CPU/eager fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError => header/library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.pa_abi_version() != 3:
        raise PlankB200Error('ABI version mismatch')
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        raise PlankB200Error(f'{what} failed ({rc}): {load().pa_last_error().decode()}')


def launch_count():
    return _LAUNCHES[0]


_LAUNCHES = [0]
PROFILE_HOOK = None     # {entry-point name: (list, work_fn | None)} -> bench.py brackets those entry points with CUDA events


def call(name, *args, launches=1):
    """Invoke a C-ABI entry point and raise on a non-zero status."""
    lib = load()
    hook = PROFILE_HOOK
    if hook is not None and name in hook:
        import torch
        rec, work = hook[name]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        rec.append((e0, e1, work(args) if work is not None else 0.0))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise PlankB200Error(f'{name} failed ({rc}): {lib.pa_last_error().decode()}')
    _LAUNCHES[0] += launches
