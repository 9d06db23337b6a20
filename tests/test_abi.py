"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every
symbol include/plank_b200.h declares (no compute calls without a GPU), and the host-side mirror
keeps the reference's parameter names and shapes."""
import ctypes
import math
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from plankassembly_b200 import build as b
    b.build()
    from plankassembly_b200 import _lib
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'plank_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pa_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol(lib):
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/plank_b200.h but not exported'
    assert lib.pa_abi_version() == 3


def test_ctypes_table_matches_header(lib):
    from plankassembly_b200._lib import SIGNATURES
    assert sorted(SIGNATURES) == header_symbols()


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from plankassembly_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.PlankB200Error):
        _lib.load()


def test_state_dict_names_match_reference_listing():
    """SURVEY.md section 8b lists the reference's parameter names; ours must be identical
    (the seeded init in synthetic.init_state_dict uses the same listing and is loaded by the
    reference itself in oracle/gen_golden.py)."""
    from plankassembly_b200 import synthetic as syn
    from plankassembly_b200.models import build_model
    for cfg in (syn.tiny_cfg(), syn.config1()):
        m = build_model(cfg)
        sd = syn.init_state_dict(cfg)
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert ours == {k: tuple(v.shape) for k, v in sd.items()}
        m.load_state_dict(sd, strict=True)
    d = 256
    assert ours['encoder.layers.0.self_attn.in_proj_weight'] == (3 * d, d)
    assert ours['decoder.layers.1.multihead_attn.out_proj.weight'] == (d, d)
    assert ours['input_embeddings.input_pos.weight'] == (math.ceil(1200 / 4), d)
    assert ours['query_pos_embedding.weight'] == (math.ceil(128 / 6), d)


def test_build_model_reads_reference_cfg_fields():
    from plankassembly_b200 import synthetic as syn
    from plankassembly_b200.models import build_model
    cfg = syn.config2()
    m = build_model(cfg)
    assert m.token.END == 512 and m.token.PAD == 513 and m.vocab_size == 514
    assert m.max_output_length == 256 and m.num_output_dof == 6
    assert m.layer_eps == 1.0 and m.encoder.norm is not None
    n = sum(p.numel() for p in m.parameters())
    assert abs(n - 32.43e6) < 0.05e6


def test_parse_sequence():
    from plankassembly_b200 import synthetic as syn
    from plankassembly_b200.models import build_model
    m = build_model(syn.tiny_cfg())
    seq = torch.tensor([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 512, 5, 512])
    out = m.parse_sequence(seq)
    assert out.shape == (2, 6) and out[1, 5] == 12
    assert m.parse_sequence(torch.tensor([512, 1, 2])).shape == (0, 6)


def test_ctypes_struct_layouts_match_header(tmp_path):
    """The argument structs cross the C ABI by pointer: sizes and the offsets of the trailing fields of the ctypes mirrors
    must equal what a C compiler sees in include/plank_b200.h."""
    import ctypes
    import subprocess
    from plankassembly_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probes = [('pa_attn_fwd_args', _lib.AttnFwdArgs, ['q', 'ldq', 'B', 'scale', 'seed', 'impl', 'drop_rows', 'kv_len']),
              ('pa_attn_bwd_args', _lib.AttnBwdArgs, ['q', 'd_o', 'delta', 'lddq', 'kpm', 'scale', 'seed', 'dbias', 'kv_len']),
              ('pa_gemm_args', _lib.GemmArgs, ['a', 'b', 'c', 'bias', 'seed', 'alpha', 'batch', 'a_batch_rows', 'split_k', 'round_out', 'mask_out', 'colsum', 'mask_scale']),
              ('pa_decode_layer', _lib.DecodeLayer, ['w_sqkv', 'g1', 'w_f2', 'self_k', 'cross_kv']),
              ('pa_decode_fused_args', _lib.DecodeFusedArgs, ['B', 'end_token', 'layer_eps', 'layers', 'gf', 'kpm', 'part', 'part_bytes',
                                                              'hfin', 'first_end', 'state', 'chains', 'profile'])]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "plank_b200.h"', 'int main(void) {']
    for cname, _, fields in probes:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'probe.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'probe'
    subprocess.run(['gcc', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)], check=True)
    seen = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct, fields in probes:
        assert int(seen[cname]) == ctypes.sizeof(ct), cname
        for f in fields:
            assert int(seen[f'{cname}.{f}']) == getattr(ct, f).offset, f'{cname}.{f}'
