"""the batched device tokeniser (csrc/tokenize.cu) against a numpy restatement of the reference's per-sample
tokeniser (ref: plankassembly/datasets/line_data.py:34-83 `prepare_input_sequence`, :85-109 `prepare_output_sequence`,
data_utils.py:6-12 `quantize_values`).  TEST INFRASTRUCTURE: the restatement below follows those lines one to one."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from plankassembly_b200 import synthetic as syn, tokenizer  # noqa: E402

END, PAD, VOCAB = syn.END, syn.PAD, syn.VOCAB


def quantize_values(verts, n_bits=9):                       # ref data_utils.py:6-12
    return ((verts - (-1)) * (2 ** n_bits - 1) / (1 - (-1))).astype('long')


def ref_input_sequence(lines, views, types, max_input_length):   # ref line_data.py:34-83
    value = quantize_values(np.array(lines))
    view, typ = np.array(views), np.array(types)
    order = np.lexsort(np.concatenate((value, view[..., np.newaxis]), axis=1).T[[3, 1, 2, 0, 4]])
    value, view, typ = value[order].flatten(), view[order], typ[order]
    _, counts = np.unique(view, return_counts=True)
    pos = np.concatenate([np.arange(c) for c in counts])
    coord = np.arange(len(value)) % 4
    pos, view, typ = np.repeat(pos, 4), np.repeat(view, 4), np.repeat(typ, 4)
    value = np.append(value, END)
    pad = max_input_length - len(value)
    value = np.pad(value, (0, pad - 1), constant_values=PAD)
    return {'input_value': value, 'input_pos': np.pad(pos, (0, pad)), 'input_coord': np.pad(coord, (0, pad)),
            'input_view': np.pad(view, (0, pad)), 'input_type': np.pad(typ, (0, pad)), 'input_mask': value == PAD}


def ref_output_sequence(planks, attach, max_output_length):    # ref line_data.py:85-109
    value = np.append(quantize_values(planks), END)
    value = np.pad(value, (0, max_output_length - len(value)), constant_values=PAD)
    label = np.pad(attach, (0, max_output_length - len(attach)), constant_values=-1)
    label[label != -1] += VOCAB
    label[label == -1] = value[label == -1]
    return {'output_value': value, 'output_label': label, 'output_mask': value == PAD}


@pytest.mark.parametrize('max_input_length,seed', [(300, 0), (1200, 1), (513, 2)])
def test_tokenize_lines_matches_reference_tokeniser(max_input_length, seed):
    g = np.random.default_rng(seed)
    cfg = syn.make_cfg(max_input_length=max_input_length)
    max_lines = (max_input_length - 2) // 4
    lines, views, types = [], [], []
    for b in range(17):
        n = int(g.integers(1, max_lines + 1)) if b else max_lines            # drawing 0 fills the sequence completely
        grid = g.integers(0, 12, (n, 4)) / 6.0 - 1.0 if b % 3 == 0 else g.uniform(-1, 1, (n, 4))   # coarse grid => many ties / duplicates
        lines.append(grid)
        views.append(g.integers(0, 3, n) if b != 5 else np.full(n, 2))       # one drawing with a single view
        types.append(g.integers(0, 2, n))
    lines[1] = np.array([[-1.0, 1.0, 0.0, 0.999999]]);  views[1] = np.array([1]);  types[1] = np.array([0])   # range ends
    out = tokenizer.tokenize_lines(lines, views, types, cfg.DATA, cfg.TOKEN)
    for b in range(len(lines)):
        ref = ref_input_sequence(lines[b], views[b], types[b], max_input_length)
        for k, v in ref.items():
            assert np.array_equal(out[k][b].cpu().numpy(), v), (b, k)
        assert int(out['kv_len'][b]) == 4 * len(lines[b]) + 1
    # sideface batches have no types (ref trainer_sideface.py): the plane is simply absent
    out2 = tokenizer.tokenize_lines(lines, views, None, cfg.DATA, cfg.TOKEN)
    assert 'input_type' not in out2 and torch.equal(out2['input_value'], out['input_value'])
    with pytest.raises(Exception):
        tokenizer.tokenize_lines([g.uniform(-1, 1, (max_lines + 1, 4))], [np.zeros(max_lines + 1, dtype=int)], None, cfg.DATA, cfg.TOKEN)


def test_tokenize_planks_matches_reference_tokeniser():
    g = np.random.default_rng(3)
    cfg = syn.make_cfg(max_output_length=128)
    coords, attach = [], []
    for b in range(9):
        n = int(g.integers(1, 22)) * 6 if b else 126
        coords.append(g.uniform(-1, 1, n))
        a = np.full(n, -1)
        idx = g.random(n) < 0.4
        a[idx] = g.integers(0, np.maximum(np.arange(n), 1))[idx]
        attach.append(a)
    out = tokenizer.tokenize_planks(coords, attach, cfg.DATA, cfg.TOKEN)
    for b in range(len(coords)):
        ref = ref_output_sequence(coords[b], attach[b].copy(), 128)
        for k, v in ref.items():
            assert np.array_equal(out[k][b].cpu().numpy(), v), (b, k)
        assert int(out['out_len'][b]) == len(coords[b]) + 1


def test_tokenised_batch_feeds_the_model():
    """The synthetic generator's drawings, tokenised on the device from raw geometry, give the very batch the host path builds
    (synthetic.make_sample re-creates LineDataset's layout) and the model takes it."""
    from plankassembly_b200.models import build_model
    cfg = syn.tiny_cfg()
    host = syn.batch_for(cfg, range(6))
    # recover raw geometry of the same drawings: de-quantised bin centres reproduce the tokens exactly
    lines, views, types = [], [], []
    for b in range(6):
        n = int((~host['input_mask'][b]).sum() - 1) // 4
        q = host['input_value'][b, :4 * n].reshape(n, 4).numpy()
        lines.append((q + 0.5) * 2.0 / 511.0 - 1.0)
        views.append(host['input_view'][b, :4 * n:4].numpy())
        types.append(host['input_type'][b, :4 * n:4].numpy())
    perm = [np.random.default_rng(b).permutation(len(l)) for b, l in enumerate(lines)]      # the loader's line order is arbitrary
    dev = tokenizer.tokenize_lines([l[p] for l, p in zip(lines, perm)], [v[p] for v, p in zip(views, perm)],
                                   [t[p] for t, p in zip(types, perm)], cfg.DATA, cfg.TOKEN)
    for k in ('input_value', 'input_pos', 'input_coord', 'input_view', 'input_type', 'input_mask'):
        assert torch.equal(dev[k].cpu(), host[k]), k
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in host.items()}
    batch.update({k: v for k, v in dev.items() if k.startswith('input')})
    m = build_model(cfg)
    m.load_state_dict(syn.init_state_dict(cfg))
    out = m.cuda().train()(batch)
    assert torch.isfinite(out['loss'])
