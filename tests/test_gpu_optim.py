"""plankassembly_b200.optim.FusedAdam (one pa_adam_flat launch over a flat master buffer, TF32 shadow weights
written by the same pass) against torch.optim.Adam as the reference constructs it (ref: trainer_complete.py:127-129)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from plankassembly_b200 import synthetic as syn  # noqa: E402


def test_fused_adam_matches_torch_adam_over_10_steps():
    from plankassembly_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(514, 128), (384, 128), (384,), (1, 128), (1,), (7, 33), (128, 256)]
    ours = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    o1, o2 = FusedAdam(ours, lr=1e-3), torch.optim.Adam(ref, lr=1e-3)
    for step in range(10):
        for i, (a, b) in enumerate(zip(ours, ref)):
            gr = torch.randn(*a.shape, generator=g).cuda() * (10.0 ** ((i % 3) - 2))
            a.grad, b.grad = (None, None) if (step == 3 and i == 2) else (gr, gr.clone())       # a parameter without a gradient is skipped
        o1.step(); o2.step()
    for a, b in zip(ours, ref):
        assert (a - b).abs().max().item() <= 1e-6 * b.abs().max().item() + 1e-7
    # moments and step counts are exposed in torch.optim.Adam's state_dict format and round-trip between the two
    sd = o1.state_dict()
    st0 = sd['state'][0]
    assert set(st0) == {'step', 'exp_avg', 'exp_avg_sq'} and float(st0['step']) == 10
    assert torch.allclose(st0['exp_avg'], o2.state_dict()['state'][0]['exp_avg'], rtol=1e-5, atol=1e-9)
    o3 = torch.optim.Adam(ref, lr=1e-3)
    o3.load_state_dict(sd)
    o1b = FusedAdam([torch.nn.Parameter(p.detach().clone()) for p in ours], lr=1e-3)
    o1b.load_state_dict(o2.state_dict())
    assert o1b._t == 10 and torch.allclose(o1b.state[o1b._ps[0]]['exp_avg'], o2.state[ref[0]]['exp_avg'])
    assert o1b.state[o1b._ps[0]]['exp_avg'].data_ptr() == o1b.exp_avg.data_ptr()          # still a view of the flat buffer


def test_fused_adam_trains_the_model_like_torch_adam():
    """Through the model: the shadows written by the optimizer kernel are the ones the next forward's GEMMs read."""
    from plankassembly_b200.models import build_model
    from plankassembly_b200.optim import FusedAdam
    cfg = syn.tiny_cfg()
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.batch_for(cfg, range(4)).items()}
    losses = []
    for make in (lambda ps: FusedAdam(ps, lr=1e-3), lambda ps: torch.optim.Adam(ps, lr=1e-3)):
        m = build_model(cfg)
        m.load_state_dict(syn.init_state_dict(cfg))
        m = m.cuda().train()
        opt = make(m.parameters())
        ls = []
        for _ in range(8):
            opt.zero_grad(set_to_none=True)
            out = m(batch)
            out['loss'].backward()
            opt.step()
            ls.append(out['loss'].item())
        losses.append(ls)
        last = m
    assert losses[0][-1] < losses[0][0]
    for a, b in zip(*losses):
        assert abs(a - b) <= 2e-4 * abs(b), losses
