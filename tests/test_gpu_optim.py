"""Fused multi-tensor optimizers (pwgb_mt_*) vs the pinned oracle restatement of the reference RAdam
(optimizers/radam.py:27-99) / torch.optim.Adam and of clip_grad_norm_ (bin/train.py:289-293): parameters after
every one of 12 steps (the RAdam rectification switches on at step 6), state_dict layout."""
import copy

import pytest
import torch

from oracle import ref_optim, synth
from oracle.make_golden_optim import SHAPES, STEPS, grads_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    import __graft_entry__

    __graft_entry__.build()
    return torch.device("cuda:0")


@pytest.mark.parametrize("kind,lr,betas,eps,wd,clip", [
    ("radam", 1e-2, (0.9, 0.999), 1e-6, 0.0, None),
    ("radam", 1e-4, (0.9, 0.999), 1e-6, 0.01, 10.0),
    ("adam", 2e-3, (0.5, 0.9), 1e-8, 0.0, None),
    ("adam", 2e-4, (0.5, 0.9), 1e-8, 0.001, 3.0),
])
def test_fused_optimizer_matches_reference(dev, kind, lr, betas, eps, wd, clip):
    from parallelwavegan_b200 import optimizers

    ps_ref = [synth.randn(s, 8000 + i) for i, s in enumerate(SHAPES)]
    ms = [torch.zeros_like(p) for p in ps_ref]
    vs = [torch.zeros_like(p) for p in ps_ref]
    ps = [torch.nn.Parameter(p.clone().to(dev)) for p in ps_ref]
    cls = optimizers.RAdam if kind == "radam" else optimizers.FusedAdam
    opt = cls(ps, lr=lr, betas=betas, eps=eps, weight_decay=wd)
    for t in range(1, STEPS + 1):
        gs = grads_for(t)
        for p, g in zip(ps, gs):
            p.grad = g.clone().to(dev)
        v0 = ps[0]._version
        opt.step(max_grad_norm=clip)
        assert ps[0]._version > v0  # raw-pointer update must still advance autograd's version counter (cache keys)
        if clip:
            norm, c = ref_optim.clip_coef(gs, clip)
            got = opt.last_grad_norm.cpu()
            assert abs(float(got[0]) - float(norm)) <= 1e-5 * float(norm) and abs(float(got[1]) - float(c)) <= 1e-5
            gs = [g * c for g in gs]
        for p, g, m, v in zip(ps_ref, gs, ms, vs):
            (ref_optim.radam_step if kind == "radam" else ref_optim.adam_step)(p, g, m, v, t, lr, betas, eps, wd)
        for i, (p, r) in enumerate(zip(ps, ps_ref)):
            err = float((p.detach().cpu() - r).abs().max())
            assert err <= 3e-6 * max(1.0, float(r.abs().max())), (kind, t, i, err)
    sd = opt.state_dict()
    assert sorted(sd.keys()) == ["param_groups", "state"] and len(sd["state"]) == len(SHAPES)
    assert sorted(sd["state"][0].keys()) == ["exp_avg", "exp_avg_sq", "step"]
    assert int(sd["state"][0]["step"]) == STEPS
    assert float((sd["state"][0]["exp_avg"].cpu() - ms[0]).abs().max()) <= 1e-6
    # round trip into a fresh optimizer, one more step gives identical parameters
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt2 = cls(ps2, lr=lr, betas=betas, eps=eps, weight_decay=wd)
    opt2.load_state_dict(copy.deepcopy(sd))  # as after torch.save / torch.load: load_state_dict itself does not copy tensors
    for p, p2, g in zip(ps, ps2, grads_for(STEPS + 1)):
        p.grad = g.clone().to(dev)
        p2.grad = g.clone().to(dev)
    opt.step(max_grad_norm=clip)
    opt2.step(max_grad_norm=clip)
    for p, p2 in zip(ps, ps2):
        assert torch.equal(p.detach(), p2.detach())


def test_fused_optimizer_skips_parameters_without_grad(dev):
    from parallelwavegan_b200 import optimizers

    a = torch.nn.Parameter(torch.ones(100, device=dev))
    b = torch.nn.Parameter(torch.ones(50, device=dev))
    opt = optimizers.RAdam([a, b], lr=1e-2)
    a.grad = torch.full_like(a, 0.5)
    opt.step()
    assert torch.equal(b.detach(), torch.ones(50, device=dev)) and float(a.detach()[0]) != 1.0
    assert len(opt.state[b]) == 0 and opt.state[a]["step"] == 1
    b.grad = torch.full_like(b, 0.25)
    opt.step()  # a at step 2, b at step 1: bucketed by step count
    assert opt.state[a]["step"] == 2 and opt.state[b]["step"] == 1
