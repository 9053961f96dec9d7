"""FusedAdamW (grad-norm + clip + AdamW in two launches) against torch's clip_grad_norm_ + AdamW
(scripts/training/train_timestamps.py:1508-1522 with the defaults of :2110-2116)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g).cuda().requires_grad_(True) for s in shapes]


def test_matches_torch_adamw_with_clipping():
    from olmoasr_b200.optim import FusedAdamW

    shapes = [(1024, 384), (384,), (51865, 8), (7,), (3, 5, 3), (100003,)]   # odd sizes: unaligned tails
    ours = _make(shapes, 0)
    ref = [p.detach().clone().requires_grad_(True) for p in ours]
    kw = dict(lr=1.5e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.1)
    o1 = FusedAdamW(ours, max_grad_norm=1.0, **kw)
    o2 = torch.optim.AdamW(ref, **kw)
    for step in range(4):
        g = torch.Generator().manual_seed(100 + step)
        grads = [torch.randn(*s, generator=g).cuda() * (10.0 if step % 2 else 0.01) for s in shapes]   # clipped and unclipped steps
        for p, q, gr in zip(ours, ref, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o2.step()
        o1.step()
        assert abs(o1.grad_norm().item() - total.item()) <= 1e-4 * total.item()
        for p, q in zip(ours, ref):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (step, p.shape, (p - q).abs().max().item())
    st = o1.state[ours[0]]
    assert set(st) >= {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 4.0
    assert torch.allclose(st["exp_avg"], o2.state[ref[0]]["exp_avg"], rtol=1e-4, atol=1e-7)


def test_non_finite_gradients_skip_the_step_and_unscale():
    from olmoasr_b200.optim import FusedAdamW

    ps = _make([(64, 64), (10,)], 1)
    before = [p.detach().clone() for p in ps]
    opt = FusedAdamW(ps, lr=1e-2, max_grad_norm=1.0)
    for p in ps:
        p.grad = torch.ones_like(p)
    ps[1].grad[3] = float("inf")
    opt.step()
    assert opt._found_inf.item() == 1.0
    assert all(torch.equal(a, b) for a, b in zip(ps, before))          # GradScaler semantics: untouched
    # loss-scaled gradients: inv_scale undoes the scale before clipping and the update
    a = _make([(257,)], 2)
    b = [a[0].detach().clone().requires_grad_(True)]
    oa = FusedAdamW(a, lr=1e-2, max_grad_norm=0.0)
    ob = FusedAdamW(b, lr=1e-2, max_grad_norm=0.0)
    g = torch.randn(257, device="cuda")
    a[0].grad = g * 65536.0
    b[0].grad = g.clone()
    oa.step(inv_scale=1.0 / 65536.0)
    ob.step()
    assert torch.allclose(a[0], b[0], rtol=1e-5, atol=1e-7)
