"""Host logic of the training path above the kernels, executed on the CPU with tests/cpu_backend.py standing in for the C
ABI: the per-block autograd Functions (operand dictionaries, hand-sequenced backward), slab mode with gradients accumulated
in place by the "kernels", the tied-embedding loss head, the conv stem -- against the fp32 oracle and against each other.
(The kernels themselves are checked on the GPU: tests/test_kernels_gpu.py, tests/test_model_gpu.py.)"""
import copy

import pytest
import torch
import torch.nn.functional as F

from tests import cpu_backend
from oracle import logmel, synth
from oracle import model as OM


def _dims():
    from olmoasr_b200.config.model_dims import ModelDimensions
    return ModelDimensions(80, 1500, 64, 1, 2, 51864, 448, 64, 1, 2)


@pytest.fixture()
def setup(monkeypatch):
    from olmoasr_b200.model import OLMoASR

    cpu_backend.install(monkeypatch)
    torch.manual_seed(0)
    m = OLMoASR(_dims())
    with torch.no_grad():      # non-trivial LayerNorm affine parameters and biases so that their gradients are exercised
        for k, p in m.named_parameters():
            if k.endswith("ln.weight") or "_ln.weight" in k or "ln_post.weight" in k:
                p.add_(0.1 * torch.randn_like(p))
            elif k.endswith(".bias"):
                p.add_(0.05 * torch.randn_like(p))
    mel = torch.from_numpy(logmel.log_mel_spectrogram(synth.waveforms(2).numpy()))
    ti, ty, pm, _ = synth.text_batch(2)
    return m, mel, ti, ty, pm


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def _oracle_grads(m, mel, ti, ty, pm):
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    p = {k: v.requires_grad_(k != "encoder.positional_embedding") for k, v in sd.items()}
    dims = OM.Dims(80, 1500, 64, 1, 2, 51864, 448, 64, 1, 2)
    loss = OM.token_ce(OM.model_forward(p, dims, mel, ti, pm, train_model=True), ty)
    loss.backward()
    return loss.item(), {k: v.grad for k, v in p.items() if v.grad is not None}


def test_plain_mode_matches_the_oracle(setup):
    m, mel, ti, ty, pm = setup
    loss_ref, g_ref = _oracle_grads(m, mel, ti, ty, pm)
    loss = m(mel, ti, pm, targets=ty)
    loss.backward()
    assert abs(loss.item() - loss_ref) <= 2e-3 * abs(loss_ref)
    for k, p in m.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32, k
        assert _rel(p.grad, g_ref[k]) <= 0.12, (k, _rel(p.grad, g_ref[k]))       # bf16 activations against the fp32 oracle
    logits = m(mel, ti, pm)
    assert logits.shape == (2, 448, 51865) and logits.dtype == torch.float32
    assert abs(F.cross_entropy(logits.view(-1, 51865), ty.view(-1), ignore_index=51864).item() - loss_ref) <= 2e-3 * abs(loss_ref)


def test_slab_mode_accumulates_the_same_gradients_in_place(setup):
    m, mel, ti, ty, pm = setup
    plain = copy.deepcopy(m)
    plain(mel, ti, pm, targets=ty).backward()
    sl = m.use_slabs()
    sl.sync_shadows()
    assert sl.direct_grads
    sl.zero_grad()
    loss = m(mel, ti, pm, targets=ty)
    loss.backward()
    for (k, p), q in zip(m.named_parameters(), plain.parameters()):
        assert p.grad.data_ptr() == sl.grad(p).data_ptr(), k                    # still the slab view: nothing was re-allocated
        assert _rel(p.grad, q.grad) <= 2e-3, (k, _rel(p.grad, q.grad))
    # gaps of the bias-less key projections never receive a gradient (the fused [bq; 0; bv] vector must stay exact)
    used = torch.zeros(sl.numel, dtype=torch.bool)
    for p in m.parameters():
        used[sl.offset[id(p)]: sl.offset[id(p)] + p.numel()] = True
    assert float(sl.G[~used].abs().max()) == 0.0
    # a second backward accumulates (gradient accumulation); zero_grad is one memset
    g1 = sl.G.clone()
    m(mel, ti, pm, targets=ty).backward()
    assert _rel(sl.G, 2 * g1) <= 1e-3
    sl.zero_grad()
    assert float(sl.G.abs().max()) == 0.0
    # autograd hand-over mode on the same slabs (what torch DDP needs): identical numbers through AccumulateGrad
    sl.direct_grads = False
    m(mel, ti, pm, targets=ty).backward()
    assert _rel(sl.G, g1) <= 2e-3
    # the fp32-logits path (drop-in signature) with slabs
    sl.direct_grads = True
    sl.zero_grad()
    lg = m(mel, ti, pm)
    F.cross_entropy(lg.view(-1, 51865), ty.view(-1), ignore_index=51864).backward()
    assert _rel(sl.G, g1) <= 2e-2


def test_completion_callbacks_fire_in_slab_layout_order(setup):
    m, mel, ti, ty, pm = setup
    sl = m.use_slabs()
    sl.sync_shadows()
    fired = []
    for ps, mod in m.grad_units():
        if mod is not None:
            mod._bwd_done_cb = (lambda mod=mod: fired.append(mod))
    m(mel, ti, pm, targets=ty).backward()
    want = [mod for _, mod in m.grad_units() if mod is not None]
    assert fired == want                      # decoder blocks (last first), decoder (embedding), encoder blocks, encoder (stem)


def test_bf16_parameters_as_under_fsdp_mixed_precision(setup):
    """FSDP MixedPrecision(param_dtype=bf16) hands the blocks bf16 parameter views: no shadow caches, fp32 gradients are cast
    back by autograd (train_fsdp_timestamps.py:2588-2615)."""
    m, mel, ti, ty, pm = setup
    mb = copy.deepcopy(m).to(torch.bfloat16)
    ref = m(mel, ti, pm, targets=ty).item()
    mb.encoder.positional_embedding.data = mb.encoder.positional_embedding.data.to(torch.bfloat16)
    loss = mb(mel.to(torch.bfloat16), ti, pm.to(torch.bfloat16), targets=ty)
    loss.backward()
    assert abs(loss.item() - ref) <= 2e-2 * abs(ref)
    assert all(p.grad is not None and p.grad.dtype == torch.bfloat16 for p in mb.parameters())
    assert mb.decoder.blocks[0]._shadow.key is None          # nothing cached for transient parameter views
