"""End-to-end parity of the CUDA model against the oracle (the reference's own arithmetic, pinned in
tests/test_oracle_pin.py) on the seeded synthetic batch of SURVEY.md section 8(d).

Tolerances (north_star: "logits/loss within 1e-3 relative in bf16"):
  * loss: |ours - ref_bf16| <= 1e-3 * |ref_bf16|  (ref_bf16 = oracle under torch.autocast(bfloat16));
  * logits: bf16 end-to-end noise of the REFERENCE ITSELF against its fp32 run is ~8e-3 relative L2 (tiny), so the
    elementwise 1e-3 cannot hold between any two bf16 executions; we require ours to be as close to the fp32 oracle
    as the reference's bf16 path is (<= 1.5x its error) and within 1e-2 relative L2 of the bf16 oracle;
  * gradients: relative L2 per tensor against the fp32 oracle.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import olmoasr_b200 as ob
    from olmoasr_b200.model import OLMoASR
    from oracle import logmel, synth
    from oracle import model as OM

    dims_o = OM.variant_dims("tiny")
    sd = OM.init_state_dict(dims_o, seed=0, train=True)
    torch.manual_seed(0)
    m = OLMoASR(ob.VARIANT_TO_DIMS["tiny"])
    assert all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)  # same seed -> the reference's initial weights
    m = m.cuda()
    B = 2
    wav = synth.waveforms(B)
    mel_cpu = torch.from_numpy(logmel.log_mel_spectrogram(wav.numpy()))
    ti, ty, pm, lens = synth.text_batch(B)
    return dict(m=m, sd=sd, dims=dims_o, OM=OM, wav=wav, mel=mel_cpu, ti=ti, ty=ty, pm=pm)


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def test_forward_logits_and_loss(setup, golden_dir):
    s = setup
    OM = s["OM"]
    with torch.no_grad():
        ref32 = OM.model_forward(s["sd"], s["dims"], s["mel"], s["ti"], s["pm"], train_model=True)
        refbf = OM.model_forward(s["sd"], s["dims"], s["mel"], s["ti"], s["pm"], train_model=True, autocast_dtype=torch.bfloat16)
        got = s["m"](s["mel"].cuda(), s["ti"].cuda(), s["pm"].cuda()).cpu()
    assert got.shape == ref32.shape == (2, 448, 51865) and got.dtype == torch.float32
    e_ref = _rel(refbf, ref32)
    e_our = _rel(got, ref32)
    e_pair = _rel(got, refbf)
    print(f"rel-L2 vs fp32 oracle: reference-bf16 {e_ref:.3e}, ours {e_our:.3e}; ours vs reference-bf16 {e_pair:.3e}")
    assert e_our <= 1.5 * e_ref
    assert e_pair <= 1e-2
    loss_ref = OM.token_ce(refbf, s["ty"]).item()
    loss_got = F.cross_entropy(got.view(-1, got.shape[-1]), s["ty"].view(-1), ignore_index=51864).item()
    assert abs(loss_got - loss_ref) <= 1e-3 * abs(loss_ref)
    g = torch.load(golden_dir / "model_tiny.pt", weights_only=False)  # outputs of the unmodified reference
    assert abs(loss_got - g["loss_bf16"]) <= 1e-3 * abs(g["loss_bf16"])
    assert _rel(got[:, ::16, ::997], g["logits_bf16_sample"]) <= 1e-2


def test_fused_loss_and_gradients(setup, golden_dir):
    s = setup
    OM, m = s["OM"], s["m"]
    p = {k: v.clone().requires_grad_(k != "encoder.positional_embedding") for k, v in s["sd"].items()}
    loss_ref = OM.token_ce(OM.model_forward(p, s["dims"], s["mel"], s["ti"], s["pm"], train_model=True), s["ty"])
    loss_ref.backward()
    m.zero_grad(set_to_none=True)
    loss = m.loss(s["mel"].cuda(), s["ti"].cuda(), s["ty"].cuda(), s["pm"].cuda())
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    worst = []
    for k, prm in m.named_parameters():
        gref = p[k].grad
        r = _rel(prm.grad.cpu(), gref)
        worst.append((r, k))
    worst.sort(reverse=True)
    print("largest gradient rel-L2 errors:", [(f"{r:.3e}", k) for r, k in worst[:6]])
    # yardstick: the reference's OWN bf16-autocast gradients against its fp32 gradients (tools/make_golden.py);
    # e.g. cross-attention query weights carry ~11 % bf16 noise at initialisation in the reference itself
    noise = torch.load(golden_dir / "grad_noise_tiny.pt", weights_only=False)
    for r, k in worst:
        assert r <= max(1.5 * noise[k], 2e-2), (k, r, noise[k])
    assert sum(r for r, _ in worst) / len(worst) <= 1.5 * sum(noise.values()) / len(noise)
    assert m.decoder.token_embedding.weight.grad[51864].abs().max().item() <= 1e-6 + m.decoder.token_embedding.weight.grad.abs().max().item()
    # drop-in path (fp32 logits + F.cross_entropy + autograd) gives the same gradients as the fused head
    g_fused = {k: q.grad.clone() for k, q in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    logits = m(s["mel"].cuda(), s["ti"].cuda(), s["pm"].cuda())
    F.cross_entropy(logits.view(-1, logits.shape[-1]), s["ty"].cuda().view(-1), ignore_index=51864).backward()
    for k, q in m.named_parameters():
        assert _rel(q.grad, g_fused[k]) <= 2e-2, k


def test_waveform_to_loss_pipeline(setup):
    """mel computed on the GPU from the waveform == mel computed by the oracle, through to the loss."""
    import olmoasr_b200 as ob
    s = setup
    mel_gpu = ob.log_mel_spectrogram(s["wav"].cuda())
    assert (mel_gpu.cpu() - s["mel"]).abs().max().item() < 1e-4
    with torch.no_grad():
        a = s["m"].loss(mel_gpu, s["ti"].cuda(), s["ty"].cuda(), s["pm"].cuda()).item()
        b = s["m"].loss(s["mel"].cuda(), s["ti"].cuda(), s["ty"].cuda(), s["pm"].cuda()).item()
    assert abs(a - b) <= 1e-3 * abs(b)


def test_no_padding_mask_path_and_error_behaviour(setup):
    s = setup
    OM, m = s["OM"], s["m"]
    with torch.no_grad():
        ref = OM.model_forward(s["sd"], s["dims"], s["mel"], s["ti"][:, :20], None, train_model=True, autocast_dtype=torch.bfloat16)
        got = m(s["mel"].cuda(), s["ti"][:, :20].cuda()).cpu()
    assert _rel(got, ref) <= 1e-2
    with pytest.raises(AssertionError, match="incorrect audio shape"):   # model.py:601
        m.encoder(torch.zeros(1, 80, 2000, device="cuda"))
