"""Greedy kv-cache decode on the GPU against the oracle's restatement of the upstream loop over the pinned oracle
model (BASELINE.json config 5 / north_star: "bit-exact argmax token ids for greedy decode").

Random-init models have logit margins of ~0.01, far below bf16 noise, so -- as SURVEY.md section 8(d) prescribes -- the
token embedding is sharpened (scaled) and the minimum top-1/top-2 margin seen by the oracle is reported.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(scale):
    import olmoasr_b200 as ob
    from olmoasr_b200 import synthetic
    from olmoasr_b200.inf_model import OLMoASR
    from oracle import logmel
    from oracle import model as OM

    dims = OM.variant_dims("tiny")
    sd = OM.init_state_dict(dims, 0, train=False)
    g = torch.Generator().manual_seed(7)
    sd["decoder.positional_embedding"] = torch.randn(448, 384, generator=g) * 0.01
    sd["decoder.token_embedding.weight"] = sd["decoder.token_embedding.weight"] * scale
    m = OLMoASR(ob.VARIANT_TO_DIMS["tiny"])
    m.load_state_dict(sd)
    mel = torch.from_numpy(logmel.log_mel_spectrogram(synthetic.waveforms(3).numpy()))
    return m.cuda(), sd, dims, mel


def test_greedy_token_ids_match_oracle():
    from olmoasr_b200.decoding import DecodingOptions, decode
    from oracle import decoding as OD

    m, sd, dims, mel = _make(scale=6.0)
    n_steps = 16
    with torch.no_grad():
        want, margins = OD.greedy_decode(sd, dims, mel, sample_len=n_steps, dtype=torch.bfloat16, return_margins=True)
    # fp16=False: bf16 decode, the dtype the CPU oracle loop runs in (fp16 is covered against the reference's GPU path in
    # tests/test_decode_engine_gpu.py)
    res = decode(m, mel.cuda(), DecodingOptions(language="en", without_timestamps=True, sample_len=n_steps, fp16=False))
    got = [r.tokens for r in res]
    print("min top-1/top-2 margin (oracle, bf16):", min(margins), "tokens:", got[0][:8])
    assert got == want
    assert all(isinstance(r.avg_logprob, float) and r.avg_logprob <= 0 for r in res)
    assert all(0.0 <= r.no_speech_prob <= 1.0 for r in res)
    one = decode(m, mel[0].cuda(), DecodingOptions(without_timestamps=True, sample_len=4, fp16=False))
    assert one.tokens == want[0][:4]


def test_generic_hook_loop_matches_the_engine(monkeypatch):
    """The hook-based loop (sampling / timestamp rules / third-party callers) and the device-resident engine give the
    same greedy ids: the static-buffer kv-cache store behaves like the reference's growing dict."""
    from olmoasr_b200 import decoding as D

    m, sd, dims, mel = _make(scale=6.0)
    opts = D.DecodingOptions(language="en", without_timestamps=True, sample_len=10, fp16=False)
    a = [r.tokens for r in D.decode(m, mel.cuda(), opts)]
    monkeypatch.setattr(D, "USE_ENGINE", False)
    b = [r.tokens for r in D.decode(m, mel.cuda(), opts)]
    assert a == b


def test_kv_cache_steps_equal_full_reforward():
    """SURVEY.md section 4 invariant: cached single-token steps reproduce the full-prefix forward."""
    m, sd, dims, mel = _make(scale=1.0)
    from olmoasr_b200 import synthetic
    ti, *_ = synthetic.text_batch(3)
    ti = ti[:, :6].cuda()
    with torch.no_grad():
        xa = m.encoder(mel.cuda())
        full = m.decoder(ti, xa)
        cache, hooks = m.install_kv_cache_hooks()
        s1 = m.decoder(ti[:, :3], xa, kv_cache=cache)
        s2 = m.decoder(ti[:, 3:4], xa, kv_cache=cache)
        s3 = m.decoder(ti[:, 4:5], xa, kv_cache=cache)
        for h in hooks:
            h.remove()
    assert len(cache) == 4 * dims.n_text_layer
    assert cache[m.decoder.blocks[0].attn.key].shape == (3, 5, 384)
    assert cache[m.decoder.blocks[0].attn.key].data_ptr() == cache[m.decoder.blocks[0].attn.key][:, :1].data_ptr()
    assert cache[m.decoder.blocks[0].attn.key].stride(0) == 448 * 384      # a view of the pre-allocated (B, 448, d) buffer: no torch.cat
    assert cache[m.decoder.blocks[0].cross_attn.key].shape == (3, 1500, 384)   # stored once, never concatenated
    ref = full.float()
    for got, sl in ((s1, slice(0, 3)), (s2, slice(3, 4)), (s3, slice(4, 5))):
        rel = (got.float() - ref[:, sl]).norm() / ref[:, sl].norm()
        assert rel < 5e-3, float(rel)


def test_inference_forward_matches_reference_golden(golden_dir):
    g = torch.load(golden_dir / "inf_tiny.pt", weights_only=False)   # outputs of the unmodified inf_model.py (fp32)
    m, sd, dims, mel = _make(scale=1.0)
    from olmoasr_b200 import synthetic
    ti, *_ = synthetic.text_batch(2)
    with torch.no_grad():
        full = m(mel[:2].cuda(), ti[:, :20].cuda()).cpu()
    rel = (full[:, ::4, ::997] - g["full_logits_sample"]).norm() / g["full_logits_sample"].norm()
    assert rel < 2e-2, float(rel)   # bf16 path vs the reference's fp32 run
