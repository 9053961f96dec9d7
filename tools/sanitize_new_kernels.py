"""Small invocations of the kernels added in round 2, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_new_kernels.py`
(the reference has no sanitizer runs; out-of-bounds or racy shared-memory accesses in hand-written kernels are
exactly what parity tests can miss)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    import olmoasr_b200 as ob
    from olmoasr_b200 import _core, audio
    from olmoasr_b200 import kernels as K
    from olmoasr_b200 import synthetic as synth
    from olmoasr_b200.config.model_dims import ModelDimensions
    from olmoasr_b200.inf_model import OLMoASR as InfModel
    from olmoasr_b200.model import OLMoASR
    from olmoasr_b200.optim import FusedAdamW

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    # log-mel: odd length (end reflection inside the padded row), int16 and fp32
    wav = synth.waveforms(2, n_samples=16000 * 3 + 77)
    audio.log_mel_spectrogram(wav.to(dev))
    audio.log_mel_spectrogram(synth.waveforms(1, n_samples=6400, int16=True).to(dev))
    # mask -> key counts
    _, _, pm, _ = synth.text_batch(3)
    _core.kv_len_from_padding_mask(pm.to(dev))
    # slab training step on a narrow model (direct gradients, flat optimizer, shadow write)
    dims = ModelDimensions(80, 1500, 64, 1, 1, 51864, 448, 64, 1, 1)
    m = OLMoASR(dims).to(dev)
    slabs = m.use_slabs()
    opt = FusedAdamW(m.parameters(), slabs=slabs)
    mel = ob.log_mel_spectrogram(synth.waveforms(1).to(dev))
    ti, ty, pmask, _ = (t.to(dev) for t in synth.text_batch(1))
    for _ in range(2):
        loss = m(mel, ti, pmask, targets=ty)
        opt.zero_grad()
        loss.backward()
        opt.step()
    # decode engine: direct (3 sequences) and staged (40 sequences) skinny GEMMs, split and fused attention, sampling
    im = InfModel(dims).to(dev)
    with torch.no_grad():
        im.decoder.positional_embedding.normal_(0, 0.02)
    for dtype in (torch.float16, torch.bfloat16):
        eng = im.decode_engine(dtype)
        for n in (3, 40):
            xa = torch.randn(n, 1500, 64, device=dev).bfloat16()
            eng.greedy(xa, [50257, 50362], 3, suppress=(7, 9))
    torch.cuda.synchronize()
    print("sanitize run complete; loss", float(loss))


if __name__ == "__main__":
    main()
