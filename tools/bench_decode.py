"""Greedy kv-cache decode real-time factor (BASELINE.json config 5: small model, inf_model, without_timestamps,
N concurrent 30 s clips, up to 224 steps):  RTF = wall time / (N * 30 s), encoder included.

    python tools/bench_decode.py [--variant small] [--clips 1,8,64] > gpurun_out/decode_rtf.json
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="small")
    ap.add_argument("--clips", default="1,8,64")
    ap.add_argument("--steps", type=int, default=224)
    args = ap.parse_args()
    import olmoasr_b200 as ob
    from olmoasr_b200 import synthetic
    from olmoasr_b200.decoding import DecodingOptions, decode
    from olmoasr_b200.inf_model import OLMoASR

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with torch.device(dev):
        model = OLMoASR(ob.VARIANT_TO_DIMS[args.variant])
        torch.nn.init.normal_(model.decoder.positional_embedding, std=0.01)
    out = []
    for n in [int(x) for x in args.clips.split(",")]:
        wav = synthetic.waveforms(n, int16=True).to(dev)
        opts = DecodingOptions(language="en", without_timestamps=True, sample_len=args.steps, suppress_tokens="-1")
        for _ in range(2):  # warm-up (shadows, allocator)
            decode(model, ob.log_mel_spectrogram(wav), DecodingOptions(without_timestamps=True, sample_len=4))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mel = ob.log_mel_spectrogram(wav)
        res = decode(model, mel, opts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n_tok = sum(len(r.tokens) for r in res)
        out.append({"clips": n, "seconds": dt, "rtf": dt / (n * 30.0), "decoded_tokens": n_tok, "tokens_per_s": n_tok / dt})
        print(json.dumps(out[-1]), flush=True)
    print(json.dumps({"metric": "greedy-decode RTF", "variant": args.variant, "steps": args.steps, "results": out}))


if __name__ == "__main__":
    main()
