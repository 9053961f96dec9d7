"""Synthetic inputs live in the product package (pure data generation); re-exported for the oracle-side tests."""
from olmoasr_b200.synthetic import *  # noqa: F401,F403
from olmoasr_b200.synthetic import text_batch, waveforms  # noqa: F401
