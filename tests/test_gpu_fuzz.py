"""GPU: a short run of the randomised GPU-vs-oracle sweep (tests/gpu_tools/fuzz_gpu.py: ragged / empty streams, terminate ops in the
middle, misaligned op buffers, per-stream context inits, both encoder formulations, every symbol profile and binarization)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu_tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_short(seed):
    assert torch.cuda.is_available()
    import fuzz_gpu
    units = 0
    for it in range(60):
        rng = np.random.default_rng([seed, it])
        units += fuzz_gpu.fuzz_ops(rng) if it % 2 == 0 else fuzz_gpu.fuzz_symbols(rng)
    assert units > 0
