"""GPU: a short run of the seeded fuzz harness (tools/fuzz_parity.py) — mixed, tiny, corrupted and truncated traffic over
random tile sizes and both execution paths, device vs oracle bit for bit.  The long runs are recorded in profiles/r1_fuzz.md."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_few_seconds(monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_parity
    monkeypatch.setenv("B2_SMALL", "on")
    batches, seeds, msgs = fuzz_parity.run(budget=float(os.environ.get("B2_FUZZ_SECONDS", "8")), base_seed=123000)     # (longer on the CPU emulator)
    assert batches >= 5 and msgs > 5000
