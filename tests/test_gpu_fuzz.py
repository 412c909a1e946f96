"""A short run of the randomised parity fuzzer (tools/fuzz_parity.py) as part of the GPU suite: random
alphabets, text shapes, sampling rates, lookup depths, storages, construction routes, text section on/off;
cursors / counts / hits / extend / single-query cursors / invalid-symbol behaviour against the oracle."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("block", range(4))
def test_fuzz_block(block):
    import fuzz_parity
    for seed in range(7_000_000 + 12 * block, 7_000_000 + 12 * (block + 1)):
        fuzz_parity.one_case(seed)
