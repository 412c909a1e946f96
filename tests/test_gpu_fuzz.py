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


def test_fuzz_with_every_batch_packed():
    """the same fuzzer with the host packer engaged for every batch however small (GDX_PACK_MIN_BYTES=0: 2-bit
    packed kernel, exception queries re-run from their IO bytes, staged uint32 results) -- the switches are read
    once per process, hence the subprocess"""
    import subprocess
    env = dict(os.environ, GDX_PACK_MIN_BYTES="0", GDX_STAGE_MIN_BYTES="0")
    code = ("import sys; sys.path.insert(0, %r); import fuzz_parity\n"
            "for seed in range(7_100_000, 7_100_030): fuzz_parity.one_case(seed)\nprint('ok')" % os.path.dirname(fuzz_parity_path()))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def fuzz_parity_path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_parity.py")
