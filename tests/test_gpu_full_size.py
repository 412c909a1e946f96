"""BASELINE.json configs at their full sizes (3.1 Gbp DNA incl. the 24-text hg38-shaped index, 500 M
residue protein index): parity against the oracle on a bounded sample plus size-independent
properties on the whole batch (tools/run_configs.py).  Needs a B200-class GPU (>= 150 GB); on smaller
devices, or with GDX_FULL_SIZE=0, the same checks run at 2 % scale."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _scale():
    import torch
    big = torch.cuda.get_device_properties(0).total_memory >= 150 * (1 << 30)
    return 1.0 if big and os.environ.get("GDX_FULL_SIZE", "1") != "0" else 0.02


@pytest.mark.parametrize("config", ["c2d13", "c3", "c4d5"])
def test_config_at_full_size(config):
    import run_configs
    res = list(run_configs.run([config], _scale()))
    assert len(res) == 1 and "parity" in res[0], res
    print(res[0])


def test_repeat_rich_text_at_two_percent():
    """C2r (tools/run_configs.py: a third of the text written by diverged repeat copies) at 2 % scale: intervals stay
    wide for most of a query that falls into a repeat, hundreds of hits per query -- counts and hits against the
    oracle, every sampled query located at its origin.  The full size runs in profiles/r2_configs.jsonl."""
    import run_configs
    res = list(run_configs.run(["c2r"], 0.02))
    assert len(res) == 1 and "parity" in res[0], res
    # (at 2 % scale a repeat family has 2 % of its copies: a few hits per query instead of the 226 of the full size)
    assert res[0]["hits"] > 2 * res[0]["queries"] and res[0]["lf_steps"] > res[0]["queries"]
