"""N > 1 on real GPUs (needs >= 2 devices; skipped otherwise): components sharded over 2 NCCL ranks, merged output
byte-identical to the compiled reference."""
import os

import pytest

import helpers
from test_dist_emu import run_world

pytestmark = pytest.mark.gpu


def test_nccl_two_ranks_multi_comp(gpu, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path)
    text = run_world("cuda", helpers.fixture_path("multi_comp"), "S001#1#chr1", 100, out)
    assert text == open(os.path.join(helpers.GOLDEN, "expected", "multi_comp.vcf")).read()
    forest = os.path.join(out, "forest")
    files = {int(f[:-5]): open(os.path.join(forest, f)).read() for f in os.listdir(forest)}
    assert files == helpers.expected_pvsts("multi_comp")
