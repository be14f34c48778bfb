"""Device VCF text writer on the B200 (pv_text.cuh): number formatting against the C library, genotype slot / prefix /
order edge cases against the compiled reference.  The byte-exact fixture / LPA / synthetic comparisons of
test_call_gpu.py all go through the same writer."""
import pytest

import text_checks as tc

pytestmark = pytest.mark.gpu


def test_af_and_integers(gpu):
    tc.check_af_and_integers(gpu)


def test_shared_slots_and_prefix_files(gpu, tmp_path):
    tc.check_shared_slots_and_prefix_files(gpu, str(tmp_path))


def test_fd_write(gpu, tmp_path):
    tc.check_fd_write(gpu, str(tmp_path))
