"""The fixed-point quantisation shortcut of the apply kernel (csrc/kmeans_tc.cu, quant_sum) must give the
same integers as the plain conversion rne(x * 2^40) for every float32 in [-1, 1].  The identity is IEEE
arithmetic only (fp32 fused multiply-add, round to nearest even), so it is checked on the host: 20 M random
bit patterns plus the edge values (ties, denormals, +-1)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_magic_add_quantisation_matches_plain_conversion(tmp_path):
    exe = str(tmp_path / "quant_identity")
    subprocess.check_call(["g++", "-O2", "-x", "c++", os.path.join(HERE, "quant_identity.c"), "-o", exe, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    assert "mismatches: 0" in out.stdout
