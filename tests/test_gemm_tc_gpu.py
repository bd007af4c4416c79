"""tcgen05 GEMM parity, isolated in a subprocess (a trapped kernel must not poison the other tests)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gemm_tcgen05_parity_all_operand_layouts():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_gemm_tc.py")], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0, "tcgen05 GEMM parity failed"
    assert "ALL OK" in r.stdout
