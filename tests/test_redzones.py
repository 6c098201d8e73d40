"""Memory safety without compute-sanitizer: the parity workloads with every device table between red zones (NS_REDZONE=1,
ns_dev_check / ns_ens_check), in a separate process because the switch is read once per process."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_no_table_is_written_out_of_bounds(built):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "redzone_run.py")], capture_output=True, text=True, timeout=1200)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    assert "violations: []" in p.stdout
