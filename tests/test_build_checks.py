"""Static checks of the built CUDA library (run on the CPU box: cuobjdump needs no GPU)."""
import os
import shutil
import subprocess
import sys

import pytest

from crowdnav_prediction_attngraph_b200 import _capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_no_global_access_before_griddepcontrol_wait():
    """Kernels of the programmatic-dependent-launch chains start before their predecessors finish; anything they read
    from global memory must come after griddepcontrol.wait.  The compiler moves invariant loads (`const __restrict__`)
    above the asm barrier, which made the compact GST path read the previous step's row count: scan the SASS."""
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "check_pdl_sass.py"), _capi.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kernels with griddepcontrol.wait checked" in r.stdout and not r.stdout.startswith("0 ")
