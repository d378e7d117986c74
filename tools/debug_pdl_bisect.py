"""Debug aid: bisect which launch of the compact GST step misbehaves under programmatic dependent launch."""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def ok(lo, hi, pdl="1"):
    env = dict(os.environ, CN_PDL_WINDOW="%d:%d" % (lo, hi), CN_PDL=pdl, CN_GST_MODE="tcc")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_gst.py", "-m", "gpu", "-x", "-q", "-k", "lockstep and tcc"],
                       cwd=REPO, env=env, capture_output=True, text=True)
    return r.returncode == 0
print("pdl off:", ok(0, 1 << 30, "0"))
print("pdl on :", ok(0, 1 << 30))
# smallest hi such that window [0, hi) fails
lo, hi = 0, 72
if not ok(0, 0): print("fails even with empty window"); sys.exit()
while hi - lo > 1:
    mid = (lo + hi) // 2
    if ok(0, mid): lo = mid
    else: hi = mid
print("first failing window end:", hi, "(launch index %d gets PDL)" % (hi - 1))
print("only that launch:", ok(hi - 1, hi))
