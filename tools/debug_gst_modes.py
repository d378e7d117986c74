"""Debug aid: run the GST predictor kernels in two CN_GST_MODE settings on the same random observation stream and
report the first difference (python tools/debug_gst_modes.py tcc tc)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import ctypes as C
import numpy as np
import torch
from crowdnav_prediction_attngraph_b200 import _capi

def make(mode, N, H):
    os.environ["CN_GST_MODE"] = mode
    lib = _capi.load_library()
    h = C.c_void_p()
    _capi.check(lib, lib.cn_gst_create(N, H, 5, 0.3, 0.3, -20.0, 0, C.byref(h)), "create")
    p = np.load(os.path.join(REPO, "tests", "golden", "gst_params.npz"))
    for k in p.files:
        a = np.ascontiguousarray(p[k], dtype=np.float32)
        _capi.check(lib, lib.cn_gst_set_param(h, k.encode(), a.ctypes.data, a.size), k)
    _capi.check(lib, lib.cn_gst_finalize(h), "finalize")
    _capi.check(lib, lib.cn_gst_reset(h, None), "reset")
    return lib, h

N, H = (int(sys.argv[3]) if len(sys.argv) > 3 else 64), 20
a, b = sys.argv[1], sys.argv[2]
la, ha = make(a, N, H)
lb, hb = make(b, N, H)
rng = np.random.RandomState(0)
pos = rng.uniform(-6, 6, (N, H, 2)).astype(np.float32)
vel = rng.uniform(-0.25, 0.25, (N, H, 2)).astype(np.float32)
vis_p = rng.uniform(0.1, 0.9, (N, 1))
hist = []
for t in range(40):
    pos = pos + vel
    robot = np.zeros((N, 7), np.float32); robot[:, :2] = rng.uniform(-3, 3, (N, 2))
    sp2 = (pos - robot[:, None, :2]).astype(np.float32)
    vis = (rng.uniform(size=(N, H)) < vis_p).astype(np.uint8)
    hist.append(vis.copy())
    outs = []
    for lib, h in ((la, ha), (lb, hb)):
        r = torch.tensor(robot, device="cuda"); s = torch.tensor(sp2, device="cuda"); v = torch.tensor(vis, device="cuda")
        rew = torch.zeros(N, device="cuda"); pen = torch.zeros(N, device="cuda"); out = torch.zeros(N, H, 12, device="cuda")
        _capi.check(lib, lib.cn_gst_step(h, r.data_ptr(), s.data_ptr(), v.data_ptr(), rew.data_ptr(), pen.data_ptr(), out.data_ptr(), None), "step")
        torch.cuda.synchronize()
        outs.append((out.cpu().numpy(), pen.cpu().numpy()))
    d = np.abs(outs[0][0] - outs[1][0])
    dp = np.abs(outs[0][1] - outs[1][1])
    print("t=%d max |d rows| %.3e  max |d pen| %.3e" % (t, d.max(), dp.max()))
    if d.max() > 1e-3 or dp.max() > 1e-3:
        e, n, c = np.unravel_index(d.argmax(), d.shape)
        print(" first big diff at env", e, "sorted row", n, "col", c, outs[0][0][e, n], outs[1][0][e, n])
        print(" vis history of env (last 5 frames):")
        for hv in hist[-5:]:
            print("  ", hv[e])
        break
