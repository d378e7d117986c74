"""Pins oracle/gst_ref.py (GST predictor + VecPretextNormalize processing, row a16 / BASELINE config 3) against
vectors recorded from the UNMODIFIED reference (tools/make_golden_gst.py)."""
import os

import numpy as np

from oracle.gst_ref import PretextWrapperRef, gst_forward, load_params

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_gst_forward_matches_reference():
    p = load_params(os.path.join(GOLD, "gst_params.npz"))
    g = np.load(os.path.join(GOLD, "gst_io.npz"))
    out, mask = gst_forward(p, g["in_traj"], g["in_mask"].astype(np.float32))
    assert np.array_equal(mask.numpy(), g["out_mask"])
    np.testing.assert_allclose(out.numpy(), g["out_traj"], rtol=0, atol=2e-5)


def test_wrapper_processing_matches_reference():
    p = load_params(os.path.join(GOLD, "gst_params.npz"))
    g = np.load(os.path.join(GOLD, "gst_rollout.npz"))
    T1, N, H = g["raw_spatial_edges"].shape[:3]
    w = PretextWrapperRef(p, N, H)
    for t in range(T1):
        O = {k: g["raw_" + k][t] for k in ("robot_node", "temporal_edges", "spatial_edges", "detected_human_num", "visible_masks")}
        rews = g["reward_env"][t - 1] if t > 0 else None
        obs, r, pen = w.process(O, rews)
        fin = g["fin_spatial_edges"][t]
        # rows of humans at identical distance (all the unseen ones sit at (15, 15)) may be permuted
        np.testing.assert_allclose(obs["spatial_edges"], fin, rtol=0, atol=2e-4, err_msg="t=%d" % t)
        if t > 0:
            np.testing.assert_allclose(r.reshape(N), g["reward"][t - 1], rtol=0, atol=1e-6)
