"""Freeze runs of the reference's frame optimiser (OpenPyStruct_FrameOpt_Discrete_Beta.py, executed verbatim on
the OpenSees shim by oracle/reference_loader.run_reference_frame) as tests/golden/frame_goldens.npz.
Needs /root/reference (build container only):  python tests/golden/make_frame_golden.py
Groundwork for SURVEY 8f row 4 (frame optimiser); no product code depends on it yet."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl   # noqa: E402

RUNS = ((2, 400), (4, 250), (11, 250))      # (seed -> frame drawn by the script, epochs cap)


def main():
    out = {}
    for seed, epochs in RUNS:
        ns, tr = rl.run_reference_frame(seed, {"num_epochs  = 5000": f"num_epochs  = {epochs}"})
        key = f"s{seed}"
        out[key + "_shape"] = np.array([ns["num_bays"], ns["num_stories"], epochs, len(ns["loss_history"])], np.int64)
        out[key + "_loss"] = np.array(ns["loss_history"], np.float64)          # total_loss.item() per epoch
        out[key + "_I"] = np.asarray(ns["opt_I"], np.float32)                  # after the last step
        out[key + "_I_trace"] = np.array(tr.I[:len(ns["loss_history"])], np.float64)[[0, 1, -1]]
        out[key + "_M_trace"] = np.array(tr.M[:len(ns["loss_history"])], np.float64)[[0, 1, -1]]
        out[key + "_V_trace"] = np.array(tr.V[:len(ns["loss_history"])], np.float64)[[0, 1, -1]]
        out[key + "_consts"] = np.array([ns[k] for k in ("E", "G", "A", "I0", "alpha_moment", "alpha_shear", "k",
                                                          "lateral_load", "vertical_load", "lr", "tolerance",
                                                          "bay_width", "story_height")], np.float64)
        out[key + "_patience"] = np.array([ns["patience"]], np.int64)
        print(key, out[key + "_shape"], "final loss", out[key + "_loss"][-1])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frame_goldens.npz"), **out)


if __name__ == "__main__":
    main()
