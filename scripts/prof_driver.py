"""Small driver for ncu: one gather per leaf mode, sized to keep the profiled kernels short."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthesis_b200 as s
from synthesis_b200 import _lib as L
games = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
explores = int(sys.argv[2]) if len(sys.argv) > 2 else 200
gl = int(sys.argv[3]) if len(sys.argv) > 3 else 32
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["rollout", "nn"]
cfg = s.study_connect4_rollout_cfg(num_explores=explores)
eng = s.Engine(0, 148 * 32, explores)
eng.set_group_lanes(gl)
eng.set_weights(s.Connect4Net.new(0).blob())
for m in modes:
    leaf = L.LEAF_ROLLOUT if m == "rollout" else L.LEAF_NN
    eng.gather_launch(cfg, leaf, 0, games, 0)
    st = eng.gather_wait(None)
    print(m, st["explores"], st["device_ns"] / 1e6, "ms", st["explores"] / st["device_ns"] * 1e3, "M explores/s")
