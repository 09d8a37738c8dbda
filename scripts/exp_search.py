"""Experiment driver: thread-per-game NN search on realistic mid-game roots, one tree per root.

usage: exp_search.py TEAMS EXPLORES WAVES [REPEAT]
  TEAMS     teams of 128 threads per CTA (SYN_TPG_TEAMS), games in flight = 148*128*TEAMS
  EXPLORES  num_explores per tree
  WAVES     roots searched = WAVES * games in flight
Prints explores/s from the engine's CUDA-event time and the per-warp phase clocks.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
teams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
explores = int(sys.argv[2]) if len(sys.argv) > 2 else 800
waves = float(sys.argv[3]) if len(sys.argv) > 3 else 2
repeat = int(sys.argv[4]) if len(sys.argv) > 4 else 2
os.environ["SYN_TPG_TEAMS"] = str(teams)
import numpy as np

import synthesis_b200 as s
from synthesis_b200 import _lib as L

in_flight = 148 * 128 * teams
n = int(in_flight * waves)
with s.Engine(0, 148 * 128, 30) as e0:  # realistic roots: the rows of a small rollout-mode gather
    cfg0 = s.study_connect4_rollout_cfg(num_explores=30, sample_actions_until=30)
    a, st0, _ = e0.gather(cfg0, L.LEAF_ROLLOUT, 0, 4096, 7)
rng = np.random.default_rng(0)
idx = rng.integers(0, len(a["my_bb"]), n)
my, op = a["my_bb"][idx], a["op_bb"][idx]
seeds = np.arange(n, dtype=np.uint64)
cfg = s.study_connect4_rollout_cfg(num_explores=explores)
with s.Engine(0, in_flight, explores) as eng:
    eng.set_weights(s.Connect4Net.new(0).blob())
    for r in range(repeat):
        out, st = eng.search(cfg, L.LEAF_NN, my, op, seeds)
        d = eng.debug_counters()
        tot = max(1, d["t_total"])
        tag = "prof" if os.environ.get("SYN_TPG_PROF") == "1" else "    "
        B = (st["select_levels"] * 20 + st["children_scanned"] * 18 + st["children_created"] * 47 + st["backprop_levels"] * 36 + st["leaf_evals"] * 64) / max(1, st["explores"])
        print(tag + " teams %d E %d roots %d: %.1f M explores/s (%.1f ms) | depth %.2f alg B/explore %.0f | advance %.1f%% teamwait %.1f%% mlp %.1f%% finish %.1f%% | cycles/round %.0f"
              % (teams, explores, n, st["explores"] / st["device_ns"] * 1e3, st["device_ns"] / 1e6, st["select_levels"] / max(1, st["explores"]), B,
                 100 * d["t_advance"] / tot, 100 * d["t_teamwait"] / tot, 100 * d["t_mlp"] / tot, 100 * d["t_finish"] / tot,
                 (d["t_advance"] + d["t_teamwait"] + d["t_mlp"] + d["t_finish"]) / max(1, d["rounds"])), flush=True)
