"""Small driver for ncu: one gather per leaf mode, sized to keep the profiled kernels short."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthesis_b200 as s
from synthesis_b200 import _lib as L
games = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
explores = int(sys.argv[2]) if len(sys.argv) > 2 else 200
gl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["rollout", "nn"]
in_flight = int(sys.argv[5]) if len(sys.argv) > 5 else 0  # games in flight (default: as many as the GPU seats)
cfg = s.study_connect4_rollout_cfg(num_explores=explores)
if len(sys.argv) > 6 and sys.argv[6] == "normalfpu":  # the shipped first-play urgency, study-connect4/src/main.rs:43-47
    cfg = s.study_connect4_rollout_cfg(num_explores=explores, mcts_cfg=s.study_connect4_mcts_cfg(fpu=s.Fpu.Normal(1.0, 0.1)), sample_actions_until=30)
if len(sys.argv) > 6 and sys.argv[6] == "config2":  # bench.py --config 2: Dirichlet(1.0, 0.25) root noise, sample_actions_until = 30
    m = s.study_connect4_mcts_cfg()
    m.root_policy_noise = s.PolicyNoise.Dirichlet(1.0, 0.25)
    cfg = s.study_connect4_rollout_cfg(num_explores=explores, mcts_cfg=m, sample_actions_until=30)
eng = s.Engine(0, in_flight or max(148 * 32, min(games, 148 * 1024)), explores)
eng.set_group_lanes(gl)
eng.set_weights(s.Connect4Net.new(0).blob())
for m in modes:
    leaf = L.LEAF_ROLLOUT if m == "rollout" else L.LEAF_NN
    eng.gather_launch(cfg, leaf, 0, games, 0)
    st = eng.gather_wait(None)
    print(m, "in flight", in_flight or "max", st["explores"], st["device_ns"] / 1e6, "ms", st["explores"] / st["device_ns"] * 1e3, "M explores/s")
    d = eng.debug_counters()
    if d["x_select"]:  # a -DSYN_LG_PROF build of the lane-group kernels: cycles per round and group
        r = max(1, d["rounds"])
        print("   lane-group clocks per round: advance %.0f (select %.0f expand %.0f end-of-move %.0f) wait %.0f leaf %.0f finish %.0f (backprop %.0f) | total %.0f | groups %d rounds/group %.0f"
              % (d["t_advance"] / r, d["x_select"] / r, d["x_expand"] / r, d["x_eom"] / r, d["t_teamwait"] / r, d["t_mlp"] / r, d["t_finish"] / r, d["x_backprop"] / r,
                 d["t_total"] / r, d["leaves"], r / max(1, d["leaves"])))
    elif gl == 1 and m == "nn":
        tot = max(1, d["t_total"])
        print("   phase share of warp time: advance %.1f%% teamwait %.1f%% mlp %.1f%% finish %.1f%% | rounds/warp %.0f leaves/round/warp %.1f | cycles/round %.0f"
              % (100 * d["t_advance"] / tot, 100 * d["t_teamwait"] / tot, 100 * d["t_mlp"] / tot, 100 * d["t_finish"] / tot,
                 d["rounds"] / (148 * 4 * int(os.environ.get("SYN_TPG_TEAMS", "5"))), d["leaves"] / max(1, d["rounds"]), (d["t_advance"] + d["t_teamwait"] + d["t_mlp"] + d["t_finish"]) / max(1, d["rounds"])))
