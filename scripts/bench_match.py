"""BASELINE.json configs[4] at scale: the evaluator's sweep (evaluator.rs:65-82) — NN MCTS::exploit (PUCT c=3, FPU 1.0,
policy_num_explores) against rollout FrozenMCTS::exploit (UCT c=2, FPU inf, no auto-extend) at explores in
{100 .. 10000}, both colours, many seeds — as two syn_engine_match launches, a thread per match, next to the oracle
playing a bounded sample of the same matches on the host cores.

usage: bench_match.py [seeds_per_setting] [policy_explores] [cpu_matches]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import synthesis_b200 as s
import synthesis_b200.evaluator as ev
from synthesis_b200 import _lib as L

seeds_per = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
pol_explores = int(sys.argv[2]) if len(sys.argv) > 2 else 800
cpu_matches = int(sys.argv[3]) if len(sys.argv) > 3 else 32
sweep = [100, 200, 400, 800, 1600, 3200, 6400, 10000]
cfg = s.EvaluationConfig(policy_num_explores=pol_explores, policy_action=s.ActionSelection.NumVisits, policy_mcts_cfg=s.study_connect4_mcts_cfg(),
                         rollout_action=s.ActionSelection.Q, rollout_num_explores=sweep, rollout_mcts_cfg=s.study_connect4_rollout_mcts_cfg(),
                         num_games_against_rollout=seeds_per)
net = s.Connect4Net.new(0)
n = len(sweep) * seeds_per
eng = s.Engine(0, min(n, 148 * 128), max(sweep))
ex = [e for e in sweep for _ in range(seeds_per)]
sd = [k for _ in sweep for k in range(seeds_per)]
ev.eval_against_rollout_mcts(eng, cfg, net, ev.RED, ex[:256], sd[:256])  # warm-up
tot = {"explores": 0, "device_ns": 0, "rows": 0}
res = {}
t0 = time.perf_counter()
for colour in (ev.RED, ev.BLACK):
    r, out, st = ev.eval_against_rollout_mcts(eng, cfg, net, colour, ex, sd, trace=True)
    res[colour] = r
    for k in tot:
        tot[k] += st[k] if k in st else 0
    tot["rows"] += int(out["n_moves"].sum())
wall = time.perf_counter() - t0
line = {"op": "evaluation sweep (configs[4])", "matches": 2 * n, "policy_explores": pol_explores, "rollout_explores": sweep,
        "explores": tot["explores"], "moves": tot["rows"], "device_s": tot["device_ns"] / 1e9, "wall_s": wall,
        "explores_per_s": tot["explores"] / tot["device_ns"] * 1e9, "matches_per_s": 2 * n / wall,
        "policy_score_as_red_by_explores": {str(e): float(np.mean(res[ev.RED][i * seeds_per:(i + 1) * seeds_per])) for i, e in enumerate(sweep)}}
if cpu_matches:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build as oracle_build
    oracle_build.build()
    import oracle_binding
    from concurrent.futures import ThreadPoolExecutor
    orc = oracle_binding.Oracle()
    cores = os.cpu_count() or 1
    players = (ev._policy_player(cfg), ev._rollout_player(cfg, max(sweep)))
    pick = np.linspace(0, n - 1, cpu_matches).astype(int)

    def one(i):
        _, st = orc.match(players, sd[i], explores2=[pol_explores, ex[i]], weights=net.blob())
        return st["explores"]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as pool:
        e = sum(pool.map(one, pick))
    dt = time.perf_counter() - t0
    line.update({"cpu_explores_per_s": e / dt, "cpu_cores": cores, "cpu_sample": "%d matches spread over the sweep, one per thread" % cpu_matches, "cpu_kind": "port"})
print(json.dumps(line))
