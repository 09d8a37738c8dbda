"""Committed golden fixtures shared by the CPU and the GPU suites (tests/golden/).

  oracle_search.json   one tree per case: position (move list), config name, explores, seed,
                       tree kind -> child visits, child solutions, root_q bits, root solution,
                       best action, nodes.len()
  oracle_gather.npz    whole self-play games: experience rows (ReplayBuffer layout) + per-row trace

Both are OUTPUTS OF THE ORACLE (tests/golden/make_golden.py is the generating script; the Rust
reference cannot run in this image), after the oracle was pinned against the reference's own
KATs (tests/test_oracle_pinning.py).  `check_*` run any implementation against them — the oracle
on CPU (drift guard) and the CUDA engine on the GPU (parity against committed vectors).
"""
import json
import os

import numpy as np

import synthesis_b200 as s
from synthesis_b200 import _lib as L

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEARCH_JSON = os.path.join(GOLD, "oracle_search.json")
GATHER_NPZ = os.path.join(GOLD, "oracle_gather.npz")


def named_cfg(name: str, explores: int) -> "s.RolloutConfig":
    """The search configurations the fixtures cover, by name."""
    cfg = s.study_connect4_rollout_cfg(num_explores=explores)
    m = cfg.mcts_cfg
    if name == "puct":  # study-connect4/src/main.rs:58-66
        pass
    elif name == "uct_q":  # the evaluator's rollout baseline (main.rs:74-82) + ActionSelection::Q
        cfg.mcts_cfg = s.study_connect4_rollout_mcts_cfg()
        cfg.action = s.ActionSelection.Q
    elif name == "uct":
        cfg.mcts_cfg = s.study_connect4_rollout_mcts_cfg()
    elif name == "no_solve":
        m.solve = False
    elif name == "no_correct":
        m.correct_values_on_solve = False
    elif name == "no_select_solved":
        m.select_solved_nodes = False
    elif name == "no_auto_extend":
        m.auto_extend = False
    elif name == "parent_q":
        m.fpu = s.Fpu.ParentQ()
    elif name == "equal_noise":
        m.root_policy_noise = s.PolicyNoise.Equal(0.25)
    elif name == "dirichlet":
        m.root_policy_noise = s.PolicyNoise.Dirichlet(1.0, 0.25)
    elif name == "fpu_normal":
        m.fpu = s.Fpu.Normal(1.0, 0.1)
    else:
        raise KeyError(name)
    return cfg


def game_from_moves(moves):
    g = s.Connect4.new()
    for m in moves:
        assert not g.step(int(m))
    return g


def f32_bits(a):
    return [int(x) for x in np.asarray(a, np.float32).view(np.uint32).reshape(-1)]


def search_case_outputs(out):
    return dict(child_visits=[int(x) for x in out["child_visits"]], child_solution=[int(x) for x in out["child_solution"]],
                root_q_bits=f32_bits(out["root_q"]), root_solution=int(out["root_solution"]), best_action=int(out["best_action"]),
                num_nodes=int(out["num_nodes"]))


def load_search_cases():
    with open(SEARCH_JSON) as f:
        return json.load(f)["cases"]


def check_search_fixture(search_fn, kinds=(L.TREE_MCTS, L.TREE_FROZEN)):
    """search_fn(cfg, leaf_kind, game, seed, tree_kind) -> dict like Engine.search()[0] for ONE position."""
    n = 0
    for c in load_search_cases():
        if c["tree_kind"] not in kinds:
            continue
        cfg = named_cfg(c["cfg"], c["explores"])
        got = search_case_outputs(search_fn(cfg, L.LEAF_ROLLOUT, game_from_moves(c["moves"]), c["seed"], c["tree_kind"]))
        for k, v in c["out"].items():
            assert got[k] == v, f"search fixture {c['cfg']} E={c['explores']} seed={c['seed']} moves={c['moves']} kind={c['tree_kind']}: {k} {got[k]} != {v}"
        n += 1
    assert n > 0
    return n


GATHER_CASES = (
    # name, cfg name, explores, sample_actions_until, value target, first game, games, seed
    ("q", "puct", 60, 30, "Q", 0, 6, 0),
    ("qtoz", "puct", 40, 10, "QtoZ", 5, 4, 3),
    ("z_stop", "puct", 150, 6, "Z_stop", 2, 4, 1),
)


def gather_case_cfg(case):
    _, cfgname, explores, sau, vt, _, _, _ = case
    cfg = named_cfg(cfgname, explores)
    cfg.sample_actions_until = sau
    if vt == "QtoZ":
        cfg.value_target = s.ValueTarget.QtoZ(0.25, 0.75)
    elif vt == "Z_stop":
        cfg.value_target = s.ValueTarget.Z()
        cfg.stop_games_when_solved = True
    return cfg


def check_gather_fixture(gather_fn):
    """gather_fn(cfg, leaf_kind, first_game, num_games, seed) -> (experience arrays, trace arrays)."""
    z = np.load(GATHER_NPZ)
    for case in GATHER_CASES:
        name, _, _, _, _, first, games, seed = case
        a, t = gather_fn(gather_case_cfg(case), L.LEAF_ROLLOUT, first, games, seed)
        for k, v in list(a.items()) + [("trace_" + k, v) for k, v in t.items()]:
            want = z[f"{name}.{k}"]
            assert v.shape == want.shape, f"gather fixture {name}: {k} shape {v.shape} != {want.shape}"
            assert v.tobytes() == want.tobytes(), f"gather fixture {name}: {k} differs"
