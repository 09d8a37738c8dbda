"""Committed golden fixtures shared by the CPU and the GPU suites (tests/golden/).

  oracle_search.json   one tree per case: position (move list), config name, explores, seed,
                       tree kind -> child visits, child solutions, root_q bits, root solution,
                       best action, nodes.len()
  oracle_match.json    evaluation matches (evaluator.rs:163-228): players, seed, explores per side ->
                       result, moves, per-move nodes.len() and root child visit counts
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
MATCH_JSON = os.path.join(GOLD, "oracle_match.json")


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


# ---------------------------------------------------------------- evaluation matches (rollout leaves only: no network involved)
def match_player(kind: str, explores: int):
    import synthesis_b200.evaluator as ev
    if kind == "frozen_rollout":  # the evaluator's baseline (main.rs:72-82)
        return ev.Player(L.TREE_FROZEN, L.LEAF_ROLLOUT, explores, s.study_connect4_rollout_mcts_cfg(), s.ActionSelection.Q)
    if kind == "mcts_rollout":  # MCTS::exploit with RolloutPolicy (mcts.rs:111-121)
        return ev.Player(L.TREE_MCTS, L.LEAF_ROLLOUT, explores, s.study_connect4_mcts_cfg(), s.ActionSelection.NumVisits)
    raise KeyError(kind)


MATCH_CASES = (
    # first player kind, second player kind, seed, explores (first, second)
    ("frozen_rollout", "frozen_rollout", 0, (100, 200)),
    ("frozen_rollout", "frozen_rollout", 1, (400, 100)),
    ("frozen_rollout", "frozen_rollout", 2, (800, 800)),
    ("mcts_rollout", "frozen_rollout", 3, (150, 300)),
    ("frozen_rollout", "mcts_rollout", 4, (300, 150)),
    ("mcts_rollout", "mcts_rollout", 5, (64, 64)),
)


def match_case_outputs(out):
    n = int(out["n_moves"])
    return dict(result=float(out["result"]), moves=[int(x) for x in out["moves"][:n]],
                tree_nodes=[int(x) for x in out["tree_nodes"][:n]],
                child_visits=[[int(v) for v in row] for row in out["child_visits"][:n]])


def check_match_fixture(match_fn):
    """match_fn(players, seed, explores2) -> dict(result, n_moves, moves[63], tree_nodes[63], child_visits[63][9]) for ONE match."""
    with open(MATCH_JSON) as f:
        cases = json.load(f)["cases"]
    assert len(cases) == len(MATCH_CASES)
    for want, (k0, k1, seed, ex) in zip(cases, MATCH_CASES):
        got = match_case_outputs(match_fn((match_player(k0, ex[0]), match_player(k1, ex[1])), seed, ex))
        for k, v in want["out"].items():
            assert got[k] == v, f"match fixture {k0} vs {k1} seed={seed} explores={ex}: {k} differs"
    return len(cases)
