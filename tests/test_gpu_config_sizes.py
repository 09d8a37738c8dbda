"""Parity at the sizes BASELINE.json's configs state (missing #3): trees of 3,200 and 10,000 explores
(configs[4], study-connect4/src/main.rs:72, synthesis/src/evaluator.rs:65-82) for MCTS and FrozenMCTS, a whole match at
10,000 explores, network-leaf self-play at 800 and 1,600 explores per move (configs[1], [2]) with the GPU's leaf outputs fed
to the oracle, and configs[0] literally: 256 rollout games at 800 explores.  Everything through the C ABI, bit for bit."""
import numpy as np
import pytest

import synthesis_b200 as s
from synthesis_b200 import _lib as L
from test_gpu_parity import _assert_match_equal, _config3, _gpu_leaf_callback, assert_rows_equal, random_positions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_engine():
    eng = s.Engine(device=0, max_games_in_flight=256, max_explores=10000)
    eng.set_group_lanes(1)  # the thread-per-game kernels (see tests/conftest.py); the last test below runs the default mapping too
    yield eng
    eng.close()


def _positions(seed, n):
    rng = np.random.default_rng(seed)
    return [s.Connect4.new()] + random_positions(rng, n - 1, max_plies=36)


@pytest.mark.parametrize("explores", [3200, 10000])
def test_mcts_search_rollout_leaves_at_sweep_sizes(big_engine, oracle, explores):
    games = _positions(explores, 8)
    cfg = s.study_connect4_rollout_cfg(num_explores=explores)
    seeds = np.arange(len(games), dtype=np.uint64) * 7 + 3
    out, stats = big_engine.search(cfg, L.LEAF_ROLLOUT, [g.my_bb for g in games], [g.op_bb for g in games], seeds)
    ccfg = cfg.to_c(L.LEAF_ROLLOUT)
    tot = 0
    for i, g in enumerate(games):
        ref, st = oracle.search(ccfg, g.my_bb, g.op_bb, int(seeds[i]))
        tot += st["explores"]
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (i, out["child_visits"][i], ref["child_visits"])
        assert np.array_equal(out["child_solution"][i], ref["child_solution"]), i
        assert np.array_equal(out["root_q"][i].view(np.uint32), ref["root_q"].view(np.uint32)), i
        assert int(out["best_action"][i]) == ref["best_action"] and int(out["num_nodes"][i]) == ref["num_nodes"], i
    assert stats["explores"] == tot
    assert int(out["child_visits"][0].sum()) == explores  # the empty board is not solved by 10,000 explores: every explore counted


@pytest.mark.parametrize("explores", [3200, 10000])
def test_frozen_mcts_search_at_sweep_sizes(big_engine, oracle, explores):
    """FrozenMCTS::exploit with the evaluator's rollout-baseline config (UCT c=2, FPU inf, no auto-extend, Q; main.rs:74-82)."""
    games = _positions(50 + explores, 8)
    cfg = s.study_connect4_rollout_cfg(num_explores=explores, mcts_cfg=s.study_connect4_rollout_mcts_cfg())
    cfg.action = s.ActionSelection.Q
    seeds = np.arange(len(games), dtype=np.uint64) * 5 + 2
    out, stats = big_engine.search(cfg, L.LEAF_ROLLOUT, [g.my_bb for g in games], [g.op_bb for g in games], seeds, tree_kind=L.TREE_FROZEN)
    ccfg = cfg.to_c(L.LEAF_ROLLOUT)
    for i, g in enumerate(games):
        ref, st = oracle.search(ccfg, g.my_bb, g.op_bb, int(seeds[i]), tree_kind=L.TREE_FROZEN)
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (i, out["child_visits"][i], ref["child_visits"])
        assert np.array_equal(out["root_q"][i].view(np.uint32), ref["root_q"].view(np.uint32)), i
        assert int(out["best_action"][i]) == ref["best_action"] and int(out["num_nodes"][i]) == ref["num_nodes"], i


def test_mcts_search_network_leaves_at_3200(big_engine, oracle):
    """NN-side MCTS::exploit at 3,200 explores: the oracle is fed the GPU's (logits, value) per leaf, the tree must be identical."""
    big_engine.set_weights(s.Connect4Net.new(11).blob())
    games = _positions(99, 3)
    cfg = s.study_connect4_rollout_cfg(num_explores=3200)
    out, _ = big_engine.search(cfg, L.LEAF_NN, [g.my_bb for g in games], [g.op_bb for g in games], np.zeros(len(games), np.uint64))
    cb = _gpu_leaf_callback(big_engine)
    ccfg = cfg.to_c(L.LEAF_NN)
    for i, g in enumerate(games):
        ref, _ = oracle.search(ccfg, g.my_bb, g.op_bb, 0, callback=cb)
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (i, out["child_visits"][i], ref["child_visits"])
        assert np.array_equal(out["root_q"][i].view(np.uint32), ref["root_q"].view(np.uint32)), i
        assert int(out["num_nodes"][i]) == ref["num_nodes"], i


def test_whole_match_at_10000_explores(big_engine, oracle):
    """eval_against_rollout_mcts (evaluator.rs:163-198) at the top of the sweep: NN MCTS::exploit (800 explores) against rollout
    FrozenMCTS::exploit (10,000 explores), both colour assignments; moves, per-move visit counts, tree sizes, result."""
    import synthesis_b200.evaluator as ev
    big_engine.set_weights(s.Connect4Net.new(6).blob())
    nn = ev.Player(L.TREE_MCTS, L.LEAF_NN, 800, s.study_connect4_mcts_cfg(), s.ActionSelection.NumVisits)
    ro = ev.Player(L.TREE_FROZEN, L.LEAF_ROLLOUT, 10000, s.study_connect4_rollout_mcts_cfg(), s.ActionSelection.Q)
    cb = _gpu_leaf_callback(big_engine)
    for players, what in (((nn, ro), "nn first"), ((ro, nn), "rollout first")):
        seeds = np.array([3], np.uint64)
        out, st = big_engine.match(players, seeds)
        ref, _ = oracle.match(players, 3, callback=cb)
        _assert_match_equal(out, 0, ref, what)


@pytest.mark.parametrize("explores,noise", [(800, False), (1600, True)])
def test_gather_network_leaves_at_config_explores(big_engine, oracle, explores, noise):
    """configs[1] (800 explores/move) and configs[2] (1,600 explores/move, Dirichlet noise, sampled actions): whole self-play
    games with network leaves; rows and per-move traces equal the oracle's when it is fed the GPU's leaf outputs."""
    big_engine.set_weights(s.Connect4Net.new(2).blob())
    cfg = _config3(explores=explores) if noise else s.study_connect4_rollout_cfg(num_explores=explores, sample_actions_until=30)
    a, st, tr = big_engine.gather(cfg, L.LEAF_NN, 5, 2, 4, trace=True)
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_NN), 4, 5, 2, callback=_gpu_leaf_callback(big_engine), threads=1)
    assert_rows_equal(tr, rtr, "trace")
    assert_rows_equal(a, ra, "experience")
    assert st["explores"] == rst["explores"] and st["nodes"] == rst["nodes"]


def test_config0_literally_256_rollout_games_at_800_explores(big_engine, oracle):
    """BASELINE.json configs[0]: 256 games, rollout leaves, MCTS-Solver, 800 explores/move — every row and every counter."""
    cfg = s.study_connect4_rollout_cfg(num_explores=800, sample_actions_until=30)
    a, st, tr = big_engine.gather(cfg, L.LEAF_ROLLOUT, 0, 256, 0, trace=True)
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_ROLLOUT), 0, 0, 256, threads=16)
    assert_rows_equal(tr, rtr, "trace")
    assert_rows_equal(a, ra, "experience")
    for k in ("explores", "leaf_evals", "rows", "trees", "nodes", "select_levels", "children_scanned", "expansions", "children_created", "backprop_levels", "rollout_plies"):
        assert st[k] == rst[k], (k, st[k], rst[k])
    big_engine.set_group_lanes(0)  # the mapping the engine itself picks for 256 games: a warp per game
    try:
        assert big_engine.launch_geometry(256, L.LEAF_ROLLOUT)[2] == 32
        b, st2, tr2 = big_engine.gather(cfg, L.LEAF_ROLLOUT, 0, 256, 0, trace=True)
    finally:
        big_engine.set_group_lanes(1)
    assert_rows_equal(tr2, rtr, "trace (warp per game)")
    assert_rows_equal(b, ra, "experience (warp per game)")
