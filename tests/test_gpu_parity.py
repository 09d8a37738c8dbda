"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Integer / index / visit-count work must be bit-exact; network outputs within 1e-3.
"""
import numpy as np
import pytest

import synthesis_b200 as s
from synthesis_b200 import _lib as L

pytestmark = pytest.mark.gpu

EXP_FIELDS = ("game_ids", "my_bb", "op_bb", "height", "player", "states", "pis", "vs")


def assert_rows_equal(a, b, what):
    for k in a:
        assert a[k].shape == b[k].shape, f"{what}: {k} shape {a[k].shape} vs {b[k].shape}"
        if a[k].dtype.kind == "f":
            same = a[k].view(np.uint32) == b[k].view(np.uint32)
        else:
            same = a[k] == b[k]
        if not np.all(same):
            bad = np.argwhere(~same)[0]
            raise AssertionError(f"{what}: {k} differs first at {tuple(bad)}: gpu={a[k][tuple(bad)]!r} oracle={b[k][tuple(bad)]!r}")


def random_positions(rng, n, max_plies=40):
    """Non-terminal positions reached by random legal play (host mirror of Game::step)."""
    out = []
    while len(out) < n:
        g = s.Connect4.new()
        k = int(rng.integers(0, max_plies))
        ok = True
        for _ in range(k):
            acts = list(g.iter_actions())
            if g.step(acts[int(rng.integers(0, len(acts)))]):
                ok = False
                break
        if ok:
            out.append(g)
    return out


# ---------------------------------------------------------------- A1/A2: the game kernel
def test_game_kernel_matches_oracle_on_random_playouts(engine, oracle):
    rng = np.random.default_rng(7)
    lists = []
    for _ in range(4000):
        g = s.Connect4.new()
        ms = []
        for _ in range(int(rng.integers(0, 64))):
            acts = list(g.iter_actions())
            a = acts[int(rng.integers(0, len(acts)))]
            ms.append(a)
            if g.step(a):
                break
        lists.append(ms)
    lists.append([])  # the empty board
    out = engine.play(lists)
    for i, ms in enumerate(lists):
        ref = oracle.c4_play(ms)
        assert int(out["my_bb"][i]) == ref["my_bb"] and int(out["op_bb"][i]) == ref["op_bb"], (i, ms)
        assert np.array_equal(out["height"][i], ref["height"])
        assert int(out["legal_mask"][i]) == ref["legal_mask"]
        assert int(out["status"][i]) == ref["status"], (i, ms, out["status"][i], ref["status"])
        assert np.array_equal(out["features"][i].view(np.uint32), ref["features"].view(np.uint32))


def test_game_kernel_reference_unit_tests(engine):
    # connect4.rs:300-334 (first/second player wins) and :337-447 (the 63-ply draw)
    first = [0, 1, 0, 1, 0, 1, 0]
    second = [0, 1, 2, 1, 2, 1, 2, 1]
    draw = []
    for pair in range(4):
        a, b = 2 * pair, 2 * pair + 1
        draw += [a, b, a, b, b, a, b, a, a, b, a, b, a, b]
    draw += [8] * 7
    out = engine.play([first, first[:-1], second, second[:-1], draw, draw[:-1], [0] * 8])
    assert list(out["status"]) == [3, 0, 3, 0, 1, 0, 255]
    assert int(out["legal_mask"][4]) == 0 and int(out["legal_mask"][5]) == 1 << 8
    assert list(out["height"][4]) == [7] * 9


# ---------------------------------------------------------------- A4-A10: one tree, rollout leaves
@pytest.mark.parametrize("explores", [1, 50, 800])
def test_search_rollout_bit_exact(engine, oracle, explores):
    rng = np.random.default_rng(explores)
    games = [s.Connect4.new()] + random_positions(rng, 47)
    # the survey's solved-root positions
    for ms in ([4, 4, 3, 3], [4, 3, 4, 3, 4, 3]):
        g = s.Connect4.new()
        for m in ms:
            g.step(m)
        games.append(g)
    cfg = s.study_connect4_rollout_cfg(num_explores=explores)
    seeds = np.arange(len(games), dtype=np.uint64) * 3 + 1
    seeds[0] = 0
    out, stats = engine.search(cfg, L.LEAF_ROLLOUT, [g.my_bb for g in games], [g.op_bb for g in games], seeds)
    ccfg = cfg.to_c(L.LEAF_ROLLOUT)
    tot = dict(explores=0, nodes=0, select_levels=0, children_scanned=0, expansions=0, leaf_evals=0, rollout_plies=0, backprop_levels=0)
    for i, g in enumerate(games):
        ref, st = oracle.search(ccfg, g.my_bb, g.op_bb, int(seeds[i]))
        for k in tot:
            tot[k] += st[k]
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (i, out["child_visits"][i], ref["child_visits"])
        assert np.array_equal(out["child_solution"][i], ref["child_solution"]), i
        assert np.array_equal(out["root_q"][i].view(np.uint32), ref["root_q"].view(np.uint32)), i
        assert int(out["root_solution"][i]) == ref["root_solution"], i
        assert int(out["best_action"][i]) == ref["best_action"], i
        assert int(out["num_nodes"][i]) == ref["num_nodes"], i
    for k in tot:
        assert stats[k] == tot[k], (k, stats[k], tot[k])
    if explores == 800:  # SURVEY.md §8(c) vector for the empty board, seed 0
        assert list(out["child_visits"][0]) == [47, 38, 87, 76, 41, 343, 24, 125, 19]
        assert int(out["num_nodes"][0]) == 7210


def test_search_uct_config_bit_exact(engine, oracle):
    """The evaluator's rollout-baseline config (UCT c=2, FPU=+inf, no auto-extend): exercises ln(),
    inf/NaN comparisons and ActionSelection::Q."""
    rng = np.random.default_rng(5)
    games = [s.Connect4.new()] + random_positions(rng, 31)
    cfg = s.study_connect4_rollout_cfg(num_explores=300, mcts_cfg=s.study_connect4_rollout_mcts_cfg())
    cfg.action = s.ActionSelection.Q
    seeds = np.arange(len(games), dtype=np.uint64) + 100
    out, _ = engine.search(cfg, L.LEAF_ROLLOUT, [g.my_bb for g in games], [g.op_bb for g in games], seeds)
    ccfg = cfg.to_c(L.LEAF_ROLLOUT)
    for i, g in enumerate(games):
        ref, _ = oracle.search(ccfg, g.my_bb, g.op_bb, int(seeds[i]))
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (i, out["child_visits"][i], ref["child_visits"])
        assert int(out["best_action"][i]) == ref["best_action"], i
        assert int(out["num_nodes"][i]) == ref["num_nodes"], i


@pytest.mark.parametrize("variant", ["no_solve", "no_correct", "no_select_solved", "parent_q", "equal_noise", "dirichlet", "fpu_normal"])
def test_search_config_variants_bit_exact(engine, oracle, variant):
    rng = np.random.default_rng(11)
    games = [s.Connect4.new()] + random_positions(rng, 23, max_plies=30)
    m = s.study_connect4_mcts_cfg()
    if variant == "no_solve":
        m.solve = False
    elif variant == "no_correct":
        m.correct_values_on_solve = False
    elif variant == "no_select_solved":
        m.select_solved_nodes = False
    elif variant == "parent_q":
        m.fpu = s.Fpu.ParentQ()
    elif variant == "equal_noise":
        m.root_policy_noise = s.PolicyNoise.Equal(0.25)
    elif variant == "dirichlet":
        m.root_policy_noise = s.PolicyNoise.Dirichlet(1.0, 0.25)
    elif variant == "fpu_normal":
        m.fpu = s.Fpu.Normal(1.0, 0.1)
    cfg = s.study_connect4_rollout_cfg(num_explores=200, mcts_cfg=m)
    seeds = np.arange(len(games), dtype=np.uint64) + 9
    out, _ = engine.search(cfg, L.LEAF_ROLLOUT, [g.my_bb for g in games], [g.op_bb for g in games], seeds)
    ccfg = cfg.to_c(L.LEAF_ROLLOUT)
    for i, g in enumerate(games):
        ref, _ = oracle.search(ccfg, g.my_bb, g.op_bb, int(seeds[i]))
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (variant, i, out["child_visits"][i], ref["child_visits"])
        assert np.array_equal(out["child_solution"][i], ref["child_solution"]), (variant, i)
        assert np.array_equal(out["root_q"][i].view(np.uint32), ref["root_q"].view(np.uint32)), (variant, i)
        assert int(out["num_nodes"][i]) == ref["num_nodes"], (variant, i)


# ---------------------------------------------------------------- A11/A12: whole games, rollout leaves
@pytest.mark.parametrize("explores,games,sample_until,vt", [(50, 64, 0, "Q"), (200, 24, 30, "Z"), (800, 8, 30, "QtoZ"), (100, 16, 10, "QZ")])
def test_gather_rollout_bit_exact(engine, oracle, explores, games, sample_until, vt):
    cfg = s.study_connect4_rollout_cfg(num_explores=explores, sample_actions_until=sample_until)
    cfg.value_target = {"Q": s.ValueTarget.Q(), "Z": s.ValueTarget.Z(), "QtoZ": s.ValueTarget.QtoZ(0.25, 0.75),
                        "QZ": s.ValueTarget.QZaverage(0.3)}[vt]
    a, st, tr = engine.gather(cfg, L.LEAF_ROLLOUT, first_game_index=3, num_games=games, seed=0, trace=True)
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_ROLLOUT), 0, 3, games, threads=8)
    assert_rows_equal(tr, rtr, "trace")
    assert_rows_equal(a, ra, "experience")
    for k in ("explores", "leaf_evals", "rows", "games", "trees", "nodes", "select_levels", "children_scanned", "expansions",
              "children_created", "backprop_levels", "rollout_plies"):
        assert st[k] == rst[k], (k, st[k], rst[k])


def test_gather_survey_vectors(engine):
    """g=0..2 at E=50 with random_actions_until=1, sample_actions_until=0."""
    cfg = s.study_connect4_rollout_cfg(num_explores=50, sample_actions_until=0)
    a, st, tr = engine.gather(cfg, L.LEAF_ROLLOUT, 0, 3, 0, trace=True)
    moves = ["743641716112556277605", "542042434480676008355502302", "15462382653550447302"]
    sums = [8130, 10374, 8021]
    off = 0
    for g in range(3):
        n = len(moves[g])
        assert "".join(str(int(x)) for x in tr["action"][off:off + n]) == moves[g]
        assert int(tr["tree_nodes"][off:off + n].sum()) == sums[g]
        assert np.all(a["game_ids"][off:off + n] == g + 1)
        off += n
    assert off == len(a["vs"])
    assert a["vs"][0].tolist() == [np.float32(15) / np.float32(51), 0.0, np.float32(36) / np.float32(51)]


def test_gather_stop_games_when_solved(engine, oracle):
    cfg = s.study_connect4_rollout_cfg(num_explores=400, sample_actions_until=8)
    cfg.stop_games_when_solved = True
    cfg.value_target = s.ValueTarget.Z()
    a, st, tr = engine.gather(cfg, L.LEAF_ROLLOUT, 0, 24, 5, trace=True)
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_ROLLOUT), 5, 0, 24, threads=8)
    assert_rows_equal(tr, rtr, "trace")
    assert_rows_equal(a, ra, "experience")


@pytest.mark.parametrize("leaf", ["rollout", "nn", "nn_split"])
def test_group_lanes_do_not_change_results(engine, leaf):
    """Thread-per-game (1), half-warp-per-game (16) and warp-per-game (32) are three schedules of the
    same serial per-tree algorithm: experience and visit counts must be identical bit for bit — with rollout leaves and
    with network leaves in either tensor-core chain (single fp16: tpg2 / selfplay_nn_team_kernel; split fp16: tpg2s /
    selfplay_nn_team_split_kernel)."""
    cfg = s.study_connect4_rollout_cfg(num_explores=100)
    kind = L.LEAF_ROLLOUT if leaf == "rollout" else L.LEAF_NN
    if leaf != "rollout":
        engine.set_weights(s.Connect4Net.new(3).blob())
        engine.set_mlp_mode(1 if leaf == "nn" else 2)
    res = {}
    try:
        for gl in (32, 16, 1):
            engine.set_group_lanes(gl)
            res[gl] = engine.gather(cfg, kind, 0, 200, 1, trace=True)
    finally:
        engine.set_group_lanes(1)
        engine.set_mlp_mode(3)
    for gl in (16, 1):
        assert_rows_equal(res[gl][0], res[32][0], f"GL{gl} vs GL32 experience")
        assert_rows_equal(res[gl][2], res[32][2], f"GL{gl} vs GL32 trace")
        for k in ("explores", "leaf_evals", "rows", "nodes", "select_levels", "children_scanned", "expansions", "children_created", "backprop_levels",
                  "rollout_plies"):
            assert res[gl][1][k] == res[32][1][k], (gl, k)


@pytest.mark.parametrize("leaf", ["rollout", "nn", "nn_split"])
def test_group_lanes_with_the_shipped_normal_fpu(engine, leaf):
    """study-connect4/src/main.rs:43-47 ships Fpu::Func(Normal(1.0, 0.1)): the closure runs for every unvisited child every
    time its parent is selected through.  The lane-group kernels keep that stream's key and current block in shared memory
    between calls; a thread per game re-derives them.  Same draws, same rows, same counters."""
    m = s.study_connect4_mcts_cfg(fpu=s.Fpu.Normal(1.0, 0.1))
    cfg = s.study_connect4_rollout_cfg(num_explores=120, mcts_cfg=m, sample_actions_until=12)
    kind = L.LEAF_ROLLOUT if leaf == "rollout" else L.LEAF_NN
    if leaf != "rollout":
        engine.set_weights(s.Connect4Net.new(3).blob())
        engine.set_mlp_mode(1 if leaf == "nn" else 2)
    res = {}
    try:
        for gl in (32, 16, 1):
            engine.set_group_lanes(gl)
            res[gl] = engine.gather(cfg, kind, 3, 150, 4, trace=True)
    finally:
        engine.set_group_lanes(1)
        engine.set_mlp_mode(3)
    for gl in (16, 32):
        assert_rows_equal(res[gl][0], res[1][0], f"GL{gl} vs thread per game: experience")
        assert_rows_equal(res[gl][2], res[1][2], f"GL{gl} vs thread per game: trace")
        assert res[gl][1]["explores"] == res[1][1]["explores"] and res[gl][1]["nodes"] == res[1][1]["nodes"]


@pytest.mark.parametrize("threads", [512, 640, 768, 896])
def test_rollout_threads_per_cta_do_not_change_results(threads):
    """selfplay_rollout_tpg2_kernel is instantiated for 1024 (default), 896, 768, 640 and 512 games per CTA (different ring
    sizes, path-table depths and child batches): identical rows, traces and counters."""
    cfg = s.study_connect4_rollout_cfg(num_explores=150, sample_actions_until=12)
    def run():
        with s.Engine(0, 2048, 150) as e:
            return e.gather(cfg, L.LEAF_ROLLOUT, 5, 300, 9, trace=True)
    a = run()
    b = _with_env("SYN_ROLLOUT_THREADS", str(threads), run)
    assert_rows_equal(a[0], b[0], "experience")
    assert_rows_equal(a[2], b[2], "trace")
    for k in ("explores", "leaf_evals", "rows", "nodes", "select_levels", "children_scanned", "expansions", "children_created", "backprop_levels",
              "rollout_plies"):
        assert a[1][k] == b[1][k], k


@pytest.mark.parametrize("teams", [4, 6])
def test_nn_teams_per_cta_do_not_change_results(teams):
    """selfplay_nn_tpg2_kernel runs 5 teams of 128 games per CTA by default (96 registers, three children per trip, a
    10-level path table); 4 teams (128 registers, five per trip, 12 levels) and 6 teams sharing 4 MLP slots must give
    identical rows, traces and counters."""
    cfg = s.study_connect4_rollout_cfg(num_explores=150, sample_actions_until=12)
    blob = s.Connect4Net.new(8).blob()
    def run():
        with s.Engine(0, 2048, 150) as e:
            e.set_weights(blob)
            return e.gather(cfg, L.LEAF_NN, 5, 900, 9, trace=True)
    a = run()
    b = _with_env("SYN_TPG_TEAMS", str(teams), run)
    assert_rows_equal(a[0], b[0], "experience")
    assert_rows_equal(a[2], b[2], "trace")
    for k in ("explores", "leaf_evals", "rows", "nodes", "select_levels", "children_scanned", "expansions", "children_created", "backprop_levels"):
        assert a[1][k] == b[1][k], k


def test_sharding_is_invisible(engine):
    """Games are seeded by their global index: two shards concatenated == one call."""
    cfg = s.study_connect4_rollout_cfg(num_explores=60)
    whole, _, _ = engine.gather(cfg, L.LEAF_ROLLOUT, 0, 48, 2)
    lo, _, _ = engine.gather(cfg, L.LEAF_ROLLOUT, 0, 20, 2)
    hi, _, _ = engine.gather(cfg, L.LEAF_ROLLOUT, 20, 28, 2)
    cat = {k: np.concatenate([lo[k], hi[k]]) for k in whole}
    assert_rows_equal(cat, whole, "shards vs whole")


# ---------------------------------------------------------------- A9: the network
_MLP_MODES = {"split_fp16_tensor_cores": 2, "fp16_tensor_cores": 1, "fp32_cuda_cores": 0}


@pytest.fixture(params=list(_MLP_MODES))
def mlp_mode(request, engine):
    engine.set_mlp_mode(_MLP_MODES[request.param])
    yield request.param
    engine.set_mlp_mode(3)


def test_nn_eval_within_tolerance(engine, oracle, mlp_mode):
    net = s.Connect4Net.new(0)
    engine.set_weights(net.blob())
    rng = np.random.default_rng(3)
    games = [s.Connect4.new()] + random_positions(rng, 999, max_plies=60)
    my = [g.my_bb for g in games]
    op = [g.op_bb for g in games]
    lg, pr = engine.eval(my, op)
    rl, rp = oracle.mlp_eval(net.blob(), my, op)
    # BASELINE.json north_star: within 1e-3 abs/rel of the fp32 forward
    np.testing.assert_allclose(lg, rl, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(pr, rp, rtol=1e-3, atol=1e-3)
    assert np.allclose(pr.sum(1), 1.0, atol=1e-5)
    print(f"{mlp_mode}: max |logit err| = {np.abs(lg - rl).max():.3e}, max |prob err| = {np.abs(pr - rp).max():.3e}")
    if mlp_mode != "fp16_tensor_cores":  # the product forward (split-fp16 operands) is fp32-grade: two orders below the contract
        np.testing.assert_allclose(lg, rl, rtol=1e-5, atol=1e-5)


def test_forward_chain_is_chosen_by_measurement(engine, oracle):
    """Auto mode (the default): random-init weights keep the single-fp16 chain (its measured error on 1,024 reachable
    positions is a few percent of the tolerance); the same network with every weight tripled (logits of tens, the scale
    training reaches) fails the quarter-tolerance test and gets the split chain; either way the outputs meet the
    UNRELAXED 1e-3 abs / rel against the fp32 oracle forward."""
    rng = np.random.default_rng(8)
    games = random_positions(rng, 500, max_plies=60)
    my = np.array([g.my_bb for g in games], np.uint64)
    op = np.array([g.op_bb for g in games], np.uint64)
    seen = {}
    for name, blob in (("init", s.Connect4Net.new(5).blob()), ("tripled", (s.Connect4Net.new(5).blob() * np.float32(3.0)).astype(np.float32))):
        engine.set_weights(blob)
        chain, ratio = engine.mlp_in_use()
        lg, pr = engine.eval(my, op)
        rl, rp = oracle.mlp_eval(blob, my, op)
        print(f"{name}: chain {chain}, fast chain's measured error {ratio:.3f} of the tolerance, max |logit| {np.abs(rl).max():.1f}, max |logit err| {np.abs(lg - rl).max():.2e}")
        np.testing.assert_allclose(lg, rl, rtol=1e-3, atol=1e-3)
        np.testing.assert_allclose(pr, rp, rtol=1e-3, atol=1e-3)
        seen[name] = (chain, ratio)
    assert seen["init"][0] == 1 and 0.0 <= seen["init"][1] <= 0.25
    assert seen["tripled"][0] == 2 and seen["tripled"][1] > 0.25
    engine.set_mlp_mode(2)
    assert engine.mlp_in_use() == (2, -1.0)
    engine.set_mlp_mode(3)


def test_nn_eval_large_weights_and_batch_independence(engine, oracle):
    """Trained-scale weights (up to 4x the init range per weight, logits of tens) at BASELINE.json's UNRELAXED tolerance
    (1e-3 abs / rel against the fp32 forward; round 1 had to loosen it to 2e-3 x max|logit| for the single-fp16 chain), and: a
    position's output must not depend on which tile row / batch it is evaluated in (that is what lets the oracle replay the
    GPU's leaf outputs)."""
    rng = np.random.default_rng(21)
    blob = (s.Connect4Net.new(5).blob() * rng.uniform(0.5, 4.0, size=L.N_WEIGHTS)).astype(np.float32)
    engine.set_weights(blob)
    games = random_positions(rng, 300, max_plies=60)
    my = np.array([g.my_bb for g in games], np.uint64)
    op = np.array([g.op_bb for g in games], np.uint64)
    lg, pr = engine.eval(my, op)
    rl, rp = oracle.mlp_eval(blob, my, op)
    print(f"large weights: max |logit| = {np.abs(rl).max():.1f}, max |logit err| = {np.abs(lg - rl).max():.3e}")
    np.testing.assert_allclose(lg, rl, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(pr, rp, rtol=1e-3, atol=1e-3)
    blob10 = (s.Connect4Net.new(5).blob() * np.float32(3.0)).astype(np.float32)  # every layer 3x: logits ~ 3^5 times the init scale
    engine.set_weights(blob10)
    lg10, _ = engine.eval(my, op)
    rl10, _ = oracle.mlp_eval(blob10, my, op)
    np.testing.assert_allclose(lg10, rl10, rtol=1e-3, atol=1e-3)
    engine.set_weights(blob)
    perm = rng.permutation(len(games))
    lg2, pr2 = engine.eval(my[perm], op[perm])
    assert np.array_equal(lg2.view(np.uint32), lg[perm].view(np.uint32))
    assert np.array_equal(pr2.view(np.uint32), pr[perm].view(np.uint32))
    one_l, one_p = engine.eval(my[:1], op[:1])
    assert np.array_equal(one_l.view(np.uint32), lg[:1].view(np.uint32))


def _gpu_leaf_callback(engine):
    def cb(ctx, my, op, logits, probs):
        lg, pr = engine.eval([my], [op])
        for i in range(9):
            logits[i] = float(lg[0, i])
        for i in range(3):
            probs[i] = float(pr[0, i])
    return cb


def test_search_nn_tree_bit_exact_given_gpu_leaf_outputs(engine, oracle, mlp_mode):
    """Tree logic is bit-exact when the oracle is fed the GPU's (logits, value) per leaf."""
    net = s.Connect4Net.new(1)
    engine.set_weights(net.blob())
    rng = np.random.default_rng(13)
    games = [s.Connect4.new()] + random_positions(rng, 5, max_plies=30)
    cfg = s.study_connect4_rollout_cfg(num_explores=150)
    seeds = np.zeros(len(games), np.uint64)
    out, _ = engine.search(cfg, L.LEAF_NN, [g.my_bb for g in games], [g.op_bb for g in games], seeds)
    ccfg = cfg.to_c(L.LEAF_NN)
    cb = _gpu_leaf_callback(engine)
    for i, g in enumerate(games):
        ref, _ = oracle.search(ccfg, g.my_bb, g.op_bb, 0, callback=cb)
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (i, out["child_visits"][i], ref["child_visits"])
        assert np.array_equal(out["child_solution"][i], ref["child_solution"]), i
        assert np.array_equal(out["root_q"][i].view(np.uint32), ref["root_q"].view(np.uint32)), i
        assert int(out["num_nodes"][i]) == ref["num_nodes"], i


def test_gather_nn_bit_exact_given_gpu_leaf_outputs(engine, oracle, mlp_mode):
    net = s.Connect4Net.new(2)
    engine.set_weights(net.blob())
    cfg = s.study_connect4_rollout_cfg(num_explores=60, sample_actions_until=12)
    a, st, tr = engine.gather(cfg, L.LEAF_NN, 0, 6, 4, trace=True)
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_NN), 4, 0, 6, callback=_gpu_leaf_callback(engine), threads=1)
    assert_rows_equal(tr, rtr, "trace")
    assert_rows_equal(a, ra, "experience")


def test_gather_nn_close_to_fp32_oracle(engine, oracle, mlp_mode):
    """Against the oracle's OWN fp32 forward the trees may differ after a near-tie, but Q values of
    the first ply (identical position, 800 explores) must agree to 1e-3-ish and games must be legal."""
    net = s.Connect4Net.new(0)
    engine.set_weights(net.blob())
    cfg = s.study_connect4_rollout_cfg(num_explores=200)
    a, st, tr = engine.gather(cfg, L.LEAF_NN, 0, 32, 0, trace=True)
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_NN), 0, 0, 32, weights=net.blob(), threads=8)
    first_gpu = a["vs"][np.r_[True, a["game_ids"][1:] != a["game_ids"][:-1]]]
    first_ref = ra["vs"][np.r_[True, ra["game_ids"][1:] != ra["game_ids"][:-1]]]
    np.testing.assert_allclose(first_gpu, first_ref, rtol=1e-3, atol=1e-3)
    n = min(len(a["vs"]), len(ra["vs"]))
    same = np.mean(np.all(a["pis"][:n] == ra["pis"][:n], axis=1))
    print(f"{mlp_mode}: rows identical to the fp32-oracle run: {same:.3f}")
    if mlp_mode != "fp16_tensor_cores":
        # fp32-grade leaves: the search almost never meets a near-tie that 1e-6 of a logit could flip (round 1 only printed this)
        assert same >= 0.99, same
    assert np.allclose(a["pis"].sum(1), 1.0, atol=1e-5)


# ---------------------------------------------------------------- errors
def test_error_behaviour(engine):
    cfg = s.study_connect4_rollout_cfg(num_explores=10)
    cfg.mcts_cfg.fpu = s.Fpu.Func(lambda: 1.0)
    with pytest.raises(L.EngineError) as e:
        engine.gather(cfg, L.LEAF_ROLLOUT, 0, 4, 0)
    assert e.value.code == L.SYN_ERR_UNSUPPORTED
    cfg = s.study_connect4_rollout_cfg(num_explores=100000)
    with pytest.raises(L.EngineError) as e:
        engine.gather(cfg, L.LEAF_ROLLOUT, 0, 4, 0)
    assert e.value.code == L.SYN_ERR_CAPACITY
    fresh = s.Engine(0, 64, 50)
    try:
        with pytest.raises(L.EngineError) as e:
            fresh.gather(s.study_connect4_rollout_cfg(num_explores=10), L.LEAF_NN, 0, 4, 0)
        assert e.value.code == L.SYN_ERR_NO_WEIGHTS
    finally:
        fresh.close()


# ---------------------------------------------------------------- committed golden fixtures (tests/golden/)
def test_engine_matches_committed_search_fixtures(engine):
    import golden_fixtures as G

    def run(cfg, kind, g, seed, tree_kind):
        out, _ = engine.search(cfg, kind, [g.my_bb], [g.op_bb], [seed], tree_kind=tree_kind)
        return {k: v[0] for k, v in out.items()}

    assert G.check_search_fixture(run) >= 90


def test_engine_matches_committed_gather_fixtures(engine):
    import golden_fixtures as G

    def run(cfg, kind, first, n, seed):
        a, _, t = engine.gather(cfg, kind, first, n, seed, trace=True)
        return a, t

    G.check_gather_fixture(run)


# ---------------------------------------------------------------- configs[4]: FrozenMCTS and evaluation matches
@pytest.mark.parametrize("explores", [1, 50, 800])
def test_frozen_search_rollout_bit_exact(engine, oracle, explores):
    """FrozenMCTS (evaluator.rs:299-534) with the evaluator's rollout-baseline config (main.rs:74-82)."""
    rng = np.random.default_rng(100 + explores)
    games = [s.Connect4.new()] + random_positions(rng, 39)
    for ms in ([4, 4, 3, 3], [4, 3, 4, 3, 4, 3]):
        g = s.Connect4.new()
        for m in ms:
            g.step(m)
        games.append(g)
    cfg = s.study_connect4_rollout_cfg(num_explores=explores, mcts_cfg=s.study_connect4_rollout_mcts_cfg())
    cfg.action = s.ActionSelection.Q
    seeds = np.arange(len(games), dtype=np.uint64) * 5 + 2
    out, stats = engine.search(cfg, L.LEAF_ROLLOUT, [g.my_bb for g in games], [g.op_bb for g in games], seeds, tree_kind=L.TREE_FROZEN)
    ccfg = cfg.to_c(L.LEAF_ROLLOUT)
    tot = dict(explores=0, nodes=0, select_levels=0, children_scanned=0, expansions=0, leaf_evals=0, rollout_plies=0, backprop_levels=0)
    for i, g in enumerate(games):
        ref, st = oracle.search(ccfg, g.my_bb, g.op_bb, int(seeds[i]), tree_kind=L.TREE_FROZEN)
        for k in tot:
            tot[k] += st[k]
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (i, out["child_visits"][i], ref["child_visits"])
        assert np.array_equal(out["child_solution"][i], ref["child_solution"]), i
        assert np.array_equal(out["root_q"][i].view(np.uint32), ref["root_q"].view(np.uint32)), i
        assert int(out["root_solution"][i]) == ref["root_solution"], i
        assert int(out["best_action"][i]) == ref["best_action"], i
        assert int(out["num_nodes"][i]) == ref["num_nodes"], i
    for k in tot:
        assert stats[k] == tot[k], (k, stats[k], tot[k])


def test_frozen_search_nn_bit_exact_given_gpu_leaf_outputs(engine, oracle):
    net = s.Connect4Net.new(4)
    engine.set_weights(net.blob())
    rng = np.random.default_rng(17)
    games = [s.Connect4.new()] + random_positions(rng, 5, max_plies=30)
    m = s.study_connect4_rollout_mcts_cfg()
    m.fpu = s.Fpu.Const(1.0)
    cfg = s.study_connect4_rollout_cfg(num_explores=150, mcts_cfg=m)
    out, _ = engine.search(cfg, L.LEAF_NN, [g.my_bb for g in games], [g.op_bb for g in games], np.zeros(len(games), np.uint64),
                           tree_kind=L.TREE_FROZEN)
    ccfg = cfg.to_c(L.LEAF_NN)
    cb = _gpu_leaf_callback(engine)
    for i, g in enumerate(games):
        ref, _ = oracle.search(ccfg, g.my_bb, g.op_bb, 0, tree_kind=L.TREE_FROZEN, callback=cb)
        assert np.array_equal(out["child_visits"][i], ref["child_visits"]), (i, out["child_visits"][i], ref["child_visits"])
        assert np.array_equal(out["root_q"][i].view(np.uint32), ref["root_q"].view(np.uint32)), i
        assert int(out["best_action"][i]) == ref["best_action"], i
        assert int(out["num_nodes"][i]) == ref["num_nodes"], i


def _assert_match_equal(out, i, ref, what):
    n = int(ref["n_moves"])
    assert int(out["n_moves"][i]) == n, (what, i, int(out["n_moves"][i]), n)
    assert list(out["moves"][i][:n]) == list(ref["moves"][:n]), (what, i)
    assert float(out["result"][i]) == ref["result"], (what, i)
    assert np.array_equal(out["tree_nodes"][i][:n], ref["tree_nodes"][:n]), (what, i)
    assert np.array_equal(out["child_visits"][i][:n], ref["child_visits"][:n]), (what, i)


def test_match_rollout_mcts_vs_mcts_bit_exact(engine, oracle):
    """mcts_vs_mcts (evaluator.rs:200-228): both sides rollout FrozenMCTS on ONE shared stream, explores per side."""
    import synthesis_b200.evaluator as ev
    ro = ev.Player(L.TREE_FROZEN, L.LEAF_ROLLOUT, 400, s.study_connect4_rollout_mcts_cfg(), s.ActionSelection.Q)
    seeds = np.arange(12, dtype=np.uint64)
    ex = np.array([[100, 200], [200, 100], [400, 50], [50, 400]] * 3, np.uint32)
    out, st = engine.match((ro, ro), seeds, ex)
    tot = 0
    for i in range(len(seeds)):
        ref, rst = oracle.match((ro, ro), int(seeds[i]), ex[i])
        _assert_match_equal(out, i, ref, "mcts_vs_mcts")
        tot += rst["explores"]
    assert st["explores"] == tot and st["games"] == len(seeds)


def test_match_nn_mcts_vs_rollout_frozen_bit_exact(engine, oracle):
    """eval_against_rollout_mcts (evaluator.rs:163-198), both colours: NN MCTS::exploit (PUCT c=3, Fpu 1.0,
    NumVisits) against rollout FrozenMCTS::exploit (UCT c=2, Fpu inf, Q); the oracle is fed the GPU's
    leaf outputs so every move's visit counts must be identical."""
    import synthesis_b200.evaluator as ev
    net = s.Connect4Net.new(6)
    engine.set_weights(net.blob())
    nn = ev.Player(L.TREE_MCTS, L.LEAF_NN, 120, s.study_connect4_mcts_cfg(), s.ActionSelection.NumVisits)
    ro = ev.Player(L.TREE_FROZEN, L.LEAF_ROLLOUT, 300, s.study_connect4_rollout_mcts_cfg(), s.ActionSelection.Q)
    cb = _gpu_leaf_callback(engine)
    for players, what in (((nn, ro), "nn first"), ((ro, nn), "rollout first")):
        seeds = np.arange(4, dtype=np.uint64) + 7
        out, st = engine.match(players, seeds)
        for i in range(len(seeds)):
            ref, _ = oracle.match(players, int(seeds[i]), callback=cb)
            _assert_match_equal(out, i, ref, what)


def test_match_two_networks_bit_exact(engine, oracle):
    """eval_against_old(p1, p2) with p1 != p2 (evaluator.rs:131-161, called at :87-94): both weight images resident, a
    forward per image in rounds where a team holds leaves of both players.  The oracle is fed each side's GPU leaf
    outputs (a second small engine holds p2), so moves, per-move visit counts and tree sizes must be identical; and the
    result must change sides correctly when the two networks swap colours.
    Two split-fp16 weight images (2 x 130 KB) do not fit one SM's shared memory, so matches between two DIFFERENT networks
    run the single-fp16 chain (mlp mode 1); the leaf outputs fed to the oracle are taken in the same mode."""
    import ctypes
    import synthesis_b200.evaluator as ev
    engine.set_mlp_mode(1)
    try:
        _match_two_networks(engine, oracle, ev, ctypes)
    finally:
        engine.set_mlp_mode(3)


def _match_two_networks(engine, oracle, ev, ctypes):
    p1, p2 = s.Connect4Net.new(21), s.Connect4Net.new(22)
    nn = ev.Player(L.TREE_MCTS, L.LEAF_NN, 150, s.study_connect4_mcts_cfg(), s.ActionSelection.NumVisits)
    mover = ctypes.c_uint32(0)
    with s.Engine(0, 1024, 8) as e2:
        e2.set_mlp_mode(1)
        for first, second, what in ((p1, p2, "p1 first"), (p2, p1, "p2 first")):
            engine.set_weights(first.blob())
            engine.set_opponent_weights(second.blob())
            e2.set_weights(second.blob())
            try:
                out, st = engine.match((nn, nn), np.zeros(3, np.uint64))
            finally:
                engine.set_opponent_weights(None)

            def cb(ctx, my, op, logits, probs):
                lg, pr = (e2 if mover.value else engine).eval([my], [op])
                for i in range(9):
                    logits[i] = float(lg[0, i])
                for i in range(3):
                    probs[i] = float(pr[0, i])
            ref, rst = oracle.match((nn, nn), 0, callback=cb, mover=mover)
            for i in range(3):  # the game is deterministic: every copy is the same game
                _assert_match_equal(out, i, ref, what)
            assert st["explores"] == 3 * rst["explores"]
    # the host mirror: same games through eval_against_old, and one network against itself is unchanged by the detour
    cfg = s.EvaluationConfig(policy_num_explores=150, policy_action=s.ActionSelection.NumVisits, policy_mcts_cfg=s.study_connect4_mcts_cfg(),
                             rollout_action=s.ActionSelection.Q, rollout_num_explores=[100], rollout_mcts_cfg=s.study_connect4_rollout_mcts_cfg(),
                             num_games_against_rollout=1)
    r12, o12, _ = ev.eval_against_old(engine, cfg, p1, p2, trace=True)
    r21, o21, _ = ev.eval_against_old(engine, cfg, p2, p1, trace=True)
    assert np.array_equal(o21["moves"][0][:int(o21["n_moves"][0])], out["moves"][0][:int(out["n_moves"][0])]) and float(r21[0]) == float(out["result"][0])
    r11a = ev.eval_against_old(engine, cfg, p1)
    r11b = ev.eval_against_old(engine, cfg, p1, p1)
    assert float(r11a[0]) == float(r11b[0])
    recs = ev.evaluate_against_old_models(engine, cfg, p1, "model_1.ot", [("model_0.ot", p2)])
    assert recs == [("model_1.ot", "model_0.ot", float(r12[0])), ("model_0.ot", "model_1.ot", float(r21[0]))]


def test_evaluator_sweep_and_pgn(engine):
    """configs[4] through the host mirror of evaluator.rs:65-82: results are +-1/0 and the PGN text is the reference's."""
    import io
    import synthesis_b200.evaluator as ev
    cfg = s.EvaluationConfig(policy_num_explores=100, policy_action=s.ActionSelection.NumVisits, policy_mcts_cfg=s.study_connect4_mcts_cfg(),
                             rollout_action=s.ActionSelection.Q, rollout_num_explores=[100, 200, 400], rollout_mcts_cfg=s.study_connect4_rollout_mcts_cfg(),
                             num_games_against_rollout=2)
    pgn = io.StringIO()
    rec = ev.evaluate_against_rollout_sweep(engine, cfg, s.Connect4Net.new(0), "model_0.ot", pgn)
    assert len(rec) == 12 and all(r in (1.0, 0.0, -1.0) for _, _, r in rec)
    lines = pgn.getvalue().splitlines()
    assert lines[0] == '[White "model_0.ot"]' and lines[1] == '[Black "VanillaMCTS100"]' and lines[2].startswith('[Result "')
    assert lines[4] == '[White "VanillaMCTS100"]' and lines[5] == '[Black "model_0.ot"]'


def test_match_rejects_what_the_reference_panics_on(engine):
    import synthesis_b200.evaluator as ev
    bad = ev.Player(L.TREE_FROZEN, L.LEAF_ROLLOUT, 10, s.study_connect4_mcts_cfg(), s.ActionSelection.Q)  # PUCT in FrozenMCTS
    with pytest.raises(L.EngineError) as e:
        engine.match((bad, bad), [0])
    assert e.value.code == L.SYN_ERR_UNSUPPORTED


def test_engine_matches_committed_match_fixtures(engine):
    import golden_fixtures as G

    def run(players, seed, ex):
        out, _ = engine.match(players, [seed], [list(ex)])
        return {k: v[0] for k, v in out.items()}

    assert G.check_match_fixture(run) == len(G.MATCH_CASES)


# ---------------------------------------------------------------- configs[2]: 1600 explores, Dirichlet root noise, sampled actions
def _config3(explores=1600):
    m = s.study_connect4_mcts_cfg()
    m.root_policy_noise = s.PolicyNoise.Dirichlet(1.0, 0.25)
    return s.study_connect4_rollout_cfg(num_explores=explores, mcts_cfg=m, sample_actions_until=30)


def test_gather_config3_rollout_bit_exact(engine, oracle):
    cfg = _config3()
    a, st, tr = engine.gather(cfg, L.LEAF_ROLLOUT, first_game_index=11, num_games=6, seed=9, trace=True)
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_ROLLOUT), 9, 11, 6, threads=6)
    assert_rows_equal(tr, rtr, "trace")
    assert_rows_equal(a, ra, "experience")
    for k in ("explores", "leaf_evals", "rows", "trees", "nodes", "select_levels", "children_scanned", "expansions", "children_created", "backprop_levels"):
        assert st[k] == rst[k], (k, st[k], rst[k])


def test_gather_config3_nn_bit_exact_given_gpu_leaf_outputs(engine, oracle):
    net = s.Connect4Net.new(8)
    engine.set_weights(net.blob())
    cfg = _config3(explores=400)
    a, st, tr = engine.gather(cfg, L.LEAF_NN, 0, 3, 5, trace=True)
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_NN), 5, 0, 3, callback=_gpu_leaf_callback(engine), threads=1)
    assert_rows_equal(tr, rtr, "trace")
    assert_rows_equal(a, ra, "experience")


def test_gather_many_games_properties(engine):
    """Size-independent properties at a size the oracle cannot replay: every game ends, rows are
    ordered by game then ply, pi rows sum to 1 over legal moves only, value rows are distributions,
    and the whole gather is reproducible bit for bit (a second run with the same seed)."""
    net = s.Connect4Net.new(1)
    engine.set_weights(net.blob())
    cfg = s.study_connect4_rollout_cfg(num_explores=64)
    a, st, _ = engine.gather(cfg, L.LEAF_NN, 0, 4736, 3)
    b, st2, _ = engine.gather(cfg, L.LEAF_NN, 0, 4736, 3)
    assert_rows_equal(a, b, "same seed, second run")
    assert st["games"] == 4736 and st["rows"] == len(a["vs"])
    ids = a["game_ids"].astype(np.int64)
    assert ids[0] == 1 and ids[-1] == 4736 and np.all(np.diff(ids) >= 0) and np.all(np.diff(ids) <= 1)
    stones = np.array([bin(int(x) | int(y)).count("1") for x, y in zip(a["my_bb"][:5000], a["op_bb"][:5000])])
    first = np.r_[True, ids[1:5000] != ids[:4999]]
    assert np.all(stones[first] == 0) and np.all(stones[~first] == stones[np.flatnonzero(~first) - 1] + 1)
    assert np.allclose(a["pis"].sum(1), 1.0, atol=1e-5) and np.allclose(a["vs"].sum(1), 1.0, atol=1e-5)
    full = a["height"] >= 7
    assert np.all(a["pis"][full] == 0.0)


# ---------------------------------------------------------------- backprop by L2 reductions (tpg2.cuh red_stat) vs load / add / store
def _with_env(name, value, fn):
    import os
    old = os.environ.get(name)
    os.environ[name] = value
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.parametrize("leaf", ["nn", "rollout"])
def test_backprop_reductions_equal_load_add_store(engine, leaf):
    """The thread-per-game kernels add a leaf's value into the path's nodes with one REDG.ADD.F32x4 per level; with
    SYN_TPG_NO_RED=1 they load, add and store like the CPU.  Same trees, same rows, same counters."""
    engine.set_weights(s.Connect4Net.new(4).blob())
    cfg = _config3(explores=300)
    kind = L.LEAF_NN if leaf == "nn" else L.LEAF_ROLLOUT
    run = lambda: engine.gather(cfg, kind, 0, 700, 2, trace=True)
    a, st, tr = run()
    b, st2, tr2 = _with_env("SYN_TPG_NO_RED", "1", run)
    assert_rows_equal(a, b, "experience")
    assert_rows_equal(tr, tr2, "trace")
    for k in ("explores", "leaf_evals", "rows", "nodes", "select_levels", "children_scanned", "expansions", "children_created", "backprop_levels"):
        assert st[k] == st2[k], k


def test_backprop_with_subnormal_values_is_bit_exact(engine, oracle):
    """The L2's adder flushes subnormals, the CPU's does not.  A net whose value head yields a subnormal draw probability
    (zero weights, l_5.bias = [.., 0, -95, -60]) must still give the oracle's outcome sums bit for bit: the kernel notices
    the value and keeps that tree on load / add / store (tpg2.cuh red_exact)."""
    p = {}
    for l in range(5):
        i, o = s.policies.LAYER_DIMS[l], s.policies.LAYER_DIMS[l + 1]
        p[f"l_{l + 1}.weight"] = np.zeros((o, i), np.float32)
        p[f"l_{l + 1}.bias"] = np.zeros((o,), np.float32)
    p["l_5.bias"][9:12] = [0.0, -95.0, -60.0]
    net = s.Connect4Net(p)
    engine.set_weights(net.blob())
    _, pr = engine.eval([0], [0])
    assert 0.0 < float(pr[0, 1]) < 1.1754944e-38, pr  # a subnormal probability reaches the tree
    cfg = s.study_connect4_rollout_cfg(num_explores=120)  # ValueTarget::Q: the rows carry the root sums
    a, st, tr = engine.gather(cfg, L.LEAF_NN, 0, 40, 6, trace=True)
    b, st2, tr2 = _with_env("SYN_TPG_NO_RED", "1", lambda: engine.gather(cfg, L.LEAF_NN, 0, 40, 6, trace=True))
    assert_rows_equal(a, b, "experience (reductions vs load/add/store)")
    ra, rst, rtr = oracle.gather(cfg.to_c(L.LEAF_NN), 6, 0, 4, callback=_gpu_leaf_callback(engine), threads=1)
    n = len(ra["vs"])
    assert np.any((a["vs"][:n, 1] > 0) & (a["vs"][:n, 1] < 1.1754944e-38)), "no subnormal Q reached the experience rows"
    for k in a:
        assert a[k][:n].tobytes() == ra[k].tobytes(), k


# ---------------------------------------------------------------- f1: ReplayBuffer::deduplicate (data.rs:196-235)
def _dedup_inputs(rng, n_rows, n_keys, heavy=0):
    """Rows over `n_keys` distinct reachable positions; `heavy` extra rows all hold the empty board (the group every
    self-play game contributes to), targets are arbitrary f32 so that the order of the additions matters."""
    pos = random_positions(rng, n_keys, max_plies=30)
    which = rng.integers(0, n_keys, n_rows)
    my = np.array([pos[k].my_bb for k in which] + [0] * heavy, np.uint64)
    op = np.array([pos[k].op_bb for k in which] + [0] * heavy, np.uint64)
    n = n_rows + heavy
    pis = (rng.random((n, 9), np.float32) * np.float32(3.0)).astype(np.float32)
    vs = rng.standard_normal((n, 3)).astype(np.float32)
    perm = rng.permutation(n)
    return my[perm], op[perm], pis[perm], vs[perm]


@pytest.mark.parametrize("n_rows,n_keys,heavy", [(1, 1, 0), (17, 3, 0), (5000, 5000, 0), (40000, 900, 0), (30000, 200, 9000), (2049, 1, 0)])
def test_deduplicate_bit_exact(engine, oracle, n_rows, n_keys, heavy):
    """Every merged row equals the oracle's bit for bit (sums in buffer order, f32), groups in order of first occurrence;
    sizes cover one row, one group, all rows distinct, a group larger than the per-CTA threshold (dedup.cuh BIG) and
    multi-tile sorts."""
    rng = np.random.default_rng(n_rows * 31 + n_keys)
    my, op, pis, vs = _dedup_inputs(rng, n_rows, n_keys, heavy)
    got, st = engine.deduplicate(my, op, pis, vs)
    want = oracle.deduplicate(my, op, pis, vs)
    assert st["rows"] == len(want["num"]) and st["kernel_launches"] > 0
    assert_rows_equal(got, want, "deduplicate")
    assert int(got["num"].sum()) == len(my)


def test_deduplicate_of_a_gather_and_edge_cases(engine, oracle):
    """The call the training loop makes (alpha_zero.rs:53-58): deduplicate what gather_experience produced; plus the
    empty buffer, a too-small destination (SYN_ERR_CAPACITY) and the ReplayBuffer mirror."""
    cfg = s.study_connect4_rollout_cfg(num_explores=50, sample_actions_until=10)
    a, _, _ = engine.gather(cfg, L.LEAF_ROLLOUT, 0, 300, 11)
    got, st = engine.deduplicate(a["my_bb"], a["op_bb"], a["pis"], a["vs"])
    want = oracle.deduplicate(a["my_bb"], a["op_bb"], a["pis"], a["vs"])
    assert_rows_equal(got, want, "deduplicate(gather)")
    assert got["num"][0] == 300 and got["my_bb"][0] == 0 and got["op_bb"][0] == 0  # every game starts from Connect4::new()
    assert len(got["num"]) < len(a["vs"])
    buf = s.ReplayBuffer.from_arrays(300, a)
    fb = buf.deduplicate(engine)
    assert fb.states.shape == (len(want["num"]), 1, 7, 9) and fb.pis.tobytes() == want["pis"].tobytes() and fb.vs.tobytes() == want["vs"].tobytes()
    assert fb.states.reshape(-1, 63).tobytes() == want["states"].tobytes()
    empty, st0 = engine.deduplicate(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros((0, 9), np.float32), np.zeros((0, 3), np.float32))
    assert len(empty["num"]) == 0 and st0["rows"] == 0
    import ctypes as C
    fbc = L.SynFlatBatch()
    fbc.capacity = 5
    small = np.zeros((5, 9), np.float32)
    fbc.pis = small.ctypes.data
    rc = engine._lib.syn_engine_deduplicate(engine._h, a["my_bb"].ctypes.data, a["op_bb"].ctypes.data, a["pis"].ctypes.data,
                                            a["vs"].ctypes.data, len(a["vs"]), C.byref(fbc), None)
    assert rc == L.SYN_ERR_CAPACITY and fbc.len == len(want["num"])


def test_deduplicate_large_properties(engine):
    """At a size the oracle is not run on: 4M rows over 300k keys.  Group sizes add up, every group's key is distinct,
    groups come in order of first occurrence, a row-order-independent statistic (num-weighted mean of exactly
    representable targets) is reproduced exactly, and a second run is bit-identical."""
    rng = np.random.default_rng(5)
    n, k = 4_000_000, 300_000
    keys_my = rng.integers(0, 1 << 62, k, dtype=np.uint64)
    keys_op = rng.integers(0, 1 << 62, k, dtype=np.uint64)
    which = rng.integers(0, k, n)
    which[:100_000] = 7  # one very large group
    which = which[rng.permutation(n)]
    my, op = keys_my[which], keys_op[which]
    pis = (rng.integers(0, 16, (n, 9)) / 16.0).astype(np.float32)  # sums of these are exact in f32 in any order
    vs = (rng.integers(0, 16, (n, 3)) / 16.0).astype(np.float32)
    got, st = engine.deduplicate(my, op, pis, vs, states=False)
    again, _ = engine.deduplicate(my, op, pis, vs, states=False)
    for f in ("pis", "vs", "my_bb", "op_bb", "num"):
        assert got[f].tobytes() == again[f].tobytes()
    uniq, first, counts = np.unique(which, return_index=True, return_counts=True)
    order = np.argsort(first)
    assert len(got["num"]) == len(uniq) and np.array_equal(got["num"], counts[order].astype(np.uint32))
    assert np.array_equal(got["my_bb"], keys_my[uniq[order]]) and np.array_equal(got["op_bb"], keys_op[uniq[order]])
    sums = np.zeros((k, 9), np.float64)
    np.add.at(sums, which, pis.astype(np.float64))
    want = (sums[uniq[order]].astype(np.float32) / counts[order].astype(np.float32)[:, None]).astype(np.float32)
    assert got["pis"].tobytes() == want.tobytes()


# ---------------------------------------------------------------- f2: the learner's batch loop (alpha_zero.rs:73-92)
def _training_rows(engine, games=64, explores=40, seed=2):
    cfg = s.study_connect4_rollout_cfg(num_explores=explores, sample_actions_until=20)
    a, _, _ = engine.gather(cfg, L.LEAF_ROLLOUT, 0, games, seed)
    d, _ = engine.deduplicate(a["my_bb"], a["op_bb"], a["pis"], a["vs"])
    return d


@pytest.mark.parametrize("cluster", ["2", "1", "0"])
@pytest.mark.parametrize("weight_decay,pw,vw", [(0.0, 1.0, 1.0), (1e-2, 0.7, 1.3)])
def test_train_matches_torch_fp32(engine, weight_decay, pw, vw, cluster):
    """All three schedules of the learner: the 8-CTA cluster with asynchronous remote stores (default), the cluster with
    cluster.sync() exchanges (train_cluster.cuh) and the single-CTA kernel (train.cuh)."""
    _with_env("SYN_TRAIN_CLUSTER", cluster, lambda: _train_matches_torch_fp32(engine, weight_decay, pw, vw))


def _train_matches_torch_fp32(engine, weight_decay, pw, vw):
    """Floating-point kernel: per-step losses and the weights after 1, 4 and 40 Adam steps against the same steps in
    PyTorch fp32 on the CPU (the ops the reference runs through tch).  Tolerance (north_star): 1e-3 abs/rel; observed
    differences are ~1e-6 (summation order)."""
    from torch_learner import TorchLearner
    d = _training_rows(engine, games=160)
    n = len(d["num"])
    rng = np.random.default_rng(0)
    net = s.Connect4Net.new(4)
    batches = np.concatenate([s.BatchRandSampler(n, 32, True, rng).all_batches() for _ in range(2)])[:40]
    assert len(batches) == 40
    ref = TorchLearner(net.blob(), 1e-3, weight_decay, pw, vw)
    engine.set_weights(net.blob())
    engine.reset_optimizer()
    done = 0
    for upto in (1, 4, 40):
        want_losses = ref.run(d["states"], d["pis"], d["vs"], batches[done:upto])
        got_losses, st = engine.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], batches[done:upto], 1e-3, weight_decay, pw, vw)
        assert st["rows"] == 32 * (upto - done) and st["kernel_launches"] == 2
        assert np.allclose(got_losses, want_losses, rtol=1e-3, atol=1e-3), (upto, np.abs(got_losses - want_losses).max())
        got, want = engine.get_weights(), ref.blob()
        print("train parity after %d steps: max |dw| %.3g, max |dloss| %.3g" % (upto, np.abs(got - want).max(), np.abs(got_losses - want_losses).max()))
        assert np.allclose(got, want, rtol=1e-3, atol=1e-3), (upto, np.abs(got - want).max())
        done = upto
    assert np.abs(engine.get_weights() - net.blob()).max() > 1e-3  # it did move
    # the search kernels see the trained weights without a set_weights call
    lg, pr = engine.eval(d["my_bb"][:64], d["op_bb"][:64])
    import torch
    with torch.no_grad():
        pl, vl = ref.forward(torch.from_numpy(d["states"][:64]))
    # after 40 Adam steps, at the unrelaxed tolerance: the forward sees the GPU-trained weights, torch its own (<= 1e-3 apart)
    assert np.allclose(lg, pl.numpy(), rtol=1e-3, atol=2e-3) and np.allclose(pr, torch.softmax(vl, -1).numpy(), rtol=1e-3, atol=2e-3)
    from oracle_binding import Oracle
    ol, op_ = Oracle().mlp_eval(engine.get_weights(), d["my_bb"][:64], d["op_bb"][:64])  # same weights on both sides
    np.testing.assert_allclose(lg, ol, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(pr, op_, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("batch_size", [64, 160])
def test_train_batch_sizes_other_than_32(engine, batch_size):
    """LearningConfig::batch_size is free in the reference (config.rs:76-94).  The learner's kernels are tiled for 32 rows; a
    multiple of 32 runs as micro-batches whose gradients are summed before ONE Adam update per batch.  Against PyTorch fp32
    with the same batch size: per-batch losses and all weights after 12 steps within 1e-3 abs/rel; other sizes are refused."""
    from torch_learner import TorchLearner
    d = _training_rows(engine, games=160)
    n = len(d["num"])
    rng = np.random.default_rng(1)
    net = s.Connect4Net.new(6)
    batches = np.concatenate([s.BatchRandSampler(n, batch_size, True, rng).all_batches() for _ in range(3)])[:12]
    assert batches.shape == (12, batch_size)
    ref = TorchLearner(net.blob(), 1e-3, 1e-2, 0.7, 1.3, batch_size=batch_size)
    engine.set_weights(net.blob())
    engine.reset_optimizer()
    want_losses = ref.run(d["states"], d["pis"], d["vs"], batches)
    got_losses, st = engine.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], batches, 1e-3, 1e-2, 0.7, 1.3, batch_size=batch_size)
    assert st["rows"] == batch_size * 12
    got, want = engine.get_weights(), ref.blob()
    print("batch %d: max |dw| %.3g, max |dloss| %.3g" % (batch_size, np.abs(got - want).max(), np.abs(got_losses - want_losses).max()))
    assert np.allclose(got_losses, want_losses, rtol=1e-3, atol=1e-3)
    assert np.allclose(got, want, rtol=1e-3, atol=1e-3)
    assert np.abs(got - net.blob()).max() > 1e-3
    with pytest.raises(L.EngineError):  # not a multiple of 32
        engine.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], batches[:, :48], 1e-3, batch_size=48)
    # the two variant kernels take batches of 32 only
    with pytest.raises(L.EngineError):
        _with_env("SYN_TRAIN_CLUSTER", "1", lambda: engine.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], batches, 1e-3, batch_size=batch_size))
    engine.set_weights(net.blob())


def test_train_loss_decreases_and_edge_cases(engine):
    """Twenty epochs over a small FlatBatch through the host mirror of the epoch loop: the KL losses fall; bad
    arguments are refused the way the header says."""
    d = _training_rows(engine, games=48)
    fb = s.FlatBatch(d["states"].reshape(-1, 1, 7, 9), d["pis"], d["vs"], d["my_bb"], d["op_bb"])
    cfg = s.LearningConfig(seed=0, logs="", lr_schedule=[(1, 1e-3)], weight_decay=1e-6, num_iterations=1, num_epochs=20, batch_size=32,
                           policy_weight=1.0, value_weight=1.0, games_to_keep=100, games_per_train=48,
                           rollout_cfg=s.study_connect4_rollout_cfg(num_explores=40))
    engine.set_weights(s.Connect4Net.new(0).blob())
    engine.reset_optimizer()
    epochs = s.train_on(cfg, fb, s.lr_for_iteration(cfg, 0), engine, np.random.default_rng(1))
    assert len(epochs) == 20 and epochs[-1][0] < 0.97 * epochs[0][0] and epochs[-1][1] < 0.5 * epochs[0][1], epochs
    assert np.all(np.isfinite(engine.get_weights()))
    n = len(d["num"])
    with pytest.raises(L.EngineError) as ei:
        engine.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], np.full((2, 32), n, np.uint32), 1e-3)
    assert ei.value.code == L.SYN_ERR_INVALID_ARGUMENT
    with pytest.raises(L.EngineError) as ei:
        engine.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], np.zeros((2, 16), np.uint32), 1e-3, batch_size=16)
    assert ei.value.code == L.SYN_ERR_UNSUPPORTED
    losses, st = engine.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], np.zeros((0, 32), np.uint32), 1e-3)
    assert losses.shape == (0, 2)


def test_alpha_zero_loop_end_to_end(engine, tmp_path):
    """gather -> deduplicate -> train -> gather with the weights never leaving the device: two iterations of the host
    mirror of alpha_zero (alpha_zero.rs:16-118), model files written where the reference writes them and read back."""
    cfg = s.LearningConfig(seed=3, logs=str(tmp_path), lr_schedule=[(1, 1e-3), (2, 5e-4)], weight_decay=1e-6, num_iterations=2, num_epochs=3,
                           batch_size=32, policy_weight=1.0, value_weight=1.0, games_to_keep=96, games_per_train=64,
                           rollout_cfg=s.study_connect4_rollout_cfg(num_explores=32))
    seen = []

    def on_iteration(i, eng, buffer, dedup, epochs):
        s.Connect4Net.from_blob(eng.get_weights()).save_ot(str(tmp_path / f"model_{i + 1}.ot"))
        seen.append((i, buffer.total_games_played(), buffer.curr_games(), len(dedup), epochs))

    net0 = s.Connect4Net.new(cfg.seed)
    net = s.alpha_zero(cfg, net0, engine=engine, on_iteration=on_iteration)
    assert [x[0] for x in seen] == [0, 1] and seen[0][1] == 64 and seen[1][1] == 128
    # keep_last_n_games(96 - 64) keeps ids >= 64 - 32, i.e. 33 games (data.rs:172-194 compares with >=), then extend adds 64
    assert seen[1][2] == 97
    assert all(len(x[4]) == 3 and np.all(np.isfinite(x[4])) for x in seen) and seen[1][3] > seen[0][3] > 64
    assert np.abs(net.blob() - net0.blob()).max() > 1e-3
    assert s.Connect4Net.load_ot(str(tmp_path / "model_2.ot")).blob().tobytes() == net.blob().tobytes()
    # the engine's search weights ARE the trained ones
    lg, _ = engine.eval([0], [0])
    engine.set_weights(net.blob())
    lg2, _ = engine.eval([0], [0])
    assert lg.tobytes() == lg2.tobytes()


def test_alpha_zero_distributed_world_of_one_equals_local(engine):
    """The multi-GPU form of the loop (syn_engine_broadcast_weights -> sharded gather -> syn_engine_gather_experience ->
    dedup + train on the trainer rank) with a communicator of ONE rank reproduces the local loop bit for bit: same games,
    same rows, same batches, same kernels — NCCL collectives of one rank included."""
    from synthesis_b200 import distributed as D
    cfg = s.LearningConfig(seed=5, logs="", lr_schedule=[(1, 1e-3)], weight_decay=0.0, num_iterations=2, num_epochs=2, batch_size=32,
                           policy_weight=1.0, value_weight=1.0, games_to_keep=200, games_per_train=48,
                           rollout_cfg=s.study_connect4_rollout_cfg(num_explores=24))
    local = s.alpha_zero(cfg, s.Connect4Net.new(cfg.seed), engine=engine)
    comm = s.Comm(s.Comm.unique_id(), 1, 0, 0)
    try:
        net = D.alpha_zero_distributed(cfg, engine, comm, policy=s.Connect4Net.new(cfg.seed))
    finally:
        comm.close()
    assert net.blob().tobytes() == local.blob().tobytes()


def test_gather_experience_through_a_communicator_of_one(engine, oracle):
    """syn_engine_gather_experience on a single rank: rows cross the wire format (72 bytes: ids, bitboards, pi, v) and the
    height / player / features columns are rebuilt from the bitboards — every column must equal syn_engine_gather's."""
    cfg = s.study_connect4_rollout_cfg(num_explores=60, sample_actions_until=12)
    a, st, _ = engine.gather(cfg, L.LEAF_ROLLOUT, 7, 40, 3)
    comm = s.Comm(s.Comm.unique_id(), 1, 0, 0)
    try:
        b, st2 = engine.gather_experience(comm, cfg, L.LEAF_ROLLOUT, 7, 40, 3)
        c, _ = engine.gather_experience(comm, cfg, L.LEAF_ROLLOUT, 7, 0, 3)  # an empty shard is allowed
    finally:
        comm.close()
    assert_rows_equal(a, b, "communicator of one vs plain gather")
    assert st["explores"] == st2["explores"] and len(c["vs"]) == 0


def _two_rank_worker(rank, world, id_bytes, q):
    try:
        import numpy as np
        import synthesis_b200 as s
        from synthesis_b200 import _lib as L
        from synthesis_b200 import distributed as D
        cfg = s.study_connect4_rollout_cfg(num_explores=80, sample_actions_until=12)
        comm = s.Comm(id_bytes, world, rank, rank)
        with s.Engine(rank, 2048, 80) as e:
            blob = s.Connect4Net.new(9).blob() if rank == 0 else None
            e.broadcast_weights(comm, blob, root=0)
            w = e.get_weights()
            merged, st = D.gather_experience_distributed(e, comm, cfg, L.LEAF_NN, 301, 5, None, 0, root=0, first_game_index=11)
            out = dict(wsum=float(np.abs(w).sum()), explores=st["explores"], chain=e.mlp_in_use()[0])
            if rank == 0:
                out["merged"] = {k: v.copy() for k, v in merged.items()}
        comm.close()
        q.put((rank, out))
    except Exception as ex:  # pragma: no cover - surfaced in the parent
        import traceback
        q.put((rank, "ERROR " + repr(ex) + "\n" + traceback.format_exc()))


def test_two_gpus_concatenated_shards_equal_one_gpu(engine):
    """301 NN-leaf games sharded over TWO GPUs (150 + 151, alpha_zero.rs:138) through the library's own
    collectives (one NCCL broadcast of the weights, one gather of 72-byte rows to rank 0) == the same games on one GPU,
    every column of every row bit for bit.  Needs two visible GPUs (gpurun --gpus 2); fails loudly otherwise."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    id_bytes = s.Comm.unique_id()
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, id_bytes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r, out = q.get(timeout=600)
        assert not isinstance(out, str), out
        res[r] = out
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    blob = s.Connect4Net.new(9).blob()
    assert res[0]["wsum"] == res[1]["wsum"] == float(np.abs(blob).sum())  # the broadcast reached rank 1
    cfg = s.study_connect4_rollout_cfg(num_explores=80, sample_actions_until=12)
    engine.set_weights(blob)
    # a broadcast measures the forward chains like set_weights does: every rank picks what one GPU picks
    assert res[0]["chain"] == res[1]["chain"] == engine.mlp_in_use()[0] == 1
    whole, st, _ = engine.gather(cfg, L.LEAF_NN, 11, 301, 5)
    assert_rows_equal(res[0]["merged"], whole, "two GPUs vs one")
    assert res[0]["explores"] + res[1]["explores"] == st["explores"]


# ---------------------------------------------------------------- tpg2 (32-byte records, the product kernels) vs tpg4 (family blocks), and seating
_COUNTERS = ("explores", "leaf_evals", "rows", "nodes", "select_levels", "children_scanned", "expansions", "children_created", "backprop_levels")


@pytest.mark.parametrize("leaf", ["nn", "rollout"])
@pytest.mark.parametrize("variant", ["default", "parent_q", "fpu_normal", "uct"])
def test_node_layouts_do_not_change_results(leaf, variant):
    """The product kernels (tpg2.cuh: 32-byte node records, three children per memory trip, q divided at selection time,
    backprop by L2 reductions or, with SYN_TPG_NO_RED=1, by load/add/store) and the family-block layout (tpg4.cuh,
    SYN_TPG_VER=4: select_best_child reads one 128-byte line per level, -child.q() memoised, u16 visit counts, children in
    column slots, backprop that reads what it updates) are layouts of the same serial per-tree algorithm: identical rows,
    per-move visit counts, tree sizes and counters."""
    cfg = _config3(explores=300)
    if variant == "parent_q":
        cfg.mcts_cfg.fpu = s.Fpu.ParentQ()
    elif variant == "fpu_normal":
        cfg.mcts_cfg.fpu = s.Fpu.Normal(1.0, 0.1)
    elif variant == "uct":
        cfg.mcts_cfg.exploration = s.Exploration.Uct(2.0)
        cfg.mcts_cfg.fpu = s.Fpu.Const(float("inf"))
        cfg.mcts_cfg.auto_extend = False
    kind = L.LEAF_NN if leaf == "nn" else L.LEAF_ROLLOUT
    blob = s.Connect4Net.new(4).blob()

    def run():
        with s.Engine(0, 2048, 300) as e:
            e.set_weights(blob)
            return e.gather(cfg, kind, 3, 700, 2, trace=True)
    a = run()
    b = _with_env("SYN_TPG_NO_RED", "1", run)
    c = _with_env("SYN_TPG_VER", "4", run)
    for other, name in ((b, "tpg2 load/add/store"), (c, "tpg4")):
        assert_rows_equal(a[0], other[0], f"experience vs {name}")
        assert_rows_equal(a[2], other[2], f"trace vs {name}")
        for k in _COUNTERS:
            assert a[1][k] == other[1][k], (name, k)


@pytest.mark.parametrize("leaf", ["nn", "rollout"])
@pytest.mark.parametrize("in_flight", [1, 37, 1000, 4096])
def test_games_in_flight_do_not_change_results(leaf, in_flight):
    """A launch that holds fewer games than the GPU has thread slots seats them over ALL SMs (tp2::seat_of): 1000 games
    (the reference's games_per_train, study-connect4/src/main.rs:26) run as 6-7 games on each of 148 SMs instead of 1000
    on two.  Results do not depend on the seating: games are seeded by their global index."""
    cfg = s.study_connect4_rollout_cfg(num_explores=80, sample_actions_until=12)
    kind = L.LEAF_NN if leaf == "nn" else L.LEAF_ROLLOUT
    blob = s.Connect4Net.new(5).blob()

    def run(n):
        with s.Engine(0, n, 80) as e:
            e.set_weights(blob)
            return e.gather(cfg, kind, 11, 1200, 4, trace=True)
    a = run(148 * 640)
    b = run(in_flight)
    assert_rows_equal(a[0], b[0], "experience")
    assert_rows_equal(a[2], b[2], "trace")
    for k in _COUNTERS:
        assert a[1][k] == b[1][k], k


def test_a_thousand_games_occupy_every_sm():
    """the public gather_experience sized its engine so that the reference's own configuration
    (games_per_train = 1000) ran on ONE CTA.  Now the engine holds all of an iteration's games in flight and the kernels
    seat them over all SMs: 1,000 games = 148 CTAs of at most 7 games, 4,096 = 148 x 28, the bench's 94,720 = 148 x 640."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for n, leaf in ((1000, L.LEAF_NN), (1000, L.LEAF_ROLLOUT), (4096, L.LEAF_NN), (256, L.LEAF_ROLLOUT), (148 * 640, L.LEAF_NN), (5, L.LEAF_NN)):
        with s.Engine(0, s.games_in_flight_for(n), 50) as e:
            e.set_group_lanes(1)
            ctas, per, lanes = e.launch_geometry(n, leaf)
        assert lanes == 1 and ctas == min(sms, n), (n, ctas)
        assert per == -(-n // ctas), (n, per)
    # left to itself the engine gives few games a lane group each (the shortest time per explore) and many games a thread each
    with s.Engine(0, 148 * 640, 50) as e:
        e.set_weights(s.Connect4Net.new(0).blob())
        assert e.launch_geometry(1000, L.LEAF_NN)[2] == 32 and e.launch_geometry(4096, L.LEAF_NN)[2] == 16
        assert e.launch_geometry(256, L.LEAF_ROLLOUT)[2] == 32 and e.launch_geometry(4096, L.LEAF_ROLLOUT)[2] == 32
        assert e.launch_geometry(148 * 640, L.LEAF_NN)[2] == 1 and e.launch_geometry(60000, L.LEAF_ROLLOUT)[2] == 1
        assert e.launch_geometry(10000, L.LEAF_NN)[2] == 16 and e.launch_geometry(20000, L.LEAF_ROLLOUT)[2] == 32  # seats refilled
        e.set_mlp_mode(2)  # fp32-grade leaves: the team kernels carry the split chain too
        assert e.launch_geometry(1000, L.LEAF_NN)[2] == 32 and e.launch_geometry(148 * 640, L.LEAF_NN)[2] == 1
        e.set_mlp_mode(0)  # the fp32 CUDA-core forward has no small-batch mapping of its own: thread per game's seating
        assert e.launch_geometry(1000, L.LEAF_NN)[2] == 1
        # a lane-group launch is dealt over all SMs too: 1,000 games are 148 CTAs of at most 7 groups
        e.set_mlp_mode(1)
        ctas, per, lanes = e.launch_geometry(1000, L.LEAF_NN)
        assert (ctas, per, lanes) == (sms, 7, 32)


@pytest.mark.parametrize("leaf", ["nn", "nn_split", "rollout"])
@pytest.mark.parametrize("games", [100, 3000, 12000])
def test_mapping_chosen_per_launch_does_not_change_results(leaf, games):
    """The default engine picks the mapping per launch (32 lanes, 16 lanes or a thread per game by the games in flight); whatever
    it picks, rows, per-move visit counts and counters equal the thread-per-game kernels'."""
    cfg = s.study_connect4_rollout_cfg(num_explores=40, sample_actions_until=12)
    kind = L.LEAF_ROLLOUT if leaf == "rollout" else L.LEAF_NN
    blob = s.Connect4Net.new(5).blob()

    def run(lanes):
        with s.Engine(0, 148 * 640, 40) as e:
            e.set_weights(blob)
            e.set_mlp_mode(2 if leaf == "nn_split" else 1)  # a pinned chain: auto could choose differently per weights, never per mapping
            e.set_group_lanes(lanes)
            return e.gather(cfg, kind, 2, games, 6, trace=True)
    a, b = run(0), run(1)
    assert_rows_equal(a[0], b[0], "experience")
    assert_rows_equal(a[2], b[2], "trace")
    for k in _COUNTERS:
        assert a[1][k] == b[1][k], k
