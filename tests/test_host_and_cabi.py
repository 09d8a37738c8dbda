"""CPU suite, part 2: host-side mirror of the reference interface, and the C-ABI library.

No compute call is made here (there is no GPU in the authoring container): the library must load,
export every symbol include/synthesis_b200.h declares, and FAIL LOUDLY without a B200 — there is
no CPU fallback behind any entry point.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import synthesis_b200 as s
from synthesis_b200 import _lib as L
from synthesis_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


# ------------------------------------------------------------------ the C ABI
def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "synthesis_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(syn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    lib = L.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.syn_abi_version() == int(re.search(r"#define SYN_ABI_VERSION (\d+)", header).group(1))
    assert b"sm_100a" in lib.syn_build_info()


def test_struct_layouts_match_the_header():
    # sizes as laid out by the C compiler for include/synthesis_b200.h (checked by a static_assert-like test)
    assert C.sizeof(L.SynMctsCfg) == 36
    assert C.sizeof(L.SynRolloutCfg) == 72
    assert C.sizeof(L.SynExperience) == 3 * C.sizeof(C.c_size_t) + 8 * C.sizeof(C.c_void_p)
    assert C.sizeof(L.SynStats) == 8 * len(L.SynStats._fields_)
    assert C.sizeof(L.SynFlatBatch) == 2 * C.sizeof(C.c_size_t) + 6 * C.sizeof(C.c_void_p)
    assert C.sizeof(L.SynTrainCfg) == 32
    src = r'''
#include "synthesis_b200.h"
#include <stdio.h>
int main(void) { printf("%zu %zu %zu %zu %zu %zu\n", sizeof(syn_mcts_cfg), sizeof(syn_rollout_cfg), sizeof(syn_experience), sizeof(syn_stats), sizeof(syn_flat_batch), sizeof(syn_train_cfg)); return 0; }
'''
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "sz.c")
        with open(p, "w") as f:
            f.write(src)
        exe = os.path.join(d, "sz")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), p, "-o", exe])  # the header is plain C
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(L.SynMctsCfg), C.sizeof(L.SynRolloutCfg), C.sizeof(L.SynExperience), C.sizeof(L.SynStats),
                     C.sizeof(L.SynFlatBatch), C.sizeof(L.SynTrainCfg)]


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_engine_creation_fails_loudly():
    with pytest.raises(L.EngineError) as e:
        s.Engine(0, 64, 100)
    assert e.value.code == L.SYN_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value) or "sm_100a" in str(e.value)
    lib = L.load()
    h = C.c_void_p()
    assert lib.syn_engine_create(0, 0, 0, C.byref(h)) != 0 and not h.value
    assert lib.syn_engine_create(0, 64, 100, None) == L.SYN_ERR_INVALID_ARGUMENT
    # NULL engines are rejected, not dereferenced
    assert lib.syn_engine_set_weights(None, None, 0) == L.SYN_ERR_INVALID_ARGUMENT
    assert lib.syn_engine_gather_wait(None, None, None) == L.SYN_ERR_INVALID_ARGUMENT
    lib.syn_engine_destroy(None)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "synthesis_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                for needle in ("liboracle", "oracle_binding", "import oracle", "from oracle", '"oracle"', "oracle/", "orc_"):
                    assert needle not in txt, f"{fn} mentions {needle!r}: the product must not reach into oracle/"


# ------------------------------------------------------------------ config.rs mirror
def test_config_maps_to_the_c_structs():
    cfg = s.study_connect4_rollout_cfg(num_explores=800, sample_actions_until=30)
    c = cfg.to_c(L.LEAF_NN)
    assert (c.num_explores, c.random_actions_until, c.sample_actions_until, c.stop_games_when_solved) == (800, 1, 30, 0)
    assert c.value_target_kind == L.VALUE_Q and c.action_selection == L.ACTION_NUM_VISITS and c.leaf_eval_kind == L.LEAF_NN
    m = c.mcts
    assert (m.exploration_kind, m.c) == (L.EXPLORATION_POLYNOMIAL_UCT, 3.0)
    assert (m.solve, m.correct_values_on_solve, m.select_solved_nodes, m.auto_extend) == (1, 1, 1, 1)
    assert (m.fpu_kind, m.fpu_a, m.noise_kind) == (L.FPU_CONST, 1.0, L.NOISE_NONE)
    r = s.study_connect4_rollout_mcts_cfg().to_c()
    assert (r.exploration_kind, r.c, r.auto_extend, r.fpu_kind) == (L.EXPLORATION_UCT, 2.0, 0, L.FPU_CONST) and np.isinf(r.fpu_a)
    cfg.value_target = s.ValueTarget.QtoZ(0.25, 0.75)
    cfg.mcts_cfg.root_policy_noise = s.PolicyNoise.Dirichlet(1.0, 0.25)
    cfg.mcts_cfg.fpu = s.Fpu.Normal(1.0, 0.1)
    c = cfg.to_c(L.LEAF_ROLLOUT)
    assert (c.value_target_kind, c.vt_a, c.vt_b) == (L.VALUE_Q_TO_Z, 0.25, 0.75)
    assert (c.mcts.noise_kind, c.mcts.noise_alpha, c.mcts.noise_weight) == (L.NOISE_DIRICHLET, 1.0, 0.25)
    assert (c.mcts.fpu_kind, c.mcts.fpu_a) == (L.FPU_NORMAL, 1.0) and abs(c.mcts.fpu_b - 0.1) < 1e-7
    cfg.mcts_cfg.fpu = s.Fpu.Func(lambda: 1.0)  # carried to the engine, which rejects it (SYN_ERR_UNSUPPORTED)
    assert cfg.to_c(L.LEAF_NN).mcts.fpu_kind == L.FPU_FUNC
    cfg.mcts_cfg.fpu = "nonsense"
    with pytest.raises(TypeError):
        cfg.to_c(L.LEAF_NN)


def test_connect4net_blob_round_trip():
    net = s.Connect4Net.new(0)
    blob = net.blob()
    assert blob.shape == (L.N_WEIGHTS,) and blob.dtype == np.float32
    assert L.N_WEIGHTS == 63 * 128 + 128 + 128 * 96 + 96 + 96 * 64 + 64 + 64 * 48 + 48 + 48 * 12 + 12
    again = s.Connect4Net.from_blob(blob)
    for k, v in net.params.items():
        assert np.array_equal(v, again.params[k])
    assert np.abs(net.params["l_1.weight"]).max() <= 1 / np.sqrt(63) and np.abs(net.params["l_5.bias"]).max() <= 1 / np.sqrt(48)
    with pytest.raises(ValueError):
        s.Connect4Net.from_blob(blob[:-1])


def test_host_connect4_mirror_against_the_oracle(oracle):
    rng = np.random.default_rng(3)
    for _ in range(300):
        g, ms = s.Connect4.new(), []
        for _ in range(int(rng.integers(0, 64))):
            acts = list(g.iter_actions())
            a = acts[int(rng.integers(0, len(acts)))]
            ms.append(a)
            if g.step(a):
                break
        r = oracle.c4_play(ms)
        assert (g.my_bb, g.op_bb, list(g.height)) == (r["my_bb"], r["op_bb"], list(r["height"]))
        assert g.is_over() == bool(r["status"] & 1)
        assert g.reward(g.player()) == r["reward_to_move"]
        assert sum(1 << c for c in g.iter_actions()) == r["legal_mask"]
        assert np.array_equal(g.features().reshape(63).view(np.uint32), r["features"].view(np.uint32))
        assert s.Connect4.from_bitboards(g.my_bb, g.op_bb) == g
    with pytest.raises(ValueError):
        g = s.Connect4.new()
        for _ in range(8):
            g.step(0)


# ------------------------------------------------------------------ data.rs mirror
def _rows(ids):
    n = len(ids)
    return dict(game_ids=np.asarray(ids, np.uint64), my_bb=np.arange(n, dtype=np.uint64), op_bb=np.zeros(n, np.uint64),
                height=np.zeros((n, 9), np.uint8), player=np.zeros(n, np.uint8), states=np.zeros((n, 63), np.float32),
                pis=np.zeros((n, 9), np.float32), vs=np.zeros((n, 3), np.float32))


def test_replay_buffer_extend_and_keep_last_n_games(oracle):
    buf = s.ReplayBuffer()
    w1 = s.ReplayBuffer.from_arrays(3, _rows([1, 1, 2, 3, 3, 3]))
    buf.keep_last_n_games(10 - 3)
    buf.extend(w1)
    assert buf.total_games_played() == 3 and buf.curr_steps() == 6 and buf.total_steps() == 6 and buf.curr_games() == 3
    assert len(w1.vs) == 0  # drained like Vec::drain(..) (data.rs:165-169)
    w2 = s.ReplayBuffer.from_arrays(2, _rows([1, 2, 2]))
    buf.extend(w2)  # ids are re-based by the running game counter (data.rs:162-163)
    assert list(buf.game_ids) == [1, 1, 2, 3, 3, 3, 4, 5, 5] and buf.game_id == 5
    # keep_last_n_games(n): drop every row whose id < game_id - n (data.rs:172-194), checked against the oracle's restatement
    for n in (0, 1, 2, 3, 4, 5, 9):
        b = s.ReplayBuffer.from_arrays(5, _rows([1, 1, 2, 3, 3, 3, 4, 5, 5]))
        ids = b.game_ids.copy()
        want = oracle.lib.orc_keep_last_n_games_prefix(ids.ctypes.data, len(ids), 5, n)
        b.keep_last_n_games(n)
        assert len(ids) - len(b.game_ids) == want, n
        assert b.total_games_played() == 5
    g = s.Connect4.new()
    g.step(4)
    buf.new_game()
    buf.add(g, np.full(9, 1 / 9, np.float32), np.zeros(3, np.float32))
    assert buf.game_ids[-1] == 6 and buf.curr_steps() == 10 and buf.games[-1] == g


def test_oracle_deduplicate_follows_the_reference_loop(oracle):
    """orc_deduplicate against a line-by-line Python restatement of data.rs:196-235 (a dict keyed by the position —
    Python dicts iterate in insertion order — f32 sums in buffer order, then sum / num as f32), on the rows of an
    oracle gather and on hand-made rows whose sums depend on the order of the additions."""
    cfg = s.study_connect4_rollout_cfg(num_explores=30, sample_actions_until=10)
    a, _, _ = oracle.gather(cfg.to_c(L.LEAF_ROLLOUT), 3, 0, 12, threads=2)
    rng = np.random.default_rng(0)
    hand_my = rng.integers(0, 3, 200).astype(np.uint64)
    hand = dict(my_bb=hand_my, op_bb=np.zeros(200, np.uint64), pis=(rng.random((200, 9)) * 1e3).astype(np.float32) ** 3,
                vs=rng.standard_normal((200, 3)).astype(np.float32))
    for rows in (a, hand):
        got = oracle.deduplicate(rows["my_bb"], rows["op_bb"], rows["pis"], rows["vs"])
        stats = {}
        for i in range(len(rows["vs"])):
            key = (int(rows["my_bb"][i]), int(rows["op_bb"][i]))
            st = stats.setdefault(key, dict(sum_pi=np.zeros(9, np.float32), sum_v=np.zeros(3, np.float32), num=0))
            st["sum_pi"] = st["sum_pi"] + rows["pis"][i]
            st["sum_v"] = st["sum_v"] + rows["vs"][i]
            st["num"] += 1
        assert len(got["num"]) == len(stats)
        for g, (key, st) in enumerate(stats.items()):
            assert (int(got["my_bb"][g]), int(got["op_bb"][g])) == key and got["num"][g] == st["num"]
            assert got["pis"][g].tobytes() == (st["sum_pi"] / np.float32(st["num"])).astype(np.float32).tobytes()
            assert got["vs"][g].tobytes() == (st["sum_v"] / np.float32(st["num"])).astype(np.float32).tobytes()
    # StateStatistics::state is the features of the position (data.rs:203)
    got = oracle.deduplicate(a["my_bb"], a["op_bb"], a["pis"], a["vs"])
    first = {}
    for i in range(len(a["vs"])):
        first.setdefault((int(a["my_bb"][i]), int(a["op_bb"][i])), i)
    assert got["states"].tobytes() == a["states"][list(first.values())].tobytes()
    assert got["num"][0] == 12  # the empty board, once per game


def test_split_games_follows_the_reference_schedule():
    # alpha_zero.rs:132-154: num_games = remaining / workers_left; 1000 games over 7 workers -> 142, 143 x 6
    parts = D.split_games(1000, 7)
    assert [n for _, n in parts] == [142] + [143] * 6
    assert parts[0][0] == 0 and all(parts[i + 1][0] == parts[i][0] + parts[i][1] for i in range(6))
    assert D.split_games(3, 8) == [(0, 0)] * 5 + [(0, 1), (1, 1), (2, 1)]
    assert sum(n for _, n in D.split_games(32768, 8)) == 32768 and {n for _, n in D.split_games(32768, 8)} == {4096}


def test_stream_seeds(oracle):
    """include/syn_streams.h: seed 0 gives rollout stream 2g and action stream 2g+1 (SURVEY §8c convention)."""
    for g in (0, 1, 5, 4095):
        assert oracle.lib.orc_stream_seed(0, g, 0) == 2 * g and oracle.lib.orc_stream_seed(0, g, 1) == 2 * g + 1
    seen = {oracle.lib.orc_stream_seed(sd, g, k) for sd in range(3) for g in range(50) for k in range(4)}
    assert len(seen) == 3 * 50 * 4


def test_add_pgn_result_text_is_the_reference_s():
    """utils.rs:32-53: three tag lines and the result line per game, '1-0' / '0-1' / '1/2-1/2' by white's reward; any other
    reward trips the reference's assert_eq!."""
    import io
    import synthesis_b200.evaluator as ev
    out = io.StringIO()
    ev.add_pgn_result(out, "model_3.ot", "VanillaMCTS800", 1.0)
    ev.add_pgn_result(out, "VanillaMCTS800", "model_3.ot", -1.0)
    ev.add_pgn_result(out, "model_3.ot", "model_2.ot", 0.0)
    assert out.getvalue() == ('[White "model_3.ot"]\n[Black "VanillaMCTS800"]\n[Result "1-0"]\n1-0\n'
                              '[White "VanillaMCTS800"]\n[Black "model_3.ot"]\n[Result "0-1"]\n0-1\n'
                              '[White "model_3.ot"]\n[Black "model_2.ot"]\n[Result "1/2-1/2"]\n1/2-1/2\n')
    with pytest.raises(AssertionError):
        ev.add_pgn_result(io.StringIO(), "a", "b", 0.5)


def test_two_shift_won_equals_the_reference_form():
    """csrc/c4.cuh won(): `t = bb & bb>>s; t & t>>2s & MASK` per direction is the reference's
    `bb & bb>>s & bb>>2s & bb>>3s & MASK` (connect4.rs:77-83) for every 64-bit pattern — checked on random boards, on
    boards built from played games, and on all single lines of four."""
    import random
    ROW0 = sum(1 << (7 * c) for c in range(9)); C05 = (1 << 42) - 1
    masks = {1: ROW0 * 0x0F, 7: C05, 8: C05 & (ROW0 * 0x0F), 6: C05 & (ROW0 * 0x78)}
    ref = lambda bb: any(bb & (bb >> s) & (bb >> 2 * s) & (bb >> 3 * s) & m for s, m in masks.items())

    def two(bb):
        out = 0
        for s, m in masks.items():
            t = bb & (bb >> s)
            out |= t & (t >> 2 * s) & m
        return out != 0
    rnd = random.Random(5)
    boards = [rnd.getrandbits(63) & rnd.getrandbits(63) for _ in range(20000)] + [rnd.getrandbits(63) for _ in range(5000)]
    for s in masks:
        for start in range(63):
            boards.append(sum(1 << (start + k * s) for k in range(4) if start + k * s < 63))
    assert all(ref(b) == two(b) for b in boards)
    assert any(ref(b) for b in boards) and not all(ref(b) for b in boards)


def test_winning_cells_algebra_equals_won_per_column():
    """csrc/c4.cuh winning_cells(): the mover's winning cells for all nine columns from one pass of masked
    neighbour shifts.  Mirrored here in Python integers and checked against Connect4::won (connect4.rs:77-83)
    applied to every candidate move of random positions."""
    import random
    ROW0 = sum(1 << (7 * c) for c in range(9)); ROW6 = ROW0 << 6; ALL = (1 << 63) - 1; M64 = (1 << 64) - 1
    C05 = (1 << 42) - 1
    VM, HM, D2M, D1M = ROW0 * 0x0F, C05, C05 & (ROW0 * 0x0F), C05 & (ROW0 * 0x78)

    def won(bb):
        return ((bb & (bb >> 6) & (bb >> 12) & (bb >> 18) & D1M) | (bb & (bb >> 8) & (bb >> 16) & (bb >> 24) & D2M) |
                (bb & (bb >> 7) & (bb >> 14) & (bb >> 21) & HM) | (bb & (bb >> 1) & (bb >> 2) & (bb >> 3) & VM)) != 0

    up = lambda x: ((x & ~ROW6) << 1) & M64
    right, left = (lambda x: (x << 7) & ALL), (lambda x: x >> 7)
    ur, dl = (lambda x: ((x & ~ROW6) << 8) & ALL), (lambda x: (x & ~ROW0) >> 8)
    dr, ul = (lambda x: ((x & ~ROW0) << 6) & ALL), (lambda x: (x & ~ROW6) >> 6)

    def line(p, F, B):
        P1 = B(p); P2 = B(P1); P3 = B(P2); M1 = F(p); M2 = F(M1); M3 = F(M2)
        return (P1 & P2 & P3) | (M1 & P1 & P2) | (M2 & M1 & P1) | (M3 & M2 & M1)

    def winning_cells(p):
        b1 = up(p); b2 = up(b1); b3 = up(b2)
        return (b1 & b2 & b3) | line(p, right, left) | line(p, ur, dl) | line(p, dr, ul)

    rnd = random.Random(5)
    checked = wins = 0
    for _ in range(20000):
        my = op = 0
        h = [0] * 9
        alive = True
        for _ in range(rnd.randint(0, 60)):
            cols = [c for c in range(9) if h[c] < 7]
            if not cols:
                break
            c = rnd.choice(cols)
            mover = my | (1 << (h[c] + 7 * c))
            h[c] += 1
            my, op = op, mover
            if won(mover):
                alive = False
                break
        if not alive:
            continue
        w = winning_cells(my)
        for c in range(9):
            if h[c] < 7:
                bit = 1 << (h[c] + 7 * c)
                assert won(my | bit) == ((w & bit) != 0), (hex(my), hex(op), c)
                checked += 1
                wins += won(my | bit)
    assert checked > 50000 and wins > 1000


def test_batch_rand_sampler_and_lr_schedule():
    """BatchRandSampler (data.rs:6-64): a permutation cut into batches, the short tail dropped iff drop_last; and the
    learning-rate lookup of alpha_zero.rs:62-70."""
    rng = np.random.default_rng(0)
    sm = s.BatchRandSampler(70, 32, True, rng)
    got = list(sm)
    assert [len(b) for b in got] == [32, 32] and len(set(np.concatenate(got).tolist())) == 64
    got = list(s.BatchRandSampler(70, 32, False, rng))
    assert [len(b) for b in got] == [32, 32, 6] and sorted(np.concatenate(got).tolist()) == list(range(70))
    assert list(s.BatchRandSampler(10, 32, True, rng)) == []
    sm = s.BatchRandSampler(100, 32, True, np.random.default_rng(3))
    ref = list(s.BatchRandSampler(100, 32, True, np.random.default_rng(3)))
    assert np.array_equal(sm.all_batches(), np.stack(ref)) and sm.all_batches().shape == (0, 32)
    cfg = s.LearningConfig(seed=0, logs="", lr_schedule=[(1, 1e-3), (20, 5e-4), (40, 1e-4)], weight_decay=0.0, num_iterations=1, num_epochs=1,
                           batch_size=32, policy_weight=1.0, value_weight=1.0, games_to_keep=1, games_per_train=1,
                           rollout_cfg=s.study_connect4_rollout_cfg())
    assert [s.lr_for_iteration(cfg, i) for i in (0, 18, 19, 38, 39, 100)] == [1e-3, 1e-3, 5e-4, 5e-4, 1e-4, 1e-4]


def test_rows_from_bitboards_matches_the_game_mirror():
    """Compact 72-byte rows (ids, bitboards, pi, v) are what crosses NVLink and PCIe; height / player / features are rebuilt
    from the bitboards.  The vectorised rebuild must equal Connect4's own accessors (connect4.rs:108-114, 237-258) on
    random playouts, and ReplayBuffer.from_arrays must accept rows without those columns."""
    from synthesis_b200.data import rows_from_bitboards
    rng = np.random.default_rng(3)
    my, op, hs, ps, fs = [], [], [], [], []
    for _ in range(300):
        g = s.Connect4.new()
        for _ in range(int(rng.integers(0, 60))):
            acts = list(g.iter_actions())
            if not acts or g.is_over():
                break
            g.step(int(rng.choice(acts)))
        my.append(g.my_bb); op.append(g.op_bb); hs.append(list(g.height)); ps.append(g.player()); fs.append(np.asarray(g.features(), np.float32).reshape(-1))
    h, p, st = rows_from_bitboards(np.array(my, np.uint64), np.array(op, np.uint64))
    assert np.array_equal(h, np.array(hs, np.uint8)) and np.array_equal(p, np.array(ps, np.uint8))
    assert st.tobytes() == np.array(fs, np.float32).tobytes()
    n = len(my)
    compact = dict(game_ids=np.arange(1, n + 1, dtype=np.uint64), my_bb=np.array(my, np.uint64), op_bb=np.array(op, np.uint64),
                   pis=np.zeros((n, 9), np.float32), vs=np.zeros((n, 3), np.float32))
    buf = s.ReplayBuffer.from_arrays(n, compact)
    assert buf.states.tobytes() == st.tobytes() and np.array_equal(buf.height, h) and np.array_equal(buf.player, p)
