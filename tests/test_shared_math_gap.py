"""How much does sharing ONE exp/ln between kernels and oracle (include/syn_detmath.h) hide?  The reference calls the platform's
libm (`ln` in UCT, mcts.rs:364 / evaluator.rs:420; `exp` in the prior softmax, mcts.rs:418).  The oracle can run either
(TreeOptions::libm); here both forms are run on the CPU over >= 1,000 UCT searches and >= 100 network-prior games and the
divergence is asserted against a stated bound — so "UCT / softmax parity rests on a shared function" becomes a number:
measured 0 of 1,000 searches and 0 of 100 games differ (glibc 2.39, x86-64)."""
import ctypes as C
import math

import numpy as np

import oracle_binding as B
import synthesis_b200 as s
from synthesis_b200 import _lib as L


def _positions(rng, n, max_plies=36):
    out = []
    while len(out) < n:
        g = s.Connect4.new()
        ok = True
        for _ in range(int(rng.integers(0, max_plies))):
            acts = list(g.iter_actions())
            if g.step(acts[int(rng.integers(0, len(acts)))]):
                ok = False
                break
        if ok:
            out.append(g)
    return out


def test_uct_visit_counts_with_libm_ln(oracle):
    """1,000 FrozenMCTS searches with the evaluator's rollout-baseline config (UCT c=2, FPU inf, main.rs:74-82), 800 explores:
    root child visit counts and best action with std::log against syn_logf.  Bound: at most 1 % of the searches may differ."""
    rng = np.random.default_rng(0)
    cfg = s.study_connect4_rollout_cfg(num_explores=800, mcts_cfg=s.study_connect4_rollout_mcts_cfg())
    cfg.action = s.ActionSelection.Q
    ccfg = cfg.to_c(L.LEAF_ROLLOUT)
    differ = best_differ = 0
    games = _positions(rng, 1000)
    for i, g in enumerate(games):
        a, _ = oracle.search(ccfg, g.my_bb, g.op_bb, i, tree_kind=L.TREE_FROZEN)
        b, _ = oracle.search(ccfg, g.my_bb, g.op_bb, i, tree_kind=L.TREE_FROZEN, flags=B.FLAG_LIBM)
        differ += not np.array_equal(a["child_visits"], b["child_visits"])
        best_differ += a["best_action"] != b["best_action"]
    print(f"UCT, 1000 searches at 800 explores: {differ} visit vectors and {best_differ} best actions differ between libm ln and syn_logf")
    assert differ <= 10 and best_differ <= 10


def test_network_prior_games_with_libm_exp(oracle):
    """100 self-play games with Connect4Net priors (PUCT, 200 explores/move, sampled actions): moves and every per-move visit
    vector with std::exp against syn_expf in the prior softmax and the value softmax.  Bound: at least 99 games identical."""
    net = s.Connect4Net.new(0)
    cfg = s.study_connect4_rollout_cfg(num_explores=200, sample_actions_until=30)
    ccfg = cfg.to_c(L.LEAF_NN)
    a, _, ta = oracle.gather(ccfg, 0, 0, 100, weights=net.blob(), threads=8)
    b, _, tb = oracle.gather(ccfg, 0, 0, 100, weights=net.blob(), threads=8, flags=B.FLAG_LIBM)
    same = 0
    for g in range(1, 101):
        ia, ib = a["game_ids"] == g, b["game_ids"] == g
        same += bool(ia.sum() == ib.sum() and np.array_equal(ta["action"][ia], tb["action"][ib]) and np.array_equal(ta["child_visits"][ia], tb["child_visits"][ib]))
    print(f"network priors, 100 games at 200 explores/move: {same} identical between libm exp and syn_expf")
    assert same >= 99


def test_syn_logf_and_expf_against_libm_pointwise(oracle):
    """Where the two could part: ln(N) for every visit count a tree can hold and exp(l - max) over the softmax's range.
    The shared functions agree with the correctly rounded f32 result (numpy float64 -> float32) to <= 1 ulp everywhere; syn_logf
    is bit-identical on > 99 % of the visit counts, syn_expf on ~91 % of the softmax's range (a 1-ulp prior moves a PUCT
    argmax so rarely that 100 whole games above do not differ)."""
    lib = oracle.lib
    lib.orc_detmath.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    n = np.arange(1, 20001, dtype=np.float32)
    x = np.linspace(-30.0, 0.0, 200001, dtype=np.float32)
    for kind, arg, ref in ((0, n, np.log(n.astype(np.float64))), (1, x, np.exp(x.astype(np.float64)))):
        out = np.zeros_like(arg)
        lib.orc_detmath(kind, arg.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), arg.size)
        want = ref.astype(np.float32)
        ulp = np.abs(out.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        exact = float((ulp == 0).mean())
        print(f"{'syn_logf' if kind == 0 else 'syn_expf'}: max {int(ulp.max())} ulp from the correctly rounded result, bit-identical on {100 * exact:.2f} % of {arg.size} arguments")
        assert int(ulp.max()) <= 1, (kind, int(ulp.max()))
        assert exact > (0.99 if kind == 0 else 0.85), (kind, exact)
