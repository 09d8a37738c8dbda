#!/usr/bin/env python
"""Regenerates tests/golden/oracle_search.json, oracle_match.json and oracle_gather.npz FROM THE ORACLE.

    python tests/golden/make_golden.py

The Rust reference cannot be built or imported in this image (no cargo/rustc, tch needs libtorch
1.8), so these vectors are outputs of oracle/ — after oracle/ itself was pinned against the
reference's own known answers (tests/golden/reference_kats.json, tests/test_oracle_pinning.py).
Positions are stored as move lists so the fixture does not depend on numpy's generator.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path[:0] = [ROOT, TESTS, os.path.join(ROOT, "oracle")]

import build as oracle_build  # noqa: E402
import golden_fixtures as G  # noqa: E402
import oracle_binding  # noqa: E402
import synthesis_b200 as s  # noqa: E402
from synthesis_b200 import _lib as L  # noqa: E402


def random_move_lists(rng, n, max_plies):
    out = []
    while len(out) < n:
        g, ms, ok = s.Connect4.new(), [], True
        for _ in range(int(rng.integers(0, max_plies))):
            acts = list(g.iter_actions())
            a = int(acts[int(rng.integers(0, len(acts)))])
            if g.step(a):
                ok = False
                break
            ms.append(a)
        if ok:
            out.append(ms)
    return out


def main():
    oracle_build.build()
    orc = oracle_binding.Oracle()
    rng = np.random.default_rng(20261017)
    positions = [[], [4, 4, 3, 3], [4, 3, 4, 3, 4, 3]] + random_move_lists(rng, 9, 45)
    cases = []
    plan = [("puct", L.TREE_MCTS, (1, 64, 800)), ("uct_q", L.TREE_MCTS, (200,)), ("uct", L.TREE_FROZEN, (1, 150, 600)),
            ("no_solve", L.TREE_MCTS, (120,)), ("no_correct", L.TREE_MCTS, (120,)), ("no_select_solved", L.TREE_MCTS, (120,)),
            ("no_auto_extend", L.TREE_MCTS, (120,)), ("parent_q", L.TREE_MCTS, (120,)), ("equal_noise", L.TREE_MCTS, (120,)),
            ("dirichlet", L.TREE_MCTS, (120,)), ("fpu_normal", L.TREE_MCTS, (120,))]
    for name, kind, explores_list in plan:
        for explores in explores_list:
            for i, moves in enumerate(positions if name in ("puct", "uct") else positions[:6]):
                seed = 0 if i == 0 else 7 * i + explores
                g = G.game_from_moves(moves)
                out, _ = orc.search(G.named_cfg(name, explores).to_c(L.LEAF_ROLLOUT), g.my_bb, g.op_bb, seed, tree_kind=kind)
                cases.append(dict(cfg=name, tree_kind=kind, explores=explores, seed=seed, moves=moves, out=G.search_case_outputs(out)))
    with open(G.SEARCH_JSON, "w") as f:
        json.dump({"_about": "oracle outputs; generator tests/golden/make_golden.py; see tests/golden_fixtures.py", "cases": cases}, f,
                  separators=(",", ":"))
    mcases = []
    for k0, k1, seed, ex in G.MATCH_CASES:
        out, _ = orc.match((G.match_player(k0, ex[0]), G.match_player(k1, ex[1])), seed, ex)
        mcases.append(dict(players=[k0, k1], seed=seed, explores=list(ex), out=G.match_case_outputs(out)))
    with open(G.MATCH_JSON, "w") as f:
        json.dump({"_about": "oracle outputs (orc_match); generator tests/golden/make_golden.py", "cases": mcases}, f, separators=(",", ":"))
    arrays = {}
    for case in G.GATHER_CASES:
        name, _, _, _, _, first, games, seed = case
        a, _, t = orc.gather(G.gather_case_cfg(case).to_c(L.LEAF_ROLLOUT), seed, first, games, threads=4)
        for k, v in a.items():
            arrays[f"{name}.{k}"] = v
        for k, v in t.items():
            arrays[f"{name}.trace_{k}"] = v
    np.savez_compressed(G.GATHER_NPZ, **arrays)
    print(len(cases), "search cases;", sum(v.nbytes for v in arrays.values()), "gather bytes")


if __name__ == "__main__":
    main()
