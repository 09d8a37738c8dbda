"""Host mirror of the game loops of synthesis/src/evaluator.rs — the callers on the far side of the
search path (SURVEY.md §8 f3 / BASELINE.json configs[4]).

The reference plays these games one at a time on one thread; here a whole sweep is ONE
`syn_engine_match` call, a thread per match (csrc/match.cuh):

    eval_against_rollout_mcts  evaluator.rs:163-198   NN `MCTS::exploit` vs rollout `FrozenMCTS::exploit`
    mcts_vs_mcts               evaluator.rs:200-228   rollout `FrozenMCTS` vs rollout `FrozenMCTS`
    eval_against_old           evaluator.rs:131-161   NN `MCTS` vs NN `MCTS`, two networks resident (called at :87-94)
    add_pgn_result             utils.rs:32-53         results.pgn in the exact text bayeselo reads

Names, argument meaning and return values follow the reference: every function returns
`game.reward(first_player)` (+1 / 0 / -1) per game.
"""
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from . import _lib as L
from .config import EvaluationConfig, MCTSConfig
from .engine import Engine
from .policies import Connect4Net

RED, BLACK = 0, 1  # connect4.rs:101-106; Red moves first (connect4.rs:182-189)


@dataclass
class Player:
    """One side of a match: the argument list of MCTS::exploit (mcts.rs:111-121) /
    FrozenMCTS::exploit (evaluator.rs:308-318) minus the game."""
    tree_kind: int        # L.TREE_MCTS | L.TREE_FROZEN
    leaf_eval_kind: int   # L.LEAF_NN | L.LEAF_ROLLOUT
    num_explores: int
    mcts_cfg: MCTSConfig
    action_selection: int

    def to_c(self) -> L.SynPlayerCfg:
        c = L.SynPlayerCfg()
        c.tree_kind, c.leaf_eval_kind = int(self.tree_kind), int(self.leaf_eval_kind)
        c.num_explores, c.action_selection = int(self.num_explores), int(self.action_selection)
        c.mcts = self.mcts_cfg.to_c()
        return c


def _policy_player(cfg: EvaluationConfig) -> Player:
    return Player(L.TREE_MCTS, L.LEAF_NN, cfg.policy_num_explores, cfg.policy_mcts_cfg, cfg.policy_action)


def _rollout_player(cfg: EvaluationConfig, explores: int) -> Player:
    return Player(L.TREE_FROZEN, L.LEAF_ROLLOUT, explores, cfg.rollout_mcts_cfg, cfg.rollout_action)


def eval_against_rollout_mcts(engine: Engine, cfg: EvaluationConfig, policy: Connect4Net, player: int,
                              opponent_explores: Sequence[int], seeds: Sequence[int], trace: bool = False):
    """evaluator.rs:163-198 for a batch: game i is policy (as colour `player`) against
    FrozenMCTS(opponent_explores[i]) with rollout stream seed seeds[i].  Returns rewards for the
    first player (and the engine's per-move trace when trace=True)."""
    engine.set_weights(policy.blob())
    n = len(seeds)
    nn, ro = _policy_player(cfg), _rollout_player(cfg, max(opponent_explores))
    players = (nn, ro) if player == RED else (ro, nn)
    ex = np.zeros((n, 2), np.uint32)
    ex[:, 0 if player == RED else 1] = cfg.policy_num_explores
    ex[:, 1 if player == RED else 0] = np.asarray(opponent_explores, np.uint32)
    out, stats = engine.match(players, seeds, ex, trace=trace)
    return (out["result"], out, stats) if trace else out["result"]


def mcts_vs_mcts(engine: Engine, cfg: EvaluationConfig, player: int, p1_explores: Sequence[int], p2_explores: Sequence[int],
                 seeds: Sequence[int], trace: bool = False):
    """evaluator.rs:200-228 for a batch: both sides are rollout FrozenMCTS sharing one rollout stream;
    `player` uses p1_explores, the other colour p2_explores."""
    n = len(seeds)
    ro = _rollout_player(cfg, max(max(p1_explores), max(p2_explores)))
    ex = np.zeros((n, 2), np.uint32)
    ex[:, 0 if player == RED else 1] = np.asarray(p1_explores, np.uint32)
    ex[:, 1 if player == RED else 0] = np.asarray(p2_explores, np.uint32)
    out, stats = engine.match((ro, ro), seeds, ex, trace=trace)
    return (out["result"], out, stats) if trace else out["result"]


def eval_against_old(engine: Engine, cfg: EvaluationConfig, p1: Connect4Net, p2: Connect4Net = None, n_games: int = 1, trace: bool = False):
    """evaluator.rs:129-160: p1 (moves first) against p2, both `MCTS::exploit` with the policy's settings; returns
    game.reward(first_player) per game (the game is deterministic: the reference plays it once per colour,
    evaluator.rs:87-94).  p2 = None plays p1 against itself.  Both weight images stay resident in the kernel."""
    engine.set_weights(p1.blob())
    engine.set_opponent_weights(None if p2 is None else p2.blob())
    try:
        nn = _policy_player(cfg)
        out, stats = engine.match((nn, nn), np.zeros(n_games, np.uint64), None, trace=trace)
    finally:
        engine.set_opponent_weights(None)
    return (out["result"], out, stats) if trace else out["result"]


def evaluate_against_old_models(engine: Engine, cfg: EvaluationConfig, policy: Connect4Net, name: str, old, pgn=None):
    """evaluator.rs:87-94: the newest model against each of the best older ones, once as each colour.  `old` = [(name, net)].
    Returns [(white, black, reward)] in the reference's order and writes the PGN records."""
    records = []
    for old_name, old_net in old:
        records.append((name, old_name, float(eval_against_old(engine, cfg, policy, old_net)[0])))
        records.append((old_name, name, float(eval_against_old(engine, cfg, old_net, policy)[0])))
    if pgn is not None:
        for w, b, r in records:
            add_pgn_result(pgn, w, b, r)
    return records


def add_pgn_result(pgn, white_name: str, black_name: str, white_reward: float) -> None:
    """utils.rs:32-53, byte for byte."""
    if white_reward == 1.0:
        result = "1-0"
    elif white_reward == -1.0:
        result = "0-1"
    else:
        assert white_reward == 0.0
        result = "1/2-1/2"
    pgn.write(f'[White "{white_name}"]\n[Black "{black_name}"]\n[Result "{result}"]\n{result}\n')


def evaluate_against_rollout_sweep(engine: Engine, cfg: EvaluationConfig, policy: Connect4Net, name: str, pgn=None):
    """The inner loops of evaluator.rs:65-82 as two launches (policy as Red, policy as Black): for every
    explores in cfg.rollout_num_explores and seed in 0..num_games_against_rollout.  Writes the PGN
    records in the reference's order and returns [(white, black, reward)]."""
    ex: List[int] = []
    sd: List[int] = []
    for explores in cfg.rollout_num_explores:
        for seed in range(cfg.num_games_against_rollout):
            ex.append(explores)
            sd.append(seed)
    as_red = eval_against_rollout_mcts(engine, cfg, policy, RED, ex, sd)
    as_black = eval_against_rollout_mcts(engine, cfg, policy, BLACK, ex, sd)
    records = []
    for i, explores in enumerate(ex):
        op_name = f"VanillaMCTS{explores}"
        records.append((name, op_name, float(as_red[i])))
        records.append((op_name, name, float(as_black[i])))
    if pgn is not None:
        for w, b, r in records:
            add_pgn_result(pgn, w, b, r)
    return records
