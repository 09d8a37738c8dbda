"""Drop-in for the self-play half of synthesis/src/alpha_zero.rs: `gather_experience`
(alpha_zero.rs:120-169) with the same argument meaning, plus `MCTS.exploit` (mcts.rs:111-121).

    gather_experience(cfg, policy, buffer, seed)

plays `cfg.games_per_train` games on the engine, then — exactly like the reference —
`buffer.keep_last_n_games(cfg.games_to_keep - cfg.games_per_train)` and `buffer.extend(new)`.
`policy` is a `Connect4Net` (the reference loads `models/<policy_name>.ot` in every worker,
alpha_zero.rs:192-194) or a `RolloutPolicy` (the faithful `run_game + RolloutPolicy` composition
of SURVEY.md §0.3).  `seed` is the iteration index like the reference's call site
(alpha_zero.rs:49); game g of the call draws from streams derived from (seed, g) — see
include/syn_streams.h — instead of one stream per worker thread.
"""
import numpy as np

from . import _lib as L
from .config import LearningConfig, MCTSConfig, RolloutConfig, ValueTarget
from .connect4 import Connect4
from .data import BatchRandSampler, FlatBatch, ReplayBuffer
from .engine import Engine
from .policies import Connect4Net, RolloutPolicy

_engines = {}


def games_in_flight_for(num_games: int) -> int:
    """How many games an engine made for `num_games` games per iteration should hold in flight: all of them, up to what one
    B200 seats (148 SMs x 640 threads).  The kernels seat a launch's games evenly over ALL SMs (tp2::seat_of), so the
    reference's 1,000 games per iteration (study-connect4/src/main.rs:26) run 6-7 per SM, not 640 on two SMs."""
    return int(min(max(int(num_games), 1), 148 * 640))


def engine_for(device: int, max_games_in_flight: int, max_explores: int) -> Engine:
    """One cached engine per (device, capacity): arenas are allocated once, not per iteration."""
    key = (device, max_games_in_flight, max_explores)
    if key not in _engines:
        _engines[key] = Engine(device, max_games_in_flight, max_explores)
    return _engines[key]


def _leaf_kind(policy) -> int:
    if isinstance(policy, Connect4Net):
        return L.LEAF_NN
    if isinstance(policy, RolloutPolicy):
        return L.LEAF_ROLLOUT
    raise TypeError("policy must be a Connect4Net or a RolloutPolicy: user-defined Policy/Game objects are host code "
                    "and cannot run inside the search kernel")


def gather_experience(cfg: LearningConfig, policy, buffer: ReplayBuffer, seed: int, *, engine: Engine = None,
                      device: int = 0, first_game_index: int = 0, return_stats: bool = False):
    kind = _leaf_kind(policy)
    n = int(cfg.games_per_train)
    eng = engine or engine_for(device, games_in_flight_for(n), int(cfg.rollout_cfg.num_explores))
    if kind == L.LEAF_NN:
        eng.set_weights(policy.blob())
    arrays, stats, _ = eng.gather(cfg.rollout_cfg, kind, first_game_index, n, seed)
    # run_n_games returns a buffer whose ids are 1..n (alpha_zero.rs:201-203)
    arrays["game_ids"] = arrays["game_ids"] - np.uint64(first_game_index)
    worker = ReplayBuffer.from_arrays(n, arrays)
    buffer.keep_last_n_games(cfg.games_to_keep - cfg.games_per_train)
    buffer.extend(worker)
    return stats if return_stats else None


def lr_for_iteration(cfg: LearningConfig, i_iter: int) -> float:
    """alpha_zero.rs:62-70: the last schedule entry whose iteration number is <= i_iter + 1."""
    return [lr for it, lr in cfg.lr_schedule if it <= i_iter + 1][-1]


def train_on(cfg: LearningConfig, dedup: FlatBatch, lr: float, engine: Engine, rng):
    """The epoch loop of alpha_zero (alpha_zero.rs:73-100) on the engine's current weights: per epoch a fresh
    BatchRandSampler(drop_last) over the deduplicated rows, one Adam step per batch, all on the device
    (syn_engine_train; the weights stay in HBM for the next gather).  Returns the per-epoch [pi_loss, v_loss] the
    reference prints (sum over batches * batch_size / n)."""
    n = len(dedup)
    epochs = []
    for _ in range(cfg.num_epochs):
        batches = BatchRandSampler(n, cfg.batch_size, True, rng).all_batches()
        if len(batches) == 0:
            epochs.append([0.0, 0.0])
            continue
        losses, _ = engine.train(dedup.my_bb, dedup.op_bb, dedup.pis, dedup.vs, batches, lr, weight_decay=cfg.weight_decay,
                                 policy_weight=cfg.policy_weight, value_weight=cfg.value_weight, batch_size=cfg.batch_size)
        tot = losses.astype(np.float32).sum(0, dtype=np.float32) * np.float32(cfg.batch_size) / np.float32(n)
        epochs.append([float(tot[0]), float(tot[1])])
    return epochs


def alpha_zero(cfg: LearningConfig, policy: Connect4Net = None, *, engine: Engine = None, device: int = 0, on_iteration=None):
    """`alpha_zero::<Connect4, Connect4Net, 9>(cfg)` (alpha_zero.rs:16-118) with every stage on the GPU: gather ->
    deduplicate -> train, the weights never leaving HBM between iterations.  File outputs (model_{i}.ot, latest_*.npy,
    git metadata) are the caller's business: `on_iteration(i_iter, engine, buffer, dedup, epoch_losses)` is called where
    the reference saves them.  Returns the trained Connect4Net."""
    n = int(cfg.games_per_train)
    eng = engine or engine_for(device, games_in_flight_for(n), int(cfg.rollout_cfg.num_explores))
    policy = policy or Connect4Net.new(cfg.seed)
    eng.set_weights(policy.blob())
    eng.reset_optimizer()
    rng = np.random.default_rng(cfg.seed)
    buffer = ReplayBuffer(256_000)
    for i_iter in range(cfg.num_iterations):
        arrays, _, _ = eng.gather(cfg.rollout_cfg, L.LEAF_NN, 0, n, i_iter)
        buffer.keep_last_n_games(cfg.games_to_keep - cfg.games_per_train)
        buffer.extend(ReplayBuffer.from_arrays(n, arrays))
        dedup = buffer.deduplicate(eng)
        epochs = train_on(cfg, dedup, lr_for_iteration(cfg, i_iter), eng, rng)
        if on_iteration is not None:
            on_iteration(i_iter, eng, buffer, dedup, epochs)
    return Connect4Net.from_blob(eng.get_weights())


class MCTS:
    """`MCTS::exploit` (mcts.rs:111-121): build a tree with `explores` explores and return best_action."""

    @staticmethod
    def exploit(explores: int, cfg: MCTSConfig, policy, game: Connect4, action_selection: int, *, engine: Engine = None,
                device: int = 0, details: bool = False):
        kind = _leaf_kind(policy)
        eng = engine or engine_for(device, 32, int(explores))
        if kind == L.LEAF_NN:
            eng.set_weights(policy.blob())
        rc = RolloutConfig(num_workers=0, num_explores=int(explores), random_actions_until=0, sample_actions_until=0,
                           stop_games_when_solved=False, value_target=ValueTarget.Q(), action=action_selection, mcts_cfg=cfg)
        seed = policy.seed if isinstance(policy, RolloutPolicy) else 0
        out, stats = eng.search(rc, kind, [game.my_bb], [game.op_bb], [seed])
        if details:
            return int(out["best_action"][0]), {k: v[0] for k, v in out.items()}, stats
        return int(out["best_action"][0])
