"""Multi-GPU self-play: games shard across ranks, one process per GPU.

The reference fans games out over `num_workers + 1` OS threads and joins their buffers in worker
order (alpha_zero.rs:132-168); weights reach the workers through `models/model_{i}.ot`
(alpha_zero.rs:192-194).  Here a rank is a worker: contiguous ranges of the global game index
(seeds derive from the index, so the result is independent of the rank count), ONE broadcast of the
weight blob from the trainer rank and ONE gather of experience to it per iteration — both outside
the search, which needs no collective at all.

The collectives live INSIDE the library (`syn_comm_*`, `syn_engine_broadcast_weights`,
`syn_engine_gather_experience`, include/synthesis_b200.h: NCCL on the engine's stream, 72-byte rows on
the wire, features rebuilt on the root), so a Rust host drives them exactly like this module does.
What stays here is host logic: how games are split over ranks (`split_games`) and how the root folds
the gathered rows into its ReplayBuffer.  `gather_rows_over_group` is the same fold over any
torch.distributed group (gloo on CPU): it is what the world_size-2 CPU tests exercise.
"""
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .data import ReplayBuffer

FIELDS = (("game_ids", np.uint64, ()), ("my_bb", np.uint64, ()), ("op_bb", np.uint64, ()), ("height", np.uint8, (9,)),
          ("player", np.uint8, ()), ("states", np.float32, (63,)), ("pis", np.float32, (9,)), ("vs", np.float32, (3,)))


def split_games(num_games: int, world_size: int) -> List[Tuple[int, int]]:
    """(first, count) per rank with the reference's rule `remaining / workers_left`
    (alpha_zero.rs:138,152-153): 1000 games over 7 workers -> 142, 143 x 6."""
    out, first, remaining, left = [], 0, int(num_games), int(world_size)
    for _ in range(world_size):
        n = remaining // left
        out.append((first, n))
        first += n
        remaining -= n
        left -= 1
    assert remaining == 0
    return out


def _torch_dtype(torch, dt):
    # torch has no uint64 arithmetic but can carry the bits as int64
    return {np.uint64: torch.int64, np.uint8: torch.uint8, np.float32: torch.float32}[dt]


def fold_into_buffer(buffer: Optional[ReplayBuffer], merged: Dict[str, np.ndarray], num_games: int, games_to_keep: int, first_game_index: int = 0):
    """What gather_experience does with the joined workers' rows (alpha_zero.rs:164-168): keep_last_n_games, then extend."""
    if buffer is None:
        return
    worker_arrays = dict(merged)
    worker_arrays["game_ids"] = merged["game_ids"] - np.uint64(first_game_index)
    worker = ReplayBuffer.from_arrays(num_games, worker_arrays)
    buffer.keep_last_n_games(games_to_keep - num_games)
    buffer.extend(worker)


def gather_experience_distributed(engine, comm, cfg, leaf_eval_kind: int, num_games: int, seed: int, buffer: Optional[ReplayBuffer], games_to_keep: int,
                                  *, root: int = 0, first_game_index: int = 0):
    """`gather_experience` (alpha_zero.rs:120-169) over `comm.size` GPUs through the library's own collective
    (syn_engine_gather_experience): rank r plays shard r of `split_games`, the root receives every rank's rows in rank
    order = game order and folds them into `buffer`.  Returns (merged arrays, stats) on the root, (None, stats) elsewhere."""
    first, count = split_games(num_games, comm.size)[comm.rank]
    merged, stats = engine.gather_experience(comm, cfg, leaf_eval_kind, first_game_index + first, count, seed, root=root,
                                             capacity_rows=63 * num_games)
    if comm.rank == root:
        fold_into_buffer(buffer, merged, num_games, games_to_keep, first_game_index)
    return merged, stats


def gather_rows_over_group(play_fn: Callable[[int, int], Dict[str, np.ndarray]], num_games: int, buffer: Optional[ReplayBuffer],
                           games_to_keep: int, *, group=None, device=None, dst: int = 0, first_game_index: int = 0):
    """The same fan-out and join over a torch.distributed group of any backend (gloo on CPU): shard `num_games`, play
    each shard with `play_fn(first, count)` (which returns the syn_experience arrays of that shard with GLOBAL 1-based game
    ids), gather all rows to rank `dst` in rank order and fold them into `buffer`.  Returns the merged arrays on `dst`,
    None elsewhere.  The GPU path does not use this (see gather_experience_distributed): it exists so that the host logic
    can be tested with world_size 2 on a machine without GPUs."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = split_games(num_games, world)[rank]
    arrays = play_fn(first_game_index + first, count) if count > 0 else {n: np.zeros((0,) + sh, dt) for n, dt, sh in FIELDS}
    rows = len(arrays["vs"])
    dev = device if device is not None else torch.device("cpu")
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([rows], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts_h = counts.cpu().numpy()
    cap = int(counts_h.max())
    merged = {}
    for name, dt, shape in FIELDS:
        a = np.ascontiguousarray(arrays[name], dtype=dt).reshape((rows,) + shape)
        if dt == np.uint64:
            a = a.view(np.int64)
        t = torch.zeros((cap,) + shape, dtype=_torch_dtype(torch, dt), device=dev)
        if rows:
            t[:rows] = torch.from_numpy(a).to(dev)
        if rank == dst:
            recv = [torch.zeros_like(t) for _ in range(world)]
            dist.gather(t, recv, dst=dst, group=group)
            parts = [recv[r][: int(counts_h[r])].cpu().numpy() for r in range(world)]
            m = np.concatenate(parts) if parts else a
            merged[name] = m.view(np.uint64) if dt == np.uint64 else m
        else:
            dist.gather(t, None, dst=dst, group=group)
    if rank != dst:
        return None
    fold_into_buffer(buffer, merged, num_games, games_to_keep, first_game_index)
    return merged


def alpha_zero_distributed(cfg, engine, comm, *, policy=None, root: int = 0, on_iteration=None):
    """`alpha_zero` (alpha_zero.rs:16-118) over `comm.size` GPUs, one engine per rank: every iteration the trainer rank
    broadcasts its current weights (ONE collective on the engine's stream — replaces the model_{i}.ot round trip; the
    library synchronises before any rank searches with them), all ranks play their shard of `games_per_train`, the rows are
    gathered to the trainer (ONE collective), which deduplicates and trains on its own GPU.  Returns the trained
    Connect4Net on the root, None elsewhere."""
    from . import _lib as L
    from .alpha_zero import lr_for_iteration, train_on
    from .policies import Connect4Net

    if comm.rank == root:
        engine.set_weights((policy or Connect4Net.new(cfg.seed)).blob())
        engine.reset_optimizer()
    rng = np.random.default_rng(cfg.seed)
    buffer = ReplayBuffer(256_000) if comm.rank == root else None
    for i_iter in range(cfg.num_iterations):
        engine.broadcast_weights(comm, None, root=root)
        gather_experience_distributed(engine, comm, cfg.rollout_cfg, L.LEAF_NN, cfg.games_per_train, i_iter, buffer, cfg.games_to_keep, root=root)
        if comm.rank == root:
            dedup = buffer.deduplicate(engine)
            epochs = train_on(cfg, dedup, lr_for_iteration(cfg, i_iter), engine, rng)
            if on_iteration is not None:
                on_iteration(i_iter, engine, buffer, dedup, epochs)
    return Connect4Net.from_blob(engine.get_weights()) if comm.rank == root else None
