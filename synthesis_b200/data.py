"""`ReplayBuffer`, the experience-buffer output format of the reference
(synthesis/src/data.rs:106-194), as numpy struct-of-arrays.  Same fields (`game_ids`, `games`
— split into the four fields of `Connect4` —, `states`, `pis`, `vs`), same counters
(`game_id`, `steps`) and the same `new_game / add / extend / keep_last_n_games` behaviour.
"""
import numpy as np

from .connect4 import Connect4

_FIELDS = (("game_ids", np.uint64, ()), ("my_bb", np.uint64, ()), ("op_bb", np.uint64, ()), ("height", np.uint8, (9,)),
           ("player", np.uint8, ()), ("states", np.float32, (63,)), ("pis", np.float32, (9,)), ("vs", np.float32, (3,)))


def rows_from_bitboards(my_bb, op_bb):
    """What a ReplayBuffer row holds beyond (game id, bitboards, pi, v) is a function of the bitboards: Connect4::height,
    Connect4::player (connect4.rs:108-114, 191) and Game::features (connect4.rs:237-258).  Rows reach the host as 72 bytes
    (syn_engine_gather_experience, bench.py's e2e leg); this rebuilds the other three columns, vectorised.
    Returns (height uint8[n][9], player uint8[n], states float32[n][63])."""
    my = np.ascontiguousarray(my_bb, np.uint64).reshape(-1)
    op = np.ascontiguousarray(op_bb, np.uint64).reshape(-1)
    n = my.size
    cell = (np.arange(7, dtype=np.uint64)[:, None] + np.uint64(7) * np.arange(9, dtype=np.uint64)[None, :]).reshape(-1)  # row*9+col -> bit row + 7*col
    mine = ((my[:, None] >> cell[None, :]) & np.uint64(1)).astype(bool)
    theirs = ((op[:, None] >> cell[None, :]) & np.uint64(1)).astype(bool)
    occ = (mine | theirs).reshape(n, 7, 9)
    height = occ.sum(axis=1).astype(np.uint8)                      # stones stack from the bottom
    player = (occ.reshape(n, -1).sum(axis=1) & 1).astype(np.uint8)  # Red (0) moves first
    states = np.where(mine, np.float32(1.0), np.where(theirs, np.float32(-1.0), np.float32(-0.1))).astype(np.float32).reshape(n, 7, 9)
    rows = np.arange(7, dtype=np.uint8)[None, :, None]
    nxt = (rows == height[:, None, :]) & ~occ                      # the next playable cell of every column with room
    states[nxt] = np.float32(0.1)
    return height, player, states.reshape(n, 63)


class BatchRandSampler:
    """data.rs:6-64: batches of a random permutation of 0..n, the last partial batch dropped when `drop_last`.  Yields
    index arrays (the reference index_selects its three tensors with them).  The permutation comes from `rng`
    (numpy Generator); torch's randperm stream (tch::manual_seed, alpha_zero.rs:28) is not reproduced."""

    def __init__(self, n: int, batch_size: int, drop_last: bool, rng):
        self.inds = rng.permutation(int(n)).astype(np.uint32)
        self.size, self.batch_size, self.index, self.drop_last = int(n), int(batch_size), 0, bool(drop_last)

    def __iter__(self):
        return self

    def __next__(self):
        nxt = min(self.index + self.batch_size, self.size)
        if self.index >= self.size or (self.drop_last and nxt - self.index < self.batch_size):
            raise StopIteration
        out = self.inds[self.index:nxt]
        self.index = nxt
        return out

    def all_batches(self) -> np.ndarray:
        """Every remaining full batch as one [n_batches][batch_size] array (what syn_engine_train takes)."""
        full = (self.size - self.index) // self.batch_size
        out = self.inds[self.index:self.index + full * self.batch_size].reshape(full, self.batch_size)
        self.index += full * self.batch_size
        return out


class FlatBatch:
    """data.rs:80-85: `states` [n][1][7][9] (Connect4::DIMS, connect4.rs:235), `pis` [n][9], `vs` [n][3]; the positions
    themselves (my_bb, op_bb) ride along so that the device learner can synthesise `states` on chip."""

    def __init__(self, states, pis, vs, my_bb=None, op_bb=None):
        self.states, self.pis, self.vs, self.my_bb, self.op_bb = states, pis, vs, my_bb, op_bb

    def __len__(self):
        return len(self.vs)


class ReplayBuffer:
    def __init__(self, n: int = 0):  # data.rs:116-126 (n is only a capacity hint there too)
        self.game_id = 0
        self.steps = 0
        for name, dt, shape in _FIELDS:
            setattr(self, name, np.zeros((0,) + shape, dtype=dt))

    # ---- data.rs:128-149
    def new_game(self):
        self.game_id += 1

    def total_games_played(self) -> int:
        return self.game_id

    def curr_games(self) -> int:
        if len(self.game_ids) == 0:
            return 0
        return int(1 + np.count_nonzero(self.game_ids[1:] != self.game_ids[:-1]))

    def total_steps(self) -> int:
        return self.steps

    def curr_steps(self) -> int:
        return len(self.vs)

    @property
    def games(self):
        """`Vec<G>` view: rebuilds the Connect4 objects on demand."""
        return [Connect4(int(m), int(o), [int(x) for x in h], int(p))
                for m, o, h, p in zip(self.my_bb, self.op_bb, self.height, self.player)]

    # ---- data.rs:151-158
    def add(self, game: Connect4, pi, v):
        row = dict(game_ids=self.game_id, my_bb=game.my_bb, op_bb=game.op_bb, height=game.height, player=game.player_,
                   states=game.features().reshape(63), pis=pi, vs=v)
        self.steps += 1
        for name, dt, shape in _FIELDS:
            arr = np.asarray(row[name], dtype=dt).reshape((1,) + shape)
            setattr(self, name, np.concatenate([getattr(self, name), arr]))

    # ---- data.rs:160-170
    def extend(self, other: "ReplayBuffer"):
        self.steps += other.steps
        start = self.game_id
        self.game_ids = np.concatenate([self.game_ids, other.game_ids + np.uint64(start)])
        self.game_id += other.game_id
        for name, dt, shape in _FIELDS[1:]:
            setattr(self, name, np.concatenate([getattr(self, name), getattr(other, name)]))
            setattr(other, name, np.zeros((0,) + shape, dtype=dt))  # drain(..)
        # like the reference, other.game_ids is left untouched (only the drained Vecs are emptied)

    # ---- data.rs:172-194
    def keep_last_n_games(self, n: int):
        if self.game_id <= n:
            return
        min_game_id = self.game_id - n
        remove = int(np.searchsorted(self.game_ids, np.uint64(min_game_id), side="left")) if self._sorted() else self._prefix(min_game_id)
        if remove:
            for name, _, _ in _FIELDS:
                setattr(self, name, getattr(self, name)[remove:])
            assert len(self.game_ids) == 0 or self.game_ids[0] >= min_game_id

    def _sorted(self) -> bool:
        return bool(np.all(self.game_ids[1:] >= self.game_ids[:-1]))

    def _prefix(self, min_game_id: int) -> int:
        ge = np.nonzero(self.game_ids >= np.uint64(min_game_id))[0]
        return int(ge[0]) if len(ge) else len(self.game_ids)

    # ---- data.rs:196-235
    def deduplicate(self, engine) -> "FlatBatch":
        """One row per distinct position, policy and value targets averaged over the rows that hold it.  Runs on the
        GPU of `engine` (syn_engine_deduplicate); values are bit-identical to the reference's, rows come in order of
        first occurrence where the reference's order is HashMap-random."""
        out, _ = engine.deduplicate(self.my_bb, self.op_bb, self.pis, self.vs)
        return FlatBatch(out["states"].reshape(-1, 1, 7, 9), out["pis"], out["vs"], out["my_bb"], out["op_bb"])

    # ---- construction from the C ABI's syn_experience arrays
    @staticmethod
    def from_arrays(games_played: int, arrays: dict) -> "ReplayBuffer":
        """A worker buffer as run_n_games returns it (alpha_zero.rs:181-209): game ids 1..games_played."""
        b = ReplayBuffer()
        n = len(arrays["vs"])
        b.game_id = int(games_played)
        b.steps = n
        if any(k not in arrays for k in ("height", "player", "states")):  # compact rows: rebuild what the bitboards imply
            arrays = dict(arrays)
            h, p, st = rows_from_bitboards(arrays["my_bb"], arrays["op_bb"])
            arrays.setdefault("height", h); arrays.setdefault("player", p); arrays.setdefault("states", st)
        for name, dt, shape in _FIELDS:
            setattr(b, name, np.ascontiguousarray(arrays[name], dtype=dt).reshape((n,) + shape))
        return b
