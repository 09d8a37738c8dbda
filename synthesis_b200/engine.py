"""`Engine`: one B200, one `syn_engine` (include/synthesis_b200.h).  Thin, typed access to the C ABI;
numpy arrays in and out (or raw device pointers for the multi-GPU path).
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .config import RolloutConfig
from .data import ReplayBuffer


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Comm:
    """One rank of a multi-GPU job (`syn_comm`, include/synthesis_b200.h): NCCL inside the library.  The host's only task is
    to hand rank 0's 128-byte id to the other ranks; `Comm.from_torch` does that over an initialised torch.distributed
    group (any backend — torch is plumbing here), `Comm(id, n, rank, device)` takes the bytes from anywhere else."""

    def __init__(self, comm_id: bytes, n_ranks: int, rank: int, device: int):
        self._lib = L.load()
        if len(comm_id) != L.COMM_ID_BYTES:
            raise ValueError(f"communicator id must be {L.COMM_ID_BYTES} bytes")
        buf = (C.c_uint8 * L.COMM_ID_BYTES).from_buffer_copy(comm_id)
        h = C.c_void_p()
        L.check(self._lib.syn_comm_create(buf, int(n_ranks), int(rank), int(device), C.byref(h)))
        self._h = h
        self.rank, self.size, self.device = int(rank), int(n_ranks), int(device)

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * L.COMM_ID_BYTES)()
        L.check(L.load().syn_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch(cls, device: int, group=None):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        backend = dist.get_backend(group)
        dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
        t = torch.zeros(L.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0, group=group)
        return cls(bytes(t.cpu().numpy().tobytes()), world, rank, device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.syn_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, device: int = 0, max_games_in_flight: int = 4736, max_explores: int = 1600):
        self._lib = L.load()
        h = C.c_void_p()
        L.check(self._lib.syn_engine_create(int(device), int(max_games_in_flight), int(max_explores), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.max_explores = int(max_explores)
        self.max_games_in_flight = int(max_games_in_flight)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.syn_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- weights (replaces vs.load(models/<name>.ot), alpha_zero.rs:192-194)
    def set_weights(self, blob):
        """blob: numpy float32[30492] (host) or an int device pointer to 30492 floats."""
        if isinstance(blob, int):
            L.check(self._lib.syn_engine_set_weights(self._h, C.c_void_p(blob), L.N_WEIGHTS))
            return
        b = np.ascontiguousarray(blob, dtype=np.float32).reshape(-1)
        L.check(self._lib.syn_engine_set_weights(self._h, _ptr(b), b.size))

    def set_opponent_weights(self, blob):
        """players[1]'s network in matches between two Connect4Net players (eval_against_old with p1 != p2,
        evaluator.rs:129-160); None: both players use set_weights' network again."""
        if blob is None:
            L.check(self._lib.syn_engine_set_opponent_weights(self._h, None, 0))
            return
        b = np.ascontiguousarray(blob, dtype=np.float32).reshape(-1)
        L.check(self._lib.syn_engine_set_opponent_weights(self._h, _ptr(b), b.size))

    def set_mlp_mode(self, mode):
        """How Connect4Net is evaluated (include/synthesis_b200.h): 3 / True = auto (default: the fast single-fp16 chain while
        its measured error stays below a quarter of the 1e-3 tolerance, else the split chain), 2 = split-fp16 operands
        (fp32-grade), 1 = single fp16 operands, 0 / False = fp32 CUDA-core kernel."""
        L.check(self._lib.syn_engine_set_mlp_mode(self._h, 3 if mode is True else int(mode)))

    def mlp_in_use(self):
        """(chain the next launch uses: 0 / 1 / 2, the fast chain's measured error in units of the tolerance or -1)."""
        chain, ratio = C.c_int(), C.c_float()
        L.check(self._lib.syn_engine_mlp_in_use(self._h, C.byref(chain), C.byref(ratio)))
        return chain.value, float(ratio.value)

    def debug_counters(self):
        """Per-warp phase clocks of the last thread-per-game launch (profiling aid)."""
        out = np.zeros(11, np.uint64)
        L.check(self._lib.syn_engine_debug_counters(self._h, _ptr(out), 11))
        names = ("t_advance", "t_teamwait", "t_mlp", "t_finish", "rounds", "leaves", "t_total", "x_select", "x_expand", "x_eom", "x_backprop")
        return {k: int(v) for k, v in zip(names, out)}

    def launch_geometry(self, num_games: int, leaf_eval_kind: int):
        """(CTAs, most games per CTA, lanes per game) a gather of num_games games would be launched as (diagnostic)."""
        ctas, per, lanes = C.c_uint32(), C.c_uint32(), C.c_uint32()
        L.check(self._lib.syn_engine_launch_geometry(self._h, int(num_games), int(leaf_eval_kind), C.byref(ctas), C.byref(per), C.byref(lanes)))
        return ctas.value, per.value, lanes.value

    def set_group_lanes(self, lanes: int):
        L.check(self._lib.syn_engine_set_group_lanes(self._h, int(lanes)))

    # ---- gather (replaces gather_experience/run_n_games/run_game, alpha_zero.rs:120-268)
    @staticmethod
    def _alloc(rows: int, trace: bool):
        a = dict(game_ids=np.zeros(rows, np.uint64), my_bb=np.zeros(rows, np.uint64), op_bb=np.zeros(rows, np.uint64),
                 height=np.zeros((rows, 9), np.uint8), player=np.zeros(rows, np.uint8),
                 states=np.zeros((rows, 63), np.float32), pis=np.zeros((rows, 9), np.float32), vs=np.zeros((rows, 3), np.float32))
        t = dict(action=np.zeros(rows, np.uint8), tree_nodes=np.zeros(rows, np.uint32),
                 child_visits=np.zeros((rows, 9), np.float32)) if trace else None
        return a, t

    def gather(self, cfg: RolloutConfig, leaf_eval_kind: int, first_game_index: int, num_games: int, seed: int, trace: bool = False):
        """Plays games [first_game_index, first_game_index+num_games).  Returns (arrays, stats, trace):
        arrays = dict of the syn_experience fields trimmed to the rows written."""
        ccfg = cfg.to_c(leaf_eval_kind)
        rows = L.MAX_TURNS * int(num_games)
        a, t = self._alloc(rows, trace)
        exp = L.SynExperience()
        exp.capacity = rows
        for k in ("game_ids", "my_bb", "op_bb", "height", "player", "states", "pis", "vs"):
            setattr(exp, k, a[k].ctypes.data)
        if trace:
            L.check(self._lib.syn_engine_set_trace(self._h, _ptr(t["action"]), _ptr(t["tree_nodes"]), _ptr(t["child_visits"])))
        else:
            L.check(self._lib.syn_engine_set_trace(self._h, None, None, None))
        stats = L.SynStats()
        try:
            L.check(self._lib.syn_engine_gather(self._h, C.byref(ccfg), int(first_game_index), int(num_games), int(seed),
                                                C.byref(exp), C.byref(stats)))
        finally:
            self._lib.syn_engine_set_trace(self._h, None, None, None)
        n = int(exp.len)
        a = {k: v[:n] for k, v in a.items()}
        if trace:
            t = {k: v[:n] for k, v in t.items()}
        return a, stats.as_dict(), t

    # ---- multi-GPU (replaces the thread fan-out / join of gather_experience and the model_{i}.ot hand-off, alpha_zero.rs:132-168, 192-194)
    def broadcast_weights(self, comm: "Comm", blob=None, root: int = 0):
        """ONE NCCL broadcast of the weight blob from rank `root` into every rank's engine.  blob (root only): numpy
        float32[30492], an int device pointer, or None for the root engine's current weights."""
        if blob is None or comm.rank != root:
            p = None
        elif isinstance(blob, int):
            p = C.c_void_p(blob)
        else:
            b = np.ascontiguousarray(blob, dtype=np.float32).reshape(-1)
            p = _ptr(b)
        L.check(self._lib.syn_engine_broadcast_weights(self._h, comm._h, p, L.N_WEIGHTS, int(root)))

    def gather_experience(self, comm: "Comm", cfg: RolloutConfig, leaf_eval_kind: int, first_game_index: int, num_games: int, seed: int,
                          root: int = 0, capacity_rows: int = None, exp: "L.SynExperience" = None):
        """Every rank plays its shard [first_game_index, +num_games); rank `root` receives all ranks' rows in rank order.
        Returns (arrays or None, stats): arrays on the root only.  `exp` (root): caller-provided destinations (pinned host or
        device pointers) instead of fresh numpy arrays; then arrays is None and exp.len holds the row count."""
        ccfg = cfg.to_c(leaf_eval_kind)
        stats = L.SynStats()
        if comm.rank != root:
            L.check(self._lib.syn_engine_gather_experience(self._h, comm._h, int(root), C.byref(ccfg), int(first_game_index), int(num_games), int(seed),
                                                           None, C.byref(stats)))
            return None, stats.as_dict()
        if exp is not None:
            L.check(self._lib.syn_engine_gather_experience(self._h, comm._h, int(root), C.byref(ccfg), int(first_game_index), int(num_games), int(seed),
                                                           C.byref(exp), C.byref(stats)))
            return None, stats.as_dict()
        rows = int(capacity_rows) if capacity_rows is not None else L.MAX_TURNS * int(num_games) * comm.size
        a, _ = self._alloc(rows, False)
        exp = L.SynExperience()
        exp.capacity = rows
        for k in ("game_ids", "my_bb", "op_bb", "height", "player", "states", "pis", "vs"):
            setattr(exp, k, a[k].ctypes.data)
        L.check(self._lib.syn_engine_gather_experience(self._h, comm._h, int(root), C.byref(ccfg), int(first_game_index), int(num_games), int(seed),
                                                       C.byref(exp), C.byref(stats)))
        n = int(exp.len)
        return {k: v[:n] for k, v in a.items()}, stats.as_dict()

    def gather_launch(self, cfg: RolloutConfig, leaf_eval_kind: int, first_game_index: int, num_games: int, seed: int):
        ccfg = cfg.to_c(leaf_eval_kind)
        L.check(self._lib.syn_engine_gather_launch(self._h, C.byref(ccfg), int(first_game_index), int(num_games), int(seed)))

    def gather_wait(self, exp: "L.SynExperience" = None):
        """Waits for the launched gather.  exp = None leaves the experience in HBM (bench `value` leg)."""
        stats = L.SynStats()
        L.check(self._lib.syn_engine_gather_wait(self._h, C.byref(exp) if exp is not None else None, C.byref(stats)))
        return stats.as_dict()

    def gather_into(self, cfg: RolloutConfig, leaf_eval_kind: int, first_game_index: int, num_games: int, seed: int, exp: "L.SynExperience"):
        """gather with caller-owned (host or device) destination arrays already set in `exp`."""
        ccfg = cfg.to_c(leaf_eval_kind)
        stats = L.SynStats()
        L.check(self._lib.syn_engine_gather(self._h, C.byref(ccfg), int(first_game_index), int(num_games), int(seed),
                                            C.byref(exp), C.byref(stats)))
        return stats.as_dict()

    # ---- search (replaces MCTS::exploit / FrozenMCTS::exploit on a batch of roots)
    def search(self, cfg: RolloutConfig, leaf_eval_kind: int, my_bb, op_bb, seeds, tree_kind: int = L.TREE_MCTS):
        my = np.ascontiguousarray(my_bb, dtype=np.uint64).reshape(-1)
        op = np.ascontiguousarray(op_bb, dtype=np.uint64).reshape(-1)
        sd = np.ascontiguousarray(seeds, dtype=np.uint64).reshape(-1)
        n = my.size
        if op.size != n or sd.size != n:
            raise ValueError("my_bb, op_bb and seeds must have the same length")
        ccfg = cfg.to_c(leaf_eval_kind)
        out = dict(child_visits=np.zeros((n, 9), np.float32), child_solution=np.zeros((n, 9), np.uint8),
                   root_q=np.zeros((n, 3), np.float32), root_solution=np.zeros(n, np.uint8),
                   best_action=np.zeros(n, np.uint8), num_nodes=np.zeros(n, np.uint32))
        stats = L.SynStats()
        L.check(self._lib.syn_engine_search(self._h, C.byref(ccfg), int(tree_kind), _ptr(my), _ptr(op), _ptr(sd), n,
                                            _ptr(out["child_visits"]), _ptr(out["child_solution"]), _ptr(out["root_q"]),
                                            _ptr(out["root_solution"]), _ptr(out["best_action"]), _ptr(out["num_nodes"]),
                                            C.byref(stats)))
        return out, stats.as_dict()

    # ---- evaluation matches (replaces the game loops of evaluator.rs:129-228)
    def match(self, players, seeds, explores=None, trace: bool = True):
        """players: two `Player`s, players[0] moves first.  seeds[i] seeds match i's RolloutPolicy stream.
        explores: optional int array [n][2] overriding the players' num_explores per match.
        Returns (dict(result, n_moves, moves[, tree_nodes, child_visits]), stats)."""
        sd = np.ascontiguousarray(seeds, dtype=np.uint64).reshape(-1)
        n = sd.size
        pc = (L.SynPlayerCfg * 2)(players[0].to_c(), players[1].to_c())
        ex = None
        if explores is not None:
            ex = np.ascontiguousarray(explores, dtype=np.uint32).reshape(n, 2)
        out = dict(result=np.zeros(n, np.float32), n_moves=np.zeros(n, np.uint8), moves=np.zeros((n, L.MAX_TURNS), np.uint8))
        if trace:
            out["tree_nodes"] = np.zeros((n, L.MAX_TURNS), np.uint32)
            out["child_visits"] = np.zeros((n, L.MAX_TURNS, 9), np.float32)
        stats = L.SynStats()
        L.check(self._lib.syn_engine_match(self._h, pc, _ptr(sd), _ptr(ex), n, _ptr(out["result"]), _ptr(out["n_moves"]),
                                           _ptr(out["moves"]), _ptr(out.get("tree_nodes")), _ptr(out.get("child_visits")),
                                           C.byref(stats)))
        return out, stats.as_dict()

    # ---- ReplayBuffer::deduplicate (data.rs:196-235)
    def deduplicate(self, my_bb, op_bb, pis, vs, states: bool = True):
        """Merges rows with equal positions; sums in row order in f32, like the reference's loop.  Returns
        (dict(states, pis, vs, my_bb, op_bb, num), stats); rows are ordered by first occurrence."""
        my = np.ascontiguousarray(my_bb, dtype=np.uint64).reshape(-1)
        op = np.ascontiguousarray(op_bb, dtype=np.uint64).reshape(-1)
        n = my.size
        pi = np.ascontiguousarray(pis, dtype=np.float32).reshape(n, 9)
        v = np.ascontiguousarray(vs, dtype=np.float32).reshape(n, 3)
        if op.size != n:
            raise ValueError("my_bb and op_bb must have the same length")
        out = dict(states=np.zeros((n, 63), np.float32) if states else None, pis=np.zeros((n, 9), np.float32),
                   vs=np.zeros((n, 3), np.float32), my_bb=np.zeros(n, np.uint64), op_bb=np.zeros(n, np.uint64),
                   num=np.zeros(n, np.uint32))
        fb = L.SynFlatBatch()
        fb.capacity = n
        for k, a in out.items():
            setattr(fb, k, a.ctypes.data if a is not None else None)
        stats = L.SynStats()
        L.check(self._lib.syn_engine_deduplicate(self._h, _ptr(my), _ptr(op), _ptr(pi), _ptr(v), n, C.byref(fb), C.byref(stats)))
        u = int(fb.len)
        return {k: (a[:u] if a is not None else None) for k, a in out.items()}, stats.as_dict()

    # ---- the learner's batch loop (alpha_zero.rs:73-92)
    def train(self, my_bb, op_bb, pis, vs, batch_index, lr: float, weight_decay: float = 0.0, policy_weight: float = 1.0,
              value_weight: float = 1.0, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, batch_size: int = 32):
        """One optimizer step per row of `batch_index` ([n_batches][batch_size] row indices) on the engine's current
        weights.  Returns (losses [n_batches][2] = pi_loss, v_loss per step, stats)."""
        my = np.ascontiguousarray(my_bb, dtype=np.uint64).reshape(-1)
        op = np.ascontiguousarray(op_bb, dtype=np.uint64).reshape(-1)
        n = my.size
        pi = np.ascontiguousarray(pis, dtype=np.float32).reshape(n, 9)
        v = np.ascontiguousarray(vs, dtype=np.float32).reshape(n, 3)
        bi = np.ascontiguousarray(batch_index, dtype=np.uint32).reshape(-1, int(batch_size))
        cfg = L.SynTrainCfg(lr=lr, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, policy_weight=policy_weight,
                            value_weight=value_weight, batch_size=int(batch_size))
        losses = np.zeros((bi.shape[0], 2), np.float32)
        stats = L.SynStats()
        L.check(self._lib.syn_engine_train(self._h, C.byref(cfg), _ptr(my), _ptr(op), _ptr(pi), _ptr(v), n, _ptr(bi), bi.shape[0],
                                           _ptr(losses), C.byref(stats)))
        return losses, stats.as_dict()

    def reset_optimizer(self):
        L.check(self._lib.syn_engine_reset_optimizer(self._h))

    def get_weights(self) -> np.ndarray:
        blob = np.zeros(L.N_WEIGHTS, np.float32)
        L.check(self._lib.syn_engine_get_weights(self._h, _ptr(blob), L.N_WEIGHTS))
        return blob

    # ---- Policy::eval for Connect4Net on a batch
    def eval(self, my_bb, op_bb):
        my = np.ascontiguousarray(my_bb, dtype=np.uint64).reshape(-1)
        op = np.ascontiguousarray(op_bb, dtype=np.uint64).reshape(-1)
        n = my.size
        logits = np.zeros((n, 9), np.float32)
        probs = np.zeros((n, 3), np.float32)
        L.check(self._lib.syn_engine_eval(self._h, _ptr(my), _ptr(op), n, _ptr(logits), _ptr(probs)))
        return logits, probs

    # ---- Game::step on move lists
    def play(self, move_lists, features: bool = True):
        n = len(move_lists)
        stride = max(1, max((len(m) for m in move_lists), default=1))
        moves = np.zeros((n, stride), np.uint8)
        nm = np.zeros(n, np.uint32)
        for i, m in enumerate(move_lists):
            nm[i] = len(m)
            moves[i, :len(m)] = np.asarray(m, dtype=np.uint8)
        out = dict(my_bb=np.zeros(n, np.uint64), op_bb=np.zeros(n, np.uint64), height=np.zeros((n, 9), np.uint8),
                   legal_lo=np.zeros(n, np.uint8), legal_hi=np.zeros(n, np.uint8), status=np.zeros(n, np.uint8),
                   features=np.zeros((n, 63), np.float32) if features else None)
        L.check(self._lib.syn_engine_play(self._h, _ptr(moves), _ptr(nm), stride, n, _ptr(out["my_bb"]), _ptr(out["op_bb"]),
                                          _ptr(out["height"]), _ptr(out["legal_lo"]), _ptr(out["legal_hi"]), _ptr(out["status"]),
                                          _ptr(out["features"])))
        out["legal_mask"] = out["legal_lo"].astype(np.uint32) | (out["legal_hi"].astype(np.uint32) << 8)
        return out
