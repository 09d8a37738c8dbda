"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (the checker, never the product)."""
import ctypes as C
import os

import numpy as np

from synthesis_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAG_LEGACY, FLAG_LIBM, FLAG_NO_CACHE = 1, 2, 4
EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        l = self.lib
        l.orc_expf.restype = C.c_float
        l.orc_expf.argtypes = [C.c_float]
        l.orc_logf.restype = C.c_float
        l.orc_logf.argtypes = [C.c_float]
        l.orc_stream_seed.restype = C.c_uint64
        l.orc_stream_seed.argtypes = [C.c_uint64, C.c_uint64, C.c_uint]
        l.orc_outcome_value.restype = C.c_float
        l.orc_outcome_from_f32.argtypes = [C.c_float]
        l.orc_outcome_from_f32.restype = C.c_uint8
        l.orc_outcome_reversed.restype = C.c_uint8
        l.orc_keep_last_n_games_prefix.restype = C.c_size_t
        l.orc_keep_last_n_games_prefix.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64]
        l.orc_search.argtypes = [C.POINTER(L.SynRolloutCfg), C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_uint32] + [C.c_void_p] * 6 + [C.POINTER(L.SynStats)]
        l.orc_gather.argtypes = [C.POINTER(L.SynRolloutCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32,
                                 C.c_int, C.c_uint32, C.POINTER(L.SynExperience), C.POINTER(L.SynStats), C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_gather_reference.argtypes = [C.POINTER(L.SynRolloutCfg), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                           C.POINTER(L.SynExperience), C.POINTER(L.SynStats), C.c_void_p, C.c_void_p]
        l.orc_match.argtypes = [C.POINTER(L.SynPlayerCfg), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32] + \
                               [C.c_void_p] * 5 + [C.POINTER(L.SynStats)]
        l.orc_match2.argtypes = [C.POINTER(L.SynPlayerCfg), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_uint32] + [C.c_void_p] * 5 + [C.POINTER(L.SynStats)]
        l.orc_mlp_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        l.orc_c4_play.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 8
        l.orc_c4_won.argtypes = [C.c_uint64]
        l.orc_ttt_kat.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_stdrng_words.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
        l.orc_gen_range_u8.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        l.orc_weighted_index.argtypes = [C.c_uint64, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
        l.orc_dirichlet.argtypes = [C.c_uint64, C.c_float, C.c_int, C.c_uint32, C.c_void_p]
        l.orc_normal.argtypes = [C.c_uint64, C.c_float, C.c_float, C.c_uint32, C.c_void_p]
        l.orc_chacha_block.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        l.orc_seed_key.argtypes = [C.c_uint64, C.c_void_p]
        l.orc_outcome_cmp.argtypes = [C.c_uint8, C.c_uint8]
        l.orc_outcome_reversed.argtypes = [C.c_uint8]
        l.orc_outcome_value.argtypes = [C.c_uint8]
        l.orc_c4_player.argtypes = [C.c_uint64, C.c_uint64]

    # ---- streams
    def stdrng_words(self, seed, n):
        out = np.zeros(n, np.uint32)
        self.lib.orc_stdrng_words(seed, n, _p(out))
        return out

    def gen_range_u8(self, seed, n, count):
        out = np.zeros(count, np.uint8)
        self.lib.orc_gen_range_u8(seed, n, count, _p(out))
        return out

    def weighted_index(self, seed, w, count):
        w = np.ascontiguousarray(w, np.float32)
        out = np.zeros(count, np.int32)
        self.lib.orc_weighted_index(seed, _p(w), len(w), count, _p(out))
        return out

    def dirichlet(self, seed, alpha, k, count):
        out = np.zeros((count, k), np.float32)
        self.lib.orc_dirichlet(seed, alpha, k, count, _p(out))
        return out

    def normal(self, seed, mean, std, count):
        out = np.zeros(count, np.float32)
        self.lib.orc_normal(seed, mean, std, count, _p(out))
        return out

    # ---- Connect4
    def c4_play(self, moves):
        m = np.ascontiguousarray(moves, np.uint8)
        my, op = C.c_uint64(), C.c_uint64()
        h = np.zeros(9, np.uint8)
        lm, st = C.c_uint32(), C.c_uint8()
        f = np.zeros(63, np.float32)
        rw = C.c_float()
        so = np.zeros(max(1, len(m)), np.uint8)
        rc = self.lib.orc_c4_play(_p(m), len(m), C.byref(my), C.byref(op), _p(h), C.byref(lm), C.byref(st), _p(f), C.byref(rw), _p(so))
        if rc:
            return None
        return dict(my_bb=my.value, op_bb=op.value, height=h, legal_mask=lm.value, status=st.value, features=f,
                    reward_to_move=rw.value, step_over=so[:len(m)])

    # ---- trees
    def search(self, ccfg, my_bb, op_bb, seed, tree_kind=0, weights=None, callback=None, flags=0):
        cv, cs, rq = np.zeros(9, np.float32), np.zeros(9, np.uint8), np.zeros(3, np.float32)
        rs, ba, nn = C.c_uint8(), C.c_uint8(), C.c_uint32()
        st = L.SynStats()
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        cb = EVAL_FN(callback) if callback else None
        rc = self.lib.orc_search(C.byref(ccfg), tree_kind, int(my_bb), int(op_bb), int(seed), _p(w), C.cast(cb, C.c_void_p) if cb else None,
                                 None, flags, _p(cv), _p(cs), _p(rq), C.cast(C.byref(rs), C.c_void_p), C.cast(C.byref(ba), C.c_void_p),
                                 C.cast(C.byref(nn), C.c_void_p), C.byref(st))
        assert rc == 0, rc
        return dict(child_visits=cv, child_solution=cs, root_q=rq, root_solution=rs.value, best_action=ba.value, num_nodes=nn.value), st.as_dict()

    def match(self, players, seed, explores2=None, weights=None, callback=None, flags=0, weights2=None, mover=None):
        """One evaluation match (evaluator.rs:129-228); players = two objects with .to_c() -> SynPlayerCfg.
        weights2 = players[1]'s network (two-network eval_against_old); mover = a ctypes.c_uint32 the oracle sets to the
        index of the player to move before every search (for callbacks that stand in for two networks)."""
        pc = (L.SynPlayerCfg * 2)(players[0].to_c(), players[1].to_c())
        res, nm = C.c_float(), C.c_uint8()
        moves, nodes, cv = np.zeros(63, np.uint8), np.zeros(63, np.uint32), np.zeros((63, 9), np.float32)
        ex = None if explores2 is None else np.ascontiguousarray(explores2, np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        w2 = None if weights2 is None else np.ascontiguousarray(weights2, np.float32)
        cb = EVAL_FN(callback) if callback else None
        st = L.SynStats()
        rc = self.lib.orc_match2(pc, int(seed), _p(ex), _p(w), _p(w2), C.cast(C.byref(mover), C.c_void_p) if mover is not None else None,
                                 C.cast(cb, C.c_void_p) if cb else None, None, flags,
                                 C.cast(C.byref(res), C.c_void_p), C.cast(C.byref(nm), C.c_void_p), _p(moves), _p(nodes), _p(cv), C.byref(st))
        assert rc == 0, rc
        return dict(result=res.value, n_moves=nm.value, moves=moves, tree_nodes=nodes, child_visits=cv), st.as_dict()

    def gather(self, ccfg, seed, first_game, num_games, weights=None, callback=None, threads=1, flags=0, trace=True):
        rows = 63 * num_games
        a = dict(game_ids=np.zeros(rows, np.uint64), my_bb=np.zeros(rows, np.uint64), op_bb=np.zeros(rows, np.uint64),
                 height=np.zeros((rows, 9), np.uint8), player=np.zeros(rows, np.uint8), states=np.zeros((rows, 63), np.float32),
                 pis=np.zeros((rows, 9), np.float32), vs=np.zeros((rows, 3), np.float32))
        t = dict(action=np.zeros(rows, np.uint8), tree_nodes=np.zeros(rows, np.uint32), child_visits=np.zeros((rows, 9), np.float32))
        exp = L.SynExperience()
        exp.capacity = rows
        for k in a:
            setattr(exp, k, a[k].ctypes.data)
        st = L.SynStats()
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        cb = EVAL_FN(callback) if callback else None
        rc = self.lib.orc_gather(C.byref(ccfg), _p(w), C.cast(cb, C.c_void_p) if cb else None, None, int(seed), int(first_game),
                                 int(num_games), int(threads), flags, C.byref(exp), C.byref(st),
                                 _p(t["action"]) if trace else None, _p(t["tree_nodes"]) if trace else None,
                                 _p(t["child_visits"]) if trace else None)
        assert rc == 0, rc
        n = int(exp.len)
        return {k: v[:n] for k, v in a.items()}, st.as_dict(), {k: v[:n] for k, v in t.items()}

    def gather_reference(self, ccfg, weights, num_workers, games_per_train, seed, flags=0, want_rows=False):
        st = L.SynStats()
        w = np.ascontiguousarray(weights, np.float32)
        hits, misses = C.c_uint64(), C.c_uint64()
        rc = self.lib.orc_gather_reference(C.byref(ccfg), _p(w), num_workers, games_per_train, int(seed), flags, None, C.byref(st),
                                           C.cast(C.byref(hits), C.c_void_p), C.cast(C.byref(misses), C.c_void_p))
        assert rc == 0, rc
        d = st.as_dict()
        d["cache_hits"], d["cache_misses"] = hits.value, misses.value
        return d

    def deduplicate(self, my_bb, op_bb, pis, vs):
        my = np.ascontiguousarray(my_bb, np.uint64)
        op = np.ascontiguousarray(op_bb, np.uint64)
        n = my.size
        pi = np.ascontiguousarray(pis, np.float32).reshape(n, 9)
        v = np.ascontiguousarray(vs, np.float32).reshape(n, 3)
        out = dict(states=np.zeros((n, 63), np.float32), pis=np.zeros((n, 9), np.float32), vs=np.zeros((n, 3), np.float32),
                   my_bb=np.zeros(n, np.uint64), op_bb=np.zeros(n, np.uint64), num=np.zeros(n, np.uint32))
        self.lib.orc_deduplicate.restype = C.c_size_t
        self.lib.orc_deduplicate.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_size_t] + [C.c_void_p] * 6
        u = self.lib.orc_deduplicate(_p(my), _p(op), _p(pi), _p(v), n, n, _p(out["states"]), _p(out["pis"]), _p(out["vs"]),
                                     _p(out["my_bb"]), _p(out["op_bb"]), _p(out["num"]))
        return {k: a[:u] for k, a in out.items()}

    def mlp_eval(self, weights, my_bb, op_bb, flags=0):
        w = np.ascontiguousarray(weights, np.float32)
        my = np.ascontiguousarray(my_bb, np.uint64)
        op = np.ascontiguousarray(op_bb, np.uint64)
        n = my.size
        lg, pr = np.zeros((n, 9), np.float32), np.zeros((n, 3), np.float32)
        self.lib.orc_mlp_eval(_p(w), _p(my), _p(op), n, flags, _p(lg), _p(pr))
        return lg, pr

    def ttt_kat(self, which, flags):
        nodes, best, rs = C.c_uint32(), C.c_int(), C.c_uint8()
        cs = np.zeros(9, np.uint8)
        self.lib.orc_ttt_kat(which, flags, C.cast(C.byref(nodes), C.c_void_p), C.cast(C.byref(best), C.c_void_p), _p(cs),
                             C.cast(C.byref(rs), C.c_void_p))
        return nodes.value, best.value, cs, rs.value
