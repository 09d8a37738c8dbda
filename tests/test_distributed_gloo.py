"""CPU suite, part 3: the N > 1 path on `gloo`, world_size 2.

Games shard by contiguous ranges of the global game index; one broadcast of the weight blob and
one gather of experience rows per iteration.  The host logic (synthesis_b200/distributed.py) is
backend-agnostic: NCCL on the GPU box, gloo here.  The per-shard "engine" in this CPU test is the
ORACLE (test infrastructure standing in for the CUDA engine, which needs a B200) — what is under
test is the sharding, the collectives and the ReplayBuffer fold, not the search.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.dirname(os.path.abspath(__file__))]

GAMES, EXPLORES, SEED, FIRST = 11, 40, 3, 7


def _free_port():
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def _cfg():
    import synthesis_b200 as s
    return s.study_connect4_rollout_cfg(num_explores=EXPLORES, sample_actions_until=12)


def _play_with_oracle(first, count):
    import oracle_binding
    from synthesis_b200 import _lib as L
    orc = oracle_binding.Oracle()
    a, _, _ = orc.gather(_cfg().to_c(L.LEAF_ROLLOUT), SEED, first, count, threads=2)
    return a


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist
        import synthesis_b200 as s
        from synthesis_b200 import distributed as D
        dist.init_process_group("gloo", rank=rank, world_size=world)
        # the weights reach rank 1 (on GPUs this is syn_engine_broadcast_weights inside the library; here a plain broadcast)
        import torch
        w = torch.from_numpy(s.Connect4Net.new(5).blob().copy()) if rank == 0 else torch.zeros(30492)
        dist.broadcast(w, src=0)
        wsum = float(np.abs(w.numpy()).sum())
        buf = s.ReplayBuffer()
        old = s.ReplayBuffer.from_arrays(4, {k: v for k, v in _play_with_oracle(100, 4).items()} | {"game_ids": _play_with_oracle(100, 4)["game_ids"] - np.uint64(100)})
        buf.extend(old)  # 4 older games already in the buffer
        merged = D.gather_rows_over_group(_play_with_oracle, GAMES, buf if rank == 0 else None, games_to_keep=GAMES + 2,
                                                 dst=0, first_game_index=FIRST)
        out = None
        if rank == 0:
            out = dict(merged={k: v.copy() for k, v in merged.items()}, wsum=wsum, buf_ids=buf.game_ids.copy(), buf_game_id=buf.game_id,
                       buf_rows=buf.curr_steps(), buf_my=buf.my_bb.copy())
        else:
            assert merged is None
            out = dict(wsum=wsum)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception as e:  # pragma: no cover - surfaced in the parent
        import traceback
        q.put((rank, "ERROR " + repr(e) + "\n" + traceback.format_exc()))


def test_two_ranks_equal_one(oracle):
    import torch.multiprocessing as mp
    import synthesis_b200 as s
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r, out = q.get(timeout=240)
        assert not isinstance(out, str), out
        res[r] = out
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the broadcast reached rank 1
    assert res[0]["wsum"] == res[1]["wsum"] == float(np.abs(s.Connect4Net.new(5).blob()).sum())
    # two shards gathered in rank order == one call over the whole range
    whole = _play_with_oracle(FIRST, GAMES)
    for k, v in whole.items():
        assert res[0]["merged"][k].shape == v.shape, k
        assert res[0]["merged"][k].tobytes() == v.tobytes(), k
    # folded into the trainer's buffer like gather_experience: keep_last_n_games(games_to_keep - n) then extend
    old_rows = len(_play_with_oracle(100, 4)["vs"])
    ids = res[0]["buf_ids"]
    assert res[0]["buf_game_id"] == 4 + GAMES
    kept_old = ids[ids <= 4]
    # keep_last_n_games(2) keeps ids >= game_id - 2, i.e. THREE games - the reference's own arithmetic (data.rs:177-181)
    assert set(kept_old.tolist()) == {2, 3, 4}
    new_ids = ids[ids > 4]
    assert np.array_equal(new_ids, whole["game_ids"] - np.uint64(FIRST) + np.uint64(4))
    assert res[0]["buf_rows"] == len(kept_old) + len(whole["vs"]) and len(kept_old) < old_rows
    assert np.array_equal(res[0]["buf_my"][len(kept_old):], whole["my_bb"])
