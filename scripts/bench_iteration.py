"""One training iteration of the reference AS SHIPPED (study-connect4/src/main.rs:13-56: 1,000 games per iteration at 1,600
explores/move with the Normal(1.0, 0.1) first-play urgency, 20 epochs of batches of 32 over the deduplicated buffer, weight
decay 1e-6) through this repo's public API — gather -> ReplayBuffer -> deduplicate -> train, every stage on the GPU — with
each phase timed, and beside it what the same work costs on the host: the CPU oracle on all cores for a bounded sample of
the games, PyTorch fp32 on one thread (main.rs:85-86) for a bounded sample of the optimizer steps.

usage: python scripts/bench_iteration.py [iterations=3] [games_per_train=1000]
Prints one JSON line.  (The CPU figures are rates measured on a sample and scaled; they are estimates of the reference's
iteration, not a run of it — the Rust reference cannot be built in this image.)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synthesis_b200 as s  # noqa: E402
from synthesis_b200 import _lib as L  # noqa: E402
from synthesis_b200.alpha_zero import games_in_flight_for, lr_for_iteration, train_on  # noqa: E402


def shipped_cfg(games_per_train):
    m = s.study_connect4_mcts_cfg(fpu=s.Fpu.Normal(1.0, 0.1))
    rc = s.study_connect4_rollout_cfg(num_explores=1600, mcts_cfg=m, sample_actions_until=30)
    return s.LearningConfig(seed=0, logs="", lr_schedule=[(1, 1e-3), (20, 5e-4), (40, 1e-4), (60, 5e-5), (80, 1e-5)], weight_decay=1e-6,
                            num_iterations=200, num_epochs=20, batch_size=32, policy_weight=1.0, value_weight=1.0, games_to_keep=20000,
                            games_per_train=games_per_train, rollout_cfg=rc)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    cfg = shipped_cfg(n)
    eng = s.Engine(0, games_in_flight_for(n), 1600)
    net = s.Connect4Net.new(0)
    eng.set_weights(net.blob())
    eng.reset_optimizer()
    rng = np.random.default_rng(0)
    buffer = s.ReplayBuffer(256_000)
    eng.gather(cfg.rollout_cfg, L.LEAF_NN, 0, 64, 999)  # warm-up: kernels loaded, staging sized
    per_iter = []
    for i in range(iters):
        t0 = time.perf_counter()
        arrays, st, _ = eng.gather(cfg.rollout_cfg, L.LEAF_NN, 0, n, i)
        t1 = time.perf_counter()
        buffer.keep_last_n_games(cfg.games_to_keep - cfg.games_per_train)
        buffer.extend(s.ReplayBuffer.from_arrays(n, arrays))
        t2 = time.perf_counter()
        dedup = buffer.deduplicate(eng)
        t3 = time.perf_counter()
        epochs = train_on(cfg, dedup, lr_for_iteration(cfg, i), eng, rng)
        t4 = time.perf_counter()
        steps = cfg.num_epochs * (len(dedup) // cfg.batch_size)
        per_iter.append({"gather_s": t1 - t0, "buffer_s": t2 - t1, "dedup_s": t3 - t2, "train_s": t4 - t3, "total_s": t4 - t0,
                         "explores": st["explores"], "rows": int(len(arrays["vs"])), "positions": len(dedup), "adam_steps": steps,
                         "gather_device_s": st["device_ns"] * 1e-9, "last_epoch_loss": epochs[-1],
                         "mapping": "%d lanes per game" % eng.launch_geometry(n, L.LEAF_NN)[2], "forward_chain": eng.mlp_in_use()[0]})
    # the host's rates on bounded samples of the same work
    import oracle_binding
    from torch_learner import TorchLearner
    orc = oracle_binding.Oracle()
    cores = os.cpu_count() or 1
    sample_games = 3 * cores
    t0 = time.perf_counter()
    ost = orc.gather_reference(cfg.rollout_cfg.to_c(L.LEAF_NN), net.blob(), cores - 1, sample_games, 0)  # workers + 1 threads, memo cache
    cpu_explores_per_s = ost["explores"] / (time.perf_counter() - t0)
    d = dedup
    tl = TorchLearner(net.blob(), 1e-3, cfg.weight_decay)
    batches = s.BatchRandSampler(len(d), 32, True, np.random.default_rng(1)).all_batches()[:300]
    t0 = time.perf_counter()
    tl.run(d.states, d.pis, d.vs, batches)
    torch_us_per_step = 1e6 * (time.perf_counter() - t0) / len(batches)
    last = per_iter[-1]
    est_cpu = {"gather_s": last["explores"] / cpu_explores_per_s, "train_s": last["adam_steps"] * torch_us_per_step * 1e-6}
    print(json.dumps({"workload": "study-connect4/src/main.rs:13-56 as shipped: %d games/iteration, 1600 explores/move, Fpu Normal(1.0, 0.1), "
                                  "20 epochs x batches of 32" % n,
                      "iterations": per_iter,
                      "host_estimate_for_last_iteration": {**est_cpu, "total_s": est_cpu["gather_s"] + est_cpu["train_s"],
                                                           "cpu_explores_per_s": cpu_explores_per_s, "cores": cores,
                                                           "torch_us_per_adam_step_one_thread": torch_us_per_step,
                                                           "how": "oracle on %d games with %d threads + per-worker memo cache; PyTorch fp32, one thread, 300 steps" % (sample_games, cores)},
                      "speedup_last_iteration": (est_cpu["gather_s"] + est_cpu["train_s"]) / last["total_s"]}))


if __name__ == "__main__":
    main()
