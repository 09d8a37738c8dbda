"""The learner's batch loop (syn_engine_train, csrc/train.cuh) against the same steps in PyTorch fp32 on one host core
(the reference sets torch threads to 1, study-connect4/src/main.rs:85-86).

usage: bench_train.py [games] [explores] [gpu_steps] [cpu_steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import synthesis_b200 as s
from synthesis_b200 import _lib as L

games = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
explores = int(sys.argv[2]) if len(sys.argv) > 2 else 200
gpu_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
cpu_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
eng = s.Engine(0, 148 * 128, explores)
net = s.Connect4Net.new(0)
eng.set_weights(net.blob())
cfg = s.study_connect4_rollout_cfg(num_explores=explores)
a, st, _ = eng.gather(cfg, L.LEAF_NN, 0, games, 0)
d, _ = eng.deduplicate(a["my_bb"], a["op_bb"], a["pis"], a["vs"])
n = len(d["num"])
rng = np.random.default_rng(0)
batches = np.concatenate([s.BatchRandSampler(n, 32, True, rng).all_batches() for _ in range(gpu_steps * 32 // max(32, n - n % 32) + 1)])[:gpu_steps]
line = {"op": "train", "rows": n, "steps": len(batches)}
for kernel, env in (("single_cta", "0"), ("cluster8_sync", "1"), ("cluster8_async", "2")):
    os.environ["SYN_TRAIN_CLUSTER"] = env
    eng.reset_optimizer()
    eng.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], batches[:100], 1e-3, 1e-6)  # warm-up
    eng.set_weights(net.blob())
    eng.reset_optimizer()
    t0 = time.perf_counter()
    losses, ts = eng.train(d["my_bb"], d["op_bb"], d["pis"], d["vs"], batches, 1e-3, 1e-6)
    wall = time.perf_counter() - t0
    line.update({kernel + "_us_per_step": ts["device_ns"] / 1e3 / len(batches), kernel + "_e2e_us_per_step": wall * 1e6 / len(batches),
                 kernel + "_first_loss": [float(x) for x in losses[0]], kernel + "_last_loss": [float(x) for x in losses[-100:].mean(0)]})
if cpu_steps:
    from torch_learner import TorchLearner
    ref = TorchLearner(net.blob(), 1e-3, 1e-6)
    ref.run(d["states"], d["pis"], d["vs"], batches[:50])
    ref = TorchLearner(net.blob(), 1e-3, 1e-6)
    t0 = time.perf_counter()
    want = ref.run(d["states"], d["pis"], d["vs"], batches[:cpu_steps])
    dt = time.perf_counter() - t0
    line.update({"cpu_us_per_step": dt * 1e6 / cpu_steps, "cpu_kind": "PyTorch fp32, 1 thread (the libtorch ops the reference calls through tch)",
                 "max_loss_diff_first_%d_steps" % cpu_steps: float(np.abs(want - losses[:cpu_steps]).max())})
print(json.dumps(line))
