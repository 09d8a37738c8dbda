#!/usr/bin/env python
"""bench.py — self-play MCTS explores/sec (Connect4 9x7, 800 explores/move), the metric of BASELINE.json.

A "step" is one gather_experience pass: `--games` whole self-play games per GPU, every move an MCTS of
`--explores` explores whose leaves are evaluated by the reference's Connect4Net MLP (random-init
weights, synthetic: no dataset or checkpoint exists offline).  Workload = BASELINE.json configs[1].

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA engine through the C ABI)
  python bench.py --impl reference [...]                          the reference's CPU path (C++ oracle,
                                                                  reference thread schedule, all host cores)
Under torchrun (N > 1) every rank drives one GPU; rank 0 prints ONE JSON line.

Keys beyond the base contract: `roofline` (dominant kernel vs the measured HBM peak), `cpu_baseline`
(the oracle on this box's host cores, bounded sample), `e2e` (the same metric through the public API
with pinned host buffers, H2D of the weights and D2H of the experience inside the timed region),
`clocks`, `gpu_launches`, `positions_per_s`, `leaf_evals_per_s`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "self-play MCTS explores/sec (Connect4 9x7, 800 explores/move)"
UNIT = "explores/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--games", type=int, default=0, help="self-play games per GPU per step (default: --games-mult x games in flight)")
    p.add_argument("--games-mult", type=int, default=6,
                   help="games per step as a multiple of the games one GPU holds in flight (amortises the end-of-step tail)")
    p.add_argument("--explores", type=int, default=800)
    p.add_argument("--leaf", default="nn", choices=["nn", "rollout"])
    p.add_argument("--group-lanes", type=int, default=int(os.environ.get("SYN_GROUP_LANES", "1")),
                   help="lanes per game: 1 = thread per game (default), 16 / 32 = lane group per game")
    p.add_argument("--cpu-games", type=int, default=0, help="games in the CPU sample (default 96 per host thread, ~10-20 s)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def size_workload(args):
    """Games one GPU holds in flight and games per step; identical for both arms so their `config` matches."""
    if args.leaf == "nn" and args.group_lanes == 1:
        in_flight = 148 * 128 * int(os.environ.get("SYN_TPG_TEAMS", "5"))  # one CTA per SM, teams of 128 games
    elif args.leaf == "nn":
        in_flight = 148 * (512 // args.group_lanes)
    elif args.group_lanes == 1:
        in_flight = 148 * int(os.environ.get("SYN_ROLLOUT_THREADS", "1024"))  # one CTA per SM, a thread per game
    else:
        in_flight = 148 * 8 * (256 // args.group_lanes)
    args.in_flight = in_flight
    args.games = args.games or args.games_mult * in_flight
    return in_flight, args.games


def workload_cfg(args):
    import synthesis_b200 as s
    # study-connect4/src/main.rs:28-36 with the reproducible MCTS config of main.rs:58-66 and 800 explores
    return s.study_connect4_rollout_cfg(num_explores=args.explores, sample_actions_until=30)


def config_dict(args, n_gpus):
    return {
        "workload": ("BASELINE.json configs[1]: Connect4 9x7 AlphaZero self-play, study-connect4 policy/value net "
                     "(MLP 63-128-96-64-48-12, random-init weights), %d explores/move" % args.explores) if args.leaf == "nn" else
                    ("BASELINE.json configs[0] scaled up: Connect4 9x7 rollout-only MCTS with MCTS-Solver, %d explores/move" % args.explores),
        "games_per_gpu_per_step": args.games,
        "explores_per_move": args.explores,
        "leaf_eval": "Connect4Net" if args.leaf == "nn" else "RolloutPolicy",
        "mcts": "PUCT c=3, Fpu::Const(1.0), solve+correct_values+select_solved+auto_extend, no root noise (study-connect4/src/main.rs:58-66)",
        "driver": "random_actions_until=1, sample_actions_until=30, ValueTarget::Q, ActionSelection::NumVisits (main.rs:31-35)",
        "parallelism": "game-sharded x%d (no data-path collective)" % n_gpus,
        "games_in_flight_per_gpu": getattr(args, "in_flight", None),
        "l2": "inputs larger than L2: the tree arenas of the games in flight (%.2f MB/game at %d explores) total ~%.0f GB per GPU vs 126 MB L2; a new seed per step"
              % ((9 * (args.explores + 1) + 8) * 32 / 1e6, args.explores, (getattr(args, "in_flight", 0) or 0) * (9 * (args.explores + 1) + 8) * 32 / 1e9),
    }


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- roofline
def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel for ONE launch of this configuration, from the
    committed ncu capture (profiles/traffic.json, written by scripts/ncu_traffic.py); None if no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            for rec in json.load(f):
                if (rec["leaf"], rec["explores"], rec["games"], rec["group_lanes"]) == (args.leaf, args.explores, args.games, args.group_lanes):
                    return rec["dram_bytes"]
    except Exception:
        pass
    return None


def algorithmic_bytes_per_explore(st):
    """B = d*(20 + 18*C) + x*C*47 + (d+1)*36 with d = select levels per explore, C = children
    per selected parent, x = expansions per explore, all measured in this run."""
    e = max(1, st["explores"])
    d = st["select_levels"] / e
    c = st["children_scanned"] / max(1, st["select_levels"])
    x = st["expansions"] / e
    cpx = st["children_created"] / max(1, st["expansions"])
    b = d * (20 + 18 * c) + x * cpx * 47 + (d + 1) * 36
    if st.get("leaf_evals") and st.get("nn"):
        b += (st["leaf_evals"] / e) * 64  # NN mode: 16 B position in + 48 B (logits, probs) out per leaf
    return b, {"select_depth": round(d, 3), "children_per_parent": round(c, 3), "expansions_per_explore": round(x, 3)}


# ------------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_reference_run(args, games, seed):
    """The reference's CPU path as the reference schedules it (alpha_zero.rs:120-209): num_workers+1 = all host
    threads, one StdRng + one weight copy + one memo cache per worker.  C++ oracle (the Rust original cannot be
    built in this image)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import build as oracle_build
    oracle_build.build()
    import oracle_binding
    import synthesis_b200 as s
    from synthesis_b200 import _lib as L
    orc = oracle_binding.Oracle()
    cores = os.cpu_count() or 1
    cfg = workload_cfg(args)
    if args.leaf == "nn":
        net = s.Connect4Net.new(0)
        t0 = time.perf_counter()
        st = orc.gather_reference(cfg.to_c(L.LEAF_NN), net.blob(), cores - 1, games, seed)
        dt = time.perf_counter() - t0
    else:
        t0 = time.perf_counter()
        _, st, _ = orc.gather(cfg.to_c(L.LEAF_ROLLOUT), seed, 0, games, threads=cores, trace=False)
        dt = time.perf_counter() - t0
    return st, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cpu_games = args.cpu_games or (96 if args.leaf == "nn" else 768) * cores  # rollout leaves are ~8x cheaper on the CPU
    size_workload(args)  # the same `config` as our arm; every step times a bounded sample of that workload
    games = cpu_games
    for w in range(args.warmup):
        cpu_reference_run(args, max(cores, games // 8), 1000 + w)
    tot_e = tot_rows = 0
    tot_t = 0.0
    for k in range(args.steps):
        st, dt, cores = cpu_reference_run(args, games, k)
        tot_e += st["explores"]; tot_rows += st["rows"]; tot_t += dt
    value = tot_e / tot_t
    sample = "%d games/step x %d steps (%d explores), %d worker threads, per-worker memo cache" % (games, args.steps, tot_e, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (random-init weights, seeded games)", "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "positions_per_s": tot_rows / tot_t, "gpu_launches": 0,
        "note": "restated reference CPU path (C++ oracle, g++ -O3 -march=native -ffp-contract=off); the Rust reference cannot be built here",
    }
    emit(line)


# ------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import synthesis_b200 as s
    from synthesis_b200 import _lib as L
    from synthesis_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    leaf = L.LEAF_NN if args.leaf == "nn" else L.LEAF_ROLLOUT
    cfg = workload_cfg(args)
    in_flight, games = size_workload(args)
    eng = s.Engine(local_rank, in_flight, args.explores)
    eng.set_group_lanes(args.group_lanes)
    first = rank * games  # weak scaling: every rank plays `games` games of the global index space

    # weights: rank 0 owns them (the trainer); ONE broadcast per iteration when N > 1
    net = s.Connect4Net.new(0)
    blob_pinned = torch.from_numpy(net.blob()).pin_memory()

    def load_weights():
        if args.leaf != "nn":
            return 0
        if dist is None:
            eng.set_weights(blob_pinned.numpy())  # H2D from pinned host memory
            return blob_pinned.numel() * 4
        w = torch.empty(L.N_WEIGHTS, dtype=torch.float32, device=dev)
        if rank == 0:
            w.copy_(blob_pinned, non_blocking=True)
        dist.broadcast(w, src=0)
        torch.cuda.synchronize()
        eng.set_weights(int(w.data_ptr()))
        return blob_pinned.numel() * 4 if rank == 0 else 0

    load_weights()

    # ---- device-resident leg: inputs (weights, config) already in HBM, experience left in HBM
    def device_step(seed):
        eng.gather_launch(cfg, leaf, first, games, seed)
        return eng.gather_wait(None)

    for w in range(args.warmup):
        device_step(1000 + w)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    acc = {}
    dev_ns = 0
    launches = 0
    for k in range(args.steps):
        st = device_step(k)
        dev_ns += st["device_ns"]
        launches += st["kernel_launches"]
        for key, v in st.items():
            acc[key] = acc.get(key, 0) + v
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None

    def reduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    t_dev = reduce(dev_ns * 1e-9, "MAX")
    t_wall = reduce(wall, "MAX")
    explores = reduce(acc["explores"], "SUM")
    rows = reduce(acc["rows"], "SUM")
    leafs = reduce(acc["leaf_evals"], "SUM")
    launches_all = reduce(launches, "SUM")
    value = explores / t_dev

    # ---- end-to-end leg: public API, pinned host buffers, H2D weights + D2H experience every step.
    # N = 1: weights pinned-host -> HBM, experience HBM -> pinned host.  N > 1 adds the two collectives the
    # path has per iteration: ONE broadcast of the weights from rank 0 and ONE gather of the rows to rank 0
    # (NCCL over NVLink), rank 0 then copies everything to its pinned host buffers.
    # rows per game: 63 always suffices (synthesis_b200.h); this workload's games last ~28 plies, so 44 per game bounds
    # the pinned buffers (the engine reports SYN_ERR_CAPACITY instead of overrunning if that were ever too small)
    cap = min(L.MAX_TURNS, 44) * games
    host = {n: torch.zeros((cap * (world if rank == 0 else 1),) + sh, dtype=D._torch_dtype(torch, dt)).pin_memory() for n, dt, sh in D.FIELDS}
    dev_out = None
    if dist is not None:
        dev_out = {n: torch.zeros((cap,) + sh, dtype=D._torch_dtype(torch, dt), device=dev) for n, dt, sh in D.FIELDS}

    def e2e_step(seed):
        h2d = load_weights()
        exp = L.SynExperience()
        exp.capacity = cap
        for n, _, _ in D.FIELDS:
            setattr(exp, n, host[n].data_ptr() if dist is None else dev_out[n].data_ptr())
        st = eng.gather_into(cfg, leaf, first, games, seed, exp)
        d2h = st["d2h_bytes"]
        if dist is not None:
            counts = torch.zeros(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(counts, torch.tensor([int(exp.len)], dtype=torch.int64, device=dev))
            counts_h = counts.cpu().tolist()
            mx = max(counts_h)
            for n, _, _ in D.FIELDS:
                t = dev_out[n][:mx]
                if rank == 0:
                    recv = [torch.empty_like(t) for _ in range(world)]
                    dist.gather(t, recv, dst=0)
                    off = 0
                    for r in range(world):
                        c = counts_h[r]
                        host[n][off:off + c].copy_(recv[r][:c], non_blocking=True)
                        d2h += c * recv[r][0].numel() * recv[r].element_size() if c else 0
                        off += c
                else:
                    dist.gather(t, None, dst=0)
            torch.cuda.synchronize()
        return st, h2d + st["h2d_bytes"], d2h, int(exp.len)

    e2e_step(2000)
    barrier()
    t0 = time.perf_counter()
    e_expl = e_h2d = e_d2h = 0
    for k in range(args.steps):
        st, h2d, d2h, _ = e2e_step(k)
        e_expl += st["explores"]; e_h2d += h2d; e_d2h += d2h
    barrier()
    e_wall = reduce(time.perf_counter() - t0, "MAX")
    e_value = reduce(e_expl, "SUM") / e_wall

    if rank == 0:
        acc["nn"] = args.leaf == "nn"
        bpe, shape = algorithmic_bytes_per_explore(acc)
        peak, peak_src = measured_peak_hbm()
        kernel_s = dev_ns * 1e-9 / max(1, args.steps)  # rank 0's kernel, average launch duration
        achieved = bpe * (acc["explores"] / max(1, args.steps)) / kernel_s / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args),
                    "kernel": ("selfplay_nn_tpg2_kernel" if args.group_lanes == 1 else "selfplay_nn_tc_kernel") if args.leaf == "nn" else ("selfplay_rollout_tpg2_kernel" if args.group_lanes == 1 else "selfplay_rollout_kernel"),
                    "algorithmic_bytes_per_explore": round(bpe, 1), "explores_per_launch": acc["explores"] / max(1, args.steps),
                    "launch_ms": 1e3 * kernel_s, "peak_source": peak_src, **shape,
                    "note": "latency-bound pointer chasing over per-game trees; see DESIGN.md"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cgames = args.cpu_games or (96 if args.leaf == "nn" else 768) * cores
            cst, cdt, cores = cpu_reference_run(args, cgames, 0)
            cpu = {"value": cst["explores"] / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d games (%d explores, %.1f s), reference schedule: %d worker threads, per-worker memo cache"
                             % (cgames, cst["explores"], cdt, cores)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (random-init weights, seeded games)", "config": config_dict(args, world),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e_value, "unit": UNIT, "h2d_bytes_per_step": e_h2d // max(1, args.steps), "d2h_bytes_per_step": e_d2h // max(1, args.steps)},
            "clocks": clocks, "gpu_launches": int(launches_all),
            "positions_per_s": rows / t_dev, "leaf_evals_per_s": leafs / t_dev, "wall_ms_per_step": 1e3 * t_wall / max(1, args.steps),
            "group_lanes": args.group_lanes,
        }
        emit(line)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # Libraries write to fd 1 behind Python's back (NCCL prints its version line there when NCCL_DEBUG is set): everything
    # but the JSON line goes to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
