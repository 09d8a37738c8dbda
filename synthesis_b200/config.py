"""Hyper-parameter types, mirroring synthesis/src/config.rs of the reference (same names, same
meaning).  Rust enums with payloads become small frozen classes: `Exploration.PolynomialUct(c=3.0)`,
`Fpu.Const(1.0)`, `PolicyNoise.Dirichlet(alpha, weight)`, `ValueTarget.QtoZ(from_, to)`.
"""
from dataclasses import dataclass, field
from typing import Callable, List, Tuple, Union

from . import _lib as L


class ValueTarget:  # config.rs:2-8
    @dataclass(frozen=True)
    class Z:
        pass

    @dataclass(frozen=True)
    class Q:
        pass

    @dataclass(frozen=True)
    class QZaverage:
        p: float

    @dataclass(frozen=True)
    class QtoZ:
        from_: float
        to: float


class Exploration:  # config.rs:10-14
    @dataclass(frozen=True)
    class Uct:
        c: float

    @dataclass(frozen=True)
    class PolynomialUct:
        c: float


class ActionSelection:  # config.rs:16-20
    Q = L.ACTION_Q
    NumVisits = L.ACTION_NUM_VISITS


class Fpu:  # config.rs:22-27
    @dataclass(frozen=True)
    class Const:
        value: float

    @dataclass(frozen=True)
    class ParentQ:
        pass

    @dataclass(frozen=True)
    class Func:
        """Host code cannot run inside the search kernel: the engine rejects it (SYN_ERR_UNSUPPORTED),
        exactly like an unsupported Fpu panics in the reference's FrozenMCTS (evaluator.rs:410)."""
        f: Callable[[], float]

    @dataclass(frozen=True)
    class Normal:
        """Device-native form of the closure the reference ships (study-connect4/src/main.rs:43-47):
        Normal(mean, std) drawn from a seeded per-game stream instead of thread_rng."""
        mean: float
        std: float


class PolicyNoise:  # config.rs:40-45
    @dataclass(frozen=True)
    class None_:
        pass

    @dataclass(frozen=True)
    class Equal:
        weight: float

    @dataclass(frozen=True)
    class Dirichlet:
        alpha: float
        weight: float


@dataclass
class MCTSConfig:  # config.rs:29-38
    exploration: object
    solve: bool
    correct_values_on_solve: bool
    select_solved_nodes: bool
    auto_extend: bool
    fpu: object
    root_policy_noise: object = field(default_factory=PolicyNoise.None_)

    def to_c(self) -> L.SynMctsCfg:
        c = L.SynMctsCfg()
        if isinstance(self.exploration, Exploration.Uct):
            c.exploration_kind = L.EXPLORATION_UCT
        elif isinstance(self.exploration, Exploration.PolynomialUct):
            c.exploration_kind = L.EXPLORATION_POLYNOMIAL_UCT
        else:
            raise TypeError(f"exploration must be Exploration.Uct or Exploration.PolynomialUct, got {self.exploration!r}")
        c.c = float(self.exploration.c)
        c.solve, c.correct_values_on_solve = int(self.solve), int(self.correct_values_on_solve)
        c.select_solved_nodes, c.auto_extend = int(self.select_solved_nodes), int(self.auto_extend)
        f = self.fpu
        if isinstance(f, Fpu.Const):
            c.fpu_kind, c.fpu_a = L.FPU_CONST, float(f.value)
        elif isinstance(f, Fpu.ParentQ):
            c.fpu_kind = L.FPU_PARENT_Q
        elif isinstance(f, Fpu.Normal):
            c.fpu_kind, c.fpu_a, c.fpu_b = L.FPU_NORMAL, float(f.mean), float(f.std)
        elif isinstance(f, Fpu.Func):
            c.fpu_kind = L.FPU_FUNC  # rejected by the engine with a clear message
        else:
            raise TypeError(f"fpu must be an Fpu variant, got {f!r}")
        n = self.root_policy_noise
        if isinstance(n, PolicyNoise.None_) or n is None:
            c.noise_kind = L.NOISE_NONE
        elif isinstance(n, PolicyNoise.Equal):
            c.noise_kind, c.noise_weight = L.NOISE_EQUAL, float(n.weight)
        elif isinstance(n, PolicyNoise.Dirichlet):
            c.noise_kind, c.noise_alpha, c.noise_weight = L.NOISE_DIRICHLET, float(n.alpha), float(n.weight)
        else:
            raise TypeError(f"root_policy_noise must be a PolicyNoise variant, got {n!r}")
        return c


@dataclass
class RolloutConfig:  # config.rs:47-56
    num_workers: int
    num_explores: int
    random_actions_until: int
    sample_actions_until: int
    stop_games_when_solved: bool
    value_target: object
    action: int
    mcts_cfg: MCTSConfig

    def to_c(self, leaf_eval_kind: int) -> L.SynRolloutCfg:
        c = L.SynRolloutCfg()
        c.num_explores = int(self.num_explores)
        c.random_actions_until = int(self.random_actions_until)
        c.sample_actions_until = int(self.sample_actions_until)
        c.stop_games_when_solved = int(self.stop_games_when_solved)
        v = self.value_target
        if isinstance(v, ValueTarget.Z):
            c.value_target_kind = L.VALUE_Z
        elif isinstance(v, ValueTarget.Q):
            c.value_target_kind = L.VALUE_Q
        elif isinstance(v, ValueTarget.QZaverage):
            c.value_target_kind, c.vt_a = L.VALUE_QZ_AVERAGE, float(v.p)
        elif isinstance(v, ValueTarget.QtoZ):
            c.value_target_kind, c.vt_a, c.vt_b = L.VALUE_Q_TO_Z, float(v.from_), float(v.to)
        else:
            raise TypeError(f"value_target must be a ValueTarget variant, got {v!r}")
        c.action_selection = int(self.action)
        c.mcts = self.mcts_cfg.to_c()
        c.leaf_eval_kind = int(leaf_eval_kind)
        return c


@dataclass
class EvaluationConfig:  # config.rs:59-73 (only the fields the search path reads)
    policy_num_explores: int
    policy_action: int
    policy_mcts_cfg: MCTSConfig
    rollout_action: int
    rollout_num_explores: List[int]
    rollout_mcts_cfg: MCTSConfig
    num_games_against_rollout: int = 5
    num_best_policies: int = 10
    num_games_against_best_policies: int = 1
    logs: str = "./_logs"


@dataclass
class LearningConfig:  # config.rs:76-94
    seed: int
    logs: str
    lr_schedule: List[Tuple[int, float]]
    weight_decay: float
    num_iterations: int
    num_epochs: int
    batch_size: int
    policy_weight: float
    value_weight: float
    games_to_keep: int
    games_per_train: int
    rollout_cfg: RolloutConfig


def study_connect4_mcts_cfg(fpu=None) -> MCTSConfig:
    """The reproducible sibling of the shipped self-play config (study-connect4/src/main.rs:58-66):
    PUCT c=3, solve/correct/select_solved/auto_extend, Fpu::Const(1.0), no root noise."""
    return MCTSConfig(exploration=Exploration.PolynomialUct(3.0), solve=True, correct_values_on_solve=True,
                      select_solved_nodes=True, auto_extend=True, fpu=fpu or Fpu.Const(1.0),
                      root_policy_noise=PolicyNoise.None_())


def study_connect4_rollout_mcts_cfg() -> MCTSConfig:
    """The evaluator's rollout-baseline config (study-connect4/src/main.rs:74-82)."""
    return MCTSConfig(exploration=Exploration.Uct(2.0), solve=True, correct_values_on_solve=True,
                      select_solved_nodes=True, auto_extend=False, fpu=Fpu.Const(float("inf")),
                      root_policy_noise=PolicyNoise.None_())


def study_connect4_rollout_cfg(num_explores=1600, mcts_cfg=None, sample_actions_until=30) -> RolloutConfig:
    """study-connect4/src/main.rs:28-36 with the reproducible MCTS config."""
    return RolloutConfig(num_workers=6, num_explores=num_explores, random_actions_until=1,
                         sample_actions_until=sample_actions_until, stop_games_when_solved=False,
                         value_target=ValueTarget.Q(), action=ActionSelection.NumVisits,
                         mcts_cfg=mcts_cfg or study_connect4_mcts_cfg())
