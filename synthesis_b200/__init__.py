"""synthesis_b200 — B200-native self-play engine for the gather_experience hot path of
coreylowman/synthesis (Connect4 9x7 MCTS self-play).  The compute lives in
libsynthesis_b200.so (hand-written sm_100a kernels behind a C ABI, include/synthesis_b200.h);
this package is the host-side mirror of the reference's Rust surface for that path.
"""
from . import _lib
from .alpha_zero import MCTS, alpha_zero, engine_for, games_in_flight_for, gather_experience, lr_for_iteration, train_on
from .config import (ActionSelection, EvaluationConfig, Exploration, Fpu, LearningConfig, MCTSConfig, PolicyNoise,
                     RolloutConfig, ValueTarget, study_connect4_mcts_cfg, study_connect4_rollout_cfg,
                     study_connect4_rollout_mcts_cfg)
from .connect4 import Connect4
from .data import BatchRandSampler, FlatBatch, ReplayBuffer
from .engine import Comm, Engine
from .policies import Connect4Net, RolloutPolicy

__all__ = ["MCTS", "alpha_zero", "gather_experience", "engine_for", "lr_for_iteration", "train_on", "BatchRandSampler", "ActionSelection", "EvaluationConfig", "Exploration", "Fpu",
           "LearningConfig", "MCTSConfig", "PolicyNoise", "RolloutConfig", "ValueTarget", "Connect4", "FlatBatch", "ReplayBuffer",
           "Engine", "Comm", "Connect4Net", "RolloutPolicy", "study_connect4_mcts_cfg", "study_connect4_rollout_cfg",
           "study_connect4_rollout_mcts_cfg"]
