"""synthesis_b200 — B200-native self-play engine for the gather_experience hot path of
coreylowman/synthesis (Connect4 9x7 MCTS self-play).  The compute lives in
libsynthesis_b200.so (hand-written sm_100a kernels behind a C ABI, include/synthesis_b200.h);
this package is the host-side mirror of the reference's Rust surface for that path.
"""
from . import _lib
from .alpha_zero import MCTS, engine_for, gather_experience
from .config import (ActionSelection, EvaluationConfig, Exploration, Fpu, LearningConfig, MCTSConfig, PolicyNoise,
                     RolloutConfig, ValueTarget, study_connect4_mcts_cfg, study_connect4_rollout_cfg,
                     study_connect4_rollout_mcts_cfg)
from .connect4 import Connect4
from .data import FlatBatch, ReplayBuffer
from .engine import Engine
from .policies import Connect4Net, RolloutPolicy

__all__ = ["MCTS", "gather_experience", "engine_for", "ActionSelection", "EvaluationConfig", "Exploration", "Fpu",
           "LearningConfig", "MCTSConfig", "PolicyNoise", "RolloutConfig", "ValueTarget", "Connect4", "FlatBatch", "ReplayBuffer",
           "Engine", "Connect4Net", "RolloutPolicy", "study_connect4_mcts_cfg", "study_connect4_rollout_cfg",
           "study_connect4_rollout_mcts_cfg"]
