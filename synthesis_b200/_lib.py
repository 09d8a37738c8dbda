"""ctypes binding of libsynthesis_b200.so (include/synthesis_b200.h).

There is deliberately no fallback: if the CUDA library has not been built, importing the engine
raises, and if no B200-class GPU is present `syn_engine_create` fails with SYN_ERR_NO_DEVICE.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SYN_B200_LIB: another build of the same library (A/B experiments: scripts/gpu_*.sh); never a different implementation
LIB_PATH = os.environ.get("SYN_B200_LIB") or os.path.join(_HERE, "libsynthesis_b200.so")

SYN_OK = 0
SYN_ERR_INVALID_ARGUMENT = -1
SYN_ERR_NO_DEVICE = -2
SYN_ERR_CUDA = -3
SYN_ERR_UNSUPPORTED = -4
SYN_ERR_CAPACITY = -5
SYN_ERR_NO_WEIGHTS = -6
SYN_ERR_DEVICE_FAULT = -7
SYN_ERR_COMM = -8

N_ACTIONS = 9
MAX_TURNS = 63
N_FEATURES = 63
N_WEIGHTS = 30492

# enums (synthesis_b200.h)
EXPLORATION_UCT, EXPLORATION_POLYNOMIAL_UCT = 0, 1
FPU_CONST, FPU_PARENT_Q, FPU_NORMAL, FPU_FUNC = 0, 1, 2, 3
NOISE_NONE, NOISE_EQUAL, NOISE_DIRICHLET = 0, 1, 2
VALUE_Z, VALUE_Q, VALUE_QZ_AVERAGE, VALUE_Q_TO_Z = 0, 1, 2, 3
ACTION_Q, ACTION_NUM_VISITS = 0, 1
LEAF_NN, LEAF_ROLLOUT = 0, 1
TREE_MCTS, TREE_FROZEN = 0, 1


class SynMctsCfg(C.Structure):
    _fields_ = [
        ("exploration_kind", C.c_uint32), ("c", C.c_float),
        ("solve", C.c_uint8), ("correct_values_on_solve", C.c_uint8),
        ("select_solved_nodes", C.c_uint8), ("auto_extend", C.c_uint8),
        ("fpu_kind", C.c_uint32), ("fpu_a", C.c_float), ("fpu_b", C.c_float),
        ("noise_kind", C.c_uint32), ("noise_alpha", C.c_float), ("noise_weight", C.c_float),
    ]


class SynRolloutCfg(C.Structure):
    _fields_ = [
        ("num_explores", C.c_uint32), ("random_actions_until", C.c_uint32), ("sample_actions_until", C.c_uint32),
        ("stop_games_when_solved", C.c_uint8), ("_pad", C.c_uint8 * 3),
        ("value_target_kind", C.c_uint32), ("vt_a", C.c_float), ("vt_b", C.c_float),
        ("action_selection", C.c_uint32), ("mcts", SynMctsCfg), ("leaf_eval_kind", C.c_uint32),
    ]


class SynPlayerCfg(C.Structure):
    _fields_ = [
        ("tree_kind", C.c_uint32), ("leaf_eval_kind", C.c_uint32), ("num_explores", C.c_uint32),
        ("action_selection", C.c_uint32), ("mcts", SynMctsCfg),
    ]


class SynExperience(C.Structure):
    _fields_ = [
        ("capacity", C.c_size_t), ("len", C.c_size_t), ("games", C.c_size_t),
        ("game_ids", C.c_void_p), ("my_bb", C.c_void_p), ("op_bb", C.c_void_p), ("height", C.c_void_p),
        ("player", C.c_void_p), ("states", C.c_void_p), ("pis", C.c_void_p), ("vs", C.c_void_p),
    ]


class SynFlatBatch(C.Structure):
    _fields_ = [
        ("capacity", C.c_size_t), ("len", C.c_size_t), ("states", C.c_void_p), ("pis", C.c_void_p), ("vs", C.c_void_p),
        ("my_bb", C.c_void_p), ("op_bb", C.c_void_p), ("num", C.c_void_p),
    ]


class SynTrainCfg(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("policy_weight", C.c_float), ("value_weight", C.c_float), ("batch_size", C.c_uint32)]


class SynStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "explores", "leaf_evals", "rows", "games", "trees", "nodes", "select_levels", "children_scanned",
        "expansions", "children_created", "backprop_levels", "rollout_plies", "device_ns", "kernel_launches",
        "h2d_bytes", "d2h_bytes")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/synthesis_b200.h declares
EXPORTED_SYMBOLS = (
    "syn_abi_version", "syn_build_info", "syn_last_error", "syn_engine_create", "syn_engine_destroy",
    "syn_engine_set_weights", "syn_engine_set_opponent_weights", "syn_engine_gather", "syn_engine_gather_launch", "syn_engine_gather_wait",
    "syn_engine_search", "syn_engine_match", "syn_engine_eval", "syn_engine_play", "syn_engine_set_trace", "syn_engine_set_group_lanes", "syn_engine_set_mlp_mode", "syn_engine_debug_counters",
    "syn_engine_deduplicate", "syn_engine_train", "syn_engine_reset_optimizer", "syn_engine_get_weights",
    "syn_comm_unique_id", "syn_comm_create", "syn_comm_destroy", "syn_comm_rank", "syn_comm_size",
    "syn_engine_broadcast_weights", "syn_engine_gather_experience", "syn_engine_launch_geometry", "syn_engine_mlp_in_use",
)
COMM_ID_BYTES = 128

_lib = None


class EngineError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libsynthesis_b200 error {code}: {message}")
        self.code = code


def load():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). synthesis_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    lib.syn_abi_version.restype = i32
    lib.syn_build_info.restype = C.c_char_p
    lib.syn_last_error.restype = C.c_char_p
    lib.syn_engine_create.argtypes = [i32, u32, u32, C.POINTER(vp)]
    lib.syn_engine_destroy.argtypes = [vp]
    lib.syn_engine_destroy.restype = None
    lib.syn_engine_set_weights.argtypes = [vp, vp, C.c_size_t]
    lib.syn_engine_set_opponent_weights.argtypes = [vp, vp, C.c_size_t]
    lib.syn_engine_gather.argtypes = [vp, C.POINTER(SynRolloutCfg), u64, u32, u64, C.POINTER(SynExperience), C.POINTER(SynStats)]
    lib.syn_engine_gather_launch.argtypes = [vp, C.POINTER(SynRolloutCfg), u64, u32, u64]
    lib.syn_engine_gather_wait.argtypes = [vp, C.POINTER(SynExperience), C.POINTER(SynStats)]
    lib.syn_engine_search.argtypes = [vp, C.POINTER(SynRolloutCfg), u32, vp, vp, vp, u32, vp, vp, vp, vp, vp, vp, C.POINTER(SynStats)]
    lib.syn_engine_match.argtypes = [vp, C.POINTER(SynPlayerCfg), vp, vp, u32, vp, vp, vp, vp, vp, C.POINTER(SynStats)]
    lib.syn_engine_eval.argtypes = [vp, vp, vp, u32, vp, vp]
    lib.syn_engine_play.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp, vp, vp, vp, vp]
    lib.syn_engine_set_trace.argtypes = [vp, vp, vp, vp]
    lib.syn_engine_set_group_lanes.argtypes = [vp, i32]
    lib.syn_engine_set_mlp_mode.argtypes = [vp, i32]
    lib.syn_engine_debug_counters.argtypes = [vp, vp, u32]
    lib.syn_engine_train.argtypes = [vp, C.POINTER(SynTrainCfg), vp, vp, vp, vp, C.c_size_t, vp, u32, vp, C.POINTER(SynStats)]
    lib.syn_engine_reset_optimizer.argtypes = [vp]
    lib.syn_engine_get_weights.argtypes = [vp, vp, C.c_size_t]
    lib.syn_engine_deduplicate.argtypes = [vp, vp, vp, vp, vp, C.c_size_t, C.POINTER(SynFlatBatch), C.POINTER(SynStats)]
    lib.syn_engine_launch_geometry.argtypes = [vp, u32, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    lib.syn_engine_mlp_in_use.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_float)]
    lib.syn_comm_unique_id.argtypes = [vp]
    lib.syn_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
    lib.syn_comm_destroy.argtypes = [vp]
    lib.syn_comm_destroy.restype = None
    lib.syn_comm_rank.argtypes = [vp]
    lib.syn_comm_size.argtypes = [vp]
    lib.syn_engine_broadcast_weights.argtypes = [vp, vp, vp, C.c_size_t, i32]
    lib.syn_engine_gather_experience.argtypes = [vp, vp, i32, C.POINTER(SynRolloutCfg), u64, u32, u64, C.POINTER(SynExperience), C.POINTER(SynStats)]
    _lib = lib
    return lib


def check(rc):
    if rc != SYN_OK:
        raise EngineError(rc, load().syn_last_error().decode("utf-8", "replace"))
