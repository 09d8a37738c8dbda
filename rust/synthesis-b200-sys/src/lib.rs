//! Raw bindings of `include/synthesis_b200.h` — mechanical, one item per C declaration.
//! (Source only; not compiled in the build image, which has no Rust toolchain.)
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

pub const SYN_N_ACTIONS: usize = 9;
pub const SYN_MAX_TURNS: usize = 63;
pub const SYN_N_FEATURES: usize = 63;
pub const SYN_N_WEIGHTS: usize = 30492;

pub const SYN_OK: c_int = 0;
pub const SYN_ERR_UNSUPPORTED: c_int = -4;
pub const SYN_ERR_COMM: c_int = -8;
pub const SYN_COMM_ID_BYTES: usize = 128;
/// include/syn_streams.h: seeds and global game indices are limited so that (seed, game) pairs never alias
pub const SYN_MAX_SEED: u64 = (1 << 30) - 1;
pub const SYN_MAX_GAME_INDEX: u64 = (1 << 32) - 1;

pub const SYN_EXPLORATION_UCT: u32 = 0;
pub const SYN_EXPLORATION_POLYNOMIAL_UCT: u32 = 1;
pub const SYN_FPU_CONST: u32 = 0;
pub const SYN_FPU_PARENT_Q: u32 = 1;
pub const SYN_FPU_NORMAL: u32 = 2;
pub const SYN_FPU_FUNC: u32 = 3;
pub const SYN_NOISE_NONE: u32 = 0;
pub const SYN_NOISE_EQUAL: u32 = 1;
pub const SYN_NOISE_DIRICHLET: u32 = 2;
pub const SYN_VALUE_Z: u32 = 0;
pub const SYN_VALUE_Q: u32 = 1;
pub const SYN_VALUE_QZ_AVERAGE: u32 = 2;
pub const SYN_VALUE_Q_TO_Z: u32 = 3;
pub const SYN_ACTION_Q: u32 = 0;
pub const SYN_ACTION_NUM_VISITS: u32 = 1;
pub const SYN_LEAF_NN: u32 = 0;
pub const SYN_LEAF_ROLLOUT: u32 = 1;
pub const SYN_TREE_MCTS: u32 = 0;
pub const SYN_TREE_FROZEN: u32 = 1;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct syn_mcts_cfg {
    pub exploration_kind: u32,
    pub c: f32,
    pub solve: u8,
    pub correct_values_on_solve: u8,
    pub select_solved_nodes: u8,
    pub auto_extend: u8,
    pub fpu_kind: u32,
    pub fpu_a: f32,
    pub fpu_b: f32,
    pub noise_kind: u32,
    pub noise_alpha: f32,
    pub noise_weight: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct syn_rollout_cfg {
    pub num_explores: u32,
    pub random_actions_until: u32,
    pub sample_actions_until: u32,
    pub stop_games_when_solved: u8,
    pub _pad: [u8; 3],
    pub value_target_kind: u32,
    pub vt_a: f32,
    pub vt_b: f32,
    pub action_selection: u32,
    pub mcts: syn_mcts_cfg,
    pub leaf_eval_kind: u32,
}

/// One player of an evaluation match (`syn_player_cfg`): the argument list of `MCTS::exploit`
/// (mcts.rs:111-121) / `FrozenMCTS::exploit` (evaluator.rs:308-318) minus the game.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct syn_player_cfg {
    pub tree_kind: u32,      // 0 = MCTS, 1 = FrozenMCTS
    pub leaf_eval_kind: u32, // 0 = Connect4Net, 1 = RolloutPolicy
    pub num_explores: u32,
    pub action_selection: u32,
    pub mcts: syn_mcts_cfg,
}

#[repr(C)]
pub struct syn_experience {
    pub capacity: usize,
    pub len: usize,
    pub games: usize,
    pub game_ids: *mut u64,
    pub my_bb: *mut u64,
    pub op_bb: *mut u64,
    pub height: *mut u8,   // [cap][9]
    pub player: *mut u8,
    pub states: *mut f32,  // [cap][63]
    pub pis: *mut f32,     // [cap][9]
    pub vs: *mut f32,      // [cap][3]
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct syn_stats {
    pub explores: u64,
    pub leaf_evals: u64,
    pub rows: u64,
    pub games: u64,
    pub trees: u64,
    pub nodes: u64,
    pub select_levels: u64,
    pub children_scanned: u64,
    pub expansions: u64,
    pub children_created: u64,
    pub backprop_levels: u64,
    pub rollout_plies: u64,
    pub device_ns: u64,
    pub kernel_launches: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
}

/// data.rs:80-104 FlatBatch / StateStatistics — what ReplayBuffer::deduplicate returns
#[repr(C)]
pub struct syn_flat_batch {
    pub capacity: usize,
    pub len: usize,
    pub states: *mut f32, // [cap][63]
    pub pis: *mut f32,    // [cap][9]
    pub vs: *mut f32,     // [cap][3]
    pub my_bb: *mut u64,
    pub op_bb: *mut u64,
    pub num: *mut u32,
}

/// tch nn::Adam::default() + LearningConfig's loss weights / batch size (config.rs:76-94)
#[repr(C)]
pub struct syn_train_cfg {
    pub lr: f32,
    pub beta1: f32,
    pub beta2: f32,
    pub eps: f32,
    pub weight_decay: f32,
    pub policy_weight: f32,
    pub value_weight: f32,
    pub batch_size: u32,
}

#[repr(C)]
pub struct syn_engine {
    _private: [u8; 0],
}

/// One NCCL communicator, created inside the library (the reference's worker threads become ranks)
#[repr(C)]
pub struct syn_comm {
    _private: [u8; 0],
}

extern "C" {
    pub fn syn_abi_version() -> c_int;
    pub fn syn_build_info() -> *const c_char;
    pub fn syn_last_error() -> *const c_char;
    pub fn syn_engine_create(cuda_device: c_int, max_games_in_flight: u32, max_explores: u32, out: *mut *mut syn_engine) -> c_int;
    pub fn syn_engine_destroy(e: *mut syn_engine);
    pub fn syn_engine_set_weights(e: *mut syn_engine, blob: *const f32, n_floats: usize) -> c_int;
    pub fn syn_engine_set_opponent_weights(e: *mut syn_engine, blob: *const f32, n_floats: usize) -> c_int;
    pub fn syn_engine_gather(e: *mut syn_engine, cfg: *const syn_rollout_cfg, first_game_index: u64, num_games: u32, seed: u64,
                             out: *mut syn_experience, stats: *mut syn_stats) -> c_int;
    pub fn syn_engine_gather_launch(e: *mut syn_engine, cfg: *const syn_rollout_cfg, first_game_index: u64, num_games: u32, seed: u64) -> c_int;
    pub fn syn_engine_gather_wait(e: *mut syn_engine, out: *mut syn_experience, stats: *mut syn_stats) -> c_int;
    pub fn syn_engine_search(e: *mut syn_engine, cfg: *const syn_rollout_cfg, tree_kind: u32, my_bb: *const u64, op_bb: *const u64,
                             seeds: *const u64, n_positions: u32, child_visits: *mut f32, child_solution: *mut u8, root_q: *mut f32,
                             root_solution: *mut u8, best_action: *mut u8, num_nodes: *mut u32, stats: *mut syn_stats) -> c_int;
    pub fn syn_engine_match(e: *mut syn_engine, players: *const syn_player_cfg /* [2] */, seeds: *const u64, explores: *const u32 /* [n][2] or null */,
                            n_matches: u32, result: *mut f32, n_moves: *mut u8, moves: *mut u8 /* [n][63] */, tree_nodes: *mut u32,
                            child_visits: *mut f32, stats: *mut syn_stats) -> c_int;
    pub fn syn_engine_eval(e: *mut syn_engine, my_bb: *const u64, op_bb: *const u64, n_positions: u32, logits: *mut f32,
                           outcome_probs: *mut f32) -> c_int;
    pub fn syn_engine_deduplicate(e: *mut syn_engine, my_bb: *const u64, op_bb: *const u64, pis: *const f32, vs: *const f32, n_rows: usize,
                                  out: *mut syn_flat_batch, stats: *mut syn_stats) -> c_int;
    pub fn syn_engine_train(e: *mut syn_engine, cfg: *const syn_train_cfg, my_bb: *const u64, op_bb: *const u64, pis: *const f32,
                            vs: *const f32, n_rows: usize, batch_index: *const u32 /* [n_batches][batch_size] */, n_batches: u32,
                            losses: *mut f32 /* [n_batches][2] or null */, stats: *mut syn_stats) -> c_int;
    pub fn syn_engine_reset_optimizer(e: *mut syn_engine) -> c_int;
    pub fn syn_engine_get_weights(e: *mut syn_engine, blob: *mut f32, n_floats: usize) -> c_int;
    pub fn syn_engine_play(e: *mut syn_engine, moves: *const u8, n_moves: *const u32, stride: u32, n_games: u32, my_bb: *mut u64,
                           op_bb: *mut u64, height: *mut u8 /* [n][9] */, legal_mask_lo: *mut u8, legal_mask_hi: *mut u8,
                           status: *mut u8, features: *mut f32 /* [n][63] or null */) -> c_int;

    // multi-GPU: alpha_zero.rs:132-168 (fan-out / join) and :192-194 (every worker's vs.load) across ranks
    pub fn syn_comm_unique_id(id: *mut u8 /* [SYN_COMM_ID_BYTES] */) -> c_int;
    pub fn syn_comm_create(id: *const u8, n_ranks: c_int, rank: c_int, cuda_device: c_int, out: *mut *mut syn_comm) -> c_int;
    pub fn syn_comm_destroy(c: *mut syn_comm);
    pub fn syn_comm_rank(c: *const syn_comm) -> c_int;
    pub fn syn_comm_size(c: *const syn_comm) -> c_int;
    pub fn syn_engine_broadcast_weights(e: *mut syn_engine, c: *mut syn_comm, blob: *const f32, n_floats: usize, root: c_int) -> c_int;
    pub fn syn_engine_gather_experience(e: *mut syn_engine, c: *mut syn_comm, root: c_int, cfg: *const syn_rollout_cfg,
                                        first_game_index: u64, num_games: u32, seed: u64, out: *mut syn_experience,
                                        stats: *mut syn_stats) -> c_int;

    // knobs and diagnostics; results never depend on them
    pub fn syn_engine_set_trace(e: *mut syn_engine, action: *mut u8, tree_nodes: *mut u32, child_visits: *mut f32) -> c_int;
    pub fn syn_engine_set_group_lanes(e: *mut syn_engine, lanes: c_int) -> c_int;
    pub fn syn_engine_launch_geometry(e: *mut syn_engine, num_games: u32, leaf_eval_kind: u32, ctas: *mut u32, games_per_cta: *mut u32,
                                      lanes_per_game: *mut u32) -> c_int;
    pub fn syn_engine_set_mlp_mode(e: *mut syn_engine, mode: c_int) -> c_int;
    pub fn syn_engine_mlp_in_use(e: *mut syn_engine, chain: *mut c_int, calibration_ratio: *mut f32) -> c_int;
    pub fn syn_engine_debug_counters(e: *mut syn_engine, out: *mut u64, n: u32) -> c_int;
}
