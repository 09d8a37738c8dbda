/* synthesis_b200.h — C ABI of libsynthesis_b200.so, the B200 (sm_100a) self-play engine
 * that replaces ONE path of coreylowman/synthesis: gather_experience → run_game → MCTS →
 * Policy::eval on Connect4 9x7.
 *
 * The reference has no FFI: its seam is a set of Rust traits and plain-data structs.
 * Every declaration below names the reference item it stands in for (paths relative to
 * the reference checkout).  The Rust-side binding a maintainer would add is shown in
 * INTEGRATION.md and kept as source under rust/.
 *
 * Conventions
 *   - every function returns 0 (SYN_OK) or a negative syn_status; nothing throws or aborts;
 *     syn_last_error() returns a thread-local, human-readable message for the last failure.
 *     (The reference panics via unwrap(): alpha_zero.rs:161,166,194 — the Rust shim turns a
 *     non-zero status back into a panic/Err.)
 *   - all buffers are caller-owned; the engine never frees or retains caller pointers.
 *   - output/weight pointers may be host OR device pointers (detected with
 *     cudaPointerGetAttributes); device pointers must belong to the engine's device.
 *   - an engine is bound to one GPU and is NOT thread-safe; calls are serialised by the caller.
 *   - there is no CPU fallback: if no sm_100-class device is present, syn_engine_create fails.
 */
#ifndef SYNTHESIS_B200_H
#define SYNTHESIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SYN_ABI_VERSION 1

/* study-connect4/src/connect4.rs:15-16,173-177 */
#define SYN_WIDTH 9
#define SYN_HEIGHT 7
#define SYN_N_ACTIONS 9   /* Game::MAX_NUM_ACTIONS */
#define SYN_MAX_TURNS 63  /* Game::MAX_TURNS */
#define SYN_N_FEATURES 63 /* Game::DIMS = [1,1,7,9] flattened, index = row*9+col */
/* study-connect4/src/policies.rs:20-24: 63-128-96-64-48-12, weights [out][in] then bias */
#define SYN_N_WEIGHTS 30492

typedef enum {
    SYN_OK = 0,
    SYN_ERR_INVALID_ARGUMENT = -1,
    SYN_ERR_NO_DEVICE = -2,      /* no CUDA device / not compute capability 10.x */
    SYN_ERR_CUDA = -3,           /* a CUDA runtime call failed; see syn_last_error() */
    SYN_ERR_UNSUPPORTED = -4,    /* e.g. Fpu::Func — host code cannot run on the device */
    SYN_ERR_CAPACITY = -5,       /* caller buffer or engine arena too small */
    SYN_ERR_NO_WEIGHTS = -6,     /* leaf_eval = NN but syn_engine_set_weights was never called */
    SYN_ERR_DEVICE_FAULT = -7,   /* the kernel reported an internal inconsistency */
    SYN_ERR_COMM = -8            /* NCCL could not be loaded or a collective failed; see syn_last_error() */
} syn_status;

/* synthesis/src/config.rs:10-14  enum Exploration { Uct{c}, PolynomialUct{c} } */
typedef enum { SYN_EXPLORATION_UCT = 0, SYN_EXPLORATION_POLYNOMIAL_UCT = 1 } syn_exploration_kind;
/* synthesis/src/config.rs:22-27  enum Fpu { Const(f32), ParentQ, Func(fn()->f32) }.
 * Func carries host code and is rejected (SYN_ERR_UNSUPPORTED); the shipped closure
 * (study-connect4/src/main.rs:43-47, Normal(1.0, 0.1) from thread_rng) is offered natively
 * as SYN_FPU_NORMAL{mean=fpu_a, std=fpu_b}, drawn from a seeded per-game stream. */
typedef enum { SYN_FPU_CONST = 0, SYN_FPU_PARENT_Q = 1, SYN_FPU_NORMAL = 2, SYN_FPU_FUNC = 3 } syn_fpu_kind;
/* synthesis/src/config.rs:40-45  enum PolicyNoise { None, Equal{weight}, Dirichlet{alpha,weight} } */
typedef enum { SYN_NOISE_NONE = 0, SYN_NOISE_EQUAL = 1, SYN_NOISE_DIRICHLET = 2 } syn_noise_kind;
/* synthesis/src/config.rs:2-8  enum ValueTarget { Z, Q, QZaverage{p}, QtoZ{from,to} } */
typedef enum { SYN_VALUE_Z = 0, SYN_VALUE_Q = 1, SYN_VALUE_QZ_AVERAGE = 2, SYN_VALUE_Q_TO_Z = 3 } syn_value_target_kind;
/* synthesis/src/config.rs:16-20  enum ActionSelection { Q, NumVisits } */
typedef enum { SYN_ACTION_Q = 0, SYN_ACTION_NUM_VISITS = 1 } syn_action_selection;
/* which Policy<G,N> evaluates leaves (synthesis/src/policies/traits.rs:4-6):
 * Connect4Net (study-connect4/src/policies.rs:47-59) or RolloutPolicy (policies/rollout.rs:8-31) */
typedef enum { SYN_LEAF_NN = 0, SYN_LEAF_ROLLOUT = 1 } syn_leaf_eval_kind;
/* which tree: MCTS (synthesis/src/mcts.rs) or the evaluator's FrozenMCTS (evaluator.rs:299-534) */
typedef enum { SYN_TREE_MCTS = 0, SYN_TREE_FROZEN = 1 } syn_tree_kind;

/* synthesis/src/config.rs:29-38  struct MCTSConfig */
typedef struct {
    uint32_t exploration_kind; /* syn_exploration_kind */
    float c;
    uint8_t solve;
    uint8_t correct_values_on_solve;
    uint8_t select_solved_nodes;
    uint8_t auto_extend;
    uint32_t fpu_kind; /* syn_fpu_kind */
    float fpu_a;       /* Const: value; Normal: mean */
    float fpu_b;       /* Normal: std */
    uint32_t noise_kind; /* syn_noise_kind */
    float noise_alpha;
    float noise_weight;
} syn_mcts_cfg;

/* synthesis/src/config.rs:47-56  struct RolloutConfig (num_workers is a host-thread count and
 * has no device meaning: games are the unit of parallelism) */
typedef struct {
    uint32_t num_explores;
    uint32_t random_actions_until;
    uint32_t sample_actions_until;
    uint8_t stop_games_when_solved;
    uint8_t _pad[3];
    uint32_t value_target_kind; /* syn_value_target_kind */
    float vt_a;                 /* QZaverage: p ; QtoZ: from */
    float vt_b;                 /* QtoZ: to */
    uint32_t action_selection;  /* syn_action_selection */
    syn_mcts_cfg mcts;
    uint32_t leaf_eval_kind; /* syn_leaf_eval_kind */
} syn_rollout_cfg;

/* synthesis/src/data.rs:106-114  struct ReplayBuffer — struct-of-arrays, one row per
 * `buffer.add` (alpha_zero.rs:250).  Rows are ordered by game index, then ply, i.e. the
 * order `ReplayBuffer::extend` (data.rs:160-170) produces when workers are joined in
 * index order (alpha_zero.rs:165-168).  `games` is split into the four fields of
 * `Connect4` (connect4.rs:108-114).  capacity is in rows; 63*num_games always suffices. */
typedef struct {
    size_t capacity;       /* in: rows available in every array below */
    size_t len;            /* out: rows written  (ReplayBuffer::curr_steps) */
    size_t games;          /* out: games played  (ReplayBuffer::total_games_played delta) */
    uint64_t* game_ids;    /* [cap]      1-based like ReplayBuffer::new_game: first_game_index+1+i */
    uint64_t* my_bb;       /* [cap]      Connect4::my_bb  (stones of the player to move) */
    uint64_t* op_bb;       /* [cap]      Connect4::op_bb */
    uint8_t* height;       /* [cap][9]   Connect4::height */
    uint8_t* player;       /* [cap]      0 = Red (moves first), 1 = Black */
    float* states;         /* [cap][63]  Game::features(), row*9+col  (connect4.rs:237-258) */
    float* pis;            /* [cap][9]   MCTS::target_policy (mcts.rs:174-211) */
    float* vs;             /* [cap][3]   value target [Lose,Draw,Win] (alpha_zero.rs:309-338) */
} syn_experience;

/* Counters of one gather/search call.  The roofline figure in bench.py is derived from
 * select_levels / children_scanned / expansions with SURVEY.md §8(d)'s formula. */
typedef struct {
    uint64_t explores;         /* executed MCTS::explore calls (mcts.rs:145) */
    uint64_t leaf_evals;       /* Policy::eval calls (mcts.rs:407) */
    uint64_t rows;             /* experience rows (= plies played) */
    uint64_t games;
    uint64_t trees;            /* MCTS::with_capacity calls */
    uint64_t nodes;            /* sum over trees of nodes.len() */
    uint64_t select_levels;    /* select_best_child calls */
    uint64_t children_scanned; /* children scored by select_best_child */
    uint64_t expansions;       /* visit() calls that pushed children */
    uint64_t children_created; /* nodes pushed by visit() */
    uint64_t backprop_levels;  /* nodes updated by backprop */
    uint64_t rollout_plies;    /* RolloutPolicy steps */
    uint64_t device_ns;        /* CUDA-event time of the search kernels on the engine's stream */
    uint64_t kernel_launches;  /* kernels launched by this call */
    uint64_t h2d_bytes;        /* bytes copied host→device by this call */
    uint64_t d2h_bytes;        /* bytes copied device→host by this call */
} syn_stats;

typedef struct syn_engine syn_engine;

/* ABI / build information. */
int syn_abi_version(void);
const char* syn_build_info(void); /* "sm_100a; nvcc 12.9; ..." */
const char* syn_last_error(void);

/* One engine per GPU.  max_games_in_flight bounds how many games own a tree arena at once
 * (more games than that are queued and started as arenas free up); max_explores bounds
 * num_explores (arena = 1 + 9*(max_explores+1) nodes per game in flight, cf. mcts.rs:123-137).
 * Replaces: the worker threads of gather_experience (alpha_zero.rs:132-154). */
int syn_engine_create(int cuda_device, uint32_t max_games_in_flight, uint32_t max_explores, syn_engine** out);
void syn_engine_destroy(syn_engine* e);

/* Replaces `vs.load(models/<name>.ot)` per worker (alpha_zero.rs:192-194).
 * blob = l_1.weight[128*63], l_1.bias[128], l_2.weight[96*128], l_2.bias[96], l_3.weight[64*96],
 * l_3.bias[64], l_4.weight[48*64], l_4.bias[48], l_5.weight[12*48], l_5.bias[12]; PyTorch
 * nn.Linear layout ([out][in] row-major), fp32, host or device pointer. n_floats must be 30492. */
int syn_engine_set_weights(syn_engine* e, const float* blob, size_t n_floats);

/* Replaces gather_experience / run_n_games / run_game (alpha_zero.rs:120-268) for games
 * [first_game_index, first_game_index + num_games).  Game g draws its random choices from
 * private ChaCha12 streams derived from (seed, g) — see DESIGN.md "random streams" — so the
 * result does not depend on how games are sharded over GPUs.  seed < 2^30, game indices < 2^32 (syn_streams.h).  Blocks until done.
 * `out` rows: see syn_experience.  stats may be NULL. */
int syn_engine_gather(syn_engine* e, const syn_rollout_cfg* cfg, uint64_t first_game_index, uint32_t num_games,
                      uint64_t seed, syn_experience* out, syn_stats* stats);

/* The same, split in two so the search can be timed with the result left in HBM:
 * _launch enqueues everything on the engine's stream and returns; _wait blocks, fills stats
 * and (if out != NULL) copies the experience to the caller's buffers. */
int syn_engine_gather_launch(syn_engine* e, const syn_rollout_cfg* cfg, uint64_t first_game_index,
                             uint32_t num_games, uint64_t seed);
int syn_engine_gather_wait(syn_engine* e, syn_experience* out, syn_stats* stats);

/* Replaces MCTS::exploit (mcts.rs:111-121) / FrozenMCTS::exploit (evaluator.rs:308-318) on a
 * batch of independent root positions, and exposes what target_policy / target_q / solution /
 * best_action read from the finished tree.  Position i is given as the two bitboards of
 * Connect4 (player to move first).  rollout rng of position i = StdRng::seed_from_u64(seeds[i]).
 * Outputs (any may be NULL): child_visits[n][9] (0 for illegal columns), child_solution[n][9]
 * (packed outcome, see syn_outcome_*; 0 = None), root_q[n][3] = target_q, root_solution[n],
 * best_action[n], num_nodes[n] = nodes.len(). */
int syn_engine_search(syn_engine* e, const syn_rollout_cfg* cfg, uint32_t tree_kind, const uint64_t* my_bb,
                      const uint64_t* op_bb, const uint64_t* seeds, uint32_t n_positions, float* child_visits,
                      uint8_t* child_solution, float* root_q, uint8_t* root_solution, uint8_t* best_action,
                      uint32_t* num_nodes, syn_stats* stats);

/* One player of an evaluation match: which tree (MCTS, mcts.rs, or the evaluator's FrozenMCTS,
 * evaluator.rs:299-534), which Policy evaluates its leaves, how many explores per move and how the
 * move is read from the finished tree.  Mirrors the argument lists of MCTS::exploit (mcts.rs:111-121)
 * and FrozenMCTS::exploit (evaluator.rs:308-318) and the pairs (policy_*, rollout_*) of
 * EvaluationConfig (config.rs:58-74). */
typedef struct {
    uint32_t tree_kind;        /* syn_tree_kind */
    uint32_t leaf_eval_kind;   /* syn_leaf_eval_kind */
    uint32_t num_explores;
    uint32_t action_selection; /* syn_action_selection */
    syn_mcts_cfg mcts;
} syn_player_cfg;

/* eval_against_old(p1, p2) with two DIFFERENT networks (evaluator.rs:131-161, called per older model at
 * :87-94): players[1] of the following syn_engine_match calls evaluates its leaves with this blob (same
 * layout as syn_engine_set_weights, host or device) when BOTH players use leaf_eval_kind = NN; players[0]
 * keeps syn_engine_set_weights' network.  blob = NULL returns to one network for both. */
int syn_engine_set_opponent_weights(syn_engine* e, const float* blob, size_t n_floats);

/* Replaces the evaluator's game loops eval_against_rollout_mcts (evaluator.rs:163-198), mcts_vs_mcts
 * (:200-228) and eval_against_old (:129-160; p1 != p2 after syn_engine_set_opponent_weights) on a batch of independent matches: match i
 * starts from Connect4::new(), players[0] moves first, every move is `exploit` of the mover's player
 * (a fresh tree per move), and the game ends when Game::step reports is_over.  All RolloutPolicy
 * draws of match i — by either player — come from ONE stream StdRng::seed_from_u64(seeds[i]), as in
 * the reference where both FrozenMCTS players share `rollout_policy` (evaluator.rs:208-209).
 * explores (optional, [n][2]) overrides players[k].num_explores per match (the evaluator sweeps the
 * opponent's explores, evaluator.rs:65-82); every value must be <= the engine's max_explores.
 * Outputs (any may be NULL): result[n] = game.reward(first_player) in {+1, 0, -1};
 * n_moves[n]; moves[n][63] (columns played); per-move trace tree_nodes[n][63] = nodes.len() and
 * child_visits[n][63][9] = the root's child visit counts by column (0 for illegal columns). */
int syn_engine_match(syn_engine* e, const syn_player_cfg players[2], const uint64_t* seeds, const uint32_t* explores,
                     uint32_t n_matches, float* result, uint8_t* n_moves, uint8_t* moves, uint32_t* tree_nodes,
                     float* child_visits, syn_stats* stats);

/* Replaces Policy::eval for Connect4Net (study-connect4/src/policies.rs:47-59) on a batch:
 * logits[n][9] are the raw policy logits, outcome_probs[n][3] = softmax of the value head,
 * [Lose, Draw, Win] for the player to move. */
int syn_engine_eval(syn_engine* e, const uint64_t* my_bb, const uint64_t* op_bb, uint32_t n_positions, float* logits,
                    float* outcome_probs);

/* Connect4 game rules on a batch of move lists (Game::step, connect4.rs:221-233): plays
 * moves[i][0..n_moves[i]) from Connect4::new() ON THE DEVICE and reports the final state.
 * Used by the parity tests of the game kernel; 255 in `status` = an illegal move was met.
 * status[i]: bit0 = is_over, bit1 = previous mover won. */
int syn_engine_play(syn_engine* e, const uint8_t* moves, const uint32_t* n_moves, uint32_t stride, uint32_t n_games,
                    uint64_t* my_bb, uint64_t* op_bb, uint8_t* height /*[n][9]*/, uint8_t* legal_mask_lo /*[n] cols 0-7*/,
                    uint8_t* legal_mask_hi /*[n] col 8*/, uint8_t* status, float* features /*[n][63] or NULL*/);

/* synthesis/src/data.rs:80-104  struct FlatBatch {states, pis, vs} + StateStatistics — what
 * ReplayBuffer::deduplicate returns: one row per DISTINCT position of the buffer.  my_bb/op_bb/num
 * (the HashMap key and StateStatistics::num) are extra and may be NULL, like every array.
 * Pointers may be host or device memory. */
typedef struct {
    size_t capacity; /* in: rows available in every non-NULL array (n_rows always suffices) */
    size_t len;      /* out: distinct positions */
    float* states;   /* [cap][63]  StateStatistics::state = Game::features() of the position */
    float* pis;      /* [cap][9]   sum_pi / num   (data.rs:220-224) */
    float* vs;       /* [cap][3]   sum_v / num    (data.rs:225-229) */
    uint64_t* my_bb; /* [cap] */
    uint64_t* op_bb; /* [cap] */
    uint32_t* num;   /* [cap]      rows merged into this one */
} syn_flat_batch;

/* Replaces ReplayBuffer::deduplicate (data.rs:196-235), the step that follows gather_experience in
 * the training loop (alpha_zero.rs:53-58), on n_rows rows given as struct-of-arrays (the my_bb, op_bb,
 * pis, vs arrays of syn_experience; host or device pointers).  Rows with equal (my_bb, op_bb) — the
 * fields Connect4's Hash/Eq reduce to — are merged; sums run in row order in f32 like the reference's
 * loop, so every output value is bit-identical to the reference's.  Output order: by first occurrence
 * of the position in the buffer (the reference's order is HashMap-random; equal as a set).
 * stats (optional): rows = distinct positions, device_ns, kernel_launches, h2d/d2h bytes. */
int syn_engine_deduplicate(syn_engine* e, const uint64_t* my_bb, const uint64_t* op_bb, const float* pis, const float* vs,
                           size_t n_rows, syn_flat_batch* out, syn_stats* stats);

/* Hyper-parameters of one training pass: tch `nn::Adam::default()` (beta1 0.9, beta2 0.999, eps 1e-8; L2 weight decay
 * added to the gradient, alpha_zero.rs:33-36), the learning rate of the current iteration (LearningConfig::lr_schedule,
 * alpha_zero.rs:62-70) and the loss weights / batch size of synthesis/src/config.rs:40-56. */
typedef struct {
    float lr;
    float beta1, beta2, eps;
    float weight_decay;   /* LearningConfig::weight_decay */
    float policy_weight;  /* LearningConfig::policy_weight */
    float value_weight;   /* LearningConfig::value_weight */
    uint32_t batch_size;  /* LearningConfig::batch_size (config.rs:76-94): a multiple of 32 up to 4096; larger than 32 = gradients summed over micro-batches of 32 */
} syn_train_cfg;

/* Replaces the batch loop of alpha_zero (alpha_zero.rs:73-92): n_batches optimizer steps on the engine's CURRENT
 * weights, batch k = rows batch_index[k][0..batch_size) of the FlatBatch given as (my_bb, op_bb, pis, vs) — features are
 * synthesised from the bitboards.  batch_index is what BatchRandSampler (data.rs:6-64) yields: chunks of a random
 * permutation, last partial chunk dropped; the permutation is the caller's (torch's randperm stream is not
 * reproduced).  Per step: forward, log_softmax, kl_div(Reduction::Sum)/batch_size per head, policy_weight*pi_loss +
 * value_weight*v_loss, backward, Adam.  losses (optional, [n_batches][2]) = {pi_loss, v_loss} per step as the reference
 * accumulates them.  Adam's moments and step count persist in the engine across calls (syn_engine_reset_optimizer
 * clears them); afterwards gathers on this engine search with the updated weights.  fp32 throughout; results agree
 * with libtorch's CPU ops to ~1e-5 relative per step (summation order differs). */
int syn_engine_train(syn_engine* e, const syn_train_cfg* cfg, const uint64_t* my_bb, const uint64_t* op_bb, const float* pis,
                     const float* vs, size_t n_rows, const uint32_t* batch_index, uint32_t n_batches, float* losses,
                     syn_stats* stats);
int syn_engine_reset_optimizer(syn_engine* e);

/* Replaces `vs.save` (alpha_zero.rs:37, 102) as far as the engine is concerned: the current fp32 weights in the blob
 * layout of syn_engine_set_weights (host or device destination). */
int syn_engine_get_weights(syn_engine* e, float* blob, size_t n_floats);

/* ---- multi-GPU: one process (or thread) per GPU, games sharded by contiguous ranges of the global game index --------
 * The reference fans games out over num_workers+1 OS threads, each loading models/<name>.ot, and joins their
 * ReplayBuffers in worker order (alpha_zero.rs:132-168, 192-194; data.rs:160-170).  Here a rank is a worker: ONE broadcast
 * of the weight blob and ONE gather of experience rows per iteration, NCCL over NVLink inside the library (bound at run
 * time by soname, so a host that already carries NCCL — PyTorch — shares its copy).  The search itself needs no
 * collective.  The host only moves the 128-byte communicator id from rank 0 to the other ranks (any transport). */
typedef struct syn_comm syn_comm;
#define SYN_COMM_ID_BYTES 128
int syn_comm_unique_id(uint8_t id[SYN_COMM_ID_BYTES]);  /* ncclGetUniqueId: call on ONE rank, hand the bytes to all */
int syn_comm_create(const uint8_t id[SYN_COMM_ID_BYTES], int n_ranks, int rank, int cuda_device, syn_comm** out); /* collective */
void syn_comm_destroy(syn_comm* c);
int syn_comm_rank(const syn_comm* c);
int syn_comm_size(const syn_comm* c);

/* Replaces every worker's `vs.load(models/<name>.ot)` (alpha_zero.rs:192-194) across GPUs: rank `root` contributes
 * `blob` (host or device, layout of syn_engine_set_weights; NULL = the root engine's current weights, e.g. fresh from
 * syn_engine_train), every rank's engine ends up with them as if by syn_engine_set_weights.  Collective; blob is ignored
 * on the other ranks. */
int syn_engine_broadcast_weights(syn_engine* e, syn_comm* c, const float* blob, size_t n_floats, int root);

/* Replaces gather_experience's fan-out and join (alpha_zero.rs:132-168) across GPUs: every rank plays games
 * [first_game_index, first_game_index + num_games) (its shard: ranks must hold ascending contiguous ranges in rank order
 * for the root's rows to come out in game order, like `extend` in worker order; num_games may be 0), the rows travel to
 * rank `root` as 72 bytes each (game id, bitboards, pi, v) and the root rebuilds height / player / features from the
 * bitboards.  On the root `out` receives all ranks' rows in rank order (capacity must cover the sum; len is set; games is
 * not); elsewhere `out` is ignored and may be NULL.  stats (optional) are this rank's.  Collective. */
int syn_engine_gather_experience(syn_engine* e, syn_comm* c, int root, const syn_rollout_cfg* cfg, uint64_t first_game_index,
                                 uint32_t num_games, uint64_t seed, syn_experience* out, syn_stats* stats);

/* Optional per-row trace of the NEXT gather (arrays of `capacity` rows like syn_experience, host or
 * device; NULL to disable): the action played from the row's state, nodes.len() of that ply's
 * tree, and the root's child visit counts by column.  Not part of the reference's ReplayBuffer;
 * it is what the parity tests compare ("bit-exact visit counts"). */
int syn_engine_set_trace(syn_engine* e, uint8_t* action, uint32_t* tree_nodes, float* child_visits /*[cap][9]*/);

/* Lanes per game: 0 (default) = chosen per launch from the games in flight at the measured crossovers — a lane group per
 * game (network leaves: 32 lanes up to 2,368 games, 16 up to ~11,000; rollout leaves: 32 up to 40,000: children scored and
 * playouts played in parallel, the shortest time per explore; seats are refilled as games end) and a thread per game beyond
 * (128 games = one tensor-core tile of leaves: the highest throughput); 1, 16, 32 force a mapping.  The SYN_GROUP_LANES
 * environment variable sets the default at syn_engine_create.  With rollout leaves 1 is a thread per game too (the thread
 * plays the rollout itself; SYN_ROLLOUT_THREADS = 512 / 640 / 768 / 896 / 1024 games per CTA, default 1024).  Results do not
 * depend on any of these. */
int syn_engine_set_group_lanes(syn_engine* e, int lanes);

/* How the thread-per-game kernels would seat a gather of num_games games on this engine: persistent CTAs launched and the
 * most games any of them holds.  Games are dealt evenly over ALL SMs (and, inside a CTA, over its warps): 1,000 games —
 * the reference's games_per_train, study-connect4/src/main.rs:26 — are 148 CTAs of at most 7, not two CTAs of 640.
 * Diagnostic; results never depend on the seating. */
int syn_engine_launch_geometry(syn_engine* e, uint32_t num_games, uint32_t leaf_eval_kind, uint32_t* ctas, uint32_t* games_per_cta,
                               uint32_t* lanes_per_game);

/* How Connect4Net (study-connect4/src/policies.rs:28-59, fp32 through libtorch in the reference) is evaluated:
 *   3 (default) = auto: every time the weights change the engine MEASURES the single-fp16 chain against the split chain on
 *       1,024 reachable positions and uses the fast one only while its largest error stays below a quarter of the
 *       tolerance BASELINE.json states (1e-3 abs + 1e-3 rel, logits and outcome probabilities).  Random-init weights
 *       pass with a margin of 25; trained-size weights do not and get the split chain;
 *   2 = tcgen05 tensor cores, split-fp16 operands (x = hi + lo, three MMAs per K-step), fp32 accumulate, activations in
 *       tensor memory: agrees with an fp32 forward to ~1e-7 at any weight scale;
 *   1 = tcgen05 tensor cores, single fp16 operands (11-bit significands), fp32 accumulate: ~4e-5 at initialisation scale,
 *       several 1e-3 at trained scale.  Matches between two DIFFERENT networks (syn_engine_set_opponent_weights) always
 *       run this chain (two pairs of split weight images do not fit one SM);
 *   0 = fp32 CUDA-core kernel (kept as the device-side numerical reference).
 * SYN_MLP=fp32|fp16|split overrides the default at syn_engine_create.  syn_engine_mlp_in_use reports the chain the next
 * launch will use (0, 1 or 2) and, in auto mode, the measured error of the fast chain in units of the tolerance (-1 if
 * not measured). */
int syn_engine_set_mlp_mode(syn_engine* e, int mode);
int syn_engine_mlp_in_use(syn_engine* e, int* chain, float* calibration_ratio);

/* Profiling aid, not part of the reference's surface: per-warp clock totals of the last thread-per-game
 * launch, out[0..7) = cycles in {tree advance, waiting for the team, Connect4Net forward, explore
 * finish}, rounds, leaves evaluated, total cycles (each summed over warps). */
int syn_engine_debug_counters(syn_engine* e, uint64_t* out, uint32_t n);

/* Packed Option<Outcome> (synthesis/src/game.rs:9-14): 0 = None, else kind<<6 | turns with
 * kind 1 = Lose, 2 = Draw, 3 = Win; turns <= 63. */
#define SYN_OUTCOME_NONE 0u
#define SYN_OUTCOME_KIND(o) ((unsigned)(o) >> 6)
#define SYN_OUTCOME_TURNS(o) ((unsigned)(o)&63u)
#define SYN_KIND_LOSE 1u
#define SYN_KIND_DRAW 2u
#define SYN_KIND_WIN 3u

#ifdef __cplusplus
}
#endif
#endif /* SYNTHESIS_B200_H */
