/* syn_streams.h — how per-game random streams are derived from (seed, game index).
 *
 * The reference gives each WORKER THREAD one StdRng (alpha_zero.rs:189) that all of the
 * worker's games consume one after another, so game k's draws depend on how long games
 * 0..k-1 of the same worker lasted.  That is a serial dependence between games and makes the
 * result depend on the thread count.  The engine instead gives every GAME private streams
 * (run_game is generic over the rng it is handed: alpha_zero.rs:229-233), all
 * `StdRng::seed_from_u64(syn_stream_seed(seed, g, k))`:
 *   k = 0  rollout policy rng   (policies/rollout.rs:5-7)
 *   k = 1  action sampling rng  (alpha_zero.rs:270-294)
 *   k = 2  Dirichlet root noise (mcts.rs:236 uses thread_rng)
 *   k = 3  Normal FPU           (study-connect4/src/main.rs:43-47 uses thread_rng)
 * With seed = 0 the first two are 2g and 2g+1, the convention of SURVEY.md §8(c)'s vectors.
 * Results therefore do not depend on how games are sharded over GPUs or threads.
 *
 * Domain: seed < 2^30 and game_index < 2^32 (SYN_MAX_SEED, SYN_MAX_GAME_INDEX) — inside it distinct (seed, game, k) give
 * distinct stream seeds; beyond it bits would fall off the top or run into the k selector, so the ABI refuses such
 * arguments (the reference's own seed is the iteration index, alpha_zero.rs:49, 140).
 */
#ifndef SYN_STREAMS_H
#define SYN_STREAMS_H
#include <stdint.h>

#define SYN_STREAM_ROLLOUT 0u
#define SYN_STREAM_ACTION 1u
#define SYN_STREAM_NOISE 2u
#define SYN_STREAM_FPU 3u
#define SYN_MAX_SEED ((1ull << 30) - 1ull)
#define SYN_MAX_GAME_INDEX ((1ull << 32) - 1ull)

#if defined(__CUDACC__)
__host__ __device__
#endif
static inline uint64_t syn_stream_seed(uint64_t seed, uint64_t game_index, unsigned k) {
    return (seed << 33) + 2ull * game_index + (uint64_t)(k & 1u) + ((uint64_t)(k >> 1) << 63);
}

#endif
