// engine.cu — host side of libsynthesis_b200.so: the C ABI declared in include/synthesis_b200.h.
//
// Each entry point names the reference interface it replaces in the header.  This file owns
// device memory (tree arenas, row buffers), launches the kernels of selfplay.cuh on the engine's
// stream, times them with CUDA events on that stream, and moves results to the caller's buffers.
// There is no CPU implementation behind any entry point: without a compute-capability-10.x device
// syn_engine_create fails with SYN_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> // types and prototypes only: the library is bound at run time (see NcclApi)

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "selfplay.cuh"
#include "selfplay_team.cuh"
#include "match.cuh"
#include "tpg2.cuh"
#include "tpg2_rollout.cuh"
#include "tpg2_split.cuh"
#include "tpg4.cuh"
#include "tpg4_rollout.cuh"
#include "dedup.cuh"
#include "train.cuh"
#include "train_cluster.cuh"

using namespace eng;

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) return fail(SYN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------ engine
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    // The first allocation is exact (the arenas are sized by the caller); a buffer that has to GROW takes half as much again, so
    // that a sequence of slightly larger requests (the rows of a gather vary by a few percent from seed to seed) does not free
    // and allocate device memory call after call — each such pair synchronises the device and showed up as 0.4-1.1 s stalls in
    // the end-to-end leg of a 0.2 s workload.
    cudaError_t reserve(size_t want) {
        if (want <= n) return cudaSuccess;
        size_t cap = p ? (want > n + n / 2 ? want : n + n / 2) : want;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
        if (e != cudaSuccess && cap != want) { cap = want; e = cudaMalloc(&p, cap * sizeof(T)); }
        if (e == cudaSuccess) n = cap;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct syn_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t max_games = 0, max_explores = 0, arena_nodes = 0; // max_games = arena slots allocated (>= req_games)
    uint32_t req_games = 0; // max_games_in_flight as the caller asked for it
    int tpg_ver = 2;        // thread-per-game tree layout: 2 = tpg2.cuh (32-byte records, the product path), 4 = tpg4.cuh (family blocks; SYN_TPG_VER=4)
    int group_lanes = 0;   // lanes per game: 0 = chosen per launch from the games in flight (pick_group_lanes), 32, 16, or 1 (thread per game)
    int tpg_teams = 5;     // teams of 128 threads per CTA in thread-per-game mode (640 threads, 96 registers each; SYN_TPG_TEAMS)
    int rollout_threads = 1024; // threads (= games) per CTA of the thread-per-game rollout kernel: 512, 640, 768, 896 or 1024
    int rollout_cw = 3;         // child records per memory round trip at 896 / 1024 threads (SYN_ROLLOUT_CW = 3 or 5)
    int lg_teams = 4;      // SYN_LG_TEAMS: teams per CTA of the lane-group network kernel (4 or 5; 0 = the CTA-wide tile of selfplay_nn_tc_kernel)
    bool tpg_prof = false; // SYN_TPG_PROF=1: the instantiation with per-warp phase clocks
   // 2 = round-synchronous kernel (tpg2.cuh), 1 = the first thread-per-game kernel (tpg.cuh)
    DevBuf<uint32_t> slot_state;
    DevBuf<uint32_t> fpu_state; // tp2::FS_WORDS per slot
    DevBuf<uint4> nodes; // 2 x uint4 = one 32-byte record per tree node
    DevBuf<float> weights;
    DevBuf<uint8_t> weight_image; // mlptc layout (fp16 weights + fp32 biases)
    DevBuf<float> weights2;        // players[1]'s network in a match between two different networks (syn_engine_set_opponent_weights)
    DevBuf<uint8_t> weight_image2;
    bool has_weights2 = false;
    bool has_weights = false;
    float bias_host[mlptc::BIAS_FLOATS] = {}; // the weight image's biases, for KParams::mlp_bias
    bool use_tc = true;           // Connect4Net on tcgen05 tensor cores (false: fp32 CUDA-core kernel)
    int mlp_mode = 3;             // requested: 3 = auto (default: see calibrate_mlp), 2 = split-fp16 operands, fp32-grade (mlp_split.cuh), 1 = single fp16 operands (mlp_team.cuh), 0 = fp32 CUDA cores
    int mlp_eff = 2;              // the chain in use: 0, 1 or 2
    float calib_ratio = -1.0f;    // auto mode: the single-fp16 chain's largest error on the calibration positions, in units of the tolerance 1e-3 + 1e-3 |y|
    DevBuf<uint64_t> calib_pos;   // 2 x CALIB_N bitboards of reachable positions
    DevBuf<float> calib_out;      // 2 x CALIB_N x 12 outputs
    DevBuf<uint8_t> weight_image_lo; // mlp_split.cuh: fp16(W - fp16(W)) in the layout of weight_image
    DevBuf<unsigned int> next_game;
    DevBuf<unsigned long long> counters;
    DevBuf<int> error;
    // per-call row buffers
    DevBuf<uint64_t> row_my, row_op, row_off;
    DevBuf<float> row_pi, row_v, row_visits;
    DevBuf<uint8_t> row_action;
    DevBuf<uint32_t> row_nodes, game_len;
    // dense staging for host destinations
    DevBuf<uint8_t> staging;
    // search-mode buffers
    DevBuf<uint64_t> pos_my, pos_op, pos_seed;
    DevBuf<float> s_visits, s_q;
    DevBuf<uint8_t> s_csol, s_rsol, s_best;
    DevBuf<uint32_t> s_nodes;
    // learner state (train.cuh): Adam moments in the padded parameter layout, optimizer step count
    DevBuf<float> adam_m, adam_v;
    DevBuf<uint8_t> tr_io;
    uint64_t adam_t = 0;
    // deduplicate workspace (dedup.cuh)
    DevBuf<uint8_t> dd_ws, dd_io;
    // multi-GPU gather (syn_engine_gather_experience): this rank's rows in wire format, the root's landing zone, row counts
    DevBuf<uint8_t> gx_local, gx_all;
    DevBuf<unsigned long long> gx_counts;
    // pending gather
    bool pending = false;
    uint32_t pend_games = 0;
    uint64_t pend_first = 0;
    uint64_t launches = 0, h2d = 0, d2h = 0;
    // trace destination (optional)
    uint8_t* trace_action = nullptr;
    uint32_t* trace_nodes = nullptr;
    float* trace_visits = nullptr;
};

constexpr uint32_t CALIB_N = 1024;
constexpr float CALIB_MAX_RATIO = 0.25f; // the fast chain is used while its observed error stays below a quarter of the tolerance

// Connect4::won on the host (connect4.rs:77-83), for argument validation only
static bool host_won(uint64_t bb) {
    const uint64_t d1 = bb & (bb >> 6) & (bb >> 12) & (bb >> 18) & c4::D1_MASK, d2 = bb & (bb >> 8) & (bb >> 16) & (bb >> 24) & c4::D2_MASK;
    const uint64_t h = bb & (bb >> 7) & (bb >> 14) & (bb >> 21) & c4::H_MASK, v = bb & (bb >> 1) & (bb >> 2) & (bb >> 3) & c4::V_MASK;
    return (d1 | d2 | h | v) != 0;
}

// CALIB_N positions reached by random legal play from the empty board (0 .. 40 plies, never past the end of a game): what the
// network is asked about during self-play.  Deterministic (a fixed xorshift stream), so every engine and every rank calibrates alike.
static void calibration_positions(std::vector<uint64_t>& my, std::vector<uint64_t>& op) {
    my.assign(CALIB_N, 0); op.assign(CALIB_N, 0);
    uint64_t s = 0x9e3779b97f4a7c15ull;
    auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (uint32_t i = 0; i < CALIB_N; ++i) {
        for (;;) {
            uint64_t a = 0, b = 0; // a = player to move
            const uint32_t plies = (uint32_t)(next() % 41);
            bool ok = true;
            for (uint32_t k = 0; k < plies && ok; ++k) {
                const uint64_t occ = a | b;
                int cols[9], n = 0;
                for (int c = 0; c < 9; ++c)
                    if (!((occ >> (7 * c + 6)) & 1ull)) cols[n++] = c;
                if (n == 0) { ok = false; break; }
                const int c = cols[next() % (uint64_t)n];
                const uint64_t bit = (occ + (1ull << (7 * c))) & (0x7full << (7 * c));
                const uint64_t mover = a | bit;
                a = b; b = mover;
                if (host_won(mover) || (occ | bit) == c4::ALL) ok = false;
            }
            if (ok) { my[i] = a; op[i] = b; break; }
        }
    }
}

static bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static int validate_cfg(const syn_rollout_cfg* cfg, const syn_engine* e) {
    if (!cfg) return fail(SYN_ERR_INVALID_ARGUMENT, "cfg is NULL");
    if (cfg->num_explores > e->max_explores)
        return fail(SYN_ERR_CAPACITY, "num_explores %u exceeds the engine's max_explores %u", cfg->num_explores, e->max_explores);
    const syn_mcts_cfg& m = cfg->mcts;
    if (m.exploration_kind > SYN_EXPLORATION_POLYNOMIAL_UCT) return fail(SYN_ERR_INVALID_ARGUMENT, "bad exploration_kind %u", m.exploration_kind);
    if (m.fpu_kind == SYN_FPU_FUNC)
        return fail(SYN_ERR_UNSUPPORTED, "Fpu::Func carries host code and cannot run on the device; use SYN_FPU_NORMAL{mean,std} for the shipped closure");
    if (m.fpu_kind > SYN_FPU_FUNC) return fail(SYN_ERR_INVALID_ARGUMENT, "bad fpu_kind %u", m.fpu_kind);
    if (m.noise_kind > SYN_NOISE_DIRICHLET) return fail(SYN_ERR_INVALID_ARGUMENT, "bad noise_kind %u", m.noise_kind);
    if (m.noise_kind == SYN_NOISE_DIRICHLET && !(m.noise_alpha > 0.0f)) return fail(SYN_ERR_INVALID_ARGUMENT, "Dirichlet alpha must be > 0");
    if (cfg->value_target_kind > SYN_VALUE_Q_TO_Z) return fail(SYN_ERR_INVALID_ARGUMENT, "bad value_target_kind %u", cfg->value_target_kind);
    if (cfg->action_selection > SYN_ACTION_NUM_VISITS) return fail(SYN_ERR_INVALID_ARGUMENT, "bad action_selection %u", cfg->action_selection);
    if (cfg->leaf_eval_kind > SYN_LEAF_ROLLOUT) return fail(SYN_ERR_INVALID_ARGUMENT, "bad leaf_eval_kind %u", cfg->leaf_eval_kind);
    if (cfg->leaf_eval_kind == SYN_LEAF_NN && !e->has_weights)
        return fail(SYN_ERR_NO_WEIGHTS, "leaf_eval_kind = NN but syn_engine_set_weights has not been called");
    return SYN_OK;
}

template <int TEAMS, int SLOTS>
static int launch_tpg(syn_engine* e, KParams& kp, uint32_t blocks) {
    size_t smem = sizeof(mlpteam::Smem<TEAMS, SLOTS>) + (size_t)tp2::path_cap(TEAMS) * 128 * TEAMS * sizeof(uint32_t);
    if (e->tpg_prof) { // with per-warp phase clocks (syn_engine_debug_counters)
        CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tpg2_kernel<TEAMS, SLOTS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        selfplay_nn_tpg2_kernel<TEAMS, SLOTS, true><<<blocks, 128 * TEAMS, smem, e->stream>>>(kp);
        return SYN_OK;
    }
    if (TEAMS == 5 && kp.cfg.mcts.fpu_kind == SYN_FPU_NORMAL) { // the shipped first-play urgency: its own instantiation with the cached stream
        constexpr int T = TEAMS == 5 ? 5 : 1, S = TEAMS == 5 ? SLOTS : 1; // (only <5, 4> is instantiated)
        CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tpg2_kernel<T, S, false, tp2::FPU_NORMAL_CACHED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        selfplay_nn_tpg2_kernel<T, S, false, tp2::FPU_NORMAL_CACHED><<<blocks, 128 * TEAMS, smem, e->stream>>>(kp);
        return SYN_OK;
    }
    CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tpg2_kernel<TEAMS, SLOTS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selfplay_nn_tpg2_kernel<TEAMS, SLOTS, false><<<blocks, 128 * TEAMS, smem, e->stream>>>(kp);
    return SYN_OK;
}

template <int TEAMS>
static int launch_tpg_split(syn_engine* e, KParams& kp, uint32_t blocks) { // network leaves at fp32-grade accuracy (tpg2_split.cuh)
    const size_t smem = tp2s::smem_bytes<TEAMS>();
    if (e->tpg_prof) {
        CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tpg2s_kernel<TEAMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        selfplay_nn_tpg2s_kernel<TEAMS, true><<<blocks, 128 * TEAMS, smem, e->stream>>>(kp);
        return SYN_OK;
    }
    if (TEAMS == 5 && kp.cfg.mcts.fpu_kind == SYN_FPU_NORMAL) {
        constexpr int T = TEAMS == 5 ? 5 : 1;
        CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tpg2s_kernel<T, false, tp2::FPU_NORMAL_CACHED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        selfplay_nn_tpg2s_kernel<T, false, tp2::FPU_NORMAL_CACHED><<<blocks, 128 * TEAMS, smem, e->stream>>>(kp);
        return SYN_OK;
    }
    CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tpg2s_kernel<TEAMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selfplay_nn_tpg2s_kernel<TEAMS, false><<<blocks, 128 * TEAMS, smem, e->stream>>>(kp);
    return SYN_OK;
}

template <int NT, int CW>
static int launch_rollout_tpg(syn_engine* e, KParams& kp, uint32_t blocks) {
    const size_t smem = tp2r::smem_bytes(NT);
    if (NT == 1024 && CW == 3 && kp.cfg.mcts.fpu_kind == SYN_FPU_NORMAL) { // the default geometry with the cached FPU stream
        constexpr int N = (NT == 1024 && CW == 3) ? 1024 : 512, C = (NT == 1024 && CW == 3) ? 3 : 5;
        CUDA_TRY(cudaFuncSetAttribute(selfplay_rollout_tpg2_kernel<N, C, tp2::FPU_NORMAL_CACHED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        selfplay_rollout_tpg2_kernel<N, C, tp2::FPU_NORMAL_CACHED><<<blocks, NT, smem, e->stream>>>(kp);
        return SYN_OK;
    }
    if (NT == 1024 && CW == 3 && kp.cfg.mcts.fpu_kind == SYN_FPU_CONST) { // the default geometry and the default FPU: an instantiation with ONE inlined descent
        constexpr int N = (NT == 1024 && CW == 3) ? 1024 : 512, C = (NT == 1024 && CW == 3) ? 3 : 5;
        CUDA_TRY(cudaFuncSetAttribute(selfplay_rollout_tpg2_kernel<N, C, SYN_FPU_CONST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        selfplay_rollout_tpg2_kernel<N, C, SYN_FPU_CONST><<<blocks, NT, smem, e->stream>>>(kp);
        return SYN_OK;
    }
    CUDA_TRY(cudaFuncSetAttribute(selfplay_rollout_tpg2_kernel<NT, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selfplay_rollout_tpg2_kernel<NT, CW><<<blocks, NT, smem, e->stream>>>(kp);
    return SYN_OK;
}

template <int TEAMS, bool PROF, int FPU>
static int launch_tpg4_k(syn_engine* e, KParams& kp, uint32_t blocks) {
    const size_t smem = tp2s::smem_bytes<TEAMS>();
    CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tpg4_kernel<TEAMS, PROF, FPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selfplay_nn_tpg4_kernel<TEAMS, PROF, FPU><<<blocks, 128 * TEAMS, smem, e->stream>>>(kp);
    return SYN_OK;
}
template <int TEAMS>
static int launch_tpg4(syn_engine* e, KParams& kp, uint32_t blocks) {
    const uint32_t fpu = kp.cfg.mcts.fpu_kind;
    if (fpu == SYN_FPU_PARENT_Q) return launch_tpg4_k<TEAMS, false, SYN_FPU_PARENT_Q>(e, kp, blocks);
    if (fpu == SYN_FPU_NORMAL) return launch_tpg4_k<TEAMS, false, SYN_FPU_NORMAL>(e, kp, blocks);
    if (e->tpg_prof) return launch_tpg4_k<TEAMS, true, SYN_FPU_CONST>(e, kp, blocks); // per-warp phase clocks (syn_engine_debug_counters)
    return launch_tpg4_k<TEAMS, false, SYN_FPU_CONST>(e, kp, blocks);
}

template <int NT, int FPU>
static int launch_rollout_tpg4_k(syn_engine* e, KParams& kp, uint32_t blocks) {
    const size_t smem = tp2r::smem_bytes(NT);
    CUDA_TRY(cudaFuncSetAttribute(selfplay_rollout_tpg4_kernel<NT, FPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selfplay_rollout_tpg4_kernel<NT, FPU><<<blocks, NT, smem, e->stream>>>(kp);
    return SYN_OK;
}
template <int NT>
static int launch_rollout_tpg4(syn_engine* e, KParams& kp, uint32_t blocks) {
    const uint32_t fpu = kp.cfg.mcts.fpu_kind;
    if (fpu == SYN_FPU_PARENT_Q) return launch_rollout_tpg4_k<NT, SYN_FPU_PARENT_Q>(e, kp, blocks);
    if (fpu == SYN_FPU_NORMAL) return launch_rollout_tpg4_k<NT, SYN_FPU_NORMAL>(e, kp, blocks);
    return launch_rollout_tpg4_k<NT, SYN_FPU_CONST>(e, kp, blocks);
}

// Seating of a thread-per-game launch (tp2::seat_of): `want` games in flight spread over all SMs, at most team_size * teams seats per
// CTA.  Returns the grid size.
static uint32_t seat_games(const syn_engine* e, KParams& kp, uint32_t team_size, uint32_t teams) {
    uint32_t want = kp.num_games < e->req_games ? kp.num_games : e->req_games;
    if (want == 0) want = 1;
    uint32_t blocks = want < (uint32_t)e->sm_count ? want : (uint32_t)e->sm_count;
    const uint32_t cta_cap = team_size * teams;
    if ((uint64_t)blocks * cta_cap < want) want = blocks * cta_cap;
    kp.seats_q = want / blocks;
    kp.seats_rem = want % blocks;
    const uint32_t per_cta = kp.seats_q + (kp.seats_rem ? 1u : 0u);
    kp.teams_used = (per_cta + team_size - 1) / team_size;
    kp.per_team = (per_cta + kp.teams_used - 1) / kp.teams_used;
    // the arena slots of a CTA are teams_used * per_team apart: never more than the engine allocated
    while ((uint64_t)blocks * kp.teams_used * kp.per_team > e->max_games && kp.per_team > 1) {
        kp.per_team -= 1;
        const uint32_t cap = kp.teams_used * kp.per_team;
        if (kp.seats_q >= cap) { kp.seats_q = cap; kp.seats_rem = 0; }
    }
    return blocks;
}

static size_t nn_tc_smem_bytes(int gpb) { return sizeof(mlptc::Smem) + (size_t)gpb * 64 * sizeof(uint32_t); }
static size_t nn_smem_bytes(int gpb) { return (size_t)(mlp::WEIGHT_FLOATS + 2 * gpb * mlp::XS + gpb * 64) * sizeof(float); }

constexpr int ROLLOUT_THREADS = 256;
constexpr int NN_THREADS = 512;
#ifndef SYN_NN_TC_THREADS
#define SYN_NN_TC_THREADS 512
#endif
constexpr int NN_TC_THREADS = SYN_NN_TC_THREADS; // CTA-wide lane-group kernel (SYN_LG_TEAMS=0): 512 = one CTA per SM (faster: profiles/r2_lane_group_clocks.txt), 256 = two

// Which mapping a launch uses when the caller has not forced one.  A tree is a strictly serial object, so with FEW games in
// flight the time of one explore is what matters: a lane group per game scores a node's children in parallel, plays a
// playout's plies in parallel windows and reads a family with one instruction (tree.cuh, selfplay_team.cuh).  With many games
// the thread-per-game kernels win by keeping every lane busy.  The thresholds below are the measured crossovers.
static int pick_group_lanes(const syn_engine* e, const KParams& kp) {
    if (e->group_lanes != 0) return e->group_lanes;
    const bool nn = kp.cfg.leaf_eval_kind == SYN_LEAF_NN;
    const uint32_t want = kp.num_games < e->req_games ? kp.num_games : e->req_games;
    if (nn) {
        if (e->mlp_eff == 0 || (e->mlp_eff == 2 && e->lg_teams == 0)) return 1; // no lane-group kernel carries this chain: thread per game
        if (want <= (uint32_t)e->sm_count * 16u) return 32;    // one wave of warps: 2,368 games
        // half warps: 4,736 seats, refilled as games end.  Single-fp16 chain: ahead of a thread per game up to ~11 k games
        // (6,000: 193 vs 139 M explores/s, 10,000: 220 vs 208, 16,000: 247 vs 307); split chain: level at 4,096 (125 vs 124)
        if (want <= (e->mlp_eff == 1 ? 11000u : (uint32_t)e->sm_count * 32u)) return 16;
        return 1;
    }
    // rollout leaves: 9,472 seats of 32 lanes, refilled; ahead up to ~45 k games (9,000: 354 vs 158 M explores/s, 20,000: 437 vs 256,
    // 40,000: 467 vs 410), then a thread per game takes over on its way to 1.2 G at 151,552 (profiles/r2_lane_group_clocks.txt)
    if (want <= 40000u) return 32;
    return 1;
}

// Geometry of a lane-group launch: persistent CTAs, all resident, groups dealt evenly over them (lg_seated).
struct LgGeometry { int gl, threads, gpb; uint32_t blocks, seats_q, seats_rem; bool tc; };
static bool lg_geometry(const syn_engine* e, const KParams& kp, int lanes, LgGeometry& G) {
    const bool nn = kp.cfg.leaf_eval_kind == SYN_LEAF_NN;
    G.gl = lanes == 1 ? 16 : lanes; // NN leaves without tensor cores land here at "1" too
    G.tc = nn && e->mlp_eff != 0;
    // tensor-core network leaves: 256-thread CTAs, two per SM (107 KB of shared memory each) — their rounds interleave and
    // half as many groups share a CTA barrier; fp32 network leaves: one 512-thread CTA per SM (120 KB of fp32 weights)
    const int teams = e->mlp_eff == 2 ? 4 : e->lg_teams; // the split chain is instantiated for four teams
    G.threads = G.tc ? (e->lg_teams ? 128 * teams : NN_TC_THREADS) : nn ? NN_THREADS : ROLLOUT_THREADS;
    G.gpb = G.threads / G.gl;
    const uint32_t max_blocks = e->max_games / (uint32_t)G.gpb;
    if (max_blocks == 0) return false;
    const uint32_t resident = (uint32_t)e->sm_count * (G.tc ? ((e->lg_teams == 0 && NN_TC_THREADS <= 256) ? 2u : 1u) : nn ? 1u : (uint32_t)(2048 / ROLLOUT_THREADS));
    G.blocks = max_blocks < resident ? max_blocks : resident;
    if (G.blocks > kp.num_games) G.blocks = kp.num_games ? kp.num_games : 1u;
    const uint64_t seats = (uint64_t)G.blocks * (uint32_t)G.gpb;
    uint32_t want = kp.num_games < seats ? kp.num_games : (uint32_t)seats;
    if (want > e->req_games) want = e->req_games; // never more games in flight than the caller allowed
    if (want == 0) want = 1;
    if (G.blocks > want) G.blocks = want;
    G.seats_q = want / G.blocks;
    G.seats_rem = want % G.blocks;
    return true;
}

// Launches the self-play kernel for `n` games/positions.  Rows/search buffers must be set in kp.
static int launch_selfplay(syn_engine* e, KParams& kp) {
    const bool nn = kp.cfg.leaf_eval_kind == SYN_LEAF_NN;
    const int lanes = pick_group_lanes(e, kp);
    const bool tpg4_teams = e->tpg_teams == 4 || e->tpg_teams == 5 || e->tpg_teams == 6;
    const bool tpg4_nt = e->rollout_threads == 512 || e->rollout_threads == 768 || e->rollout_threads == 1024;
    const bool tpg4_ok = e->tpg_ver == 4 && e->max_explores <= tp4::MAX_EXPLORES && (e->arena_nodes >> 2) <= tp4::MAX_LINES;
    if (nn && lanes == 1 && e->mlp_eff == 2 && tpg4_ok && tpg4_teams) { // thread per game on family blocks (tpg4.cuh), SYN_TPG_VER=4
        const uint32_t blocks = seat_games(e, kp, 128u, (uint32_t)e->tpg_teams);
        CUDA_TRY(cudaMemsetAsync(e->next_game.p, 0, sizeof(unsigned int), e->stream));
        int rc = e->tpg_teams == 6 ? launch_tpg4<6>(e, kp, blocks) : e->tpg_teams == 5 ? launch_tpg4<5>(e, kp, blocks) : launch_tpg4<4>(e, kp, blocks);
        if (rc) return rc;
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
        return SYN_OK;
    }
    if (!nn && lanes == 1 && tpg4_ok && tpg4_nt) { // rollout leaves on family blocks (tpg4_rollout.cuh), SYN_TPG_VER=4
        const uint32_t blocks = seat_games(e, kp, (uint32_t)e->rollout_threads, 1u);
        CUDA_TRY(cudaMemsetAsync(e->next_game.p, 0, sizeof(unsigned int), e->stream));
        int rc = e->rollout_threads == 1024 ? launch_rollout_tpg4<1024>(e, kp, blocks)
                 : e->rollout_threads == 768 ? launch_rollout_tpg4<768>(e, kp, blocks) : launch_rollout_tpg4<512>(e, kp, blocks);
        if (rc) return rc;
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
        return SYN_OK;
    }
    if (nn && lanes == 1 && e->mlp_eff == 2 && (e->tpg_teams == 4 || e->tpg_teams == 5 || e->tpg_teams == 6)) { // the product path for network leaves
        const uint32_t blocks = seat_games(e, kp, 128u, (uint32_t)e->tpg_teams);
        CUDA_TRY(cudaMemsetAsync(e->next_game.p, 0, sizeof(unsigned int), e->stream));
        int rc = e->tpg_teams == 6 ? launch_tpg_split<6>(e, kp, blocks) : e->tpg_teams == 5 ? launch_tpg_split<5>(e, kp, blocks) : launch_tpg_split<4>(e, kp, blocks);
        if (rc) return rc;
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
        return SYN_OK;
    }
    if (nn && lanes == 1 && e->mlp_eff != 0) { // thread per game (tpg2.cuh), single-fp16 forward: one persistent CTA per SM, games seated over all SMs
        const uint32_t blocks = seat_games(e, kp, 128u, (uint32_t)e->tpg_teams);
        CUDA_TRY(cudaMemsetAsync(e->next_game.p, 0, sizeof(unsigned int), e->stream));
        int rc = e->tpg_teams == 8 ? launch_tpg<8, 4>(e, kp, blocks)
                 : e->tpg_teams == 6 ? launch_tpg<6, 4>(e, kp, blocks)
                 : e->tpg_teams == 5 ? launch_tpg<5, 4>(e, kp, blocks)
                 : e->tpg_teams == 4 ? launch_tpg<4, 4>(e, kp, blocks)
                 : e->tpg_teams == 2 ? launch_tpg<2, 2>(e, kp, blocks) : launch_tpg<1, 1>(e, kp, blocks);
        if (rc) return rc;
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
        return SYN_OK;
    }
    if (!nn && lanes == 1) { // rollout leaves, thread per game (tpg2_rollout.cuh): one persistent CTA per SM
        const uint32_t blocks = seat_games(e, kp, (uint32_t)e->rollout_threads, 1u);
        CUDA_TRY(cudaMemsetAsync(e->next_game.p, 0, sizeof(unsigned int), e->stream));
        const bool cw5 = e->rollout_cw == 5;
        int rc = e->rollout_threads == 1024 ? (cw5 ? launch_rollout_tpg<1024, 5>(e, kp, blocks) : launch_rollout_tpg<1024, 3>(e, kp, blocks))
                 : e->rollout_threads == 896 ? (cw5 ? launch_rollout_tpg<896, 5>(e, kp, blocks) : launch_rollout_tpg<896, 3>(e, kp, blocks))
                 : e->rollout_threads == 768 ? launch_rollout_tpg<768, 3>(e, kp, blocks)
                 : e->rollout_threads == 640 ? launch_rollout_tpg<640, 5>(e, kp, blocks) : launch_rollout_tpg<512, 5>(e, kp, blocks);
        if (rc) return rc;
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
        return SYN_OK;
    }
    LgGeometry G;
    if (!lg_geometry(e, kp, lanes, G)) return fail(SYN_ERR_CAPACITY, "max_games_in_flight %u is smaller than one CTA of lane groups", e->max_games);
    const int gl = G.gl, threads = G.threads, gpb = G.gpb;
    const uint32_t blocks = G.blocks;
    const bool tc = G.tc;
    kp.seats_q = G.seats_q;
    kp.seats_rem = G.seats_rem;
    CUDA_TRY(cudaMemsetAsync(e->next_game.p, 0, sizeof(unsigned int), e->stream));
    if (tc && e->lg_teams && e->mlp_eff == 2) { // teams of four warps, split-fp16 chain (fp32-grade leaves)
        if (gl == 32) {
            const size_t smem = nn_team_split_smem_bytes<4, 32>();
            CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_team_split_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            selfplay_nn_team_split_kernel<32, 4><<<blocks, threads, smem, e->stream>>>(kp);
        } else {
            const size_t smem = nn_team_split_smem_bytes<4, 16>();
            CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_team_split_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            selfplay_nn_team_split_kernel<16, 4><<<blocks, threads, smem, e->stream>>>(kp);
        }
    } else if (tc && e->lg_teams) { // teams of four warps with their own barrier and MLP slot, one shared tile (selfplay_team.cuh)
#define SYN_LAUNCH_TEAM(GLv, Tv)                                                                                                        \
    do {                                                                                                                                \
        const size_t smem = nn_team_smem_bytes<Tv, 4, GLv>();                                                                           \
        CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_team_kernel<GLv, Tv, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        selfplay_nn_team_kernel<GLv, Tv, 4><<<blocks, threads, smem, e->stream>>>(kp);                                                  \
    } while (0)
        if (e->lg_teams == 6) { if (gl == 32) SYN_LAUNCH_TEAM(32, 6); else SYN_LAUNCH_TEAM(16, 6); }
        else if (e->lg_teams == 5) { if (gl == 32) SYN_LAUNCH_TEAM(32, 5); else SYN_LAUNCH_TEAM(16, 5); }
        else { if (gl == 32) SYN_LAUNCH_TEAM(32, 4); else SYN_LAUNCH_TEAM(16, 4); }
#undef SYN_LAUNCH_TEAM
    } else if (tc) {
        size_t smem = nn_tc_smem_bytes(gpb);
        if (gl == 32) {
            CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tc_kernel<32, NN_TC_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            selfplay_nn_tc_kernel<32, NN_TC_THREADS><<<blocks, threads, smem, e->stream>>>(kp);
        } else {
            CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_tc_kernel<16, NN_TC_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            selfplay_nn_tc_kernel<16, NN_TC_THREADS><<<blocks, threads, smem, e->stream>>>(kp);
        }
    } else if (nn) {
        size_t smem = nn_smem_bytes(gpb);
        if (gl == 32) {
            CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_kernel<32, NN_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            selfplay_nn_kernel<32, NN_THREADS><<<blocks, threads, smem, e->stream>>>(kp);
        } else {
            CUDA_TRY(cudaFuncSetAttribute(selfplay_nn_kernel<16, NN_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            selfplay_nn_kernel<16, NN_THREADS><<<blocks, threads, smem, e->stream>>>(kp);
        }
    } else {
        if (gl == 32) selfplay_rollout_kernel<32, ROLLOUT_THREADS><<<blocks, threads, 0, e->stream>>>(kp);
        else selfplay_rollout_kernel<16, ROLLOUT_THREADS><<<blocks, threads, 0, e->stream>>>(kp);
    }
    CUDA_TRY(cudaGetLastError());
    e->launches += 1;
    return SYN_OK;
}

static void fill_common(syn_engine* e, KParams& kp, const syn_rollout_cfg* cfg) {
    std::memset(&kp, 0, sizeof(kp));
    kp.cfg = *cfg;
    kp.arena_nodes = e->arena_nodes;
    kp.nodes = e->nodes.p;
    kp.next_game = e->next_game.p;
    kp.slot_state = e->slot_state.p;
    kp.fpu_state = e->fpu_state.p;
    kp.counters = e->counters.p;
    kp.error = e->error.p;
    kp.weights = e->weights.p;
    kp.weight_image = e->weight_image.p;
    kp.weight_image_lo = e->weight_image_lo.p;
    std::memcpy(kp.mlp_bias, e->bias_host, sizeof(kp.mlp_bias));
    const char* nored = std::getenv("SYN_TPG_NO_RED"); // read per launch so that one test process can run both forms
    kp.no_reductions = (nored && std::atoi(nored) == 1) ? 1u : 0u;
}

static int read_stats(syn_engine* e, syn_stats* stats, float ms) {
    unsigned long long c[CNT_N];
    CUDA_TRY(cudaMemcpyAsync(c, e->counters.p, sizeof(c), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->explores = c[CNT_EXPLORES]; stats->leaf_evals = c[CNT_LEAF_EVALS]; stats->rows = c[CNT_ROWS];
        stats->games = c[CNT_GAMES]; stats->trees = c[CNT_TREES]; stats->nodes = c[CNT_NODES];
        stats->select_levels = c[CNT_SELECT_LEVELS]; stats->children_scanned = c[CNT_CHILDREN_SCANNED];
        stats->expansions = c[CNT_EXPANSIONS]; stats->children_created = c[CNT_CHILDREN_CREATED];
        stats->backprop_levels = c[CNT_BACKPROP_LEVELS]; stats->rollout_plies = c[CNT_ROLLOUT_PLIES];
        stats->device_ns = (uint64_t)((double)ms * 1e6);
        stats->kernel_launches = e->launches;
        stats->h2d_bytes = e->h2d;
        stats->d2h_bytes = e->d2h;
    }
    return SYN_OK;
}

static int check_device_error(syn_engine* e) {
    int derr = 0;
    CUDA_TRY(cudaMemcpyAsync(&derr, e->error.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (derr == DERR_ARENA_OVERFLOW) return fail(SYN_ERR_CAPACITY, "a tree outgrew its arena of %u nodes", e->arena_nodes);
    if (derr == DERR_DEPTH_OVERFLOW) return fail(SYN_ERR_DEVICE_FAULT, "a tree path exceeded 64 levels");
    if (derr == DERR_BAD_WEIGHTS) return fail(SYN_ERR_DEVICE_FAULT, "WeightedIndex met an all-zero or non-finite search policy (the reference would panic)");
    if (derr == DERR_NO_BEST_ACTION) return fail(SYN_ERR_DEVICE_FAULT, "FrozenMCTS::best_action met a root whose children are all unvisited (the reference would panic)");
    if (derr) return fail(SYN_ERR_DEVICE_FAULT, "device error %d", derr);
    return SYN_OK;
}

static int validate_mcts_cfg(const syn_mcts_cfg& m) {
    if (m.exploration_kind > SYN_EXPLORATION_POLYNOMIAL_UCT) return fail(SYN_ERR_INVALID_ARGUMENT, "bad exploration_kind %u", m.exploration_kind);
    if (m.fpu_kind == SYN_FPU_FUNC)
        return fail(SYN_ERR_UNSUPPORTED, "Fpu::Func carries host code and cannot run on the device; use SYN_FPU_NORMAL{mean,std} for the shipped closure");
    if (m.fpu_kind > SYN_FPU_FUNC) return fail(SYN_ERR_INVALID_ARGUMENT, "bad fpu_kind %u", m.fpu_kind);
    if (m.noise_kind > SYN_NOISE_DIRICHLET) return fail(SYN_ERR_INVALID_ARGUMENT, "bad noise_kind %u", m.noise_kind);
    if (m.noise_kind == SYN_NOISE_DIRICHLET && !(m.noise_alpha > 0.0f)) return fail(SYN_ERR_INVALID_ARGUMENT, "Dirichlet alpha must be > 0");
    return SYN_OK;
}

// FrozenMCTS panics on anything but Fpu::Const and Exploration::Uct (evaluator.rs:410, 424)
static int validate_frozen_cfg(const syn_mcts_cfg& m) {
    if (m.fpu_kind != SYN_FPU_CONST) return fail(SYN_ERR_UNSUPPORTED, "FrozenMCTS supports Fpu::Const only (evaluator.rs:410 panics otherwise)");
    if (m.exploration_kind != SYN_EXPLORATION_UCT) return fail(SYN_ERR_UNSUPPORTED, "FrozenMCTS supports Exploration::Uct only (evaluator.rs:424 panics otherwise)");
    return SYN_OK;
}

static int validate_player(const syn_player_cfg& pl, const syn_engine* e, int k) {
    if (pl.tree_kind > SYN_TREE_FROZEN) return fail(SYN_ERR_INVALID_ARGUMENT, "players[%d]: bad tree_kind %u", k, pl.tree_kind);
    if (pl.leaf_eval_kind > SYN_LEAF_ROLLOUT) return fail(SYN_ERR_INVALID_ARGUMENT, "players[%d]: bad leaf_eval_kind %u", k, pl.leaf_eval_kind);
    if (pl.action_selection > SYN_ACTION_NUM_VISITS) return fail(SYN_ERR_INVALID_ARGUMENT, "players[%d]: bad action_selection %u", k, pl.action_selection);
    if (pl.num_explores > e->max_explores)
        return fail(SYN_ERR_CAPACITY, "players[%d]: num_explores %u exceeds the engine's max_explores %u", k, pl.num_explores, e->max_explores);
    if (pl.leaf_eval_kind == SYN_LEAF_NN && !e->has_weights)
        return fail(SYN_ERR_NO_WEIGHTS, "players[%d]: leaf_eval_kind = NN but syn_engine_set_weights has not been called", k);
    int rc = validate_mcts_cfg(pl.mcts);
    if (rc) return rc;
    if (pl.tree_kind == SYN_TREE_FROZEN) return validate_frozen_cfg(pl.mcts);
    return SYN_OK;
}

// Launches the thread-per-match kernel (match.cuh) over kp.num_games matches or search roots.
constexpr int MATCH_TEAMS = 4;
static int launch_match(syn_engine* e, KParams& kp, mtc::MParams& mp) {
    const uint32_t per_cta = 128u * MATCH_TEAMS;
    uint32_t blocks = kp.num_games < (uint32_t)e->sm_count ? kp.num_games : (uint32_t)e->sm_count;
    if (blocks == 0) blocks = 1;
    uint32_t active = (kp.num_games + blocks - 1) / blocks;
    if (active > per_cta) active = per_cta;
    if ((uint64_t)blocks * active > e->max_games) active = e->max_games / blocks;
    if (active == 0) { blocks = e->max_games; active = 1; }
    if (mp.weight_image2) { // two resident weight images leave room for two teams' activation tiles
        constexpr int T2 = 2;
        const uint32_t per_cta2 = 128u * T2;
        if (active > per_cta2) active = per_cta2;
        mp.active_per_block = active;
        size_t smem2 = sizeof(mlpteam::Smem<T2, T2>) + mlptc::IMG_BYTES;
        CUDA_TRY(cudaFuncSetAttribute(match_tpg_kernel<T2, T2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        match_tpg_kernel<T2, T2, true><<<blocks, per_cta2, smem2, e->stream>>>(kp, mp);
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
        return SYN_OK;
    }
    mp.active_per_block = active;
    if (e->mlp_eff == 2 && (mp.players[0].leaf_eval_kind == SYN_LEAF_NN || mp.players[1].leaf_eval_kind == SYN_LEAF_NN)) { // one network, fp32-grade forward
        const size_t smem_s = sizeof(mlps::Smem<MATCH_TEAMS, 2>);
        CUDA_TRY(cudaFuncSetAttribute(match_tpg_split_kernel<MATCH_TEAMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        match_tpg_split_kernel<MATCH_TEAMS><<<blocks, per_cta, smem_s, e->stream>>>(kp, mp);
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
        return SYN_OK;
    }
    // no Connect4Net player: the kernel never touches the MLP state, and the shared memory it would take comes out of the L1
    const bool any_nn = mp.players[0].leaf_eval_kind == SYN_LEAF_NN || mp.players[1].leaf_eval_kind == SYN_LEAF_NN;
    size_t smem = any_nn ? sizeof(mlpteam::Smem<MATCH_TEAMS, MATCH_TEAMS>) : 0;
    CUDA_TRY(cudaFuncSetAttribute(match_tpg_kernel<MATCH_TEAMS, MATCH_TEAMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_tpg_kernel<MATCH_TEAMS, MATCH_TEAMS, false><<<blocks, per_cta, smem, e->stream>>>(kp, mp);
    CUDA_TRY(cudaGetLastError());
    e->launches += 1;
    return SYN_OK;
}

// ------------------------------------------------------------------ multi-GPU: one process (or thread) per GPU
// NCCL is bound at run time by its soname: inside a process that already carries an NCCL (PyTorch's bundled copy) the same
// library instance is used — two copies of NCCL in one process do not share their bootstrap state — and a plain C / Rust
// host gets the system's libnccl.so.2.  Nothing else in this library depends on NCCL being present.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = {std::getenv("SYN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return nullptr;
    bool ok = true;
    auto bind = [&](auto& fn, const char* name) {
        fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.handle, name));
        ok = ok && fn != nullptr;
    };
    bind(api.GetUniqueId, "ncclGetUniqueId"); bind(api.CommInitRank, "ncclCommInitRank"); bind(api.CommDestroy, "ncclCommDestroy");
    bind(api.Broadcast, "ncclBroadcast"); bind(api.AllGather, "ncclAllGather"); bind(api.Send, "ncclSend"); bind(api.Recv, "ncclRecv");
    bind(api.GroupStart, "ncclGroupStart"); bind(api.GroupEnd, "ncclGroupEnd"); bind(api.GetErrorString, "ncclGetErrorString");
    if (!ok) { dlclose(api.handle); api.handle = nullptr; return nullptr; }
    return &api;
}
#define NCCL_TRY(api, expr)                                                                                                   \
    do {                                                                                                                      \
        ncclResult_t _r = (expr);                                                                                             \
        if (_r != ncclSuccess) return fail(SYN_ERR_COMM, "%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

struct syn_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1, device = 0;
};

static_assert(SYN_COMM_ID_BYTES == sizeof(ncclUniqueId), "syn_comm id is an ncclUniqueId");

// Wire format of one experience row between GPUs: game id, the two bitboards, pi[9], v[3] = 72 bytes in five arrays;
// height, player and the 63 features are functions of the bitboards and are rebuilt on the root (expand_rows_kernel).
static const size_t GX_ELT[5] = {8, 8, 8, 36, 12};
static size_t gx_field_off(int f, size_t rows) { // 256-byte aligned arrays of `rows` elements each
    size_t off = 0;
    for (int i = 0; i < f; ++i) off += (GX_ELT[i] * rows + 255) / 256 * 256;
    return off;
}

static int launch_eval(syn_engine* e, int chain, const uint64_t* my, const uint64_t* op, uint32_t n, float* logits, float* probs) {
    if (chain == 2) {
        const size_t smem = sizeof(mlps::Smem<1, 1>);
        CUDA_TRY(cudaFuncSetAttribute(eval_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint32_t blocks = (n + 127) / 128;
        if (blocks > (uint32_t)e->sm_count) blocks = (uint32_t)e->sm_count;
        mlps::Bias bias;
        std::memcpy(bias.b, e->bias_host, sizeof(bias.b));
        eval_split_kernel<<<blocks, 128, smem, e->stream>>>(e->weight_image.p, e->weight_image_lo.p, bias, my, op, n, logits, probs);
    } else if (chain == 1) {
        size_t smem = sizeof(mlptc::Smem);
        CUDA_TRY(cudaFuncSetAttribute(eval_tc_kernel<NN_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint32_t blocks = (n + 127) / 128;
        if (blocks > (uint32_t)e->sm_count) blocks = (uint32_t)e->sm_count;
        eval_tc_kernel<NN_THREADS><<<blocks, NN_THREADS, smem, e->stream>>>(e->weight_image.p, my, op, n, logits, probs);
    } else {
        size_t smem = (size_t)(mlp::WEIGHT_FLOATS + 2 * 32 * mlp::XS) * sizeof(float);
        CUDA_TRY(cudaFuncSetAttribute(eval_kernel<NN_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint32_t blocks = (n + 31) / 32;
        if (blocks > (uint32_t)e->sm_count) blocks = (uint32_t)e->sm_count;
        eval_kernel<NN_THREADS><<<blocks, NN_THREADS, smem, e->stream>>>(e->weights.p, my, op, n, logits, probs);
    }
    CUDA_TRY(cudaGetLastError());
    return SYN_OK;
}

// Which Connect4Net chain the kernels use.  Requested modes 0 / 1 / 2 are taken as they are.  In auto mode (the default) the
// engine MEASURES, every time the weights change, the single-fp16 chain against the fp32-grade split chain on CALIB_N reachable
// positions: the fast chain is used only while its largest error stays below a quarter of BASELINE.json's tolerance
// (1e-3 abs + 1e-3 rel, logits and outcome probabilities).  Random-init weights pass with a margin of 25 (4e-5); trained-size
// weights do not (7x the tolerance) and get the split chain.  Needs the weights and both images on the device and bias_host filled.
static int calibrate_mlp(syn_engine* e) {
    e->calib_ratio = -1.0f;
    if (e->mlp_mode != 3 || !e->has_weights) { e->mlp_eff = e->mlp_mode == 3 ? 2 : e->mlp_mode; return SYN_OK; }
    if (!e->calib_pos.p) {
        std::vector<uint64_t> my, op;
        calibration_positions(my, op);
        CUDA_TRY(e->calib_pos.reserve(2 * (size_t)CALIB_N));
        CUDA_TRY(e->calib_out.reserve(2 * (size_t)CALIB_N * 12));
        CUDA_TRY(cudaMemcpyAsync(e->calib_pos.p, my.data(), CALIB_N * 8, cudaMemcpyHostToDevice, e->stream));
        CUDA_TRY(cudaMemcpyAsync(e->calib_pos.p + CALIB_N, op.data(), CALIB_N * 8, cudaMemcpyHostToDevice, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    float* o1 = e->calib_out.p;
    float* o2 = e->calib_out.p + (size_t)CALIB_N * 12;
    int rc;
    if ((rc = launch_eval(e, 1, e->calib_pos.p, e->calib_pos.p + CALIB_N, CALIB_N, o1, o1 + (size_t)CALIB_N * 9))) return rc;
    if ((rc = launch_eval(e, 2, e->calib_pos.p, e->calib_pos.p + CALIB_N, CALIB_N, o2, o2 + (size_t)CALIB_N * 9))) return rc;
    std::vector<float> h(2 * (size_t)CALIB_N * 12);
    CUDA_TRY(cudaMemcpyAsync(h.data(), e->calib_out.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    float worst = 0.0f;
    const size_t n = (size_t)CALIB_N * 12;
    for (size_t i = 0; i < n; ++i) {
        const float a = h[i], b = h[n + i];
        const float r = std::fabs(a - b) / (1e-3f + 1e-3f * std::fabs(b));
        if (!(r <= worst)) worst = r; // a NaN ratio (a chain that overflowed fp16) also lands here
    }
    e->calib_ratio = worst;
    e->mlp_eff = (worst <= CALIB_MAX_RATIO) ? 1 : 2;
    return SYN_OK;
}

extern "C" {

int syn_abi_version(void) { return SYN_ABI_VERSION; }
const char* syn_build_info(void) {
#define SYN_STR2(x) #x
#define SYN_STR(x) SYN_STR2(x)
    return "libsynthesis_b200: sm_100a, nvcc " SYN_STR(__CUDACC_VER_MAJOR__) "." SYN_STR(__CUDACC_VER_MINOR__) ", -fmad=false, built " __DATE__;
}
const char* syn_last_error(void) { return g_err; }

int syn_engine_create(int cuda_device, uint32_t max_games_in_flight, uint32_t max_explores, syn_engine** out) {
    if (!out) return fail(SYN_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SYN_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU path");
    }
    if (cuda_device < 0 || cuda_device >= ndev) return fail(SYN_ERR_INVALID_ARGUMENT, "cuda_device %d out of range (%d devices)", cuda_device, ndev);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cuda_device));
    if (prop.major != 10)
        return fail(SYN_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", cuda_device, prop.major, prop.minor);
    if (max_games_in_flight == 0 || max_explores == 0) return fail(SYN_ERR_INVALID_ARGUMENT, "max_games_in_flight and max_explores must be > 0");
    CUDA_TRY(cudaSetDevice(cuda_device));
    // node records are 16/32-byte random accesses: ask L2 to fetch single sectors from DRAM instead of pairs
    const char* fenv = std::getenv("SYN_L2_FETCH");
    if (fenv && std::atoi(fenv) > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(fenv));
    syn_engine* e = new syn_engine();
    e->device = cuda_device;
    e->sm_count = prop.multiProcessorCount;
    const char* mlpenv = std::getenv("SYN_MLP");
    e->use_tc = !(mlpenv && std::strcmp(mlpenv, "fp32") == 0);
    e->mlp_mode = !e->use_tc ? 0 : ((mlpenv && std::strcmp(mlpenv, "fp16") == 0) ? 1 : ((mlpenv && std::strcmp(mlpenv, "split") == 0) ? 2 : 3));
    e->mlp_eff = e->mlp_mode == 3 ? 2 : e->mlp_mode;
    const char* glenv = std::getenv("SYN_GROUP_LANES");
    e->group_lanes = (glenv && std::atoi(glenv) == 16) ? 16 : ((glenv && std::atoi(glenv) == 32) ? 32 : ((glenv && std::atoi(glenv) == 1) ? 1 : 0));
    const char* tenv = std::getenv("SYN_TPG_TEAMS");
    const char* penv = std::getenv("SYN_TPG_PROF");
    e->tpg_prof = penv && std::atoi(penv) == 1;
    if (const char* lt = std::getenv("SYN_LG_TEAMS")) { const int v = std::atoi(lt); if (v == 0 || v == 4 || v == 5 || v == 6) e->lg_teams = v; }
    if (tenv && (std::atoi(tenv) == 1 || std::atoi(tenv) == 2 || std::atoi(tenv) == 4 || std::atoi(tenv) == 5 || std::atoi(tenv) == 6 || std::atoi(tenv) == 8)) e->tpg_teams = std::atoi(tenv);
    const char* venv = std::getenv("SYN_TPG_VER");
    if (venv && std::atoi(venv) == 4) e->tpg_ver = 4;
    const char* renv = std::getenv("SYN_ROLLOUT_THREADS");
    if (renv && (std::atoi(renv) == 512 || std::atoi(renv) == 640 || std::atoi(renv) == 768 || std::atoi(renv) == 896 || std::atoi(renv) == 1024)) e->rollout_threads = std::atoi(renv);
    const char* cwenv = std::getenv("SYN_ROLLOUT_CW");
    if (cwenv && std::atoi(cwenv) == 5) e->rollout_cw = 5;
    // round the in-flight game count up to whole CTAs of every kernel, plus the slack of the even seating (tp2::seat_of)
    // (teams_used * per_team slots per CTA: at most 8 more than the CTA's share of the games)
    uint32_t unit = 1024;
    e->req_games = max_games_in_flight;
    e->max_games = (uint32_t)((((uint64_t)max_games_in_flight + 8ull * (uint64_t)prop.multiProcessorCount + unit - 1) / unit) * unit);
    e->max_explores = max_explores;
    // nodes.len() <= 1 + 9 * (explores + 1): every visit pushes at most 9 nodes (mcts.rs:384-397)
    e->arena_nodes = 1u + 9u * (max_explores + 1u) + 7u;
    // tpg4 cuts the arena into 128-byte lines: an expansion takes three (two or more children) or one (an only child: every
    // step of an auto-extended chain), so 768 bytes per explore leave room for a chain of three behind every expansion
    if (e->tpg_ver == 4 && e->arena_nodes < 24u * (max_explores + 1u) + 256u) e->arena_nodes = 24u * (max_explores + 1u) + 256u;
    e->arena_nodes = (e->arena_nodes + 7u) & ~7u;
    cudaError_t ce;
    size_t total = (size_t)e->max_games * e->arena_nodes;
    if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaEventCreate(&e->ev0)) != cudaSuccess || (ce = cudaEventCreate(&e->ev1)) != cudaSuccess ||
        (ce = e->nodes.reserve(2 * total)) != cudaSuccess || (ce = e->slot_state.reserve((size_t)e->max_games * tp2::SS_WORDS)) != cudaSuccess || (ce = e->fpu_state.reserve((size_t)e->max_games * tp2::FS_WORDS)) != cudaSuccess ||
        (ce = e->weights.reserve(SYN_N_WEIGHTS)) != cudaSuccess || (ce = e->weight_image.reserve(mlptc::IMG_BYTES)) != cudaSuccess || (ce = e->weight_image_lo.reserve(mlptc::W_TOTAL)) != cudaSuccess || (ce = e->next_game.reserve(1)) != cudaSuccess ||
        (ce = e->counters.reserve(CNT_ALL)) != cudaSuccess || (ce = e->error.reserve(1)) != cudaSuccess) {
        int rc = fail(SYN_ERR_CUDA, "engine allocation failed (%zu arena nodes = %.1f MiB): %s", total,
                      (double)total * 32.0 / 1048576.0, cudaGetErrorString(ce));
        syn_engine_destroy(e);
        return rc;
    }
    *out = e;
    return SYN_OK;
}

void syn_engine_destroy(syn_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->nodes.release(); e->slot_state.release(); e->fpu_state.release(); e->weights.release(); e->weight_image.release(); e->weight_image_lo.release(); e->calib_pos.release(); e->calib_out.release(); e->weights2.release(); e->weight_image2.release(); e->next_game.release(); e->counters.release(); e->error.release();
    e->row_my.release(); e->row_op.release(); e->row_off.release(); e->row_pi.release(); e->row_v.release(); e->row_visits.release();
    e->row_action.release(); e->row_nodes.release(); e->game_len.release(); e->staging.release();
    e->pos_my.release(); e->pos_op.release(); e->pos_seed.release(); e->s_visits.release(); e->s_q.release();
    e->s_csol.release(); e->s_rsol.release(); e->s_best.release(); e->s_nodes.release();
    e->gx_local.release(); e->gx_all.release(); e->gx_counts.release();
    e->dd_ws.release(); e->dd_io.release(); e->adam_m.release(); e->adam_v.release(); e->tr_io.release();
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int syn_engine_set_group_lanes(syn_engine* e, int lanes) {
    if (!e || (lanes != 0 && lanes != 1 && lanes != 16 && lanes != 32)) return fail(SYN_ERR_INVALID_ARGUMENT, "group lanes must be 0 (chosen per launch), 1, 16 or 32");
    e->group_lanes = lanes;
    return SYN_OK;
}

int syn_engine_debug_counters(syn_engine* e, uint64_t* out, uint32_t n) {
    if (!e || !out) return fail(SYN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (e->pending) return fail(SYN_ERR_INVALID_ARGUMENT, "a gather is in flight");
    unsigned long long c[CNT_ALL];
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaMemcpy(c, e->counters.p, sizeof(c), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; ++i) out[i] = i < (uint32_t)(CNT_ALL - CNT_N) ? c[CNT_N + i] : 0;
    return SYN_OK;
}

int syn_engine_mlp_in_use(syn_engine* e, int* chain, float* calibration_ratio) {
    if (!e) return fail(SYN_ERR_INVALID_ARGUMENT, "engine is NULL");
    if (chain) *chain = e->mlp_eff;
    if (calibration_ratio) *calibration_ratio = e->calib_ratio;
    return SYN_OK;
}

int syn_engine_launch_geometry(syn_engine* e, uint32_t num_games, uint32_t leaf_eval_kind, uint32_t* ctas, uint32_t* games_per_cta, uint32_t* lanes_per_game) {
    if (!e) return fail(SYN_ERR_INVALID_ARGUMENT, "engine is NULL");
    KParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.num_games = num_games;
    kp.cfg.leaf_eval_kind = leaf_eval_kind;
    const bool nn = leaf_eval_kind == SYN_LEAF_NN;
    const int lanes = pick_group_lanes(e, kp);
    if (lanes_per_game) *lanes_per_game = (uint32_t)lanes;
    if (lanes == 1) {
        const uint32_t blocks = nn ? seat_games(e, kp, 128u, (uint32_t)e->tpg_teams) : seat_games(e, kp, (uint32_t)e->rollout_threads, 1u);
        if (ctas) *ctas = blocks;
        if (games_per_cta) *games_per_cta = kp.seats_q + (kp.seats_rem ? 1u : 0u);
        return SYN_OK;
    }
    LgGeometry G;
    if (!lg_geometry(e, kp, lanes, G)) return fail(SYN_ERR_CAPACITY, "max_games_in_flight %u is smaller than one CTA of lane groups", e->max_games);
    if (ctas) *ctas = G.blocks;
    if (games_per_cta) *games_per_cta = G.seats_q + (G.seats_rem ? 1u : 0u);
    return SYN_OK;
}

int syn_engine_set_mlp_mode(syn_engine* e, int tensor_cores) {
    if (!e) return fail(SYN_ERR_INVALID_ARGUMENT, "engine is NULL");
    if (tensor_cores < 0 || tensor_cores > 3) return fail(SYN_ERR_INVALID_ARGUMENT, "mlp mode must be 0 (fp32 CUDA cores), 1 (fp16 operands), 2 (split-fp16 operands) or 3 (auto)");
    if (e->pending) return fail(SYN_ERR_INVALID_ARGUMENT, "a gather is in flight");
    e->use_tc = tensor_cores != 0;
    e->mlp_mode = tensor_cores;
    CUDA_TRY(cudaSetDevice(e->device));
    return calibrate_mlp(e);
}

int syn_engine_set_weights(syn_engine* e, const float* blob, size_t n_floats) {
    if (!e || !blob) return fail(SYN_ERR_INVALID_ARGUMENT, "engine or blob is NULL");
    if (n_floats != SYN_N_WEIGHTS) return fail(SYN_ERR_INVALID_ARGUMENT, "expected %d floats (63-128-96-64-48-12 MLP), got %zu", SYN_N_WEIGHTS, n_floats);
    CUDA_TRY(cudaSetDevice(e->device));
    bool dev = is_device_ptr(blob);
    CUDA_TRY(cudaMemcpyAsync(e->weights.p, blob, n_floats * sizeof(float), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
    mlptc::build_weight_image<<<32, 256, 0, e->stream>>>(e->weights.p, e->weight_image.p);
    mlps::build_weight_image_lo<<<32, 256, 0, e->stream>>>(e->weights.p, e->weight_image_lo.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(e->bias_host, e->weight_image.p + mlptc::BIAS_OFF, sizeof(e->bias_host), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (!dev) e->h2d += n_floats * sizeof(float);
    e->has_weights = true;
    return calibrate_mlp(e);
}

int syn_engine_set_opponent_weights(syn_engine* e, const float* blob, size_t n_floats) {
    if (!e) return fail(SYN_ERR_INVALID_ARGUMENT, "engine is NULL");
    if (!blob) { e->has_weights2 = false; return SYN_OK; } // back to one network for both players
    if (n_floats != SYN_N_WEIGHTS) return fail(SYN_ERR_INVALID_ARGUMENT, "expected %d floats (63-128-96-64-48-12 MLP), got %zu", SYN_N_WEIGHTS, n_floats);
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(e->weights2.reserve(SYN_N_WEIGHTS));
    CUDA_TRY(e->weight_image2.reserve(mlptc::IMG_BYTES));
    bool dev = is_device_ptr(blob);
    CUDA_TRY(cudaMemcpyAsync(e->weights2.p, blob, n_floats * sizeof(float), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
    mlptc::build_weight_image<<<32, 256, 0, e->stream>>>(e->weights2.p, e->weight_image2.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (!dev) e->h2d += n_floats * sizeof(float);
    e->has_weights2 = true;
    return SYN_OK;
}

int syn_engine_gather_launch(syn_engine* e, const syn_rollout_cfg* cfg, uint64_t first_game_index, uint32_t num_games, uint64_t seed) {
    if (!e) return fail(SYN_ERR_INVALID_ARGUMENT, "engine is NULL");
    if (e->pending) return fail(SYN_ERR_INVALID_ARGUMENT, "a gather is already in flight; call syn_engine_gather_wait first");
    int rc = validate_cfg(cfg, e);
    if (rc) return rc;
    if (num_games == 0) return fail(SYN_ERR_INVALID_ARGUMENT, "num_games must be > 0");
    // the domain on which (seed, game, stream) -> ChaCha12 seed is injective (include/syn_streams.h)
    if (seed > SYN_MAX_SEED) return fail(SYN_ERR_INVALID_ARGUMENT, "seed %llu exceeds 2^30 - 1 (syn_streams.h)", (unsigned long long)seed);
    if (first_game_index + num_games - 1ull > SYN_MAX_GAME_INDEX)
        return fail(SYN_ERR_INVALID_ARGUMENT, "game indices must stay below 2^32 (first %llu + %u games)", (unsigned long long)first_game_index, num_games);
    CUDA_TRY(cudaSetDevice(e->device));
    size_t rows = (size_t)num_games * 63;
    CUDA_TRY(e->row_my.reserve(rows)); CUDA_TRY(e->row_op.reserve(rows)); CUDA_TRY(e->row_pi.reserve(rows * 9));
    CUDA_TRY(e->row_v.reserve(rows * 3)); CUDA_TRY(e->row_visits.reserve(rows * 9)); CUDA_TRY(e->row_action.reserve(rows));
    CUDA_TRY(e->row_nodes.reserve(rows)); CUDA_TRY(e->game_len.reserve(num_games)); CUDA_TRY(e->row_off.reserve(num_games));
    KParams kp;
    fill_common(e, kp, cfg);
    kp.seed = seed; kp.first_game = first_game_index; kp.num_games = num_games; kp.search_mode = 0;
    kp.row_my = e->row_my.p; kp.row_op = e->row_op.p; kp.row_pi = e->row_pi.p; kp.row_v = e->row_v.p;
    kp.row_action = e->row_action.p; kp.row_nodes = e->row_nodes.p; kp.row_visits = e->row_visits.p; kp.game_len = e->game_len.p;
    e->launches = 0;
    CUDA_TRY(cudaMemsetAsync(e->counters.p, 0, CNT_ALL * sizeof(unsigned long long), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->error.p, 0, sizeof(int), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->game_len.p, 0, num_games * sizeof(uint32_t), e->stream));
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    rc = launch_selfplay(e, kp);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    e->pending = true;
    e->pend_games = num_games;
    e->pend_first = first_game_index;
    return SYN_OK;
}

// copies a dense device array to a caller pointer (host or device)
static int deliver(syn_engine* e, void* dst, const void* src_dev, size_t bytes) {
    if (!dst || bytes == 0) return SYN_OK;
    bool dev = is_device_ptr(dst);
    CUDA_TRY(cudaMemcpyAsync(dst, src_dev, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
    if (!dev) e->d2h += bytes;
    return SYN_OK;
}

int syn_engine_set_trace(syn_engine* e, uint8_t* action, uint32_t* tree_nodes, float* child_visits) {
    if (!e) return fail(SYN_ERR_INVALID_ARGUMENT, "engine is NULL");
    e->trace_action = action;
    e->trace_nodes = tree_nodes;
    e->trace_visits = child_visits;
    return SYN_OK;
}

int syn_engine_gather_wait(syn_engine* e, syn_experience* out, syn_stats* stats) {
    if (!e) return fail(SYN_ERR_INVALID_ARGUMENT, "engine is NULL");
    if (!e->pending) return fail(SYN_ERR_INVALID_ARGUMENT, "no gather in flight");
    CUDA_TRY(cudaSetDevice(e->device));
    e->pending = false;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    int rc = check_device_error(e);
    if (rc) return rc;
    const uint32_t n = e->pend_games;
    if (out) {
        std::vector<uint32_t> len(n);
        CUDA_TRY(cudaMemcpyAsync(len.data(), e->game_len.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        e->d2h += n * sizeof(uint32_t);
        std::vector<uint64_t> off(n);
        uint64_t total = 0;
        for (uint32_t i = 0; i < n; ++i) { off[i] = total; total += len[i]; }
        out->len = (size_t)total;
        out->games = n;
        if (total > out->capacity) return fail(SYN_ERR_CAPACITY, "experience needs %llu rows, caller provided %zu", (unsigned long long)total, out->capacity);
        CUDA_TRY(cudaMemcpyAsync(e->row_off.p, off.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
        e->h2d += n * sizeof(uint64_t);
        // dense staging layout (only for host destinations; device destinations are written in place)
        struct Field { void* dst; size_t elt; size_t off; };
        Field f[11] = {{out->game_ids, 8, 0}, {out->my_bb, 8, 0}, {out->op_bb, 8, 0}, {out->height, 9, 0}, {out->player, 1, 0},
                       {out->states, 63 * 4, 0}, {out->pis, 9 * 4, 0}, {out->vs, 3 * 4, 0},
                       {e->trace_action, 1, 0}, {e->trace_nodes, 4, 0}, {e->trace_visits, 9 * 4, 0}};
        size_t need = 0;
        for (auto& x : f) {
            if (x.dst && !is_device_ptr(x.dst)) { x.off = need; need += ((x.elt * total + 255) / 256) * 256; }
        }
        CUDA_TRY(e->staging.reserve(need ? need + need / 4 : 256)); // headroom: the next seed's gather has a few percent more rows
        auto target = [&](int i) -> void* {
            if (!f[i].dst) return nullptr;
            return is_device_ptr(f[i].dst) ? f[i].dst : (void*)(e->staging.p + f[i].off);
        };
        CompactParams c;
        std::memset(&c, 0, sizeof(c));
        c.num_games = n; c.first_game = e->pend_first; c.game_len = e->game_len.p; c.row_off = e->row_off.p;
        c.row_my = e->row_my.p; c.row_op = e->row_op.p; c.row_pi = e->row_pi.p; c.row_v = e->row_v.p;
        c.row_action = e->row_action.p; c.row_nodes = e->row_nodes.p; c.row_visits = e->row_visits.p;
        c.game_ids = (uint64_t*)target(0); c.my_bb = (uint64_t*)target(1); c.op_bb = (uint64_t*)target(2);
        c.height = (uint8_t*)target(3); c.player = (uint8_t*)target(4); c.states = (float*)target(5);
        c.pis = (float*)target(6); c.vs = (float*)target(7); c.t_action = (uint8_t*)target(8);
        c.t_nodes = (uint32_t*)target(9); c.t_visits = (float*)target(10);
        uint64_t warps = (uint64_t)n * 63;
        uint32_t blocks = (uint32_t)((warps * 32 + 255) / 256);
        compact_kernel<<<blocks, 256, 0, e->stream>>>(c);
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
        for (int i = 0; i < 11; ++i)
            if (f[i].dst && !is_device_ptr(f[i].dst)) {
                rc = deliver(e, f[i].dst, e->staging.p + f[i].off, f[i].elt * total);
                if (rc) return rc;
            }
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return read_stats(e, stats, ms);
}

int syn_engine_gather(syn_engine* e, const syn_rollout_cfg* cfg, uint64_t first_game_index, uint32_t num_games, uint64_t seed,
                      syn_experience* out, syn_stats* stats) {
    if (!out) return fail(SYN_ERR_INVALID_ARGUMENT, "out is NULL");
    if (e) { e->h2d = 0; e->d2h = 0; }
    int rc = syn_engine_gather_launch(e, cfg, first_game_index, num_games, seed);
    if (rc) return rc;
    return syn_engine_gather_wait(e, out, stats);
}


int syn_comm_unique_id(uint8_t id[SYN_COMM_ID_BYTES]) {
    if (!id) return fail(SYN_ERR_INVALID_ARGUMENT, "id is NULL");
    NcclApi* api = nccl_api();
    if (!api) return fail(SYN_ERR_COMM, "libnccl.so.2 could not be loaded: %s", dlerror());
    ncclUniqueId u;
    NCCL_TRY(api, api->GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return SYN_OK;
}

int syn_comm_create(const uint8_t id[SYN_COMM_ID_BYTES], int n_ranks, int rank, int cuda_device, syn_comm** out) {
    if (!out) return fail(SYN_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(SYN_ERR_INVALID_ARGUMENT, "bad communicator arguments (n_ranks %d, rank %d)", n_ranks, rank);
    NcclApi* api = nccl_api();
    if (!api) return fail(SYN_ERR_COMM, "libnccl.so.2 could not be loaded: %s", dlerror());
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SYN_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU path");
    }
    if (cuda_device < 0 || cuda_device >= ndev) return fail(SYN_ERR_INVALID_ARGUMENT, "cuda_device %d out of range (%d devices)", cuda_device, ndev);
    CUDA_TRY(cudaSetDevice(cuda_device));
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    syn_comm* c = new syn_comm();
    c->rank = rank; c->size = n_ranks; c->device = cuda_device;
    ncclResult_t r = api->CommInitRank(&c->comm, n_ranks, u, rank);
    if (r != ncclSuccess) {
        int rc = fail(SYN_ERR_COMM, "ncclCommInitRank failed: %s", api->GetErrorString(r));
        delete c;
        return rc;
    }
    *out = c;
    return SYN_OK;
}

void syn_comm_destroy(syn_comm* c) {
    if (!c) return;
    NcclApi* api = nccl_api();
    if (api && c->comm) { cudaSetDevice(c->device); api->CommDestroy(c->comm); }
    delete c;
}

int syn_comm_rank(const syn_comm* c) { return c ? c->rank : -1; }
int syn_comm_size(const syn_comm* c) { return c ? c->size : 0; }

int syn_engine_broadcast_weights(syn_engine* e, syn_comm* c, const float* blob, size_t n_floats, int root) {
    if (!e || !c) return fail(SYN_ERR_INVALID_ARGUMENT, "engine or communicator is NULL");
    if (root < 0 || root >= c->size) return fail(SYN_ERR_INVALID_ARGUMENT, "root %d out of range (%d ranks)", root, c->size);
    if (c->device != e->device) return fail(SYN_ERR_INVALID_ARGUMENT, "communicator is on device %d, engine on %d", c->device, e->device);
    if (n_floats != SYN_N_WEIGHTS) return fail(SYN_ERR_INVALID_ARGUMENT, "expected %d floats (63-128-96-64-48-12 MLP), got %zu", SYN_N_WEIGHTS, n_floats);
    NcclApi* api = nccl_api();
    if (!api) return fail(SYN_ERR_COMM, "libnccl.so.2 could not be loaded");
    CUDA_TRY(cudaSetDevice(e->device));
    if (c->rank == root) {
        if (blob) { // NULL on the root: broadcast the engine's current weights (e.g. just trained by syn_engine_train)
            const bool dev = is_device_ptr(blob);
            CUDA_TRY(cudaMemcpyAsync(e->weights.p, blob, n_floats * sizeof(float), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
            if (!dev) e->h2d += n_floats * sizeof(float);
        } else if (!e->has_weights) {
            return fail(SYN_ERR_NO_WEIGHTS, "the root has neither a blob nor weights of its own to broadcast");
        }
    }
    NCCL_TRY(api, api->Broadcast(e->weights.p, e->weights.p, n_floats, ncclFloat32, root, c->comm, e->stream));
    mlptc::build_weight_image<<<32, 256, 0, e->stream>>>(e->weights.p, e->weight_image.p);
    mlps::build_weight_image_lo<<<32, 256, 0, e->stream>>>(e->weights.p, e->weight_image_lo.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(e->bias_host, e->weight_image.p + mlptc::BIAS_OFF, sizeof(e->bias_host), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->has_weights = true;
    return calibrate_mlp(e); // every rank measures the same weights on the same positions, so every rank picks the same chain
}

int syn_engine_gather_experience(syn_engine* e, syn_comm* c, int root, const syn_rollout_cfg* cfg, uint64_t first_game_index, uint32_t num_games,
                                 uint64_t seed, syn_experience* out, syn_stats* stats) {
    if (!e || !c) return fail(SYN_ERR_INVALID_ARGUMENT, "engine or communicator is NULL");
    if (root < 0 || root >= c->size) return fail(SYN_ERR_INVALID_ARGUMENT, "root %d out of range (%d ranks)", root, c->size);
    if (c->device != e->device) return fail(SYN_ERR_INVALID_ARGUMENT, "communicator is on device %d, engine on %d", c->device, e->device);
    if (c->rank == root && !out) return fail(SYN_ERR_INVALID_ARGUMENT, "out is NULL on the root");
    NcclApi* api = nccl_api();
    if (!api) return fail(SYN_ERR_COMM, "libnccl.so.2 could not be loaded");
    e->h2d = 0; e->d2h = 0;
    // ---- this rank's shard: search, then compact the rows into wire format on the device
    uint64_t my_rows = 0;
    float ms = 0.0f;
    int rc = SYN_OK;
    if (num_games > 0) {
        if ((rc = syn_engine_gather_launch(e, cfg, first_game_index, num_games, seed))) return rc;
        e->pending = false;
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        if ((rc = check_device_error(e))) return rc;
        std::vector<uint32_t> len(num_games);
        CUDA_TRY(cudaMemcpyAsync(len.data(), e->game_len.p, (size_t)num_games * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        e->d2h += (uint64_t)num_games * sizeof(uint32_t);
        std::vector<uint64_t> off(num_games);
        for (uint32_t i = 0; i < num_games; ++i) { off[i] = my_rows; my_rows += len[i]; }
        CUDA_TRY(cudaMemcpyAsync(e->row_off.p, off.data(), (size_t)num_games * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
        e->h2d += (uint64_t)num_games * sizeof(uint64_t);
        CUDA_TRY(e->gx_local.reserve(gx_field_off(5, my_rows) + 256));
        CompactParams cp;
        std::memset(&cp, 0, sizeof(cp));
        cp.num_games = num_games; cp.first_game = first_game_index; cp.game_len = e->game_len.p; cp.row_off = e->row_off.p;
        cp.row_my = e->row_my.p; cp.row_op = e->row_op.p; cp.row_pi = e->row_pi.p; cp.row_v = e->row_v.p;
        cp.row_action = e->row_action.p; cp.row_nodes = e->row_nodes.p; cp.row_visits = e->row_visits.p;
        uint8_t* b = e->gx_local.p;
        cp.game_ids = (uint64_t*)(b + gx_field_off(0, my_rows)); cp.my_bb = (uint64_t*)(b + gx_field_off(1, my_rows));
        cp.op_bb = (uint64_t*)(b + gx_field_off(2, my_rows)); cp.pis = (float*)(b + gx_field_off(3, my_rows)); cp.vs = (float*)(b + gx_field_off(4, my_rows));
        const uint64_t threads = (uint64_t)num_games * 63 * 32;
        compact_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, e->stream>>>(cp);
        CUDA_TRY(cudaGetLastError());
        e->launches += 1;
    } else {
        CUDA_TRY(cudaSetDevice(e->device));
        e->launches = 0;
        CUDA_TRY(cudaMemsetAsync(e->counters.p, 0, CNT_ALL * sizeof(unsigned long long), e->stream));
    }
    // ---- how many rows every rank holds (ONE small all-gather), then ONE grouped send/recv of the five arrays to the root
    const int n = c->size;
    CUDA_TRY(e->gx_counts.reserve((size_t)n + 1));
    unsigned long long mine = my_rows;
    CUDA_TRY(cudaMemcpyAsync(e->gx_counts.p + n, &mine, sizeof(mine), cudaMemcpyHostToDevice, e->stream));
    NCCL_TRY(api, api->AllGather(e->gx_counts.p + n, e->gx_counts.p, 1, ncclUint64, c->comm, e->stream));
    std::vector<unsigned long long> counts(n);
    CUDA_TRY(cudaMemcpyAsync(counts.data(), e->gx_counts.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    uint64_t total = 0;
    std::vector<uint64_t> first_row(n);
    for (int r = 0; r < n; ++r) { first_row[r] = total; total += counts[r]; }
    if (c->rank == root) {
        out->len = (size_t)total;
        // every rank must take part in the exchange even if the root cannot hold the result: the capacity error is reported afterwards
        CUDA_TRY(e->gx_all.reserve(gx_field_off(5, total) + 256));
    }
    NCCL_TRY(api, api->GroupStart());
    for (int f = 0; f < 5; ++f) {
        if (c->rank == root) {
            uint8_t* dst = e->gx_all.p + gx_field_off(f, total);
            for (int r = 0; r < n; ++r) {
                if (counts[r] == 0) continue;
                if (r == root) CUDA_TRY(cudaMemcpyAsync(dst + GX_ELT[f] * first_row[r], e->gx_local.p + gx_field_off(f, my_rows), GX_ELT[f] * my_rows, cudaMemcpyDeviceToDevice, e->stream));
                else NCCL_TRY(api, api->Recv(dst + GX_ELT[f] * first_row[r], GX_ELT[f] * counts[r], ncclUint8, r, c->comm, e->stream));
            }
        } else if (my_rows) {
            NCCL_TRY(api, api->Send(e->gx_local.p + gx_field_off(f, my_rows), GX_ELT[f] * my_rows, ncclUint8, root, c->comm, e->stream));
        }
    }
    NCCL_TRY(api, api->GroupEnd());
    if (c->rank == root) {
        if (total > out->capacity) {
            CUDA_TRY(cudaStreamSynchronize(e->stream));
            return fail(SYN_ERR_CAPACITY, "experience needs %llu rows, caller provided %zu", (unsigned long long)total, out->capacity);
        }
        // ---- rebuild height / player / features from the bitboards and deliver (device destinations in place, host ones staged)
        const uint8_t* a = e->gx_all.p;
        const uint64_t* all_my = (const uint64_t*)(a + gx_field_off(1, total));
        const uint64_t* all_op = (const uint64_t*)(a + gx_field_off(2, total));
        struct Field { void* dst; size_t elt; size_t off; };
        Field f3[3] = {{out->height, 9, 0}, {out->player, 1, 0}, {out->states, 63 * 4, 0}};
        size_t need = 0;
        for (auto& x : f3)
            if (x.dst && !is_device_ptr(x.dst)) { x.off = need; need += ((x.elt * total + 255) / 256) * 256; }
        CUDA_TRY(e->staging.reserve(need ? need : 256));
        auto tgt = [&](int i) -> void* { return !f3[i].dst ? nullptr : (is_device_ptr(f3[i].dst) ? f3[i].dst : (void*)(e->staging.p + f3[i].off)); };
        if (total && (f3[0].dst || f3[1].dst || f3[2].dst)) {
            expand_rows_kernel<<<(uint32_t)((total * 32 + 255) / 256), 256, 0, e->stream>>>(total, all_my, all_op, (uint8_t*)tgt(0), (uint8_t*)tgt(1), (float*)tgt(2));
            CUDA_TRY(cudaGetLastError());
            e->launches += 1;
        }
        void* dst5[5] = {out->game_ids, out->my_bb, out->op_bb, out->pis, out->vs};
        for (int f = 0; f < 5; ++f)
            if ((rc = deliver(e, dst5[f], a + gx_field_off(f, total), GX_ELT[f] * total))) return rc;
        for (int i = 0; i < 3; ++i)
            if (f3[i].dst && !is_device_ptr(f3[i].dst) && (rc = deliver(e, f3[i].dst, e->staging.p + f3[i].off, f3[i].elt * total))) return rc;
        // games = this call's games over all ranks as far as the root can tell: distinct ids are contiguous per rank
        out->games = 0;
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return read_stats(e, stats, ms);
}

int syn_engine_search(syn_engine* e, const syn_rollout_cfg* cfg, uint32_t tree_kind, const uint64_t* my_bb, const uint64_t* op_bb,
                      const uint64_t* seeds, uint32_t n, float* child_visits, uint8_t* child_solution, float* root_q,
                      uint8_t* root_solution, uint8_t* best_action, uint32_t* num_nodes, syn_stats* stats) {
    if (!e || !my_bb || !op_bb || !seeds) return fail(SYN_ERR_INVALID_ARGUMENT, "engine, bitboards and seeds must not be NULL");
    if (e->pending) return fail(SYN_ERR_INVALID_ARGUMENT, "a gather is in flight");
    if (n == 0) return fail(SYN_ERR_INVALID_ARGUMENT, "n_positions must be > 0");
    if (tree_kind > SYN_TREE_FROZEN) return fail(SYN_ERR_INVALID_ARGUMENT, "bad tree_kind %u", tree_kind);
    int rc = validate_cfg(cfg, e);
    if (rc) return rc;
    if (tree_kind == SYN_TREE_FROZEN && (rc = validate_frozen_cfg(cfg->mcts))) return rc;
    // reject positions the reference's MCTS is never built on: finished games, malformed boards
    for (uint32_t i = 0; i < n && !is_device_ptr(my_bb) && !is_device_ptr(op_bb); ++i) {
        uint64_t my = my_bb[i], op = op_bb[i];
        if ((my & op) || ((my | op) >> 63)) return fail(SYN_ERR_INVALID_ARGUMENT, "position %u: overlapping or out-of-board stones", i);
        if ((my | op) == c4::ALL) return fail(SYN_ERR_INVALID_ARGUMENT, "position %u is a full board", i);
        if (host_won(my) || host_won(op)) return fail(SYN_ERR_INVALID_ARGUMENT, "position %u is already won: MCTS::exploit is never called on a finished game", i);
    }
    CUDA_TRY(cudaSetDevice(e->device));
    e->h2d = 0; e->d2h = 0; e->launches = 0;
    CUDA_TRY(e->pos_my.reserve(n)); CUDA_TRY(e->pos_op.reserve(n)); CUDA_TRY(e->pos_seed.reserve(n));
    CUDA_TRY(e->s_visits.reserve((size_t)n * 9)); CUDA_TRY(e->s_q.reserve((size_t)n * 3)); CUDA_TRY(e->s_csol.reserve((size_t)n * 9));
    CUDA_TRY(e->s_rsol.reserve(n)); CUDA_TRY(e->s_best.reserve(n)); CUDA_TRY(e->s_nodes.reserve(n));
    CUDA_TRY(cudaMemcpyAsync(e->pos_my.p, my_bb, n * 8, cudaMemcpyDefault, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->pos_op.p, op_bb, n * 8, cudaMemcpyDefault, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->pos_seed.p, seeds, n * 8, cudaMemcpyDefault, e->stream));
    e->h2d += (uint64_t)n * 24;
    KParams kp;
    fill_common(e, kp, cfg);
    kp.num_games = n; kp.search_mode = 1;
    kp.pos_my = e->pos_my.p; kp.pos_op = e->pos_op.p; kp.pos_seed = e->pos_seed.p;
    kp.s_child_visits = e->s_visits.p; kp.s_child_sol = e->s_csol.p; kp.s_root_q = e->s_q.p; kp.s_root_sol = e->s_rsol.p;
    kp.s_best = e->s_best.p; kp.s_nodes = e->s_nodes.p;
    CUDA_TRY(cudaMemsetAsync(e->counters.p, 0, CNT_ALL * sizeof(unsigned long long), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->error.p, 0, sizeof(int), e->stream));
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    if (tree_kind == SYN_TREE_FROZEN) { // thread per root (match.cuh), the same player on every root
        mtc::MParams mp;
        std::memset(&mp, 0, sizeof(mp));
        for (int k = 0; k < 2; ++k) {
            mp.players[k].tree_kind = SYN_TREE_FROZEN; mp.players[k].leaf_eval_kind = cfg->leaf_eval_kind;
            mp.players[k].num_explores = cfg->num_explores; mp.players[k].action_selection = cfg->action_selection;
            mp.players[k].mcts = cfg->mcts;
        }
        rc = launch_match(e, kp, mp);
    } else {
        rc = launch_selfplay(e, kp);
    }
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    rc = check_device_error(e);
    if (rc) return rc;
    if ((rc = deliver(e, child_visits, e->s_visits.p, (size_t)n * 36))) return rc;
    if ((rc = deliver(e, child_solution, e->s_csol.p, (size_t)n * 9))) return rc;
    if ((rc = deliver(e, root_q, e->s_q.p, (size_t)n * 12))) return rc;
    if ((rc = deliver(e, root_solution, e->s_rsol.p, n))) return rc;
    if ((rc = deliver(e, best_action, e->s_best.p, n))) return rc;
    if ((rc = deliver(e, num_nodes, e->s_nodes.p, (size_t)n * 4))) return rc;
    return read_stats(e, stats, ms);
}

int syn_engine_match(syn_engine* e, const syn_player_cfg players[2], const uint64_t* seeds, const uint32_t* explores, uint32_t n,
                     float* result, uint8_t* n_moves, uint8_t* moves, uint32_t* tree_nodes, float* child_visits, syn_stats* stats) {
    if (!e || !players || !seeds) return fail(SYN_ERR_INVALID_ARGUMENT, "engine, players and seeds must not be NULL");
    if (e->pending) return fail(SYN_ERR_INVALID_ARGUMENT, "a gather is in flight");
    if (n == 0) return fail(SYN_ERR_INVALID_ARGUMENT, "n_matches must be > 0");
    int rc;
    for (int k = 0; k < 2; ++k)
        if ((rc = validate_player(players[k], e, k))) return rc;
    if (explores && !is_device_ptr(explores))
        for (size_t i = 0; i < 2 * (size_t)n; ++i)
            if (explores[i] > e->max_explores)
                return fail(SYN_ERR_CAPACITY, "explores[%zu][%zu] = %u exceeds the engine's max_explores %u", i / 2, i % 2, explores[i], e->max_explores);
    CUDA_TRY(cudaSetDevice(e->device));
    e->h2d = 0; e->d2h = 0; e->launches = 0;
    size_t rows = (size_t)n * 63;
    CUDA_TRY(e->pos_seed.reserve(n)); CUDA_TRY(e->row_visits.reserve(rows * 9)); CUDA_TRY(e->row_action.reserve(rows));
    CUDA_TRY(e->row_nodes.reserve(rows)); CUDA_TRY(e->s_q.reserve(n)); CUDA_TRY(e->s_rsol.reserve(n));
    if (explores) CUDA_TRY(e->s_nodes.reserve(2 * (size_t)n));
    CUDA_TRY(cudaMemcpyAsync(e->pos_seed.p, seeds, (size_t)n * 8, cudaMemcpyDefault, e->stream));
    e->h2d += (uint64_t)n * 8;
    if (explores) {
        CUDA_TRY(cudaMemcpyAsync(e->s_nodes.p, explores, (size_t)n * 8, cudaMemcpyDefault, e->stream));
        e->h2d += (uint64_t)n * 8;
    }
    CUDA_TRY(cudaMemsetAsync(e->row_action.p, 0, rows, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->row_nodes.p, 0, rows * 4, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->row_visits.p, 0, rows * 36, e->stream));
    syn_rollout_cfg dummy;
    std::memset(&dummy, 0, sizeof(dummy));
    KParams kp;
    fill_common(e, kp, &dummy);
    kp.num_games = n; kp.search_mode = 0; kp.pos_seed = e->pos_seed.p;
    kp.row_visits = e->row_visits.p; kp.row_action = e->row_action.p; kp.row_nodes = e->row_nodes.p;
    mtc::MParams mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.players[0] = players[0]; mp.players[1] = players[1];
    mp.explores = explores ? e->s_nodes.p : nullptr;
    mp.result = e->s_q.p; mp.n_moves = e->s_rsol.p;
    // two different networks (syn_engine_set_opponent_weights): only when both players ask for Connect4Net
    if (e->has_weights2 && players[0].leaf_eval_kind == SYN_LEAF_NN && players[1].leaf_eval_kind == SYN_LEAF_NN) mp.weight_image2 = e->weight_image2.p;
    CUDA_TRY(cudaMemsetAsync(e->counters.p, 0, CNT_ALL * sizeof(unsigned long long), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->error.p, 0, sizeof(int), e->stream));
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    if ((rc = launch_match(e, kp, mp))) return rc;
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    if ((rc = check_device_error(e))) return rc;
    if ((rc = deliver(e, result, e->s_q.p, (size_t)n * 4))) return rc;
    if ((rc = deliver(e, n_moves, e->s_rsol.p, n))) return rc;
    if ((rc = deliver(e, moves, e->row_action.p, rows))) return rc;
    if ((rc = deliver(e, tree_nodes, e->row_nodes.p, rows * 4))) return rc;
    if ((rc = deliver(e, child_visits, e->row_visits.p, rows * 36))) return rc;
    return read_stats(e, stats, ms);
}

int syn_engine_eval(syn_engine* e, const uint64_t* my_bb, const uint64_t* op_bb, uint32_t n, float* logits, float* outcome_probs) {
    if (!e || !my_bb || !op_bb || !logits || !outcome_probs) return fail(SYN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!e->has_weights) return fail(SYN_ERR_NO_WEIGHTS, "syn_engine_set_weights has not been called");
    if (n == 0) return SYN_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(e->pos_my.reserve(n)); CUDA_TRY(e->pos_op.reserve(n));
    CUDA_TRY(e->s_visits.reserve((size_t)n * 9)); CUDA_TRY(e->s_q.reserve((size_t)n * 3));
    CUDA_TRY(cudaMemcpyAsync(e->pos_my.p, my_bb, n * 8, cudaMemcpyDefault, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->pos_op.p, op_bb, n * 8, cudaMemcpyDefault, e->stream));
    {
        int rc0 = launch_eval(e, e->mlp_eff, e->pos_my.p, e->pos_op.p, n, e->s_visits.p, e->s_q.p);
        if (rc0) return rc0;
    }
    CUDA_TRY(cudaGetLastError());
    int rc;
    if ((rc = deliver(e, logits, e->s_visits.p, (size_t)n * 36))) return rc;
    if ((rc = deliver(e, outcome_probs, e->s_q.p, (size_t)n * 12))) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SYN_OK;
}

int syn_engine_play(syn_engine* e, const uint8_t* moves, const uint32_t* n_moves, uint32_t stride, uint32_t n_games, uint64_t* my_bb,
                    uint64_t* op_bb, uint8_t* height, uint8_t* legal_mask_lo, uint8_t* legal_mask_hi, uint8_t* status, float* features) {
    if (!e || !moves || !n_moves || !my_bb || !op_bb || !height || !legal_mask_lo || !legal_mask_hi || !status)
        return fail(SYN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_games == 0) return SYN_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    size_t mv = (size_t)n_games * stride;
    size_t off_nm = (mv + 255) / 256 * 256;
    size_t off_my = (off_nm + (size_t)n_games * 4 + 15) / 16 * 16, off_op = off_my + (size_t)n_games * 8;
    size_t off_h = off_op + (size_t)n_games * 8, off_lo = off_h + (size_t)n_games * 9, off_hi = off_lo + n_games, off_st = off_hi + n_games;
    size_t off_f = (off_st + n_games + 15) / 16 * 16, total = off_f + (features ? (size_t)n_games * 63 * 4 : 0);
    CUDA_TRY(e->staging.reserve(total));
    uint8_t* b = e->staging.p;
    CUDA_TRY(cudaMemcpyAsync(b, moves, mv, cudaMemcpyDefault, e->stream));
    CUDA_TRY(cudaMemcpyAsync(b + off_nm, n_moves, (size_t)n_games * 4, cudaMemcpyDefault, e->stream));
    uint32_t blocks = (n_games * 32 + 255) / 256;
    play_kernel<<<blocks, 256, 0, e->stream>>>(b, (const uint32_t*)(b + off_nm), stride, n_games, (uint64_t*)(b + off_my), (uint64_t*)(b + off_op),
                                               b + off_h, b + off_lo, b + off_hi, b + off_st, features ? (float*)(b + off_f) : nullptr);
    CUDA_TRY(cudaGetLastError());
    int rc;
    if ((rc = deliver(e, my_bb, b + off_my, (size_t)n_games * 8))) return rc;
    if ((rc = deliver(e, op_bb, b + off_op, (size_t)n_games * 8))) return rc;
    if ((rc = deliver(e, height, b + off_h, (size_t)n_games * 9))) return rc;
    if ((rc = deliver(e, legal_mask_lo, b + off_lo, n_games))) return rc;
    if ((rc = deliver(e, legal_mask_hi, b + off_hi, n_games))) return rc;
    if ((rc = deliver(e, status, b + off_st, n_games))) return rc;
    if (features && (rc = deliver(e, features, b + off_f, (size_t)n_games * 63 * 4))) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SYN_OK;
}

// ---- ReplayBuffer::deduplicate (data.rs:196-235) on the device, see dedup.cuh
static int dd_scan(syn_engine* e, const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, uint32_t* total) {
    const uint32_t nb = (n + dd::SCAN_TILE - 1) / dd::SCAN_TILE;
    if (nb <= 1u) {
        dd::scan_down_kernel<<<1, dd::T, 0, e->stream>>>(in, out, n, nullptr, total);
        e->launches += 1;
        return SYN_OK;
    }
    dd::scan_reduce_kernel<<<nb, dd::T, 0, e->stream>>>(in, n, tmp);
    e->launches += 1;
    int rc = dd_scan(e, tmp, tmp, nb, tmp + (nb + 63u) / 64u * 64u, nullptr);
    if (rc) return rc;
    dd::scan_down_kernel<<<nb, dd::T, 0, e->stream>>>(in, out, n, tmp, total);
    e->launches += 1;
    return SYN_OK;
}

int syn_engine_deduplicate(syn_engine* e, const uint64_t* my_bb, const uint64_t* op_bb, const float* pis, const float* vs, size_t n_rows,
                           syn_flat_batch* out, syn_stats* stats) {
    if (!e || !out) return fail(SYN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (e->pending) return fail(SYN_ERR_INVALID_ARGUMENT, "a gather is in flight on this engine");
    out->len = 0;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n_rows == 0) return SYN_OK;
    if (!my_bb || !op_bb || !pis || !vs) return fail(SYN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_rows > (1ull << 30)) return fail(SYN_ERR_CAPACITY, "deduplicate handles at most 2^30 rows per call, got %zu", n_rows);
    CUDA_TRY(cudaSetDevice(e->device));
    e->h2d = 0; e->d2h = 0; e->launches = 0;
    const uint32_t n = (uint32_t)n_rows;
    uint32_t cap = 1024;
    while (cap < 2u * n) cap <<= 1;
    const uint32_t nblk = (n + dd::RS_TILE - 1) / dd::RS_TILE;
    const uint32_t hist_n = 256u * nblk;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    // ---- workspace
    const size_t scan_m = std::max<size_t>(n, hist_n);
    const size_t scan_tmp = al((scan_m / dd::SCAN_TILE + 64) * 2 * 4 + 4096);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
    const size_t o_srow = take((size_t)cap * 4), o_sfirst = take((size_t)cap * 4);
    const size_t o_ka = take((size_t)(n + 1) * 4), o_va = take((size_t)(n + 1) * 4), o_kb = take((size_t)(n + 1) * 4), o_vb = take((size_t)(n + 1) * 4);
    const size_t o_hist = take((size_t)hist_n * 4), o_tmp = take(scan_tmp), o_big = take((size_t)(n / dd::BIG + 2) * 4), o_scal = take(256);
    const size_t o_pv = take((size_t)n * 48);
    CUDA_TRY(e->dd_ws.reserve(off));
    uint8_t* w = e->dd_ws.p;
    uint32_t *slot_row = (uint32_t*)(w + o_srow), *slot_first = (uint32_t*)(w + o_sfirst);
    uint32_t *ka = (uint32_t*)(w + o_ka), *va = (uint32_t*)(w + o_va), *kb = (uint32_t*)(w + o_kb), *vb = (uint32_t*)(w + o_vb);
    uint32_t *hist = (uint32_t*)(w + o_hist), *tmp = (uint32_t*)(w + o_tmp), *big_list = (uint32_t*)(w + o_big);
    uint32_t *n_groups = (uint32_t*)(w + o_scal), *big_count = n_groups + 1;
    float* pv = (float*)(w + o_pv);
    // ---- inputs: used in place when they already live on the device
    const void* src[4] = {my_bb, op_bb, pis, vs};
    const size_t elt[4] = {8, 8, 36, 12};
    const void* in[4];
    size_t io = 0, in_off[4];
    for (int i = 0; i < 4; ++i) { in_off[i] = io; if (!is_device_ptr(src[i])) io += al(elt[i] * n); }
    // ---- outputs: staged when the destination is host memory (sized by the caller's capacity, at most n)
    const size_t ocap = std::min<size_t>(out->capacity, n);
    void* dst[6] = {out->my_bb, out->op_bb, out->num, out->states, out->pis, out->vs};
    const size_t oelt[6] = {8, 8, 4, 252, 36, 12};
    size_t out_off[6];
    for (int i = 0; i < 6; ++i) { out_off[i] = io; if (dst[i] && !is_device_ptr(dst[i])) io += al(oelt[i] * ocap); }
    CUDA_TRY(e->dd_io.reserve(io ? io : 256));
    for (int i = 0; i < 4; ++i) {
        if (is_device_ptr(src[i])) { in[i] = src[i]; continue; }
        CUDA_TRY(cudaMemcpyAsync(e->dd_io.p + in_off[i], src[i], elt[i] * n, cudaMemcpyHostToDevice, e->stream));
        e->h2d += elt[i] * n;
        in[i] = e->dd_io.p + in_off[i];
    }
    const uint64_t* d_my = (const uint64_t*)in[0];
    const uint64_t* d_op = (const uint64_t*)in[1];
    const float* d_pis = (const float*)in[2];
    const float* d_vs = (const float*)in[3];
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    const uint32_t gridn = (n + dd::T - 1) / dd::T;
    // ---- group keys: rep[i] = first row holding row i's position
    dd::fill_kernel<<<e->sm_count * 8, dd::T, 0, e->stream>>>(slot_row, (size_t)2 * cap, dd::EMPTY); // slot_row and slot_first are adjacent (cap * 4 is a multiple of 256)
    dd::insert_kernel<<<gridn, dd::T, 0, e->stream>>>(d_my, d_op, n, slot_row, slot_first, cap - 1u, ka);
    dd::rep_kernel<<<gridn, dd::T, 0, e->stream>>>(ka, slot_first, n);
    e->launches += 3;
    // ---- stable sort of (rep, row) by rep
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (unsigned long long)n) ++bits;
    uint32_t *kin = ka, *vin = nullptr, *kout = kb, *vout = vb;
    for (int shift = 0; shift < bits; shift += 8) {
        dd::rs_hist_kernel<<<nblk, dd::T, 0, e->stream>>>(kin, n, shift, hist, nblk);
        e->launches += 1;
        int rc = dd_scan(e, hist, hist, hist_n, tmp, nullptr);
        if (rc) return rc;
        dd::rs_scatter_kernel<<<nblk, dd::T, 0, e->stream>>>(kin, vin, kout, vout, n, shift, hist, nblk);
        e->launches += 1;
        uint32_t* nk = kout; uint32_t* nv = vout;
        kout = kin; vout = (vin ? vin : va);
        kin = nk; vin = nv;
    }
    uint32_t *rep_sorted = kin, *rows_sorted = vin, *gid = kout, *gstart = vout;
    // ---- group boundaries and start table
    dd::heads_kernel<<<gridn, dd::T, 0, e->stream>>>(rep_sorted, n, gid);
    e->launches += 1;
    {
        int rc = dd_scan(e, gid, gid, n, tmp, n_groups);
        if (rc) return rc;
    }
    dd::starts_kernel<<<gridn, dd::T, 0, e->stream>>>(rep_sorted, gid, n, n_groups, gstart);
    e->launches += 1;
    CUDA_TRY(cudaMemsetAsync(big_count, 0, 4, e->stream));
    uint32_t U = 0;
    CUDA_TRY(cudaMemcpyAsync(&U, n_groups, 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaGetLastError());
    out->len = U;
    if (U > out->capacity) return fail(SYN_ERR_CAPACITY, "deduplicate found %u distinct positions, caller provided room for %zu", U, out->capacity);
    // ---- sums in row order, averages, features
    dd::Out o;
    void* tgt[6];
    for (int i = 0; i < 6; ++i) tgt[i] = !dst[i] ? nullptr : (is_device_ptr(dst[i]) ? dst[i] : (void*)(e->dd_io.p + out_off[i]));
    o.my_bb = (uint64_t*)tgt[0]; o.op_bb = (uint64_t*)tgt[1]; o.num = (uint32_t*)tgt[2];
    o.states = (float*)tgt[3]; o.pis = (float*)tgt[4]; o.vs = (float*)tgt[5];
    dd::gather_kernel<<<(uint32_t)(((size_t)n * 16 + dd::T - 1) / dd::T), dd::T, 0, e->stream>>>(rows_sorted, n, d_pis, d_vs, pv);
    dd::reduce_kernel<<<(uint32_t)(((size_t)U * 16 + dd::T - 1) / dd::T), dd::T, 0, e->stream>>>(rows_sorted, gstart, n_groups, d_my, d_op, pv, o, big_list,
                                                                                                  big_count);
    e->launches += 2;
    uint32_t nbig = 0;
    CUDA_TRY(cudaMemcpyAsync(&nbig, big_count, 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (nbig) {
        CUDA_TRY(cudaFuncSetAttribute(dd::reduce_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dd::BIG_SMEM));
        dd::reduce_big_kernel<<<nbig, dd::BIG_T, dd::BIG_SMEM, e->stream>>>(rows_sorted, gstart, big_list, d_my, d_op, pv, o);
        e->launches += 1;
    }
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    CUDA_TRY(cudaGetLastError());
    for (int i = 0; i < 6; ++i)
        if (dst[i] && !is_device_ptr(dst[i])) {
            int rc = deliver(e, dst[i], e->dd_io.p + out_off[i], oelt[i] * U);
            if (rc) return rc;
        }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (stats) {
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        stats->rows = U;
        stats->device_ns = (uint64_t)((double)ms * 1e6);
        stats->kernel_launches = e->launches;
        stats->h2d_bytes = e->h2d;
        stats->d2h_bytes = e->d2h + 8;
    }
    return SYN_OK;
}

// ---- the learner's inner loop (alpha_zero.rs:73-92) on the device, see train.cuh
int syn_engine_reset_optimizer(syn_engine* e) {
    if (!e) return fail(SYN_ERR_INVALID_ARGUMENT, "engine is NULL");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(e->adam_m.reserve(trn::P_FLOATS));
    CUDA_TRY(e->adam_v.reserve(trn::P_FLOATS));
    CUDA_TRY(cudaMemsetAsync(e->adam_m.p, 0, trn::P_FLOATS * sizeof(float), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->adam_v.p, 0, trn::P_FLOATS * sizeof(float), e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->adam_t = 0;
    return SYN_OK;
}

int syn_engine_get_weights(syn_engine* e, float* blob, size_t n_floats) {
    if (!e || !blob) return fail(SYN_ERR_INVALID_ARGUMENT, "engine or blob is NULL");
    if (n_floats != SYN_N_WEIGHTS) return fail(SYN_ERR_INVALID_ARGUMENT, "expected %d floats, got %zu", SYN_N_WEIGHTS, n_floats);
    if (!e->has_weights) return fail(SYN_ERR_NO_WEIGHTS, "syn_engine_set_weights has not been called");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaMemcpyAsync(blob, e->weights.p, n_floats * sizeof(float), cudaMemcpyDefault, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return SYN_OK;
}

int syn_engine_train(syn_engine* e, const syn_train_cfg* cfg, const uint64_t* my_bb, const uint64_t* op_bb, const float* pis, const float* vs,
                     size_t n_rows, const uint32_t* batch_index, uint32_t n_batches, float* losses, syn_stats* stats) {
    if (!e || !cfg) return fail(SYN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (e->pending) return fail(SYN_ERR_INVALID_ARGUMENT, "a gather is in flight on this engine");
    if (!e->has_weights) return fail(SYN_ERR_NO_WEIGHTS, "syn_engine_set_weights has not been called");
    // SYN_TRAIN_CLUSTER: 0 = the single-CTA kernel (train.cuh); 1 = a cluster of 8 CTAs exchanging through cluster.sync();
    // 2 (default) = the cluster with asynchronous remote stores and mbarriers (train_cluster.cuh)
    const char* tc = std::getenv("SYN_TRAIN_CLUSTER");
    const int mode = tc ? std::atoi(tc) : 2;
    // LearningConfig::batch_size is free in the reference (config.rs:76-94; 32 in study-connect4/src/main.rs:21).  The kernels
    // are tiled for 32 rows; a multiple of 32 runs as micro-batches whose gradients are summed before the Adam update.
    if (cfg->batch_size == 0 || cfg->batch_size % (uint32_t)trn::B != 0 || cfg->batch_size > 4096u)
        return fail(SYN_ERR_UNSUPPORTED, "batch_size %u: the device learner takes multiples of %d up to 4096 (study-connect4/src/main.rs:21 uses 32)", cfg->batch_size, trn::B);
    const uint32_t micro = cfg->batch_size / (uint32_t)trn::B;
    if (micro > 1 && mode != 2) return fail(SYN_ERR_UNSUPPORTED, "batch_size %u needs the default learner kernel (SYN_TRAIN_CLUSTER=2)", cfg->batch_size);
    if ((uint64_t)n_batches * micro > 0xffffffffull) return fail(SYN_ERR_CAPACITY, "too many batches");
    if (!(cfg->beta1 >= 0.0f && cfg->beta1 < 1.0f && cfg->beta2 >= 0.0f && cfg->beta2 < 1.0f && cfg->eps > 0.0f && cfg->lr >= 0.0f))
        return fail(SYN_ERR_INVALID_ARGUMENT, "bad Adam hyper-parameters");
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n_batches == 0) return SYN_OK;
    if (!my_bb || !op_bb || !pis || !vs || !batch_index || n_rows == 0) return fail(SYN_ERR_INVALID_ARGUMENT, "NULL argument");
    if (n_rows > 0xffffffffull) return fail(SYN_ERR_CAPACITY, "at most 2^32-1 rows");
    CUDA_TRY(cudaSetDevice(e->device));
    e->h2d = 0; e->d2h = 0; e->launches = 0;
    if (!e->adam_m.p) {
        int rc = syn_engine_reset_optimizer(e);
        if (rc) return rc;
    }
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const void* src[5] = {my_bb, op_bb, pis, vs, batch_index};
    const size_t bytes[5] = {8 * n_rows, 8 * n_rows, 36 * n_rows, 12 * n_rows, (size_t)n_batches * cfg->batch_size * 4};
    size_t off[5], io = 0;
    for (int i = 0; i < 5; ++i) { off[i] = io; if (!is_device_ptr(src[i])) io += al(bytes[i]); }
    const size_t o_sched = io; io += al((size_t)n_batches * sizeof(float2));
    const size_t o_loss = io; io += al((size_t)n_batches * 2 * sizeof(float));
    CUDA_TRY(e->tr_io.reserve(io));
    const void* in[5];
    for (int i = 0; i < 5; ++i) {
        if (is_device_ptr(src[i])) { in[i] = src[i]; continue; }
        CUDA_TRY(cudaMemcpyAsync(e->tr_io.p + off[i], src[i], bytes[i], cudaMemcpyHostToDevice, e->stream));
        e->h2d += bytes[i];
        in[i] = e->tr_io.p + off[i];
    }
    // bias corrections in double like torch::optim::Adam::step, one pair per optimizer step
    std::vector<float2> sched(n_batches);
    for (uint32_t k = 0; k < n_batches; ++k) {
        const double t = (double)(e->adam_t + k + 1);
        sched[k].x = (float)((double)cfg->lr / (1.0 - std::pow((double)cfg->beta1, t)));
        sched[k].y = (float)(1.0 / std::sqrt(1.0 - std::pow((double)cfg->beta2, t)));
    }
    CUDA_TRY(cudaMemcpyAsync(e->tr_io.p + o_sched, sched.data(), (size_t)n_batches * sizeof(float2), cudaMemcpyHostToDevice, e->stream));
    e->h2d += (size_t)n_batches * sizeof(float2);
    CUDA_TRY(cudaMemsetAsync(e->error.p, 0, sizeof(int), e->stream));
    trn::Params tp;
    tp.blob = e->weights.p; tp.m = e->adam_m.p; tp.v = e->adam_v.p;
    tp.my = (const uint64_t*)in[0]; tp.op = (const uint64_t*)in[1]; tp.pis = (const float*)in[2]; tp.vs = (const float*)in[3];
    tp.batch_idx = (const uint32_t*)in[4]; tp.sched = (const float2*)(e->tr_io.p + o_sched);
    tp.losses = (float*)(e->tr_io.p + o_loss); tp.error = e->error.p;
    const char* tprof = std::getenv("SYN_TRAIN_PROF"); // per-phase clocks of the async cluster kernel, printed to stderr
    tp.prof = (tprof && std::atoi(tprof) == 1) ? (unsigned long long*)e->counters.p : nullptr;
    tp.n_rows = (uint32_t)n_rows; tp.n_steps = n_batches; tp.micro = micro;
    tp.beta1 = cfg->beta1; tp.beta2 = cfg->beta2; tp.eps = cfg->eps; tp.wd = cfg->weight_decay; tp.pw = cfg->policy_weight; tp.vw = cfg->value_weight;
    if (mode == 2) CUDA_TRY(cudaFuncSetAttribute(trc::train_cluster_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(trc::SmemA)));
    else if (mode == 1) CUDA_TRY(cudaFuncSetAttribute(trc::train_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(trc::Smem)));
    else CUDA_TRY(cudaFuncSetAttribute(trn::train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(trn::Smem)));
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    if (mode == 2) trc::train_cluster_async_kernel<<<trc::NC, trc::NT, sizeof(trc::SmemA), e->stream>>>(tp);
    else if (mode == 1) trc::train_cluster_kernel<<<trc::NC, trc::NT, sizeof(trc::Smem), e->stream>>>(tp);
    else trn::train_kernel<<<1, trn::NT, sizeof(trn::Smem), e->stream>>>(tp);
    CUDA_TRY(cudaGetLastError());
    mlptc::build_weight_image<<<32, 256, 0, e->stream>>>(e->weights.p, e->weight_image.p); // the search kernels see the new weights
    mlps::build_weight_image_lo<<<32, 256, 0, e->stream>>>(e->weights.p, e->weight_image_lo.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(e->bias_host, e->weight_image.p + mlptc::BIAS_OFF, sizeof(e->bias_host), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    e->launches += 2;
    e->adam_t += n_batches;
    if (losses) {
        int rc = deliver(e, losses, e->tr_io.p + o_loss, (size_t)n_batches * 2 * sizeof(float));
        if (rc) return rc;
    }
    int derr = 0;
    CUDA_TRY(cudaMemcpyAsync(&derr, e->error.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (derr) return fail(SYN_ERR_INVALID_ARGUMENT, "batch_index holds a row >= n_rows (%zu)", n_rows);
    if (tp.prof && mode == 2) {
        unsigned long long pc[14];
        CUDA_TRY(cudaMemcpy(pc, tp.prof, sizeof(pc), cudaMemcpyDeviceToHost));
        static const char* names[13] = {"stage", "fwd0", "fwd1", "fwd2", "fwd3", "fwd4", "loss", "bwd4", "bwd3", "bwd2", "bwd1", "bwd0", "endbar"};
        std::fprintf(stderr, "train phases (cycles/step, rank 0 thread 0):");
        for (int k = 0; k < 13; ++k) std::fprintf(stderr, " %s %.0f", names[k], (double)pc[k] / ((double)n_batches * micro));
        std::fprintf(stderr, "\n");
    }
    if (stats) {
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        stats->rows = (uint64_t)n_batches * cfg->batch_size;
        stats->device_ns = (uint64_t)((double)ms * 1e6);
        stats->kernel_launches = e->launches;
        stats->h2d_bytes = e->h2d;
        stats->d2h_bytes = e->d2h + 4;
    }
    return calibrate_mlp(e); // the weights moved: which forward chain is accurate enough is measured again
}

} // extern "C"
