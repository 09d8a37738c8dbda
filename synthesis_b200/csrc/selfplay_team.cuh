// selfplay_team.cuh — lane group per game, network leaves, TEAMS of 128 threads per CTA that run their rounds independently.
//
// Same algorithm and same results as selfplay_nn_tc_kernel (selfplay.cuh: gather_experience / run_game / MCTS with
// Connect4Net leaves, synthesis/src/alpha_zero.rs:120-338, mcts.rs:29-489, study-connect4/src/policies.rs:28-59); what
// changes is who waits for whom.  In that kernel a round ends at a CTA barrier: 28 groups wait for the deepest of 28
// descents, then for a forward that nobody overlaps — per-group clocks (profiles/r2_lane_group_clocks.txt): 12 k cycles of
// descent, 13 k of waiting, 6 k of forward, 3 k of finish.  Here a CTA is TEAMS teams of four warps; a team takes an MLP slot
// (128 TMEM columns + an mbarrier) for its forward and synchronises on its own named barrier, so a group waits for the
// 128 / GL - 1 other groups of its team only, and one team's forward runs under the other teams' descents.  All teams
// share ONE A tile (see LgTeamSmem).  The group writes its leaf's features into its row (one 16-byte column entry per
// lane), the lane whose thread index in the team equals that row runs the row's epilogue like a tpg2 thread does, and
// hands the 12 outputs to the group by shuffles.
#pragma once
#include "mlp_split.cuh"
#include "mlp_team.cuh"
#include "selfplay.cuh"

namespace eng {

// Shared memory of the kernel: ONE weight image and ONE A tile for all teams.  A team's leaves occupy the rows
// gt * GL + team (gt = the group's index in its team), so the teams' rows are disjoint and every team can run its chain
// over the whole 128-row tile whenever it is ready: the rows of other teams are read as whatever they hold at that moment
// and produce accumulator rows nobody looks at.  What a team owns is an MLP slot = 128 TMEM columns + an mbarrier.
// 107 KB instead of the 202 KB of a tile per team: the difference stays L1, which at these batch sizes holds the upper
// levels of every tree of the SM (with a tile per team the descents ran 35 % slower, profiles/r2_lane_group_clocks.txt).
template <int TEAMS, int SLOTS>
struct __align__(128) LgTeamSmem {
    uint8_t img[mlptc::IMG_BYTES];
    uint8_t a[mlpteam::A_BYTES];
    uint4 col_lut[256];
    uint64_t bar_w;
    uint64_t bar_mma[SLOTS];
    uint32_t slot_busy[SLOTS];
    uint32_t slot_phase[SLOTS];
    uint32_t team_slot[TEAMS];
    uint32_t tmem_base;
    uint32_t pad;
};

template <int TEAMS, int SLOTS, int GL>
constexpr size_t nn_team_smem_bytes() { return sizeof(LgTeamSmem<TEAMS, SLOTS>) + (size_t)(128 * TEAMS / GL) * (64 + FPU_SM_WORDS) * sizeof(uint32_t); }

namespace lgteam {
using namespace mlptc;

template <int TEAMS, int SLOTS>
__device__ __forceinline__ void setup(LgTeamSmem<TEAMS, SLOTS>& s, const uint8_t* __restrict__ weight_image) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&s.bar_w, 1);
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) { mbar_init(&s.bar_mma[t], 1); s.slot_busy[t] = 0u; s.slot_phase[t] = 0u; }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&s.tmem_base, 128 * SLOTS);
    for (int i = threadIdx.x; i < mlpteam::A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(s.a)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s.col_lut[i] = mlpteam::col_lut_entry(i < 255 ? i : 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s.bar_w, IMG_BYTES);
        bulk_g2s(s.img, weight_image, IMG_BYTES, &s.bar_w);
    }
    mbar_wait(&s.bar_w, 0);
    fence_proxy_async();
    __syncthreads();
}

template <int TEAMS, int SLOTS>
__device__ __forceinline__ void teardown(LgTeamSmem<TEAMS, SLOTS>& s) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc(s.tmem_base, 128 * SLOTS);
}

template <int TEAMS, int SLOTS>
__device__ __forceinline__ int acquire_slot(LgTeamSmem<TEAMS, SLOTS>& s, int team, int r, uint32_t& phase) {
    if (SLOTS == TEAMS) { phase = s.slot_phase[team]; return team; }
    if (r == 0) {
        int got = -1;
        for (int k = team; got < 0; ++k) {
            int cand = k % SLOTS;
            if (atomicCAS(&s.slot_busy[cand], 0u, 1u) == 0u) got = cand;
            else if (cand == SLOTS - 1) __nanosleep(64);
        }
        __threadfence_block();
        s.team_slot[team] = (uint32_t)got;
    }
    mlpteam::team_sync(team);
    int slot = (int)s.team_slot[team];
    phase = s.slot_phase[slot];
    return slot;
}

template <int TEAMS, int SLOTS>
__device__ __forceinline__ void release_slot(LgTeamSmem<TEAMS, SLOTS>& s, int team, int r, int slot, uint32_t phase) {
    if (SLOTS == TEAMS) { if (r == 0) s.slot_phase[slot] = phase; return; }
    tc_fence_before();
    mlpteam::team_sync(team);
    if (r == 0) {
        s.slot_phase[slot] = phase;
        __threadfence_block();
        atomicExch(&s.slot_busy[slot], 0u);
    }
}

// mlpteam::forward_img on the shared tile: the chain of one team over all 128 rows; only `owner` threads (one per group:
// the thread whose index in the team is the group's tile row) run the arithmetic of the epilogue and write the row back.
template <int TEAMS, int SLOTS>
__device__ __forceinline__ void forward(LgTeamSmem<TEAMS, SLOTS>& s, int team, int slot, int r, bool owner, uint32_t& phase, float (&y)[12]) {
    uint8_t* a_tile = s.a;
    const uint32_t tmem = s.tmem_base + (uint32_t)(slot * 128);
    const uint32_t tlane = tmem + ((uint32_t)((r >> 5) * 32) << 16);
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int K = layer_k(l), N = layer_n(l);
        fence_proxy_async();
        tc_fence_before();
        mlpteam::team_sync(team);
        if (r == 0) {
            tc_fence_after();
            const uint32_t a_base = smem_u32(a_tile), b_base = smem_u32(s.img + w_off(l));
#pragma unroll
            for (int kk = 0; kk < K / 16; ++kk) {
                uint64_t ad = make_desc(a_base + kk * 2 * (M_TILE * 16), M_TILE * 16, 128);
                uint64_t bd = make_desc(b_base + kk * 2 * (N * 16), N * 16, 128);
                umma_f16(tmem, ad, bd, make_idesc(N), kk > 0 ? 1u : 0u);
            }
            umma_commit(&s.bar_mma[slot]);
        }
        mbar_wait(&s.bar_mma[slot], phase);
        phase ^= 1u;
        tc_fence_after();
        const float* bias = reinterpret_cast<const float*>(s.img + BIAS_OFF) + b_off(l);
        // (up to four 16-column reads in flight per tcgen05.wait::ld were measured here too: no change, 72 bytes of spills)
#pragma unroll
        for (int c16 = 0; c16 < N / 16; ++c16) {
            uint32_t v[16];
            mlpteam::tmem_ld16(tlane + (uint32_t)(c16 * 16), v);
            tmem_ld_wait();
            if (owner) {
                if (l < NL - 1) mlpteam::epilogue16(v, bias, a_tile, r, c16);
                else {
#pragma unroll
                    for (int j = 0; j < 12; ++j) y[j] = __uint_as_float(v[j]) + bias[j];
                }
            }
        }
    }
    tc_fence_before();
}
} // namespace lgteam

template <int GL, int TEAMS, int SLOTS>
__global__ void __launch_bounds__(128 * TEAMS, 1) selfplay_nn_team_kernel(const __grid_constant__ KParams p) {
    static_assert(TEAMS <= GL, "a group's tile row is gt * GL + team");
    constexpr int GPT = 128 / GL;        // groups per team
    constexpr int GPB = GPT * TEAMS;     // groups per CTA
    extern __shared__ __align__(128) uint8_t smem_raw[];
    LgTeamSmem<TEAMS, SLOTS>& ms = *reinterpret_cast<LgTeamSmem<TEAMS, SLOTS>*>(smem_raw);
    uint32_t* s_path = reinterpret_cast<uint32_t*>(smem_raw + sizeof(LgTeamSmem<TEAMS, SLOTS>));
    lgteam::setup<TEAMS, SLOTS>(ms, p.weight_image);

    Grp<GL> g;
    const int team = threadIdx.x >> 7, r = threadIdx.x & 127;
    const int gt = r / GL;               // the group's index in its team
    const int row = gt * GL + team;      // its row of the shared tile; lane `team` of the group is the row's owner (r == row)
    const bool owner = g.gl == team;
    const int seat = gt * TEAMS + team;  // seats are dealt team by team, so that every team holds its share of few games
    const size_t slot_id = (size_t)blockIdx.x * GPB + (size_t)seat;
    Tree<GL> t;
    t.stat.base = t.meta.base = p.nodes + 2 * slot_id * p.arena_nodes;
    t.path = s_path + (team * GPT + gt) * 64;
    t.fpu_sm = s_path + GPB * 64 + (team * GPT + gt) * FPU_SM_WORDS; // after the path tables
    t.cap = p.arena_nodes;
    t.cfg = &p.cfg.mcts;
    t.err = 0;
    t.nn = 1;
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) t.cnt[i] = 0u;
    GroupState<GL, true> st;
    st.phase = lg_seated(p, seat) ? PH_NEED_GAME : PH_DONE;
    st.is_init = false;
    RolloutRng<GL> rr;
    rr.buf = nullptr; rr.kA = rr.kB = 0; rr.pos = 0;
    Pending pend;
    LGP_INIT(t);
    long long lgp_t = LGP_NOW();
    for (;;) {
        const bool need = advance<GL, true>(g, p, t, st, rr, nullptr, pend);
        LGP_MARK(adv, lgp_t);
        if (!mlpteam::team_any(team, need)) break; // advance() only comes back without a leaf when the group has no game left
        uint32_t mma_phase;
        const int slot = lgteam::acquire_slot<TEAMS, SLOTS>(ms, team, r, mma_phase);
        LGP_MARK(wait, lgp_t);
        if (need) { // Game::features of the leaf into the group's row: lane c writes board column c (mlp_team.cuh's table)
            const uint64_t occ = pend.my | pend.op;
            if (g.gl < 9) {
                const uint32_t oc = (uint32_t)(occ >> (7 * g.gl)) & 0x7fu, mc = (uint32_t)(pend.my >> (7 * g.gl)) & 0x7fu;
                *reinterpret_cast<uint4*>(ms.a + g.gl * (mlptc::M_TILE * 16) + row * 16) = ms.col_lut[oc + mc];
            } else if (g.gl == 9) {
                *reinterpret_cast<uint4*>(ms.a + 9 * (mlptc::M_TILE * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u); // K 72..79
            }
        }
        float y[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) y[j] = 0.0f;
        lgteam::forward<TEAMS, SLOTS>(ms, team, slot, r, owner, mma_phase, y);
        lgteam::release_slot<TEAMS, SLOTS>(ms, team, r, slot, mma_phase);
        LGP_MARK(leaf, lgp_t);
#ifdef SYN_LG_PROF
        ++lgp.rounds;
#endif
        float logit = 0.0f;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float v = g.shfl(y[j], team);
            if (g.gl == j) logit = v;
        }
        const float y0 = g.shfl(y[9], team), y1 = g.shfl(y[10], team), y2 = g.shfl(y[11], team);
        if (need) {
            // value.softmax(-1) (study-connect4/src/policies.rs:54-56)
            const float m = fmaxf(y0, fmaxf(y1, y2));
            const float e0 = syn_expf(__fsub_rn(y0, m)), e1 = syn_expf(__fsub_rn(y1, m)), e2 = syn_expf(__fsub_rn(y2, m));
            const float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
            explore_finish(g, t, pend, false, logit, __fdiv_rn(e0, tot), __fdiv_rn(e1, tot), __fdiv_rn(e2, tot));
            after_eval(g, t, st);
        }
        LGP_MARK(fin, lgp_t);
    }
    flush_counters(g, p, t);
    LGP_FLUSH(g, p, t);
    lgteam::teardown<TEAMS, SLOTS>(ms);
}

// (One step further was measured and lost: every WARP running its chain alone — a warp can read the TMEM lanes of its own
// quarter, which are the rows its lanes stand on.  Nobody waits for anybody's descent, but 16 warps then queue for the four
// TMEM slots with a 9 k-cycle forward each: 166 against 223 M explores/s at 4,096 games in flight, 83 against 81 M at 1,000.
// A forward shared by the eight groups of a team is the better trade.  profiles/r2_lane_group_clocks.txt)

// The same schedule with the split-fp16 chain (mlp_split.cuh: x = hi + lo operands, three MMAs per K-step, activations in
// tensor memory — fp32-grade outputs, what the engine selects for weights that the single-fp16 chain cannot carry).
// Activations never touch shared memory here, so there is no tile to share: every lane of a group puts the group's leaf on
// ITS OWN tile row (the rows of a group are copies), runs its row's epilogue in lockstep with the others, and ends up
// holding all 12 outputs — no hand-over by shuffles.  Four teams take turns on two slots of 256 TMEM columns.
template <int TEAMS, int GL>
constexpr size_t nn_team_split_smem_bytes() { return sizeof(mlps::Smem<TEAMS, 2>) + (size_t)(128 * TEAMS / GL) * (64 + FPU_SM_WORDS) * sizeof(uint32_t); }

template <int GL, int TEAMS>
__global__ void __launch_bounds__(128 * TEAMS, 1) selfplay_nn_team_split_kernel(const __grid_constant__ KParams p) {
    constexpr int SLOTS = 2;
    constexpr int GPT = 128 / GL, GPB = GPT * TEAMS;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    mlps::Smem<TEAMS, SLOTS>& ms = *reinterpret_cast<mlps::Smem<TEAMS, SLOTS>*>(smem_raw);
    uint32_t* s_path = reinterpret_cast<uint32_t*>(smem_raw + sizeof(mlps::Smem<TEAMS, SLOTS>));
    mlps::setup<TEAMS, SLOTS>(ms, p.weight_image, p.weight_image_lo);

    Grp<GL> g;
    const int team = threadIdx.x >> 7, r = threadIdx.x & 127;
    const int gt = r / GL;
    const int seat = gt * TEAMS + team;
    const size_t slot_id = (size_t)blockIdx.x * GPB + (size_t)seat;
    Tree<GL> t;
    t.stat.base = t.meta.base = p.nodes + 2 * slot_id * p.arena_nodes;
    t.path = s_path + (team * GPT + gt) * 64;
    t.fpu_sm = s_path + GPB * 64 + (team * GPT + gt) * FPU_SM_WORDS; // after the path tables
    t.cap = p.arena_nodes;
    t.cfg = &p.cfg.mcts;
    t.err = 0;
    t.nn = 1;
#pragma unroll
    for (int i = 0; i < CNT_N; ++i) t.cnt[i] = 0u;
    GroupState<GL, true> st;
    st.phase = lg_seated(p, seat) ? PH_NEED_GAME : PH_DONE;
    st.is_init = false;
    RolloutRng<GL> rr;
    rr.buf = nullptr; rr.kA = rr.kB = 0; rr.pos = 0;
    Pending pend;
    pend.my = pend.op = 0ull;
    LGP_INIT(t);
    long long lgp_t = LGP_NOW();
    for (;;) {
        const bool need = advance<GL, true>(g, p, t, st, rr, nullptr, pend);
        LGP_MARK(adv, lgp_t);
        if (!mlps::team_any(team, need)) break;
        uint32_t mma_phase;
        const int slot = mlps::acquire_slot<TEAMS, SLOTS>(ms, team, r, mma_phase);
        LGP_MARK(wait, lgp_t);
        mlps::write_features<TEAMS, SLOTS>(ms, slot, r, pend.my, pend.op, need);
        float y[12];
        mlps::forward<TEAMS, SLOTS>(ms, p.mlp_bias, team, slot, r, mma_phase, y);
        mlps::release_slot<TEAMS, SLOTS>(ms, team, r, slot, mma_phase);
        LGP_MARK(leaf, lgp_t);
#ifdef SYN_LG_PROF
        ++lgp.rounds;
#endif
        if (need) {
            float logit = 0.0f;
#pragma unroll
            for (int j = 0; j < 9; ++j)
                if (g.gl == j) logit = y[j];
            // value.softmax(-1) (study-connect4/src/policies.rs:54-56)
            const float m = fmaxf(y[9], fmaxf(y[10], y[11]));
            const float e0 = syn_expf(__fsub_rn(y[9], m)), e1 = syn_expf(__fsub_rn(y[10], m)), e2 = syn_expf(__fsub_rn(y[11], m));
            const float tot = __fadd_rn(__fadd_rn(e0, e1), e2);
            explore_finish(g, t, pend, false, logit, __fdiv_rn(e0, tot), __fdiv_rn(e1, tot), __fdiv_rn(e2, tot));
            after_eval(g, t, st);
        }
        LGP_MARK(fin, lgp_t);
    }
    flush_counters(g, p, t);
    LGP_FLUSH(g, p, t);
    mlps::teardown<TEAMS, SLOTS>(ms);
}

} // namespace eng
