//! Drop-in body for `gather_experience` (synthesis/src/alpha_zero.rs:120-169) over the C ABI.
//! Source only — shown in INTEGRATION.md; the build image has no Rust toolchain.
//!
//! Goes into `synthesis/src/alpha_zero.rs` behind `#[cfg(feature = "b200")]`, next to the CPU body.
//! The accelerated path is monomorphic in `Connect4` + `Connect4Net` weights: a user-defined
//! `Game`/`Policy` or an `Fpu::Func` is host code and cannot run inside the search kernel, so the
//! `Game` impl opts in through `B200Game` (implemented for `Connect4` in study-connect4).
use synthesis_b200_sys as sys;

/// What a `Game<N>` must expose for the device path (Connect4: its two bitboards, height, player).
pub trait B200Game<const N: usize>: Game<N> {
    fn from_b200_row(my_bb: u64, op_bb: u64, height: &[u8], player: u8) -> Self;
    fn features_from_row(states: &[f32]) -> Self::Features;
}

fn to_c(cfg: &RolloutConfig) -> sys::syn_rollout_cfg {
    let m = &cfg.mcts_cfg;
    let (exploration_kind, c) = match m.exploration {
        Exploration::Uct { c } => (sys::SYN_EXPLORATION_UCT, c),
        Exploration::PolynomialUct { c } => (sys::SYN_EXPLORATION_POLYNOMIAL_UCT, c),
    };
    let (fpu_kind, fpu_a) = match m.fpu {
        Fpu::Const(v) => (sys::SYN_FPU_CONST, v),
        Fpu::ParentQ => (sys::SYN_FPU_PARENT_Q, 0.0),
        Fpu::Func(_) => (sys::SYN_FPU_FUNC, 0.0), // rejected by the engine (SYN_ERR_UNSUPPORTED) -> panic below
    };
    let (noise_kind, noise_alpha, noise_weight) = match m.root_policy_noise {
        PolicyNoise::None => (sys::SYN_NOISE_NONE, 0.0, 0.0),
        PolicyNoise::Equal { weight } => (sys::SYN_NOISE_EQUAL, 0.0, weight),
        PolicyNoise::Dirichlet { alpha, weight } => (sys::SYN_NOISE_DIRICHLET, alpha, weight),
    };
    let (value_target_kind, vt_a, vt_b) = match cfg.value_target {
        ValueTarget::Z => (sys::SYN_VALUE_Z, 0.0, 0.0),
        ValueTarget::Q => (sys::SYN_VALUE_Q, 0.0, 0.0),
        ValueTarget::QZaverage { p } => (sys::SYN_VALUE_QZ_AVERAGE, p, 0.0),
        ValueTarget::QtoZ { from, to } => (sys::SYN_VALUE_Q_TO_Z, from, to),
    };
    sys::syn_rollout_cfg {
        num_explores: cfg.num_explores as u32,
        random_actions_until: cfg.random_actions_until as u32,
        sample_actions_until: cfg.sample_actions_until as u32,
        stop_games_when_solved: cfg.stop_games_when_solved as u8,
        _pad: [0; 3],
        value_target_kind, vt_a, vt_b,
        action_selection: match cfg.action { ActionSelection::Q => sys::SYN_ACTION_Q, ActionSelection::NumVisits => sys::SYN_ACTION_NUM_VISITS },
        mcts: sys::syn_mcts_cfg {
            exploration_kind, c,
            solve: m.solve as u8, correct_values_on_solve: m.correct_values_on_solve as u8,
            select_solved_nodes: m.select_solved_nodes as u8, auto_extend: m.auto_extend as u8,
            fpu_kind, fpu_a, fpu_b: 0.0, noise_kind, noise_alpha, noise_weight,
        },
        leaf_eval_kind: sys::SYN_LEAF_NN,
    }
}

fn check(rc: i32) {
    if rc != sys::SYN_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(sys::syn_last_error()) }.to_string_lossy().into_owned();
        panic!("synthesis_b200: {} ({})", msg, rc); // the reference unwrap()s here too: alpha_zero.rs:161,166,194
    }
}

/// Same signature and meaning as the reference's `gather_experience`; `weights` is the VarStore's
/// `l_1.weight, l_1.bias, ..., l_5.bias` flattened (what `vs.load(models/<policy_name>.ot)` yields).
pub fn gather_experience_b200<G: 'static + B200Game<N>, const N: usize>(
    engine: *mut sys::syn_engine,
    cfg: &LearningConfig,
    weights: &[f32],
    buffer: &mut ReplayBuffer<G, N>,
    seed: usize,
) {
    let n = cfg.games_per_train;
    let cap = G::MAX_TURNS * n;
    let (mut ids, mut my, mut op) = (vec![0u64; cap], vec![0u64; cap], vec![0u64; cap]);
    let (mut height, mut player) = (vec![0u8; cap * 9], vec![0u8; cap]);
    let (mut states, mut pis, mut vs) = (vec![0f32; cap * 63], vec![0f32; cap * N], vec![0f32; cap * 3]);
    let mut exp = sys::syn_experience {
        capacity: cap, len: 0, games: 0,
        game_ids: ids.as_mut_ptr(), my_bb: my.as_mut_ptr(), op_bb: op.as_mut_ptr(), height: height.as_mut_ptr(),
        player: player.as_mut_ptr(), states: states.as_mut_ptr(), pis: pis.as_mut_ptr(), vs: vs.as_mut_ptr(),
    };
    let ccfg = to_c(&cfg.rollout_cfg);
    unsafe {
        check(sys::syn_engine_set_weights(engine, weights.as_ptr(), weights.len()));
        check(sys::syn_engine_gather(engine, &ccfg, 0, n as u32, seed as u64, &mut exp, std::ptr::null_mut()));
    }
    // one worker buffer holding all n games in index order == the reference's workers joined in order
    let mut worker = ReplayBuffer::<G, N>::new(exp.len);
    let mut last = 0u64;
    for i in 0..exp.len {
        if ids[i] != last { worker.new_game(); last = ids[i]; }
        let game = G::from_b200_row(my[i], op[i], &height[9 * i..9 * i + 9], player[i]);
        let mut pi = [0f32; N];
        pi.copy_from_slice(&pis[N * i..N * i + N]);
        worker.add(&game, &pi, [vs[3 * i], vs[3 * i + 1], vs[3 * i + 2]]);
    }
    buffer.keep_last_n_games(cfg.games_to_keep - cfg.games_per_train); // alpha_zero.rs:164
    buffer.extend(&mut worker);                                        // alpha_zero.rs:165-168
}

/// `l_1.weight, l_1.bias, ..., l_5.weight, l_5.bias` of the trainer's VarStore as one blob of `SYN_N_WEIGHTS` floats — the
/// layout `syn_engine_set_weights` takes.  `VarStore::variables()` (tch) is a name -> Tensor map; Connect4Net names its
/// layers "l_1".."l_5" (study-connect4/src/policies.rs:28-45) and nn::linear stores [out, in] weights row-major.
pub fn flatten_weights(vs: &tch::nn::VarStore) -> Vec<f32> {
    let vars = vs.variables();
    let mut blob = Vec::with_capacity(sys::SYN_N_WEIGHTS);
    for layer in ["l_1", "l_2", "l_3", "l_4", "l_5"] {
        for part in ["weight", "bias"] {
            let t = vars.get(&format!("{}.{}", layer, part)).expect("Connect4Net variable");
            blob.extend(Vec::<f32>::from(t.to_kind(tch::Kind::Float).flatten(0, -1)));
        }
    }
    assert_eq!(blob.len(), sys::SYN_N_WEIGHTS);
    blob
}

/// The reference's `remaining / workers_left` split (alpha_zero.rs:134-141) with ranks in the place of worker threads:
/// rank r plays `count` games starting at global game index `first`.
pub fn shard_of(games_per_train: usize, n_ranks: usize, rank: usize) -> (u64, u32) {
    let (mut first, mut left) = (0usize, games_per_train);
    for r in 0..n_ranks {
        let n = left / (n_ranks - r);
        if r == rank {
            return (first as u64, n as u32);
        }
        first += n;
        left -= n;
    }
    unreachable!()
}

/// `gather_experience` across GPUs: one process (or thread) per GPU, each with its own engine and a `syn_comm` built from
/// the 128-byte id of `syn_comm_unique_id` (rank 0 creates it; any transport carries it to the others).  Every rank calls
/// this; only `root` touches `buffer`.  Two collectives per iteration, both inside the library: the weight blob out, the
/// 72-byte experience rows in, in rank order == the order `extend` would have produced in worker order.
pub fn gather_experience_b200_ranks<G: 'static + B200Game<N>, const N: usize>(
    engine: *mut sys::syn_engine,
    comm: *mut sys::syn_comm,
    root: i32,
    cfg: &LearningConfig,
    weights: Option<&[f32]>, // Some(..) on root after a tch optimiser step; None = the root engine's own (syn_engine_train)
    buffer: Option<&mut ReplayBuffer<G, N>>,
    seed: usize,
) {
    let (rank, n_ranks) = unsafe { (sys::syn_comm_rank(comm), sys::syn_comm_size(comm)) };
    let (first, count) = shard_of(cfg.games_per_train, n_ranks as usize, rank as usize);
    let ccfg = to_c(&cfg.rollout_cfg);
    let (ptr, len) = weights.map_or((std::ptr::null(), sys::SYN_N_WEIGHTS), |w| (w.as_ptr(), w.len()));
    unsafe { check(sys::syn_engine_broadcast_weights(engine, comm, ptr, len, root)) };
    if rank != root {
        unsafe { check(sys::syn_engine_gather_experience(engine, comm, root, &ccfg, first, count, seed as u64, std::ptr::null_mut(), std::ptr::null_mut())) };
        return;
    }
    let cap = G::MAX_TURNS * cfg.games_per_train;
    let (mut ids, mut my, mut op) = (vec![0u64; cap], vec![0u64; cap], vec![0u64; cap]);
    let (mut height, mut player) = (vec![0u8; cap * 9], vec![0u8; cap]);
    let (mut states, mut pis, mut vs) = (vec![0f32; cap * 63], vec![0f32; cap * N], vec![0f32; cap * 3]);
    let mut exp = sys::syn_experience {
        capacity: cap, len: 0, games: 0,
        game_ids: ids.as_mut_ptr(), my_bb: my.as_mut_ptr(), op_bb: op.as_mut_ptr(), height: height.as_mut_ptr(),
        player: player.as_mut_ptr(), states: states.as_mut_ptr(), pis: pis.as_mut_ptr(), vs: vs.as_mut_ptr(),
    };
    unsafe { check(sys::syn_engine_gather_experience(engine, comm, root, &ccfg, first, count, seed as u64, &mut exp, std::ptr::null_mut())) };
    let buffer = buffer.expect("the root rank owns the replay buffer");
    let mut joined = ReplayBuffer::<G, N>::new(exp.len);
    let mut last = u64::MAX;
    for i in 0..exp.len {
        if ids[i] != last { joined.new_game(); last = ids[i]; }
        let game = G::from_b200_row(my[i], op[i], &height[9 * i..9 * i + 9], player[i]);
        let mut pi = [0f32; N];
        pi.copy_from_slice(&pis[N * i..N * i + N]);
        joined.add(&game, &pi, [vs[3 * i], vs[3 * i + 1], vs[3 * i + 2]]);
    }
    buffer.keep_last_n_games(cfg.games_to_keep - cfg.games_per_train);
    buffer.extend(&mut joined);
}
