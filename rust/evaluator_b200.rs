// evaluator_b200.rs — what goes into synthesis/src/evaluator.rs behind `#[cfg(feature = "b200")]`.
//
// The reference plays its evaluation games one after another on the evaluator thread
// (evaluator.rs:52-82: for every opponent strength and seed, once per colour).  With the engine the
// whole sweep is two calls of `syn_engine_match` (policy as Red, policy as Black); the PGN records are
// written by the reference's own `add_pgn_result` in the reference's order, so bayeselo and
// plot_ratings.py see the same file format.  (No Rust toolchain exists in this repository's build
// image; this file is kept as source.  The Python mirror that the tests drive is
// synthesis_b200/evaluator.py.)
use synthesis_b200_sys as ffi;

fn mcts_cfg_to_c(c: &MCTSConfig) -> ffi::syn_mcts_cfg { /* as in gather_experience_b200.rs */ unimplemented!() }

fn player(tree_kind: u32, leaf: u32, explores: usize, action: ActionSelection, cfg: &MCTSConfig) -> ffi::syn_player_cfg {
    ffi::syn_player_cfg {
        tree_kind,
        leaf_eval_kind: leaf,
        num_explores: explores as u32,
        action_selection: match action { ActionSelection::Q => 0, ActionSelection::NumVisits => 1 },
        mcts: mcts_cfg_to_c(cfg),
    }
}

/// Replaces the two nested loops of evaluator.rs:52-82.  `weights` = l_1.weight .. l_5.bias of model_{i}.ot.
pub fn eval_against_rollout_sweep_b200(
    engine: *mut ffi::syn_engine,
    cfg: &EvaluationConfig,
    weights: &[f32],
    name: &String,
    pgn: &mut std::fs::File,
) -> std::io::Result<()> {
    unsafe { assert_eq!(ffi::syn_engine_set_weights(engine, weights.as_ptr(), weights.len()), 0) };
    let nn = player(0, 0, cfg.policy_num_explores, cfg.policy_action, &cfg.policy_mcts_cfg);
    let max_ro = *cfg.rollout_num_explores.iter().max().unwrap();
    let ro = player(1, 1, max_ro, cfg.rollout_action, &cfg.rollout_mcts_cfg);
    // one match per (explores, seed), the reference's iteration order
    let mut seeds = Vec::new();
    let mut opp = Vec::new();
    for &explores in cfg.rollout_num_explores.iter() {
        for seed in 0..cfg.num_games_against_rollout {
            seeds.push(seed as u64);
            opp.push(explores as u32);
        }
    }
    let n = seeds.len();
    let mut results = [vec![0f32; n], vec![0f32; n]];
    for (side, players) in [[nn, ro], [ro, nn]].iter().enumerate() {
        // explores[i] = [first mover's, second mover's]
        let ex: Vec<u32> = opp.iter().flat_map(|&o| if side == 0 { [cfg.policy_num_explores as u32, o] } else { [o, cfg.policy_num_explores as u32] }).collect();
        let rc = unsafe {
            ffi::syn_engine_match(engine, players.as_ptr(), seeds.as_ptr(), ex.as_ptr(), n as u32, results[side].as_mut_ptr(),
                                  std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut())
        };
        assert_eq!(rc, 0, "syn_engine_match failed"); // the reference would have panicked inside exploit()
    }
    for i in 0..n {
        let op_name = format!("VanillaMCTS{}", opp[i]);
        add_pgn_result(pgn, name, &op_name, results[0][i])?; // evaluator.rs:64: policy plays first
        add_pgn_result(pgn, &op_name, name, results[1][i])?; // evaluator.rs:72: rollout MCTS plays first
    }
    Ok(())
}

/// Replaces evaluator.rs:87-94 + eval_against_old (:131-161): the new model against one of the best older models,
/// once as each colour.  Both networks stay resident in the match kernel; the game is deterministic, one match each.
pub fn eval_against_old_b200(
    engine: *mut ffi::syn_engine,
    cfg: &EvaluationConfig,
    name: &String,
    weights: &[f32],
    prev_name: &String,
    prev_weights: &[f32],
    pgn: &mut std::fs::File,
) -> std::io::Result<()> {
    let nn = player(0, 0, cfg.policy_num_explores, cfg.policy_action, &cfg.policy_mcts_cfg);
    let players = [nn, nn];
    let seed = [0u64];
    for (first, second, white, black) in [(weights, prev_weights, name, prev_name), (prev_weights, weights, prev_name, name)].iter() {
        let mut result = [0f32];
        unsafe {
            assert_eq!(ffi::syn_engine_set_weights(engine, first.as_ptr(), first.len()), 0);
            assert_eq!(ffi::syn_engine_set_opponent_weights(engine, second.as_ptr(), second.len()), 0);
            let rc = ffi::syn_engine_match(engine, players.as_ptr(), seed.as_ptr(), std::ptr::null(), 1, result.as_mut_ptr(),
                                           std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
            assert_eq!(rc, 0, "syn_engine_match failed");
            assert_eq!(ffi::syn_engine_set_opponent_weights(engine, std::ptr::null(), 0), 0);
        }
        add_pgn_result(pgn, white, black, result[0])?;
    }
    Ok(())
}
