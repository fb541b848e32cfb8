"""Canonical env configurations (BASELINE.json `configs`) plus edge cases, shared by the
golden generator and the parity tests."""
from oracle import default_config

CASES = {
    # cfg1: ImposterTrainingGround 1v1, empty grid (notebooks/experiment_1v1.ipynb "No Wall")
    "cfg1_itg_1v1_nowall": default_config("training_ground", n_crew=1, n_jobs=0, kill_reward=-3.0, sabotage_reward=0.0,
                                          game_end_reward=0.0, time_step_reward=0.0, include_walls=False),
    # cfg2: same, walled grid
    "cfg2_itg_1v1_wall": default_config("training_ground", n_crew=1, n_jobs=0, kill_reward=-3.0, sabotage_reward=0.0,
                                        game_end_reward=0.0, time_step_reward=0.0),
    # cfg3: tagging 1v2, 5 jobs, all defaults
    "cfg3_tagging_1v2": default_config("tagging", n_crew=2, n_jobs=5),
    # cfg4 (headline): FourRoomEnv 1v4, 5 jobs, all defaults
    "cfg4_base_1v4": default_config("base", n_crew=4, n_jobs=5),
    # cfg4-alt: ImposterTrainingGround 1v4 walled
    "cfg4alt_itg_1v4": default_config("training_ground", n_crew=4, n_jobs=0, kill_reward=-3.0, sabotage_reward=0.0,
                                      game_end_reward=0.0, time_step_reward=0.0),
    # edge cases
    "base_2v3_j3": default_config("base", n_imposters=2, n_crew=3, n_jobs=3, max_time_steps=60),
    "base_1v2_j0": default_config("base", n_crew=2, n_jobs=0),  # crew wins on step 1 (quirk C-4)
    "base_fixed_order_tsr": default_config("base", n_crew=3, n_jobs=2, is_action_order_random=False,
                                           shuffle_imposter_index=False, time_step_reward=-1.0, max_time_steps=40),
    "tagging_2v5_short": default_config("tagging", n_imposters=2, n_crew=5, n_jobs=4, tag_reset_interval=7,
                                        max_time_steps=50, time_step_reward=-1.0),
    "tagging_1v4": default_config("tagging", n_crew=4, n_jobs=5),
    "itg_1v3_jobs_shuffle": default_config("training_ground", n_crew=3, n_jobs=2, kill_reward=-3.0,
                                           sabotage_reward=0.0, game_end_reward=5.0, time_step_reward=-0.5,
                                           shuffle_imposter_index=True),
    "base_1v7_j8_nowall": default_config("base", n_crew=7, n_jobs=8, include_walls=False, max_time_steps=30),
}

# featurizer coverage per case: which model-ready featurizers are defined for it
GLOBAL_CASES = ["cfg4_base_1v4", "base_2v3_j3", "base_fixed_order_tsr", "base_1v7_j8_nowall", "itg_1v3_jobs_shuffle"]
# sets with "scent" (the one float-valued component) take the float-row paths; the scent-free ones exercise the byte-staged
# rows of k_step_flat / k_encode_flat on every integer-valued component (negative values, tagging fields, two one-hot
# segments in one row)
FLAT_COMPONENT_SETS = {
    "cfg4alt_itg_1v4": [["onehot_pos", "alive_crew", "closest_crew"],
                        ["coords", "l1_crew", "dist_to_imposter", "walls", "rooms", "scent", "state_alive"],
                        ["coords", "l1_crew", "dist_to_imposter", "walls", "rooms", "state_alive", "onehot_pos"]],
    "cfg2_itg_1v1_wall": [["onehot_pos"], ["coords"]],
    "cfg1_itg_1v1_nowall": [["onehot_pos"], ["walls", "rooms"]],
    "cfg4_base_1v4": [["onehot_pos", "state_alive", "state_job_status", "walls", "rooms", "l1_crew", "closest_crew",
                       "dist_to_imposter", "scent", "coords", "alive_crew"],
                      ["onehot_pos", "state_alive", "state_job_status", "walls", "rooms", "l1_crew", "closest_crew",
                       "dist_to_imposter", "coords", "alive_crew", "onehot_pos"]],
}
# CUDA vs oracle only: the reference's own featurizers raise on the tagging env (its state_fields map is inconsistent,
# SURVEY.md App. C-7), so there is nothing to pin these against
FLAT_COMPONENT_SETS_TAGGING = {
    "cfg3_tagging_1v2": [["state_used_tags", "state_tag_counts", "onehot_pos", "state_job_status", "dist_to_imposter"]],
}


def random_case(rng):
    """A random but valid constructor-argument set (any variant, sizes up to the build's limits, integer AND
    non-integer reward constants, short episodes / vote windows so that every branch fires often)."""
    variant = ("base", "tagging", "training_ground")[int(rng.integers(0, 3))]
    n_imp = 1 if variant == "training_ground" else int(rng.integers(1, 4))
    n_crew = int(rng.integers(max(1, n_imp + (variant != "training_ground")), 9 - n_imp))
    n_jobs = int(rng.integers(1 if variant == "tagging" else 0, 9))

    def reward():
        v = float(rng.integers(-6, 7))
        return v if rng.random() < 0.5 else v + float(rng.choice([0.1, 0.25, -0.3, 1.0 / 3.0]))

    cfg = default_config(
        variant, n_imposters=n_imp, n_crew=n_crew, n_jobs=n_jobs, include_walls=bool(rng.integers(0, 2)),
        is_action_order_random=bool(rng.integers(0, 2)), shuffle_imposter_index=bool(rng.integers(0, 2)),
        max_time_steps=int(rng.integers(5, 80)), tag_reset_interval=int(rng.integers(1, 12)),
        kill_reward=reward(), complete_job_reward=reward(), sabotage_reward=reward(), time_step_reward=reward(),
        game_end_reward=reward(), dead_penalty=reward(), vote_reward=reward())
    if variant == "training_ground":  # pred_prey.py:52-66 fixes these
        cfg.update(dead_penalty=0.0, is_action_order_random=False, complete_job_reward=3.0, max_time_steps=1000)
    return cfg


# extremes of the constructor-argument space (checked oracle-vs-reference by tools/check_oracle_vs_reference.py --edge
# and CUDA-vs-oracle by tests/test_gpu_parity.py)
EDGE_CASES = {
    "edge_max_time_steps_1": default_config("base", n_crew=2, n_jobs=1, max_time_steps=1),
    "edge_max_time_steps_2_tagging": default_config("tagging", n_crew=3, n_jobs=1, max_time_steps=2, tag_reset_interval=1),
    "edge_tag_interval_1": default_config("tagging", n_imposters=3, n_crew=4, n_jobs=2, tag_reset_interval=1, max_time_steps=30),
    "edge_largest_8_agents_8_jobs_tagging": default_config("tagging", n_imposters=3, n_crew=5, n_jobs=8, tag_reset_interval=3,
                                                           max_time_steps=40, include_walls=False),
    "edge_all_zero_rewards": default_config("base", n_crew=3, n_jobs=2, kill_reward=0.0, complete_job_reward=0.0,
                                            sabotage_reward=0.0, game_end_reward=0.0, dead_penalty=0.0, time_step_reward=0.0,
                                            max_time_steps=25),
    "edge_negative_zero_tagging": default_config("tagging", n_crew=2, n_jobs=1, kill_reward=0.0, complete_job_reward=0.0,
                                                 sabotage_reward=0.0, game_end_reward=0.0, dead_penalty=0.0,
                                                 time_step_reward=0.0, vote_reward=0.0, tag_reset_interval=2, max_time_steps=20),
    "edge_itg_1v7": default_config("training_ground", n_crew=7, n_jobs=8, kill_reward=-3.0, sabotage_reward=0.0,
                                   game_end_reward=2.5, time_step_reward=-0.125, shuffle_imposter_index=True),
    "edge_single_job_single_crew_itg": default_config("training_ground", n_crew=1, n_jobs=1, kill_reward=1.5,
                                                      sabotage_reward=0.0, game_end_reward=-4.0, time_step_reward=0.0,
                                                      include_walls=False),
}
