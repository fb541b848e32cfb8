"""Shared helpers for the parity tests: golden fixture loading and env adaptors."""
import glob
import os

import numpy as np

from tests.cases import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files(kind):
    return sorted(glob.glob(os.path.join(GOLDEN, f"*.{kind}.npz")))


def load(path):
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def case_of(path):
    return os.path.basename(path).split(".")[0]


def reward_bits(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.int64)


def make_cuda_env(cfg, num_envs, seed, env_id_base=0, auto_reset=True, batched=True, device="cuda"):
    """Build the sus_net_b200 env for a tests.cases config dict."""
    import sus_net_b200 as S

    common = dict(num_envs=num_envs, seed=seed, env_id_base=env_id_base, auto_reset=auto_reset, batched=batched,
                  device=device)
    v = cfg["variant"]
    if v == "training_ground":
        return S.BatchedImposterTrainingGround(
            n_crew=cfg["n_crew"], n_jobs=cfg["n_jobs"], time_step_reward=cfg["time_step_reward"],
            kill_reward=cfg["kill_reward"], sabotage_reward=cfg["sabotage_reward"],
            end_of_game_reward=cfg["game_end_reward"], shuffle_imposter_index=cfg["shuffle_imposter_index"],
            include_walls=cfg["include_walls"], **common)
    kw = dict(n_imposters=cfg["n_imposters"], n_crew=cfg["n_crew"], n_jobs=cfg["n_jobs"],
              is_action_order_random=cfg["is_action_order_random"], kill_reward=cfg["kill_reward"],
              complete_job_reward=cfg["complete_job_reward"], sabotage_reward=cfg["sabotage_reward"],
              time_step_reward=cfg["time_step_reward"], game_end_reward=cfg["game_end_reward"],
              dead_penalty=cfg["dead_penalty"], shuffle_imposter_index=cfg["shuffle_imposter_index"],
              max_time_steps=cfg["max_time_steps"], include_walls=cfg["include_walls"], **common)
    if v == "tagging":
        return S.BatchedFourRoomEnvWithTagging(**kw, tag_reset_interval=cfg["tag_reset_interval"],
                                               vote_reward=cfg["vote_reward"])
    return S.BatchedFourRoomEnv(**kw)


def flat_featurizer(env, components):
    import sus_net_b200 as S

    classes = dict(onehot_pos=S.OneHotAgentPositionFeaturizer, coords=S.CoordinateAgentPositionsFeaturizer,
                   alive_crew=S.AliveCrewFeaturizer, closest_crew=S.ClosestAliveCrewFeaturizer,
                   l1_crew=S.L1CrewFeaturizer, dist_to_imposter=S.DistanceToImposterFeaturizer,
                   walls=S.WallsFeaturizer, rooms=S.ImposterVSCrewRoomLocaionFeaturizer, scent=S.ImposterScentFeaturizer)
    fields = dict(state_alive=S.StateFields.ALIVE_AGENTS, state_job_status=S.StateFields.JOB_STATUS,
                  state_used_tags=S.StateFields.USED_TAGS, state_tag_counts=S.StateFields.TAG_COUNTS)
    parts = [classes[c](env) if c in classes else S.StateFieldFeaturizer(env, fields[c]) for c in components]
    return S.FlatFeaturizer(env, S.CompositeFeaturizer(parts))


__all__ = ["CASES", "golden_files", "load", "case_of", "reward_bits", "make_cuda_env", "flat_featurizer"]
