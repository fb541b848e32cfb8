"""sus_net_b200 -- B200-native batched simulator + observation encoder for Sus-Net's Among-Us grid environment.

Drop-in for the reference's hot path (src/environment + src/features) behind its own class and method names; the
compute is hand-written sm_100a CUDA behind a C ABI (include/susnet_b200.h).  There is no CPU fallback: the package
needs the built `libsusnet_b200.so` (python -m sus_net_b200.build) and a CUDA device to run anything.
"""
from ._lib import LIB_PATH, SusNetError, lib  # noqa: F401
from .env import (  # noqa: F401
    Action,
    BatchedFourRoomEnv,
    BatchedFourRoomEnvWithTagging,
    BatchedImposterTrainingGround,
    StateFields,
)
from .featurizers import (  # noqa: F401
    AgentPositionsFeaturizer,
    AliveCrewFeaturizer,
    ClosestAliveCrewFeaturizer,
    CompositeFeaturizer,
    CoordinateAgentPositionsFeaturizer,
    DistanceToImposterFeaturizer,
    FeaturizerType,
    FlatFeaturizer,
    GlobalFeaturizer,
    ImposterScentFeaturizer,
    ImposterVSCrewRoomLocaionFeaturizer,
    JobFeaturizer,
    L1CrewFeaturizer,
    OneHotAgentPositionFeaturizer,
    PerspectiveFeaturizer,
    StateFieldFeaturizer,
    WallsFeaturizer,
)
from .metrics import METRIC_ORDER, STAT_KEYS, EpisodicMetricHandler, SusMetrics  # noqa: F401
from .distributed import reduce_episode_stats, reduce_return_sums, shard_range  # noqa: F401
from .host_pipeline import HostStepper  # noqa: F401
from .replay_memory import Batch, ReplayBuffer  # noqa: F401
from .train import (  # noqa: F401
    BatchedActor,
    BatchedTrainingLoop,
    DQNTeamTrainer,
    ExponentialSchedule,
    FeatureSequence,
    allreduce_grads,
    train_batched,
)
from .compact import CompactProtocol  # noqa: F401
from .mlp import FusedMLP  # noqa: F401

# reference class names, for `from sus_net_b200 import FourRoomEnv` style drop-in use
FourRoomEnv = BatchedFourRoomEnv
FourRoomEnvWithTagging = BatchedFourRoomEnvWithTagging
ImposterTrainingGround = BatchedImposterTrainingGround
