from .device import DeviceRogueEnv
from .parallel import ParallelRogueEnv
from .rogue_env import DungeonType, ImageSetting, PlayerState, RogueEnv, StatusFlag
from .wrappers import FirstFloorEnv, StairRewardEnv, StairRewardParallel

__all__ = ["DeviceRogueEnv", "DungeonType", "FirstFloorEnv", "ImageSetting", "ParallelRogueEnv", "PlayerState",
           "RogueEnv", "StairRewardEnv", "StairRewardParallel", "StatusFlag"]
