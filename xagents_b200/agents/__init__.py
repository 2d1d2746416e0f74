"""Drop-in mirrors of the reference's on-policy agents: same constructors, same method surface,
arithmetic on the B200 through the C ABI."""
from .a2c import A2C
from .acer import ACER
from .base import BaseAgent, EnvMajorView, OnPolicy
from .cfg import CfgNetwork, ModelReader
from .models import AsCodedConv1dCNN, KerasModel, NatureCNN, TorchModel, adapt
from .ppo import PPO
from .tc_cnn import NatureCnnTc
from .trpo import TRPO

__all__ = ['A2C', 'ACER', 'PPO', 'TRPO', 'BaseAgent', 'OnPolicy', 'EnvMajorView', 'TorchModel', 'KerasModel', 'NatureCNN', 'NatureCnnTc', 'AsCodedConv1dCNN', 'adapt',
           'ModelReader', 'CfgNetwork']
