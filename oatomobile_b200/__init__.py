"""oatomobile_b200 — B200-native (sm_100a) drop-in for the OATomobile RIP/DIM hot path.

Public API mirrors `oatomobile.baselines.torch` (reference
oatomobile/baselines/torch/__init__.py:17-21).
"""
from oatomobile_b200.agents import CILAgent, DIMAgent, RIPAgent
from oatomobile_b200.models import BehaviouralModel, ImitativeModel
from oatomobile_b200.networks import MLP, AutoregressiveFlow, MobileNetV2
from oatomobile_b200.rip import HostRIPPipeline, RIPScorer

__all__ = ["ImitativeModel", "BehaviouralModel", "RIPAgent", "DIMAgent", "CILAgent",
           "AutoregressiveFlow", "MobileNetV2", "MLP", "RIPScorer", "HostRIPPipeline"]
