"""heli3dof-b200: batched, B200-native (sm_100a) closed-loop control tick of BartHelder/msc-thesis.

The directory name contains a hyphen; import it as `msc_thesis_b200` (alias module at the repo root) or
with importlib.import_module("msc-thesis_b200").

Reference-API classes (batched along a trailing episode axis):
    Helicopter3DOF          /root/reference/heli_models.py:109-344
    CollectivePID3DOF       /root/reference/PID.py:4-24
    RecursiveLeastSquares   /root/reference/model.py:5-96
    DHPAgent                /root/reference/agents.py:66-209
Fused fast path:
    ClosedLoop3DOF          /root/reference/main_3dof.py:122-197
"""
from . import _lib
from ._lib import H3Error, lib
from .PID import CollectivePID3DOF
from .agents import DHPActor, DHPAgent, DHPCritic
from .closed_loop import TRAJ_COLUMNS, ClosedLoop3DOF, sinusoid_ref_table
from .distributed import allreduce_stats, shard_range
from .excitation import excitation_signal, load_excitation
from .heli_models import Helicopter3DOF
from .host_batch import HostBatchPipeline, loop_result_tensors, loop_state_tensors
from .model import RecursiveLeastSquares
from . import episodes
from .episodes import EpisodeRunner
from . import sweep
from . import heli6dof

__all__ = ["Helicopter3DOF", "CollectivePID3DOF", "RecursiveLeastSquares", "DHPAgent", "DHPActor", "DHPCritic",
           "ClosedLoop3DOF", "TRAJ_COLUMNS", "sinusoid_ref_table", "excitation_signal", "load_excitation",
           "shard_range", "allreduce_stats", "H3Error", "lib", "EpisodeRunner", "episodes"]
