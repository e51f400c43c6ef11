"""Drop-in module name of the reference (`from PID import ...` as in main_3dof.py:8-12 / train.py:15): put this directory
on sys.path instead of the reference checkout and the batched B200 classes are picked up."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msc_thesis_b200 import CollectivePID3DOF  # noqa: E402,F401
from msc_thesis_b200.heli6dof import CollectivePID6DOF, LatPedPID  # noqa: E402,F401
