"""Drop-in module name of the reference (`from heli_models import ...` as in main_3dof.py:8-12 / train.py:14): put this directory
on sys.path instead of the reference checkout and the batched B200 classes are picked up."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msc_thesis_b200 import Helicopter3DOF  # noqa: E402,F401
from msc_thesis_b200.heli6dof import Helicopter6DOF  # noqa: E402,F401
