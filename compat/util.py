"""Drop-in module name of the reference (`from util import ...` as in train.py:16): put this directory
on sys.path instead of the reference checkout and the batched B200 classes are picked up."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msc_thesis_b200.heli6dof import Logger, RefGenerator, envelope_limits_reached  # noqa: E402,F401
