"""Drop-in module name of the reference (`from train import ...` as in main.py:10): put this directory
on sys.path instead of the reference checkout and the batched B200 classes are picked up."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msc_thesis_b200.heli6dof import mp_test, mp_train, train  # noqa: E402,F401
