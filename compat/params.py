"""Drop-in module name of the reference (`from params import ...` as in train.py:268-269): put this directory
on sys.path instead of the reference checkout and the batched B200 classes are picked up."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msc_thesis_b200.heli6dof import (ac_params_test, ac_params_train, env_params_test1, env_params_test2,  # noqa: E402,F401
                                      env_params_train, pid_params, rls_params)
