"""`from sys_params import SYS_PARAMS` as in run.py:4 / control.py:9 of the reference."""
from mppi_robotarm_b200.arm_params import SYS_PARAMS  # noqa: F401
