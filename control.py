"""Drop-in replacement for the reference's control.py: `from control import MPPIControllerForPathTracking`
(run.py:6) resolves to the B200 implementation.  See mppi_robotarm_b200/controller.py."""
from mppi_robotarm_b200.controller import MPPIControllerForPathTracking  # noqa: F401
