"""CPU oracle for the Vision Workbench stereo-correlation hot path.

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs -- never from the product
package ``visionworkbench_b200``.  See oracle/vw_oracle.h for what each function
restates (reference file:line) and how the oracle is pinned.
"""
from .binding import *  # noqa: F401,F403
