"""xlab_ee_fortran_b200 - B200-native (sm_100a) drop-in for the elliptic-solve hot path of
meteorologytoday/XLab-EE-fortran.  The product is the C-ABI library (include/xee_b200.h,
csrc/*.cu); this package holds the host-side mirror of the reference interface for that path.
"""
from . import _lib  # noqa: F401
from .elliptic_tools import (cal_coe, do_elliptic, err_explode, err_over_max_iteration, judge_error,  # noqa: F401
                             solve_elliptic)
from .plan import Plan, SolveParams  # noqa: F401
