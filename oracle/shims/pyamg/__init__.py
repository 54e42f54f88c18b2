"""Stand-in for the un-vendored third-party `pyamg` (pinned 5.0.1 in the reference's
requirements.txt; not installable offline).  TEST INFRASTRUCTURE ONLY -- lets
/root/reference run so goldens can be generated (oracle/gen_golden.py).

Only what the hot path calls is restated:
  pyamg.amg_core.block_gauss_seidel      <- dgfem/pyamg_relaxation.py:253
  pyamg.util.utils.get_block_diag        <- dgfem/pyamg_relaxation.py:231
"parity unpinned" against a real pyamg binary; see oracle/dgoracle/__init__.py.
"""
from . import amg_core  # noqa: F401
from . import util  # noqa: F401

__version__ = "5.0.1-restated"


def ruge_stuben_solver(*args, **kwargs):
    raise NotImplementedError("classical AMG is out of scope (SURVEY.md section 2.2)")
