"""Minimal restatement of the `airfoil_db` (>=1.4.3) API surface MachUpX calls.

Test infrastructure for the oracle harness only: lets the UNMODIFIED reference in
/root/reference import and run in this container.  The real package is a
third-party dependency (setup.py:13 of the reference) that is not vendored and
not installable here (no network).  The "linear" and "database" airfoil types are restated;
see SURVEY.md section 8(c) / Appendix C for how the formulas were pinned against
the reference's own golden tests (61 of 62 pass).
"""
from .exceptions import DatabaseBoundsError
from .airfoil import Airfoil
