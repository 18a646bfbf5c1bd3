"""Kept for the fixture generator's import path: the loader of the unmodified reference lives in oracle/refload.py
(test infrastructure shared with bench.py's reference legs)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.refload import *  # noqa: F401,F403,E402
from oracle.refload import REF, REF_SRC, load_reference, load_primal  # noqa: F401,E402
