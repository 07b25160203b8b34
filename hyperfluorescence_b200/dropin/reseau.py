"""Drop-in shim: put this directory on PYTHONPATH and the reference's scripts
(`from reseau import ...`) run unmodified on the GPU backend."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from hyperfluorescence_b200.reseau import *  # noqa: E402,F401,F403
from hyperfluorescence_b200.reseau import Lattice  # noqa: E402,F401
