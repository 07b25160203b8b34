import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.basename(p)[len("traj_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "traj_*.npz")))


def load_golden(name):
    """A golden trajectory (dict of arrays) with its lattice arrays merged in."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"traj_{name}.npz")))
    lat = np.load(os.path.join(GOLDEN_DIR, str(g["lattice"]) + ".npz"))
    g.update({k: lat[k] for k in lat.files})
    cfg = g["config"]
    g["dims"] = (int(cfg[0]), int(cfg[1]), int(cfg[2]))
    g["charges"], g["distance"], g["n_steps"] = int(cfg[3]), int(cfg[4]), int(cfg[5])
    g["trajectory"], g["realisation"] = int(cfg[6]), int(cfg[7])
    g["seed_int"] = int(g["seed"])
    g["field_f"] = float(g["field"])
    g["error_s"] = str(g["error"])
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)
