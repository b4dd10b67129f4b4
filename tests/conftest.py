import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def c2():
    """docked.mol2 ligand + xtal receptor stand-in in the lds simulation box (configs C1/C2)."""
    from mmo_b200 import workloads
    return workloads.load_c2("docked")


@pytest.fixture(scope="session")
def c2_roi_rec(c2):
    """receptor atoms that can be within 12 A of any pose centred in the ROI"""
    from mmo_b200 import workloads
    rl = workloads.lig_radius(c2["centered"])
    return workloads.carve(c2["rec"], c2["roi"][:3], c2["roi"][3] + rl + 12.0)


@pytest.fixture(scope="session")
def gpu():
    """initialised library on cuda:0 -- fails loudly (no fallback) if it cannot start"""
    import mmo_b200
    mmo_b200.init(0)
    return mmo_b200


def tol_ok(e, ref):
    """north-star tolerance: 1e-6 relative or 1e-4 kcal/mol absolute, whichever is looser"""
    e = np.asarray(e, np.float64)
    ref = np.asarray(ref, np.float64)
    return np.abs(e - ref) <= np.maximum(1e-6 * np.abs(ref), 1e-4)
