import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def expected():
    with open(os.path.join(GOLDEN, "expected.json")) as f:
        return json.load(f)


def golden_mesh_path(name):
    return os.path.join(GOLDEN, "meshes", name)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def load_osh_oracle(name):
    from oracle import osh_oracle as oo
    m = oo.read_osh(golden_mesh_path(name))
    return oo.elem_verts(m), oo.coords(m), m


def synthetic_meshes():
    """name -> (cells, coords): the small parity inputs shared by CPU and GPU tests."""
    import oracle
    out = {}
    out["one_tri"] = (np.array([[0, 1, 2]], np.int32), np.array([[0, 0], [1, 0], [0, 1]], float))
    out["fan4"] = (np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], np.int32),
                   np.array([[0, 0], [1, 0], [1, 1], [0, 1], [0.4, 0.55]], float))
    out["box2d_8x8"] = oracle.box2d(8, 8)
    out["box2d_33x17"] = oracle.box2d(33, 17)
    out["box2d_100x70"] = oracle.box2d(100, 70)
    out["box3d_5x4x3"] = oracle.box3d(5, 4, 3)
    out["box3d_12x11x10"] = oracle.box3d(12, 11, 10)
    # a sheared, non-uniform 2-D mesh: no right angles, no exact zeros
    c, x = oracle.box2d(21, 13)
    x = x.copy()
    x[:, 0] += 0.31 * x[:, 1] ** 2 + 0.05 * np.sin(7 * x[:, 1])
    x[:, 1] += 0.17 * x[:, 0]
    out["sheared_21x13"] = (c, x)
    c, x = oracle.box3d(6, 5, 7)
    x = x.copy()
    x[:, 0] += 0.2 * x[:, 2] * x[:, 1]
    x[:, 2] += 0.1 * np.sin(3 * x[:, 0])
    out["warped3d_6x5x7"] = (c, x)
    return out
