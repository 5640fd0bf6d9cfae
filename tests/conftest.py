import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.join(ROOT, "tests")
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)  # tests/ref_harness.py
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def objects_dir(tmp_path_factory, golden):
    """The reference's objects/*.obj assets re-materialised from tests/golden/meshes.npz."""
    import torch
    import ptk_b200
    d = tmp_path_factory.mktemp("objects")
    m = golden("meshes")
    for key, fname in [("vision", "vision_charts.obj"), ("touch", "touch_chart.obj"), ("obj0", "0.obj")]:
        ptk_b200.obj_io.save_obj(str(d / fname), torch.from_numpy(m[key + "_verts"]),
                                 torch.from_numpy(m[key + "_faces"].astype(np.int64)))
    ptk_b200.utils.set_object_dir(str(d))
    return str(d)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc
