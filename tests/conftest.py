import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "depth-vo-feat_b200")
for p in (REPO, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    # DVF_TEST_ORDER=reverse|shuffle:<seed> -- the kernels keep ticket counters in cached workspaces and shared memory
    # is not cleared between launches, so the suite must pass in any order (it caught a stale-NaN bug once)
    order = os.environ.get("DVF_TEST_ORDER", "")
    if order == "reverse":
        items.reverse()
    elif order.startswith("shuffle"):
        import random
        random.Random(int(order.split(":")[1]) if ":" in order else 0).shuffle(items)
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_oracle
    cpu_oracle.build()
    return cpu_oracle
