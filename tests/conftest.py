import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; never imported by the product package)."""
    from oracle import reference_ops
    return reference_ops


@pytest.fixture(scope="session")
def reference_modules():
    """The unmodified reference: /root/reference in the build container, the byte-identical
    copy staged by ``oracle/stage_reference.py`` into the git-ignored ``oracle/_ref/`` on the GPU box."""
    from oracle import reference_ops  # noqa: F401  (puts oracle/shim on sys.path)
    from oracle import stage_reference
    if os.path.isdir(REFERENCE):
        if REFERENCE not in sys.path:
            sys.path.insert(0, REFERENCE)
        import modules
        import image_model
        import video_model
        return modules, image_model, video_model
    if not stage_reference.available():
        pytest.skip("reference neither mounted nor staged under oracle/_ref")
    return stage_reference.import_reference()
