import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


# The thread-per-cluster stage is only used for large batches by default (AVK_THREAD_MIN_REGIONS); the test batches are small,
# so the suite lowers the threshold and every compare test goes through it.  test_warp_only_pipeline_vs_oracle covers the
# pipeline without it.
import os
os.environ.setdefault("AVK_THREAD_MIN_REGIONS", "0")
