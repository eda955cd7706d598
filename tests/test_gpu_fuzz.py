"""A short run of tests/fuzz_vs_oracle.py inside the GPU suite: randomised engine-vs-oracle searches (domains, A* / Q*, batch sizes,
instances added / removed mid-search, tie-heavy and sub-cut heuristics, dx_run bursts, OPEN / sort switches), every iteration compared."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [3, 4])
def test_randomised_searches_match_the_oracle(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_vs_oracle.py"), "6", str(seed)], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    tail = "\n".join(r.stdout.splitlines()[-3:])
    assert r.returncode == 0 and "all equal" in tail, tail + r.stderr[-2000:]
