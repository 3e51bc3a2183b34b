"""Re-export of the repo-root `synthdata` module (deterministic synthetic inputs and weights).

The generator is shared by the oracle, the golden scripts, the tests and `bench.py`; it lives outside `oracle/` so that the
GPU arm of the benchmark and the developer scripts under `tools/` do not import test infrastructure."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthdata import *  # noqa: F401,F403,E402
from synthdata import MODEL_CFG  # noqa: F401,E402
