import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["prod_smallfc_mixed", "prod_smallfc_edge", "prod_full", "min_smallfc", "variant_a", "reads300_ragged"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    from dl4vc_b200.config import DanConfig

    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    d = json.loads(str(g["config"]))
    d["conv_1d_pool_layers"] = tuple(d["conv_1d_pool_layers"])
    d["layer_sizes"] = tuple(d["layer_sizes"])
    g["cfg"] = DanConfig(**d)
    g["seed"] = int(g["seed"])
    g["arrays"] = tuple(g[k] for k in ("reads", "q_scores", "strands", "ref", "ref_masks", "var_masks"))
    return g


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)


def rel_err(a, b):
    """|a-b| / max(|b|, max|b| over the batch) — the logit tolerance definition of SURVEY 7.2-4a."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), np.abs(b).max())))
