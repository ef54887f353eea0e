import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def epoch_config(multiepoch=6, jcw=0.2, tsw=(0.1, 1.0), jsw=(0.25, 0.25, 0.25, 0.25)):
    """config/slt_simplified_mini.cfg:32-38,68-72,94 restricted to the search-relevant keys."""
    dims = {"lf0": 1, "mag": 60, "real": 45, "imag": 45}
    return {
        "datadims": dims, "stream_list_join": ["mag", "real", "imag", "lf0"], "datadims_join": dims,
        "stream_list_target": ["mag", "lf0"], "datadims_target": dims,
        "target_stream_weights": list(tsw), "join_stream_weights": list(jsw), "join_cost_weight": jcw,
        "target_representation": "epoch", "greedy_search": True, "multiepoch": multiepoch, "search_epsilon": 0.0,
    }


def halfphone_config(n_candidates=50, jcw=0.2, preselection="acoustic"):
    """config/hybrid_halfphone_default.cfg search-relevant keys (threepoint + duration = 184 dims)."""
    dims = {"lf0": 1, "mag": 60, "real": 45, "imag": 45}
    return {
        "datadims": dims, "stream_list_join": ["mag", "real", "imag", "lf0"], "datadims_join": dims,
        "stream_list_target": ["mag", "lf0"], "datadims_target": dims,
        "target_stream_weights": [0.5, 0.5], "join_stream_weights": [0.25] * 4, "join_cost_weight": jcw,
        "target_representation": "threepoint", "add_duration_as_target": True, "duration_target_weight": 0.5,
        "greedy_search": False, "multiepoch": 1, "n_candidates": n_candidates, "preselection_method": preselection,
        "search_epsilon": 0.0,
    }


@pytest.fixture(scope="session")
def golden_epoch():
    return np.load(os.path.join(GOLDEN, "epoch_greedy.npz"))


@pytest.fixture(scope="session")
def golden_halfphone():
    return np.load(os.path.join(GOLDEN, "halfphone_viterbi.npz"))
