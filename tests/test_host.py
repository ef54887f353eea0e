"""CPU tests of the host-side logic that mirrors the reference's Python interface."""
import numpy as np
import pytest

from conftest import epoch_config
from oracle import snickery_oracle as O
from snickery_b200 import synth


def test_read_config_accepts_python_source(tmp_path):
    p = tmp_path / "mini.cfg"
    p.write_text("datadims = {'lf0':1, 'mag': 60}\nstream_list_target = ['mag', 'lf0']\n"
                 "target_stream_weights = [1.0 / float(len(stream_list_target))] * len(stream_list_target)\n"
                 "join_cost_weight = 0.2\nmultiepoch=6\nsearch_epsilon = 10.0\n")
    cfg = synth.read_config(str(p))
    assert cfg["target_stream_weights"] == [0.5, 0.5] and cfg["multiepoch"] == 6
    assert "__builtins__" not in cfg


def test_segment_axis_cut_matches_oracle():
    a = np.arange(23 * 3, dtype=np.float64).reshape(23, 3)
    got = synth.segment_axis_cut(a, 6)
    want = O.segment_axis0(a, 6, 0).reshape(23 // 6, 18)
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        synth.segment_axis_cut(a[:5], 6)


def test_weight_vectors_match_oracle():
    cfg = epoch_config()
    s = synth.Synthesiser.__new__(synth.Synthesiser)
    s.config = cfg
    s.stream_list_target, s.stream_list_join = cfg["stream_list_target"], cfg["stream_list_join"]
    s.datadims_target, s.datadims_join = cfg["datadims_target"], cfg["datadims_join"]
    s.target_representation = "epoch"
    s.set_target_weights(np.array(cfg["target_stream_weights"]) * 0.8)
    s.set_join_weights(np.array(cfg["join_stream_weights"]) * 0.2)
    wt = O.per_coeff_weights(np.array(cfg["target_stream_weights"]) * 0.8, cfg["stream_list_target"], cfg["datadims_target"])
    wj = O.per_coeff_weights(np.array(cfg["join_stream_weights"]) * 0.2, cfg["stream_list_join"], cfg["datadims_join"])
    assert np.array_equal(s.target_weight_vector, wt) and np.array_equal(s.join_weight_vector, wj)
    with pytest.raises(AssertionError):
        s.set_join_weights([1.0])   # assert len(weights) == len(streams), synth_simple.py:235
