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


def test_rprop_balance_loop_on_a_synthetic_cost_surface():
    """Host logic of the balancing loop (balance_stream_weights.py:82-172) with an analytic evaluate()."""
    from snickery_b200 import balance

    def evaluate(jw, tw):   # stream cost grows with the square of its weight
        js = np.tile((np.array([1.0, 2.0, 3.0, 4.0]) * jw ** 2)[None, :], (5, 1))
        ts = np.tile((np.array([10.0, 0.5]) * tw ** 2)[None, :], (6, 1))
        js[0, :] = 0.0      # zeros are ignored by the mean (:99-107)
        return js, ts

    best, losses, hist = balance.rprop_balance(evaluate, 4, 2, max_epochs=60)
    assert losses[-1] < losses[0] * 0.2 and len(hist) == len(losses)
    assert np.all(best >= 0.0)
    m = balance.mean_scores_without_zeros(*evaluate(best[:4], best[4:]))
    assert abs(m[:4].sum() - m[4:].sum()) < 0.5 * m.sum()       # join and target contributions pulled together
