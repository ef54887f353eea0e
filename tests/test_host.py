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


def test_single_utterance_kernel_lane_maps():
    """Host restatement of two pieces of index arithmetic in snickery_b200/csrc/greedy_one.cu, so that an edit of the kernel's
    lane maps has something to be checked against without a GPU.
    (1) Diagonal sum.  An m16n8 accumulator tile C[row][column j] = (frame row) . (query frame j) is spread over a warp as
        lane (g, t): C[g][2t], C[g][2t+1], C[g+8][2t], C[g+8][2t+1].  key(u) needs sum_j C[u + j][j] over m consecutive
        rows: rows of this tile and the first m - 1 of the next.  The kernel fetches them with index shuffles from lane
        (((g + j) & 7) << 2) | t and picks the register by whether g + j < 8, then adds over the four lanes of a row.
    (2) Row slices.  bounds[c] = floor(ngroups * (rates before c) / (all rates)) must be monotone, start at 0, end at ngroups,
        and give a faster CTA more tiles."""
    rng = np.random.default_rng(0)
    for m in (1, 3, 4, 6):
        ck, cn = rng.standard_normal((16, 8)), rng.standard_normal((16, 8))
        cp = np.zeros((32, 4))
        cnx = np.zeros((32, 4))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            cp[lane] = [ck[g, 2 * t], ck[g, 2 * t + 1], ck[g + 8, 2 * t], ck[g + 8, 2 * t + 1]]
            cnx[lane] = [cn[g, 2 * t], cn[g, 2 * t + 1], cn[g + 8, 2 * t], cn[g + 8, 2 * t + 1]]
        ta, tb = np.zeros(32), np.zeros(32)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for par in (0, 1):
                j = 2 * t + par
                if j >= m:
                    continue
                src = (((g + j) & 7) << 2) | t
                lo = g + j < 8
                ta[lane] += cp[src, par] if lo else cp[src, 2 + par]
                tb[lane] += cp[src, 2 + par] if lo else cnx[src, par]
        both = np.vstack([ck, cn])
        for g in range(8):
            assert np.isclose(ta[4 * g:4 * g + 4].sum(), sum(both[g + j, j] for j in range(m)))
            assert np.isclose(tb[4 * g:4 * g + 4].sum(), sum(both[g + 8 + j, j] for j in range(m)))
    # (2) the re-cut, in the kernel's float32 arithmetic: lane l owns CTAs [5 l, 5 l + 5)
    grid, ngroups = 148, 43750
    speed = (0.0075 * (1 + 0.08 * rng.standard_normal(grid))).astype(np.float32)
    speed[[17, 90]] = 0.0                                   # no measurement yet: filled in with the mean rate
    fill_in = np.float32(speed[speed > 0].sum() / np.float32((speed > 0).sum()))
    eff = np.where(speed > 0, speed, fill_in).astype(np.float32)
    per = (grid + 31) // 32
    mine = np.array([eff[l * per:min(grid, (l + 1) * per)].sum(dtype=np.float32) for l in range(32)], dtype=np.float32)
    incl = np.cumsum(mine, dtype=np.float32)
    scale = np.float32(ngroups) / incl[31]
    bounds = np.zeros(grid + 1, dtype=np.int64)
    for l in range(32):
        cum = np.float32(incl[l] - mine[l])
        for c in range(l * per, min(grid, (l + 1) * per)):
            bounds[c] = min(ngroups, int(cum * scale))
            cum = np.float32(cum + eff[c])
    bounds[0], bounds[grid] = 0, ngroups
    sizes = np.diff(bounds)
    assert np.all(sizes >= 0) and sizes.sum() == ngroups
    assert np.corrcoef(sizes, eff)[0, 1] > 0.98
