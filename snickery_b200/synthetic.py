"""Seeded synthetic magphase-shaped unit databases and target utterances.

Nothing here comes from the reference's code; it manufactures arrays with the
schema the reference's trainers write:

* epoch voices   -- ``train_unit_features`` f32 [N, 61] (mag 60 + lf0 1) and
  ``join_contexts`` f32 [N+1, 151] (mag 60, real 45, imag 45, lf0 1) where row
  u+1 is the join frame of unit u and row 0 duplicates the first frame
  (reference: script/train_simple.py:203-217, 278-289).
* halfphone voices -- target f32 [N, 184] (three-point 3*61 + duration,
  reference: script/train_halfphone.py:181-184, script/const.py:17) with the same
  join layout.

Streams follow AR(1) trajectories (rho 0.95, reset per utterance) with a
per-coefficient scale decaying 1.0 -> 0.1 inside each stream; lf0 has 35 %
unvoiced runs pinned to the constant -20 sigma (reference: script/const.py:13-15,
script/data_manipulation.py:174-183) which creates exact ties on purpose.
"""
from __future__ import annotations

import numpy as np

MAG, REAL, IMAG, LF0 = 60, 45, 45, 1
EPOCH_JOIN_DIMS = {"mag": MAG, "real": REAL, "imag": IMAG, "lf0": LF0}
EPOCH_TARGET_DIMS = {"mag": MAG, "lf0": LF0}
UV_VALUE = -20.0  # standardised value of an unvoiced lf0 frame


def _ar1_fast(rng, n, dim, rho, resets):
    """Same process as _ar1 via scipy.signal.lfilter per utterance segment."""
    from scipy.signal import lfilter

    eps = rng.standard_normal((n, dim))
    c = np.sqrt(1.0 - rho * rho)
    out = np.empty_like(eps)
    bounds = list(resets) + [n]
    for a, b in zip(bounds[:-1], bounds[1:]):
        seg = eps[a:b].copy()
        seg[1:] *= c
        out[a:b] = lfilter([1.0], [1.0, -rho], seg, axis=0)
    return out


def _frames(rng, n, resets, rho=0.95):
    """[n, 151] standardised frames laid out mag | real | imag | lf0."""
    cols = []
    for width in (MAG, REAL, IMAG):
        scale = np.linspace(1.0, 0.1, width)
        cols.append(_ar1_fast(rng, n, width, rho, resets) * scale)
    lf0 = _ar1_fast(rng, n, 1, rho, resets)
    # unvoiced runs: two-state Markov chain with ~35 % occupancy, mean run ~20 frames
    uv = np.zeros(n, dtype=bool)
    state = False
    r = rng.random(n)
    p_enter, p_leave = 0.027, 0.05
    for t in range(n):
        if state:
            state = r[t] >= p_leave
        else:
            state = r[t] < p_enter
        uv[t] = state
    lf0[uv, 0] = UV_VALUE
    cols.append(lf0)
    return np.hstack(cols).astype(np.float32)


def utterance_starts(rng, n_utts, lo, hi):
    lens = rng.integers(lo, hi + 1, size=n_utts)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return starts.astype(np.int64), lens.astype(np.int64)


def make_epoch_db(n_utts=100, seed=1235, utt_len=(300, 1000), n_units=None):
    """Epoch voice: returns dict(F, Jc, starts, lens).

    If n_units is given, utterances are generated until that many rows exist
    and the arrays are cut to exactly n_units.
    """
    rng = np.random.default_rng(seed)
    if n_units is not None:
        n_utts = int(np.ceil(n_units / float(utt_len[0]))) + 1
    starts, lens = utterance_starts(rng, n_utts, *utt_len)
    n = int(lens.sum())
    X = _frames(rng, n, starts)
    if n_units is not None:
        assert n >= n_units
        X = X[:n_units]
        keep = starts < n_units
        starts, lens = starts[keep], lens[keep]
        lens[-1] = n_units - starts[-1]
        n = n_units
    F = np.ascontiguousarray(np.hstack([X[:, :MAG], X[:, -1:]]))
    Jc = np.ascontiguousarray(np.vstack([X[:1], X]))
    return {"F": F, "Jc": Jc, "starts": starts, "lens": lens,
            "stream_list_target": ["mag", "lf0"], "datadims_target": dict(EPOCH_TARGET_DIMS),
            "stream_list_join": ["mag", "real", "imag", "lf0"], "datadims_join": dict(EPOCH_JOIN_DIMS)}


def make_halfphone_db(n_units=90000, seed=1237, utt_len=(60, 100), n_phones=40):
    """Halfphone voice: three-point target (3*61) + duration column = 184 dims,
    151-dim join contexts, a monophone class per unit for label preselection."""
    rng = np.random.default_rng(seed)
    n_utts = int(np.ceil(n_units / float(utt_len[0]))) + 1
    starts, lens = utterance_starts(rng, n_utts, *utt_len)
    n = int(lens.sum())
    X = _frames(rng, n + 2, np.concatenate([starts, [n]]))[: n + 2]
    pts = [np.hstack([X[i:n + i, :MAG], X[i:n + i, -1:]]) for i in range(3)]
    dur = rng.standard_normal((n, 1)).astype(np.float32)
    F = np.hstack(pts + [dur]).astype(np.float32)[:n_units]
    Xj = X[1:n + 1][:n_units]
    Jc = np.ascontiguousarray(np.vstack([Xj[:1], Xj]))
    phones = rng.integers(0, n_phones, size=n_units).astype(np.int32)
    keep = starts < n_units
    starts, lens = starts[keep], lens[keep]
    lens[-1] = n_units - starts[-1]
    return {"F": np.ascontiguousarray(F), "Jc": Jc, "starts": starts, "lens": lens, "phones": phones,
            "stream_list_target": ["mag", "lf0"], "datadims_target": dict(EPOCH_TARGET_DIMS),
            "stream_list_join": ["mag", "real", "imag", "lf0"], "datadims_join": dict(EPOCH_JOIN_DIMS)}


def make_targets(F, n_utts, length, seed, noise=0.3):
    """Target utterances: consecutive DB trajectories + N(0, noise^2), f32 [n_utts][length, Dt].
    Unvoiced lf0 values (exact constant) are kept un-noised so ties survive."""
    rng = np.random.default_rng(seed)
    n, dt = F.shape
    out = []
    for _ in range(n_utts):
        s = int(rng.integers(0, n - length))
        seg = F[s:s + length].astype(np.float64)
        noisy = seg + noise * rng.standard_normal(seg.shape)
        uv = seg == UV_VALUE
        noisy[uv] = UV_VALUE
        out.append(noisy.astype(np.float32))
    return out


def quinphone_like_candidates(phones, target_phones, k, seed, dup_rate=0.3):
    """Label-index style candidate sets (reference: synth_halfphone.py:1305-1336):
    ids drawn from the target's phone class, with back-off duplicates and -1 padding."""
    rng = np.random.default_rng(seed)
    by_phone = {}
    order = np.argsort(phones, kind="stable")
    sp = phones[order]
    cuts = np.flatnonzero(np.diff(sp)) + 1
    for grp in np.split(order, cuts):
        by_phone[int(phones[grp[0]])] = grp
    cand = np.full((len(target_phones), k), -1, dtype=np.int64)
    for t, p in enumerate(target_phones):
        pool = by_phone.get(int(p), np.array([1]))
        n_spec = int(rng.integers(1, k + 1))
        spec = rng.choice(pool, size=min(n_spec, len(pool)), replace=False)
        row = list(spec)
        if rng.random() < dup_rate:  # back-off lists repeat the specific units first
            row += list(spec)
        if rng.random() < 0.7:
            row += list(rng.choice(pool, size=min(k, len(pool)), replace=False))
        row = row[:k]
        cand[t, :len(row)] = row
    return cand
