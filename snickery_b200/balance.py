"""Stream-weight balancing on the GPU-resident database.

Mirrors the loop of the reference's script/balance_stream_weights.py:40-172: every epoch sets new
per-stream join / target weights, "rebuilds the tree", searches the tune utterances and moves each
weight against the sign of its stream's cost error with RPROP-style step sizes.  In the reference an
epoch re-weights the whole voice in numpy and rebuilds a cKDTree; here it is one `snk_db_set_weights`
pass over resident matrices, a batched greedy search and the per-stream cost reductions of
`snk_greedy_path_scores` -- the loop becomes search-bound.
"""
from __future__ import annotations

import copy

import numpy as np


def mean_scores_without_zeros(cached_jscores, cached_tscores):
    """balance_stream_weights.py:95-110: per-stream mean over the strictly positive entries."""
    means = []
    for scores in (cached_jscores, cached_tscores):
        for column in range(scores.shape[1]):
            vals = scores[:, column]
            vals = vals[vals > 0.0]
            means.append(0.0 if vals.shape[0] == 0 else vals.sum() / vals.shape[0])
    return np.array(means)


def rprop_balance(evaluate, njoin, ntarget, max_epochs=1000, patience=5, thresh=0.001, eta=0.1, amplifier=1.2,
                  attenuator=0.5, dmax=50.0, dmin=0.000001, weight_floor=0.0, verbose=False):
    """The engine-agnostic loop.  evaluate(join_weights, target_weights) -> (jscores [*, njoin], tscores [*, ntarget])
    stacked over the tune utterances.  Returns (best_weights, losses, weight_history)."""
    weights = np.ones(njoin + ntarget)
    best_weights = copy.copy(weights)
    best_score = previous_score = float("inf")
    epochs_without_improvement = 0
    lrates = np.ones(weights.shape) * eta
    prev_directions = np.ones(weights.shape)
    losses, history, goals = [], [], None
    for i in range(max_epochs):
        jscores, tscores = evaluate(weights[:njoin], weights[njoin:])
        mean_scores = mean_scores_without_zeros(jscores, tscores)
        if i == 0:   # target and join contribute equally; streams equally within each (:113-118)
            goal_join = (mean_scores.sum() / 2.0) / njoin
            goal_target = (mean_scores.sum() / 2.0) / ntarget
            goals = np.array([goal_join] * njoin + [goal_target] * ntarget)
        errors = mean_scores - goals
        loss = np.abs(errors).sum()
        losses.append(loss)
        history.append(copy.copy(weights))
        if verbose:
            print("=== iteration %s | loss %s ===" % (i + 1, loss))
        if loss < previous_score:
            epochs_without_improvement = 0
        else:
            epochs_without_improvement += 1
        if loss < best_score:
            best_score = loss
            best_weights = copy.copy(weights)
        if epochs_without_improvement == patience or loss < thresh:
            break
        directions = np.sign(-1.0 * errors)
        direction_change = directions * prev_directions
        lrates[direction_change > 0] *= amplifier
        lrates[direction_change < 0] *= attenuator
        lrates = np.clip(lrates, dmin, dmax)
        prev_directions = copy.copy(directions)
        weights = np.maximum(weights + directions * lrates, weight_floor)
        previous_score = loss
    return best_weights, losses, history


def balance_stream_weights(synth, tune_utts_unweighted, **kwargs):
    """Runs the loop on a snickery_b200.Synthesiser.  tune_utts_unweighted: standardised, UNWEIGHTED target
    features [T, Dt] per tune utterance (the reference re-weights them every epoch, synth_simple.py:389)."""
    njoin, ntarget = len(synth.stream_list_join), len(synth.stream_list_target)
    utts = [np.asarray(u, dtype=np.float64) for u in tune_utts_unweighted]
    m = synth.db.multiepoch

    def evaluate(join_weights, target_weights):
        synth.set_join_weights(join_weights)          # balance_stream_weights.py:84-88
        synth.set_target_weights(target_weights)
        synth.get_tree_for_greedy_search()
        weighted = [u * synth.target_weight_vector[None, :] for u in utts]
        paths = synth.greedy_joint_search_batch(weighted)
        js, ts = [], []
        for u, p in zip(weighted, paths):
            t, j = synth.get_scores_per_stream(u[: len(p) * m], p)
            ts.append(t)
            js.append(j)
        return np.vstack(js), np.vstack(ts)

    return rprop_balance(evaluate, njoin, ntarget, **kwargs)
