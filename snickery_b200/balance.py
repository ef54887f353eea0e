"""Stream-weight balancing on the GPU-resident database.

What the reference's script/balance_stream_weights.py:82-172 computes: per-stream weights for which every join stream
and every target stream contributes the same mean cost along the selected paths of a set of tune utterances, found by a
sign-driven step rule (the step of a stream grows while its error keeps its sign and shrinks when it flips).  There an
epoch re-weights the whole voice in numpy, rebuilds a cKDTree and searches the utterances one by one; here an epoch is

    snk_db_set_weights  ->  ONE batched greedy search of all tune utterances  ->  snk_greedy_path_scores per utterance

over resident matrices, so the loop is bound by the search.  The step rule lives in `SignStepBalancer` (an object with
state, so a caller can also drive it epoch by epoch); `rprop_balance` and `balance_stream_weights` are the drivers.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def positive_column_means(*score_blocks):
    """Mean of the strictly positive entries of every column of every block, blocks side by side (join streams first).
    A column without positive entries -- a stream that saw no join -- contributes 0
    (balance_stream_weights.py:95-110)."""
    out = []
    for block in score_blocks:
        block = np.asarray(block, dtype=np.float64)
        pos = block > 0.0
        n = pos.sum(axis=0)
        # add only the positive entries, in row order, as a masked column sum does
        sums = np.array([block[pos[:, c], c].sum() for c in range(block.shape[1])])
        out.append(np.divide(sums, n, out=np.zeros(block.shape[1]), where=n > 0))
    return np.concatenate(out)


# name kept for callers of the first release
mean_scores_without_zeros = positive_column_means


@dataclass
class SignStepBalancer:
    """State of the balancing rule for njoin + ntarget streams.

    observe(mean_costs) takes the per-stream mean costs measured with the CURRENT weights, records them, and returns
    True while the search should go on; `weights` then holds the weights to try next.  The targets are fixed by the
    first observation: half of the total cost for the join streams, half for the target streams, split evenly inside
    each half (balance_stream_weights.py:113-118)."""
    njoin: int
    ntarget: int
    eta: float = 0.1                 # initial step
    grow: float = 1.2                # step factor while the error keeps its sign
    shrink: float = 0.5              # ... when it flips
    step_max: float = 50.0
    step_min: float = 0.000001
    floor: float = 0.0               # lower limit of a weight
    patience: int = 5                # epochs without improvement before giving up
    tolerance: float = 0.001         # loss at which the search stops
    weights: np.ndarray = field(init=False)
    best_weights: np.ndarray = field(init=False)
    losses: list = field(default_factory=list)
    history: list = field(default_factory=list)

    def __post_init__(self):
        n = self.njoin + self.ntarget
        self.weights = np.ones(n)
        self.best_weights = self.weights.copy()
        self._step = np.full(n, self.eta)
        self._last_sign = np.ones(n)
        self._targets = None
        self._best = self._previous = np.inf
        self._stale = 0

    def observe(self, mean_costs):
        mean_costs = np.asarray(mean_costs, dtype=np.float64)
        if self._targets is None:
            half = mean_costs.sum() / 2.0
            self._targets = np.concatenate([np.full(self.njoin, half / self.njoin), np.full(self.ntarget, half / self.ntarget)])
        err = mean_costs - self._targets
        loss = np.abs(err).sum()
        self.losses.append(loss)
        self.history.append(self.weights.copy())
        self._stale = 0 if loss < self._previous else self._stale + 1
        if loss < self._best:
            self._best, self.best_weights = loss, self.weights.copy()
        if self._stale == self.patience or loss < self.tolerance:
            return False
        sign = np.sign(-err)                                   # move against the error
        agree = sign * self._last_sign
        self._step = np.clip(np.where(agree > 0, self._step * self.grow, np.where(agree < 0, self._step * self.shrink, self._step)),
                             self.step_min, self.step_max)
        self._last_sign = sign
        self.weights = np.maximum(self.weights + sign * self._step, self.floor)
        self._previous = loss
        return True


def rprop_balance(evaluate, njoin, ntarget, max_epochs=1000, patience=5, thresh=0.001, eta=0.1, amplifier=1.2,
                  attenuator=0.5, dmax=50.0, dmin=0.000001, weight_floor=0.0, verbose=False):
    """Engine-agnostic driver.  evaluate(join_weights, target_weights) -> (jscores [*, njoin], tscores [*, ntarget])
    stacked over the tune utterances.  Returns (best_weights, losses, weight_history).  Keyword names follow the
    reference's command-line options."""
    b = SignStepBalancer(njoin, ntarget, eta=eta, grow=amplifier, shrink=attenuator, step_max=dmax, step_min=dmin,
                         floor=weight_floor, patience=patience, tolerance=thresh)
    for epoch in range(max_epochs):
        more = b.observe(positive_column_means(*evaluate(b.weights[:njoin], b.weights[njoin:])))
        if verbose:
            print("=== iteration %s | loss %s ===" % (epoch + 1, b.losses[-1]))
        if not more:
            break
    return b.best_weights, b.losses, b.history


def balance_stream_weights(synth, tune_utts_unweighted, **kwargs):
    """Runs the search on a snickery_b200.Synthesiser.  tune_utts_unweighted: standardised, UNWEIGHTED target
    features [T, Dt] per tune utterance (the reference re-weights them every epoch, synth_simple.py:389)."""
    njoin, ntarget = len(synth.stream_list_join), len(synth.stream_list_target)
    utts = [np.asarray(u, dtype=np.float64) for u in tune_utts_unweighted]
    m = synth.db.multiepoch

    def evaluate(join_weights, target_weights):
        synth.set_join_weights(join_weights)          # one pass over the resident matrices, no tree
        synth.set_target_weights(target_weights)
        synth.get_tree_for_greedy_search()
        weighted = [u * synth.target_weight_vector[None, :] for u in utts]
        paths = synth.greedy_joint_search_batch(weighted)     # every tune utterance in one batched search
        scores = [synth.get_scores_per_stream(u[: len(p) * m], p) for u, p in zip(weighted, paths)]
        return np.vstack([j for _, j in scores]), np.vstack([t for t, _ in scores])

    return rprop_balance(evaluate, njoin, ntarget, **kwargs)
