"""The argument behind the pruned candidate sort (prefilter.cuh: best_kernel histogram over every fourth block, cand_threshold_kernel,
pf_extract_cand_keys), checked on a numpy model: a threshold read off the score histogram of ANY subset of a query's candidates is a lower
bound of computeScoreThreshold (QueryMatcher.h:211-221), so keeping the candidates at or above it keeps every candidate the selection
(QueryMatcher.cpp:183-240: threshold, cut at max-seqs in array order) can return, and the selection over the survivors is the selection
over all candidates.  The CUDA path itself is compared with the oracle and the reference's outputs in tests/test_prefilter_gpu.py."""
import numpy as np


def score_threshold(scores, max_hits):
    """computeScoreThreshold: the score at which the count from 255 downwards reaches max_hits, 0 if it never does"""
    hist = np.bincount(scores, minlength=256)
    cum = 0
    for sc in range(255, 0, -1):
        cum += hist[sc]
        if cum >= max_hits:
            return sc
    return 0


def select(scores, order, max_hits, min_score):
    """candidates in the reference's array order (score desc, then `order`), threshold max(thr, min_score), first max_hits survive"""
    idx = np.lexsort((order, 255 - scores))
    thr = max(score_threshold(scores, max_hits), min_score)
    keep = idx[scores[idx] >= thr][:max_hits]
    return keep


def test_subset_threshold_is_a_lower_bound_and_pruning_keeps_the_selection():
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(1, 6000))
        max_hits = int(rng.choice([1, 10, 300, 1000, 2048]))
        # ungapped scores: mostly just above the minimum, a tail up to the 255 clamp, sometimes many saturated ones
        scores = np.minimum(255, 30 + rng.geometric(0.12, n) + (rng.random(n) < 0.02) * rng.integers(0, 230, n)).astype(np.int64)
        if trial % 7 == 0:
            scores[rng.random(n) < 0.5] = 255
        order = rng.permutation(n)
        full_thr = score_threshold(scores, max_hits)
        sub = rng.random(n) < rng.choice([0.25, 0.5, 1.0, 0.05])        # any subset: every fourth block in the kernel
        low = score_threshold(scores[sub], max_hits) if sub.any() else 0
        assert low <= full_thr
        survivors = np.nonzero(scores >= low)[0]
        a = select(scores, order, max_hits, 30)
        b = survivors[select(scores[survivors], order[survivors], max_hits, 30)]
        assert (a == b).all()
