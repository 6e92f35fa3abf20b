"""Pair-sharding for multi-GPU runs.

Residual blocks of different directed frame pairs are independent given the (replicated) parameter
vector, so whole pairs are assigned to ranks by longest-processing-time greedy on their constraint
count; every rank accumulates its share of J^T J / J^T r / cost and one all-reduce per evaluation
restores the full normal equations (rcvd_problem_init_comm).  Regulariser rows of frame f belong to
rank f % nranks (robust_cvd_b200/csrc/rcvd_eval.cuh, k_regularisers).
"""
import numpy as np


def lpt_partition(counts, nranks):
    """Returns a list of index arrays (one per rank), pair indices sorted ascending inside a rank."""
    counts = np.asarray(counts, np.int64)
    order = np.argsort(-counts, kind="stable")
    load = np.zeros(nranks, np.int64)
    bins = [[] for _ in range(nranks)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i)); load[r] += int(counts[i])
    return [np.array(sorted(b), np.int64) for b in bins]


def take_pairs(pair_frames, offsets, records, sel):
    """Sub-problem arrays (pair_frames, offsets, records) for the selected pair indices."""
    pair_frames = np.asarray(pair_frames, np.int32).reshape(-1, 2)
    offsets = np.asarray(offsets, np.int64)
    records = np.asarray(records, np.float32).reshape(-1, 6)
    recs, offs = [], [0]
    for i in sel:
        r = records[offsets[i]:offsets[i + 1]]
        recs.append(r); offs.append(offs[-1] + r.shape[0])
    return (pair_frames[sel].copy(), np.asarray(offs, np.int64),
            np.concatenate(recs, axis=0) if recs else np.zeros((0, 6), np.float32))


def frame_owner(frame, nranks):
    return frame % nranks
