"""CPU oracle for the brute-force Tanimoto scan + top-k path.  TEST INFRASTRUCTURE ONLY.

This module restates, in numpy, what the reference computes on its hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it; the product (``gpusimilarity_b200``) never does.

Pinning: the restatement is checked (tests/test_oracle.py) against
  * the reference's own known answers for this path (test/test_gpusim.cpp:110-113 cutoff
    counts, :98 multi-DB top id, :64-67 GPU==CPU order, :136-145 CPUSort, :151-165 fold),
  * outputs of the reference's own sources compiled verbatim (``oracle/_ref``,
    built by ``oracle/Makefile`` from /root/reference with the header shims in
    ``oracle/qt_shims``) and frozen in ``tests/golden/`` by ``tests/golden/make_golden.py``.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint32) if np.asarray(a).dtype == np.int32 \
        else np.asarray(a, dtype=np.uint32)


def popcounts(words: np.ndarray) -> np.ndarray:
    """Per-row population count of an (N, W) word matrix."""
    return np.bitwise_count(_u32(words)).sum(axis=-1, dtype=np.int64)


def common_union(query: np.ndarray, db: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``common`` and ``total - common`` of reference calculation_functors.cpp:8-16 /
    fingerprintdb_cuda.cu:91-99 for every row."""
    q = _u32(query).reshape(1, -1)
    d = _u32(db)
    common = np.bitwise_count(d & q).sum(axis=1, dtype=np.int64)
    total = popcounts(d) + int(np.bitwise_count(q).sum())
    return common, total - common


def tanimoto_scores_cpu(query: np.ndarray, db: np.ndarray) -> np.ndarray:
    """TanimotoFunctorCPU::operator() (calculation_functors.cpp:6-20): f32(common)/f32(union),
    IEEE divide, NO cutoff; 0/0 stays NaN."""
    common, union = common_union(query, db)
    with np.errstate(invalid="ignore", divide="ignore"):
        return common.astype(np.float32) / union.astype(np.float32)


def tanimoto_scores_gpu(query: np.ndarray, db: np.ndarray, cutoff: float) -> np.ndarray:
    """TanimotoFunctor::operator() (fingerprintdb_cuda.cu:89-103): the CPU score, then
    ``score >= cutoff ? score : 0`` evaluated in f32 (NaN compares false -> 0)."""
    s = tanimoto_scores_cpu(query, db)
    c = np.float32(cutoff)
    with np.errstate(invalid="ignore"):
        return np.where(s >= c, s, np.float32(0)).astype(np.float32)


def metric_scores_gpu(query: np.ndarray, db: np.ndarray, cutoff: float, metric: str = "tanimoto",
                      alpha: float = 1.0, beta: float = 1.0) -> np.ndarray:
    """Dice and Tversky variants of the scan.  The reference scores Tanimoto only, so
    these have no reference counterpart to be pinned to ("parity unpinned" for the two extra
    metrics): this function IS their definition, restated operation by operation by the kernels
    (gsb_kernels.cuh metric_score) — f32, every operation rounded, then the reference's cutoff rule
    (fingerprintdb_cuda.cu:102).
      dice      f32(2c) / f32(pq + pd)
      tversky   f32(c) / ((alpha * f32(pq - c) + beta * f32(pd - c)) + f32(c))"""
    if metric == "tanimoto":
        return tanimoto_scores_gpu(query, db, cutoff)
    q = _u32(query).reshape(1, -1)
    d = _u32(db)
    c = np.bitwise_count(d & q).sum(axis=1, dtype=np.int64)
    pq = int(np.bitwise_count(q).sum())
    pd = popcounts(d)
    with np.errstate(invalid="ignore", divide="ignore"):
        if metric == "dice":
            s = (2 * c).astype(np.float32) / (pq + pd).astype(np.float32)
        elif metric == "tversky":
            c32 = c.astype(np.float32)
            t1 = np.float32(alpha) * (pq - c).astype(np.float32)
            t2 = np.float32(beta) * (pd - c).astype(np.float32)
            s = c32 / ((t1 + t2) + c32)
        else:
            raise ValueError(metric)
        return np.where(s >= np.float32(cutoff), s, np.float32(0)).astype(np.float32)


def search_gpu_metric(query: np.ndarray, db: np.ndarray, k: int, cutoff: float, metric: str,
                      alpha: float = 1.0, beta: float = 1.0) -> Tuple[np.ndarray, np.ndarray, int]:
    """search_gpu with the score of ``metric_scores_gpu`` (same cutoff / survivor / order rules)."""
    s = metric_scores_gpu(query, db, cutoff, metric, alpha, beta)
    rows = np.arange(db.shape[0], dtype=np.int64)
    if np.float32(cutoff) > 0:
        keep = s != 0
        s, rows = s[keep], rows[keep]
    approx = int(rows.shape[0])
    order = canonical_order(s, rows)[:k]
    return rows[order], s[order], approx


def canonical_order(scores: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """Indices that sort by (score desc, row asc) — the order a stable descending sort of
    (score, row) produces from ascending rows (fingerprintdb_cuda.cu:245,280-282)."""
    return np.lexsort((rows, -scores.astype(np.float64)))


def search_gpu(query: np.ndarray, db: np.ndarray, k: int, cutoff: float,
               row_base: int = 0) -> Tuple[np.ndarray, np.ndarray, int]:
    """FingerprintDB::search semantics for an unfolded database (fingerprintdb_cuda.cu:228-381):
    score+zero (:258-262), drop zero scores only when cutoff > 0 (:265-271), survivors =
    approximate count (:272-277, :367-369), stable descending sort (:280-282), first
    min(k, survivors) (:284-290, :376-380).  Cross-chunk order is canonicalised
    (score desc, global row asc), see SURVEY App. D."""
    s = tanimoto_scores_gpu(query, db, cutoff)
    rows = np.arange(db.shape[0], dtype=np.int64)
    if np.float32(cutoff) > 0:
        keep = s != 0
        s, rows = s[keep], rows[keep]
    approx = int(rows.shape[0])
    order = canonical_order(s, rows)[:k]
    return rows[order] + row_base, s[order], approx


def top_results_bubble_sort(indices: List[int], scores: List[float], number_required: int) -> None:
    """In-place partial bubble sort, fingerprintdb_cuda.cpp:92-103 (strict '>' => stable)."""
    count = len(indices)
    for i in range(number_required):
        for j in range(count - 1, i, -1):
            if scores[j] > scores[j - 1]:
                indices[j], indices[j - 1] = indices[j - 1], indices[j]
                scores[j], scores[j - 1] = scores[j - 1], scores[j]


def search_cpu(query: np.ndarray, db: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """FingerprintDB::search_cpu (fingerprintdb_cuda.cpp:20-54): CPU scores (no cutoff, no
    approximate count), then the first k of the partial bubble sort.  For NaN-free scores the
    bubble sort's first k equal the stable (score desc, row asc) prefix, which is what is
    computed here; ``k`` is clamped to N (the reference over-reads when k > N, :49)."""
    s = tanimoto_scores_cpu(query, db)
    rows = np.arange(db.shape[0], dtype=np.int64)
    order = canonical_order(np.nan_to_num(s, nan=-1.0), rows)[:min(k, db.shape[0])]
    return rows[order], s[order]


def fold_fingerprint(fp: np.ndarray, factor: int) -> np.ndarray:
    """FoldFingerprintFunctorCPU::operator() (calculation_functors.cpp:22-41): bit ``pos`` of the
    unfolded row lands on ``pos % new_size`` with the same in-word position, i.e. the folded
    row is the OR of the ``factor`` contiguous segments of ``words/factor`` words."""
    w = _u32(fp)
    words = w.shape[-1]
    assert words % factor == 0
    seg = w.reshape(w.shape[:-1] + (factor, words // factor))
    return np.bitwise_or.reduce(seg, axis=-2).view(np.int32)


def effective_fold_factor(words: int, fold_factor: int) -> int:
    """copyToGPU bumps the factor to the next divisor of the word count (.cu:170-173)."""
    f = max(1, int(fold_factor))
    while words % f != 0:
        f += 1
    return f


def fold_candidate_count(n_survivors: int, k: int, factor: int) -> int:
    """results_to_consider (fingerprintdb_cuda.cu:284-287)."""
    return min(n_survivors, k * factor * int(math.log2(2 * factor)))


def search_gpu_folded(query: np.ndarray, db: np.ndarray, k: int, cutoff: float,
                      fold_factor: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Folded search of one storage chunk (fingerprintdb_cuda.cu:246-331): scan the folded rows
    with the folded query, take ``fold_candidate_count`` candidates in canonical order,
    re-score them with the full fingerprints (:309-316), order them (bubble sort :317 ==
    stable score-desc over the candidate order), keep k, and stop at the first re-scored
    result below the cutoff (:321-326)."""
    words = db.shape[1]
    f = effective_fold_factor(words, fold_factor)
    if f == 1:
        return search_gpu(query, db, k, cutoff)
    fq, fdb = fold_fingerprint(query, f), fold_fingerprint(db, f)
    s = tanimoto_scores_gpu(fq, fdb, cutoff)
    rows = np.arange(db.shape[0], dtype=np.int64)
    if np.float32(cutoff) > 0:
        keep = s != 0
        s, rows = s[keep], rows[keep]
    approx = int(rows.shape[0])
    cand = rows[canonical_order(s, rows)[:fold_candidate_count(approx, k, f)]]
    full = tanimoto_scores_cpu(query, db[cand])
    order = np.argsort(-np.nan_to_num(full, nan=-1.0).astype(np.float64), kind="stable")[:k]
    out_rows, out_scores = cand[order], full[order]
    with np.errstate(invalid="ignore"):
        below = np.nonzero(out_scores < np.float32(cutoff))[0]
    if below.size:
        out_rows, out_scores = out_rows[:below[0]], out_scores[:below[0]]
    return out_rows, out_scores, approx


def search_databases(per_db: Sequence[Tuple[Sequence[bytes], Sequence[bytes], Sequence[float]]],
                     results_requested: int) -> Tuple[List[bytes], List[bytes], List[float]]:
    """GPUSimServer::searchDatabases merge (gpusim.cpp:339-373): concatenate every database's
    (score, smiles, id) results, order by score descending, join the ids of identical SMILES
    with ';:;' (collecting stops once ``results_requested`` distinct SMILES were seen), and
    emit each SMILES once.  The reference breaks score ties by char* address; here ties keep
    (database order, rank) order."""
    flat = []
    for smiles, ids, scores in per_db:
        flat.extend(zip(scores, smiles, ids))
    flat.sort(key=lambda t: -float(t[0]))
    joined: Dict[bytes, bytes] = {}
    for _, smi, cid in flat:
        joined[smi] = joined[smi] + b";:;" + cid if smi in joined else cid
        if len(joined) >= results_requested:
            break
    out_s, out_i, out_f, seen = [], [], [], set()
    for score, smi, _ in flat:
        if smi in seen:
            continue
        seen.add(smi)
        out_f.append(float(score))
        out_s.append(smi)
        out_i.append(joined.get(smi, b""))
        if len(out_s) >= results_requested:
            break
    return out_s, out_i, out_f


# ---------------------------------------------------------------------------
# Synthetic database generator (host twin of csrc/synth.cuh; DESIGN.md "Synthetic data").
# Counter-based, so any row can be regenerated anywhere: word w of row r is the AND of five
# independent 32-bit hashes (bit density 1/32, ~32 of 1024 bits set, like Morgan r=2 in the
# reference fixture); one row in ``plant_period`` is instead a near-duplicate of the
# query template (template XOR a few hashed single-bit flips) so top-k has real structure.
# ---------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def _hash32(seed: int, row: np.ndarray, word: np.ndarray, salt: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (np.asarray(row, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + np.asarray(word, dtype=np.uint64) * np.uint64(0xD1B54A32D192ED03)
             + np.uint64(salt) * np.uint64(0x8CB92BA72F3D8DD7)
             + np.uint64(seed))
    return (_mix64(x) >> np.uint64(32)).astype(np.uint32)


SYNTH_TEMPLATE_ROW = 0xFFFFFFFF  # the template uses this (never a database row) id
SYNTH_MAX_FLIPS = 24


def synth_random_rows(seed: int, rows: np.ndarray, words: int) -> np.ndarray:
    r = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
    w = np.arange(words, dtype=np.uint64).reshape(1, -1)
    out = _hash32(seed, r, w, 0)
    for salt in range(1, 5):
        out &= _hash32(seed, r, w, salt)
    return out


def synth_template(seed: int, words: int) -> np.ndarray:
    return synth_random_rows(seed, np.array([SYNTH_TEMPLATE_ROW]), words)[0].view(np.int32)


def synth_rows(seed: int, rows: np.ndarray, words: int, plant_period: int) -> np.ndarray:
    """Rows ``rows`` (global ids) of the synthetic database, as (len(rows), words) int32."""
    rows = np.asarray(rows, dtype=np.uint64)
    out = synth_random_rows(seed, rows, words)
    if plant_period > 0:
        sel = _hash32(seed, rows, np.uint64(0), 7)
        planted = (sel % np.uint32(plant_period)) == 0
        if planted.any():
            pr = rows[planted]
            tmpl = synth_random_rows(seed, np.array([SYNTH_TEMPLATE_ROW]), words)[0]
            block = np.tile(tmpl, (pr.shape[0], 1))
            nflip = 1 + (_hash32(seed, pr, np.uint64(1), 7) % np.uint32(SYNTH_MAX_FLIPS))
            for j in range(SYNTH_MAX_FLIPS):
                pos = _hash32(seed, pr, np.uint64(2 + j), 7) % np.uint32(words * 32)
                active = (nflip > j)
                idx = np.nonzero(active)[0]
                block[idx, (pos[idx] >> np.uint32(5)).astype(np.int64)] ^= (
                    np.uint32(1) << (pos[idx] & np.uint32(31)))
            out[planted] = block
    return out.view(np.int32)


def synth_db(seed: int, n_rows: int, words: int = 32, plant_period: int = 0,
             row_base: int = 0) -> np.ndarray:
    return synth_rows(seed, np.arange(row_base, row_base + n_rows, dtype=np.uint64), words,
                      plant_period)
