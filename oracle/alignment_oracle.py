"""CPU oracle for the duration extraction that follows the Aligner (reference: utils/alignments.py, utils/metrics.py:5-44):
attention maps -> per-head scores -> best head -> shortest monotonic path -> integer durations.

TEST INFRASTRUCTURE ONLY (same rules as the other oracles).  PINNED: the path search calls the very routine the reference
calls -- ``scipy.sparse.csgraph.dijkstra`` on the adjacency matrix built exactly as ``utils/alignments.py:21-55`` builds it
(scipy is installed here) -- so the integer durations are the reference's own whenever the shortest path is unique.
The score formulas restate ``utils/metrics.py`` in numpy (TensorFlow is not installable).
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import dijkstra


def to_adj_matrix(mat: np.ndarray):
    """utils/alignments.py:21-55: edges right / down / down-right, weight = value of the TARGET node."""
    rows, cols = mat.shape
    row_ind, col_ind, data = [], [], []
    for i in range(rows):
        for j in range(cols):
            node = cols * i + j
            if j < cols - 1:
                row_ind.append(node); col_ind.append(cols * i + j + 1); data.append(mat[i, j + 1])
            if i < rows - 1 and j < cols:
                row_ind.append(node); col_ind.append(cols * (i + 1) + j); data.append(mat[i + 1, j])
            if i < rows - 1 and j < cols - 1:
                row_ind.append(node); col_ind.append(cols * (i + 1) + j + 1); data.append(mat[i + 1, j + 1])
    return coo_matrix((data, (row_ind, col_ind)), shape=(rows * cols, rows * cols)).tocsr()


def extract_durations_with_dijkstra(attention_map: np.ndarray) -> np.ndarray:
    """utils/alignments.py:58-91."""
    attn_max = np.max(attention_map)
    path_probs = attn_max - attention_map
    adj = to_adj_matrix(path_probs)
    dist, pred = dijkstra(csgraph=adj, directed=True, indices=0, return_predecessors=True)
    path = []
    pr = pred[-1]
    while pr != 0:
        path.append(pr)
        pr = pred[pr]
    path.reverse()
    path = [0] + path + [dist.size - 1]
    cols = path_probs.shape[1]
    mel_text = {}
    durations = np.zeros(attention_map.shape[1], dtype=np.int32)
    for node in path:
        i, j = node // cols, node % cols
        mel_text[i] = j
    for j in mel_text.values():
        durations[j] += 1
    return durations


def diagonal_mask(mel_len: int, phon_len: int, padded_shape) -> np.ndarray:
    """utils/metrics.py:59-70 (float64 ratio, |.|, cast to float32)."""
    max_m = min(int(mel_len), int(padded_shape[0]))
    max_n = int(phon_len)
    i = np.tile(np.arange(max_n)[None, :], (max_m, 1)) / max_n
    j = np.tile(np.arange(max_m)[:, None], (1, max_n)) / max_m
    out = np.zeros(tuple(padded_shape), dtype=np.float32)
    out[:max_m, :max_n] = np.sqrt(np.square(i - j)).astype(np.float32)
    return out


def attention_score(att: np.ndarray, mel_len: np.ndarray, phon_len: np.ndarray, r: int = 1):
    """utils/metrics.py:5-24 -> (loc_score, peak_score, 3 / diag_score), each (N, heads) float32."""
    N, H, Tq, Tk = att.shape
    mask = (np.arange(Tq)[None, :] < mel_len[:, None]).astype(np.int32)[:, None, :]            # (N,1,Tq)
    max_loc = np.argmax(att, axis=3)                                                            # (N,H,Tq)
    diff = np.abs(max_loc[:, :, 1:] - max_loc[:, :, :-1])
    loc = ((diff >= 0).astype(np.int32) * (diff <= r).astype(np.int32) * mask[:, :, 1:]).sum(-1)
    loc_score = (loc / (mel_len - 1)[:, None]).astype(np.float32)
    peak_score = (np.max(att, axis=3) * mask.astype(np.float32)).mean(-1).astype(np.float32)   # mean over the PADDED length
    dmask = np.stack([diagonal_mask(mel_len[b], phon_len[b], (Tq, Tk)) for b in range(N)])[:, None]
    diag_score = (att * dmask).sum((-2, -1))
    return loc_score, peak_score, (3.0 / diag_score).astype(np.float32)


def get_durations_from_alignment(batch_alignments: np.ndarray, mels: np.ndarray, phonemes: np.ndarray, weighted: bool = False):
    """utils/alignments.py:103-143 (durations and the three scores; the plotting matrix is left out)."""
    mel_len = (np.abs(mels).sum(-1) != 0).sum(-1).astype(np.int64) - 1      # mel_lengths(mels, 0.) - 1
    phon_len = (phonemes != 0).sum(-1).astype(np.int64) - 1                 # phoneme_lengths(phonemes) - 1
    jump, peak, diag = attention_score(batch_alignments, mel_len, phon_len, r=1)
    scores = diag + jump + peak
    durations = []
    for b, al in enumerate(batch_alignments):
        unpad = al[:, 1:mel_len[b], 1:phon_len[b]]
        if weighted:
            ref = np.sum(unpad * scores[b][:, None, None], axis=0)
        else:
            ref = unpad[np.argmax(scores[b])]
        d = extract_durations_with_dijkstra(ref)
        assert d.sum() == mel_len[b] - 1
        durations.append(d)
    return durations, jump, peak, diag


def durations_by_dynamic_programming(attention_map: np.ndarray) -> np.ndarray:
    """Independent restatement of the path search as the dynamic programme the CUDA kernel runs: the graph is a DAG whose edge
    weight depends on the target node only, so dist[i,j] = w[i,j] + min(dist[i,j-1], dist[i-1,j], dist[i-1,j-1]) in float64.
    Equal to the Dijkstra result whenever no two predecessor distances tie exactly."""
    a = attention_map.astype(np.float32)
    w = (np.max(a) - a).astype(np.float64)
    M, N = w.shape
    dist = np.full((M, N), np.inf)
    pred = np.zeros((M, N), dtype=np.int8)   # 0 = left, 1 = up, 2 = diagonal
    dist[0, 0] = 0.0
    for i in range(M):
        for j in range(N):
            if i == 0 and j == 0:
                continue
            best, code = np.inf, 0
            if j > 0 and dist[i, j - 1] < best:
                best, code = dist[i, j - 1], 0
            if i > 0 and dist[i - 1, j] < best:
                best, code = dist[i - 1, j], 1
            if i > 0 and j > 0 and dist[i - 1, j - 1] < best:
                best, code = dist[i - 1, j - 1], 2
            dist[i, j] = best + w[i, j]
            pred[i, j] = code
    durations = np.zeros(N, dtype=np.int32)
    i, j = M - 1, N - 1
    last_row = -1
    while True:
        if i != last_row:          # walking backwards, the first visit of a row is its right-most column on the path
            durations[j] += 1
            last_row = i
        if i == 0 and j == 0:
            break
        c = pred[i, j]
        if c == 0:
            j -= 1
        elif c == 1:
            i -= 1
        else:
            i -= 1
            j -= 1
    return durations
