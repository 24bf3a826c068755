"""CPU model of what `k_local_graph_warp` does instead of sorting n^2 weights: drop every weight
<= the cut bound, histogram the rest into 255 bins, walk the bins in chunks (skip entries whose endpoints already
share a segment, sort the rest by (w desc, flat index asc), merge with re-evaluation) and stop as soon as nothing
can change the segment of local vertex 0 any more (S0 rule: the next weight is <= Int(S0) - k/|S0|).  The model must give the reference's connect list (oracle.cut = cutGraphSegmentation,
VS.h:1913-2029) for ANY weight matrix — in particular with heavy ties and arbitrary chunk boundaries, which real
scenes hardly exercise.  (The CUDA kernel itself is compared with the oracle by the -m gpu parity tests.)"""
import numpy as np
import pytest

from oracle import oracle

F = np.float32


def chunked_cut(W, k, rng):
    n = W.shape[0]
    k = F(k)
    lb = F(1.0 - 2.0 * float(k) + float(k) / n - 4e-7 * (n + 8))
    scale = F(255) / max(F(1) - lb, F(1e-3))
    ent = []                                  # (bin, w, f, v1, v2) with f = col * 256 + row, v1 = col, v2 = row
    for a in range(n):
        for b in range(a + 1, n):
            for (row, col) in ((a, b), (b, a)):
                w = F(W[row, col])
                if w > lb:                    # NaN fails the test, like every w <= lb
                    ent.append((min(254, int((F(1) - w) * scale)), w, col * 256 + row, col, row))
    seg = list(range(n))
    size = [1] * n
    thr = [F(1) - k / F(1)] * n
    nseg = n
    ent.sort(key=lambda e: e[0])
    bins = [e[0] for e in ent]
    c0 = 0
    while c0 < 255 and nseg > 1:
        c1 = min(255, c0 + int(rng.integers(1, 40)))            # arbitrary chunk boundary (whole bins)
        chunk = [e for e, b in zip(ent, bins) if c0 <= b < c1 and seg[e[3]] != seg[e[4]]]
        c0 = c1
        if not chunk:
            continue
        chunk.sort(key=lambda e: (-float(e[1]), e[2]))
        if not (chunk[0][1] > thr[seg[0]]):
            break                                               # S0 rule: the segment of vertex 0 cannot merge any more
        for _, w, _, v1, v2 in chunk:
            s1, s2 = seg[v1], seg[v2]
            if s1 == s2:
                continue
            m1, m2 = thr[s1], thr[s2]
            keep, drop, t = (s1, s2, m1) if m1 >= m2 else (s2, s1, m2)
            if w > t:
                seg = [keep if s == drop else s for s in seg]
                size[keep] += size[drop]
                size[drop] = 0
                thr[keep] = w - k / F(size[keep])
                nseg -= 1
    return np.array([v for v in range(n) if seg[v] == seg[0]], np.int32)


@pytest.mark.parametrize("kind", ["ties", "smooth", "nan", "clusters"])
def test_chunked_cut_equals_reference_cut(kind):
    rng = np.random.default_rng({"ties": 1, "smooth": 2, "nan": 3, "clusters": 4}[kind])
    for trial in range(150):
        n = int(rng.integers(2, 28))
        k = float(rng.choice([0.1, 0.3, 0.45]))
        if kind == "ties":
            W = rng.choice(np.array([0.2, 0.5, 0.75, 0.8, 0.9, 0.95, 1.0], F), size=(n, n))
        elif kind == "smooth":
            W = rng.random((n, n), dtype=F)
        elif kind == "nan":
            W = rng.random((n, n), dtype=F)
            W[rng.random((n, n)) < 0.1] = np.nan
        else:                                  # two groups with strong inner and weak outer affinity, asymmetric noise
            g = rng.integers(0, 2, n)
            W = np.where(g[:, None] == g[None, :], 0.9, 0.35).astype(F) + (rng.random((n, n), dtype=F) - F(0.5)) * F(0.2)
        W = np.ascontiguousarray(W, F)
        np.fill_diagonal(W, 1.0)
        ref = np.sort(oracle.cut(W, k))
        got = chunked_cut(W, k, rng)
        assert np.array_equal(ref, got), (kind, trial, n, k)
