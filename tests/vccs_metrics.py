"""Segmentation-quality metrics for supervoxels on the synthetic scenes (ground truth = index of the surface a point
was sampled from): corrected undersegmentation error (Neubert & Protzel) and boundary recall on the voxel lattice
(Papon et al. 2013 evaluate VCCS with these two)."""
import numpy as np


def undersegmentation_error(sv, gt):
    """(1/N) * sum over supervoxels s and ground-truth regions g of min(|s & g|, |s - g|); unlabelled points (sv 0) are skipped."""
    m = sv > 0
    s, g = sv[m].astype(np.int64), gt[m].astype(np.int64)
    pair, cnt = np.unique(s * (int(g.max()) + 1) + g, return_counts=True)
    size = np.bincount(s)[pair // (int(g.max()) + 1)]
    return float(np.minimum(cnt, size - cnt).sum()) / float(m.sum())


def purity(sv, gt):
    m = sv > 0
    s, g = sv[m].astype(np.int64), gt[m].astype(np.int64)
    k = int(g.max()) + 1
    pair, cnt = np.unique(s * k + g, return_counts=True)
    best = np.zeros(int(s.max()) + 1, np.int64)
    np.maximum.at(best, pair // k, cnt)
    return float(best.sum()) / float(m.sum())


def _voxel_majority(point_voxel, lab, n_vox):
    m = point_voxel >= 0
    k = int(lab.max()) + 1
    pair, cnt = np.unique(point_voxel[m].astype(np.int64) * k + lab[m], return_counts=True)
    order = np.lexsort((cnt, pair // k))
    v = (pair // k)[order]
    last = np.r_[v[1:] != v[:-1], True]
    out = np.full(n_vox, -1, np.int64)
    out[v[last]] = (pair % k)[order][last]
    return out


def boundary_recall(unit_key, point_voxel, sv, gt, tolerance=0):
    """Fraction of ground-truth boundary voxels that are supervoxel boundary voxels themselves (tolerance 0) or have one
    in their 26-neighbourhood (tolerance 1).  Supervoxels of ~5x5 voxels put a boundary within one voxel of almost
    everything, so only tolerance 0 discriminates."""
    V = unit_key.shape[0]
    key = unit_key.astype(np.int64)
    packed = (key[:, 0] << 42) | (key[:, 1] << 21) | key[:, 2]
    order = np.argsort(packed)
    sp = packed[order]
    gv = _voxel_majority(point_voxel, gt, V)
    svv = _voxel_majority(point_voxel, sv, V)
    gb = np.zeros(V, bool)
    sb = np.zeros(V, bool)
    nbs = []
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                q = ((key[:, 0] + dx) << 42) | ((key[:, 1] + dy) << 21) | (key[:, 2] + dz)
                pos = np.searchsorted(sp, q)
                pos[pos >= V] = V - 1
                hit = sp[pos] == q
                nb = np.where(hit, order[pos], -1)
                nbs.append(nb)
                ok = nb >= 0
                gb |= ok & (gv[np.maximum(nb, 0)] != gv)
                sb |= ok & (svv[np.maximum(nb, 0)] != svv)
    near = sb.copy()
    if tolerance:
        for nb in nbs:
            near |= (nb >= 0) & sb[np.maximum(nb, 0)]
    return float((gb & near).sum()) / float(max(1, gb.sum()))
