// cut_lab.cpp — CPU laboratory for the local-graph stage (NOT product code, NOT the oracle): replays the per-unit
// cut with the device arithmetic of vgs_math.cuh (host build) and counts how many sorted entries each termination
// rule visits.  Used to size the CUDA design (tools/cut_lab.py drives it); results are checked against the oracle's
// connect lists by the caller.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../vgs_svgs_segmentation_b200/csrc/vgs_math.cuh"

using namespace vgs;

struct Ent { float w; int f; int v1, v2; };

extern "C" {
// rec: nu x 16 floats; adj CSR; used flags from rec.  stats (per call, int64[16]):
// 0 units, 1 sum nv, 2 sum pairs, 3 sum kept entries, 4 visited(minthr rule), 5 visited(S0 rule), 6 useful(minthr),
// 7 useful(S0), 8 singleton-shortcut units, 9 units ending nseg==1, 10 sum |S0|, 11 merges total (S0 rule)
// conn_cnt[u] / conn_idx (at adj offsets): result lists under the S0 rule (global ids ascending by local order)
void lab_run(const float* rec, const int64_t* adj_off, const int32_t* adj_idx, int64_t nu, const float* sig6, float k,
             int64_t* stats, int32_t* conn_cnt, int32_t* conn_idx, int32_t* visited_hist /*[101] percent of kept visited*/) {
  PairParams P{sig6[0], sig6[1], sig6[2], sig6[3], sig6[4], sig6[5], 0};
  std::memset(stats, 0, 16 * sizeof(int64_t));
  std::vector<Ent> ent;
  std::vector<int> gid, seg, size;
  std::vector<float> thr;
  for (int64_t u = 0; u < nu; u++) {
    conn_cnt[u] = 0;
    if (!(f2i(rec[u * 16 + REC_FLAGS]) & F_USED)) continue;
    const int64_t off = adj_off[u];
    const int n = (int)(adj_off[u + 1] - off);
    gid.clear();
    for (int i = 0; i < n; i++) {
      const int g = adj_idx[off + i];
      if (f2i(rec[(int64_t)g * 16 + REC_FLAGS]) & F_USED) gid.push_back(g);
    }
    const int nv = (int)gid.size();
    stats[0]++; stats[1] += nv; stats[2] += (int64_t)nv * (nv - 1) / 2;
    const float lb = (float)(1.0 - 2.0 * (double)k + (double)k / (double)n - 4e-7 * (double)(n + 8));
    ent.clear();
    float wmax0 = -1.f;
    for (int a = 0; a < nv; a++)
      for (int b = a + 1; b < nv; b++) {
        float wab, wba;
        pair_weights(rec + (int64_t)gid[a] * 16, rec + (int64_t)gid[b] * 16, P, wab, wba);
        // entry (row i, col j) = weight(idx[i] -> idx[j]); f = col*256+row; v1 = col, v2 = row
        if (wab > lb) ent.push_back(Ent{wab, b * 256 + a, b, a});
        if (wba > lb) ent.push_back(Ent{wba, a * 256 + b, a, b});
        if (a == 0) { wmax0 = std::max(wmax0, std::max(wab, wba)); }
      }
    stats[3] += (int64_t)ent.size();
    std::sort(ent.begin(), ent.end(), [](const Ent& x, const Ent& y) { return x.w > y.w || (x.w == y.w && x.f < y.f); });
    seg.resize(nv); size.assign(nv, 1); thr.assign(nv, 1.0f - k / 1.0f);
    for (int i = 0; i < nv; i++) seg[i] = i;
    int nseg = nv;
    if (!(wmax0 > 1.0f - k / 1.0f)) stats[8]++;
    bool stop_min = false, stop_s0 = false;
    int64_t vis_min = 0, vis_s0 = 0, use_min = 0, use_s0 = 0, merges = 0;
    std::vector<int> seg_at_s0;
    for (size_t e = 0; e < ent.size(); e++) {
      const float w = ent[e].w;
      if (nseg <= 1) { if (!stop_min) stop_min = true; if (!stop_s0) { stop_s0 = true; seg_at_s0 = seg; } break; }
      if (!stop_min) {
        float mt = 3e38f;
        for (int v = 0; v < nv; v++) if (size[v] > 0) mt = std::min(mt, thr[v]);
        if (!(w > mt)) stop_min = true;
      }
      if (!stop_s0) {
        if (!(w > thr[seg[0]])) { stop_s0 = true; seg_at_s0 = seg; }
      }
      if (stop_min) break;   // the minthr rule is the weaker one: nothing can change any more
      const int s1 = seg[ent[e].v1], s2 = seg[ent[e].v2];
      if (!stop_min) vis_min++;
      if (!stop_s0) vis_s0++;
      if (s1 == s2) continue;
      if (!stop_min) use_min++;
      if (!stop_s0) use_s0++;
      const float m1 = thr[s1], m2 = thr[s2];
      const bool a_wins = m1 >= m2;
      if (w > (a_wins ? m1 : m2)) {
        const int keep = a_wins ? s1 : s2, drop = a_wins ? s2 : s1;
        for (int v = 0; v < nv; v++) if (seg[v] == drop) seg[v] = keep;
        size[keep] += size[drop]; size[drop] = 0;
        thr[keep] = w - k / (float)size[keep];
        nseg--;
        if (!stop_s0) merges++;
      }
    }
    if (!stop_s0) seg_at_s0 = seg;
    stats[4] += vis_min; stats[5] += vis_s0; stats[6] += use_min; stats[7] += use_s0; stats[11] += merges;
    if (nseg <= 1) stats[9]++;
    // S0 under the early rule must equal S0 at the very end
    int c = 0;
    for (int v = 0; v < nv; v++) if (seg_at_s0[v] == seg_at_s0[0]) { conn_idx[off + c++] = gid[v]; }
    int c_end = 0;
    for (int v = 0; v < nv; v++) if (seg[v] == seg[0]) c_end++;
    if (c_end != c) stats[15]++;   // mismatch counter (must stay 0)
    conn_cnt[u] = c;
    stats[10] += c;
    if (!ent.empty()) visited_hist[std::min<int64_t>(100, 100 * vis_s0 / (int64_t)ent.size())]++;
  }
}
}
