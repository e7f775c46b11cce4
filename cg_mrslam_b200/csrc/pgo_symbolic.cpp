// Structure analysis of the pose-graph solver. See pgo_symbolic.h.
#include "pgo_symbolic.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <numeric>

namespace pgo {

namespace {

struct Graph {
  int n;
  std::vector<int> ptr, adj;  // CSR, symmetric, no self loops, no duplicates
};

Graph build_graph(int n, const std::vector<std::pair<int, int> >& edges) {
  Graph g;
  g.n = n;
  std::vector<int> deg(n + 1, 0);
  for (size_t e = 0; e < edges.size(); ++e) {
    const int a = edges[e].first, b = edges[e].second;
    if (a == b) continue;
    ++deg[a + 1];
    ++deg[b + 1];
  }
  g.ptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) g.ptr[i + 1] = g.ptr[i] + deg[i + 1];
  std::vector<int> tmp(g.ptr[n]);
  std::vector<int> cur(g.ptr.begin(), g.ptr.end() - 1);
  for (size_t e = 0; e < edges.size(); ++e) {
    const int a = edges[e].first, b = edges[e].second;
    if (a == b) continue;
    tmp[cur[a]++] = b;
    tmp[cur[b]++] = a;
  }
  // sort + unique per row (parallel edges between the same pair collapse)
  g.adj.clear();
  g.adj.reserve(tmp.size());
  std::vector<int> new_ptr(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    std::sort(tmp.begin() + g.ptr[i], tmp.begin() + g.ptr[i + 1]);
    int last = -1;
    for (int t = g.ptr[i]; t < g.ptr[i + 1]; ++t)
      if (tmp[t] != last) {
        g.adj.push_back(tmp[t]);
        last = tmp[t];
      }
    new_ptr[i + 1] = static_cast<int>(g.adj.size());
  }
  g.ptr.swap(new_ptr);
  return g;
}

// ---- nested dissection by breadth-first level structures -----------------------------------
// each side of a cut keeps at least this share of the vertices: tighter while ranks are being
// carved out (load balance), looser below (smaller separators, less fill)
const double kBalanceDeciding = 0.38, kBalanceFill = 0.30;
const int kLeafSize = 4;

class Dissector {
 public:
  Dissector(const Graph& g, int leaf, int world)
      : g_(g), leaf_(leaf), mark_(g.n, -1), dist_(g.n, -1), owner_(g.n, 0), load_(world, 0) {
    for (int f = 0; f < 4; ++f) d4_[f].assign(g.n, -1);
    pos_.assign(g.n, 0);
    part_.assign(g.n, 0);
    lock_.assign(g.n, 0);
    nbc_[0].assign(g.n, 0);
    nbc_[1].assign(g.n, 0);
    cut_depth_ = 0;
    while ((1 << cut_depth_) < world) ++cut_depth_;
  }

  // owner (per original vertex): rank, or -1 for the shared separators of the top cut_depth levels
  const std::vector<int>& owner() const { return owner_; }
  const std::vector<std::vector<int> >& groups() const { return groups_; }

  void run(std::vector<int>* order) {
    order_ = order;
    order_->clear();
    order_->reserve(g_.n);
    std::vector<int> all(g_.n);
    std::iota(all.begin(), all.end(), 0);
    rec(all, 0, 0);
  }

 private:
  // BFS inside the current subset (mark_ == id) from `start`; fills dist_, returns visit order.
  void bfs(int start, int id, std::vector<int>* visit) {
    visit->clear();
    visit->push_back(start);
    dist_[start] = 0;
    for (size_t h = 0; h < visit->size(); ++h) {
      const int v = (*visit)[h];
      for (int t = g_.ptr[v]; t < g_.ptr[v + 1]; ++t) {
        const int w = g_.adj[t];
        if (mark_[w] == id && dist_[w] < 0) {
          dist_[w] = dist_[v] + 1;
          visit->push_back(w);
        }
      }
    }
  }

  // Everything in `s` goes to one rank of the range this branch still owns: the least loaded.
  void assign(const std::vector<int>& s, int depth, int path) {
    const int span = 1 << (cut_depth_ - depth), first = path << (cut_depth_ - depth);
    int best = first;
    for (int r = first; r < first + span; ++r)
      if (load_[r] < load_[best]) best = r;
    load_[best] += static_cast<long long>(s.size());
    for (size_t i = 0; i < s.size(); ++i) owner_[s[i]] = best;
  }

  // depth < cut_depth_: this call still decides ownership; path = branch index at that depth.
  // depth == cut_depth_: s already belongs to a rank. depth > cut_depth_ never occurs.
  void rec(std::vector<int>& s, int depth, int path) {
    const bool deciding = depth < cut_depth_;
    if (static_cast<int>(s.size()) <= leaf_) {
      if (deciding) assign(s, depth, path);
      emit_leaf(s);
      return;
    }
    const int id = next_id_++;
    for (size_t i = 0; i < s.size(); ++i) {
      mark_[s[i]] = id;
      dist_[s[i]] = -1;
    }
    // connected components first
    std::vector<int> visit;
    bfs(s[0], id, &visit);
    if (visit.size() != s.size()) {
      std::vector<std::vector<int> > comps;
      comps.push_back(visit);
      for (size_t i = 0; i < s.size(); ++i)
        if (dist_[s[i]] < 0) {
          bfs(s[i], id, &visit);
          comps.push_back(visit);
        }
      if (!deciding) {
        for (size_t c = 0; c < comps.size(); ++c) rec(comps[c], depth, path);
        return;
      }
      // still deciding ownership: a dominant component keeps being dissected across this
      // branch's ranks; the others go (whole) to the least loaded rank of the branch
      size_t big = 0;
      for (size_t c = 1; c < comps.size(); ++c)
        if (comps[c].size() > comps[big].size()) big = c;
      const bool dominant = comps[big].size() * 10 >= s.size() * 6;
      if (dominant) rec(comps[big], depth, path);
      for (size_t c = 0; c < comps.size(); ++c) {
        if (dominant && c == big) continue;
        assign(comps[c], depth, path);
        rec(comps[c], cut_depth_, 0);
      }
      return;
    }
    const double bal = deciding ? kBalanceDeciding : kBalanceFill;
    std::vector<int> a, b, sep;
    if (!separate(s, id, bal, &a, &b, &sep) || a.empty() || b.empty()) {  // (nearly) a clique
      if (deciding) assign(s, depth, path);
      order_->insert(order_->end(), s.begin(), s.end());
      return;
    }
    { std::vector<int>().swap(s); }  // release before recursing
    if (deciding) {
      for (size_t i = 0; i < sep.size(); ++i) owner_[sep[i]] = -1;  // shared separator
      if (depth + 1 == cut_depth_) {  // the two halves are whole ranks' interiors
        assign(a, depth + 1, 2 * path);
        assign(b, depth + 1, 2 * path + 1);
      }
      rec(a, depth + 1, 2 * path);
      rec(b, depth + 1, 2 * path + 1);
    } else {
      rec(a, depth, path);
      rec(b, depth, path);
    }
    order_->insert(order_->end(), sep.begin(), sep.end());
  }


  // A leaf of the dissection is eliminated as ONE supernode: its connected pieces are recorded as
  // groups, and analyse() makes every column of a group carry the union of the group's boundary
  // (explicit zero blocks inside the leaf's columns; no fill is added above the leaf, whose boundary
  // becomes a clique either way). One dense trapezoid per leaf instead of a handful of one-column
  // panels: a fraction of the tasks and of the scattered atomic updates.
  void emit_leaf(const std::vector<int>& s) {
    const int id = next_id_++;
    for (size_t i = 0; i < s.size(); ++i) {
      mark_[s[i]] = id;
      dist_[s[i]] = -1;
    }
    std::vector<int> visit;
    for (size_t i = 0; i < s.size(); ++i) {
      if (dist_[s[i]] >= 0) continue;
      bfs(s[i], id, &visit);
      order_->insert(order_->end(), visit.begin(), visit.end());
      if (visit.size() > 1) groups_.push_back(visit);
    }
  }

  // BFS inside subset `id` from `start` writing hop counts into `d` (entries of the subset must
  // be -1 on entry); returns the last vertex reached.
  int bfs_dist(int start, int id, std::vector<int>& d) {
    queue_.clear();
    queue_.push_back(start);
    d[start] = 0;
    for (size_t h = 0; h < queue_.size(); ++h) {
      const int v = queue_[h];
      for (int t = g_.ptr[v]; t < g_.ptr[v + 1]; ++t) {
        const int w = g_.adj[t];
        if (mark_[w] == id && d[w] < 0) {
          d[w] = d[v] + 1;
          queue_.push_back(w);
        }
      }
    }
    return queue_.back();
  }

  // Vertex separator of the connected subset s (mark_ == id).
  //   1. four hop-distance fields: from a pseudo-peripheral pair (p, q) and from a second pair
  //      (u, w) chosen off the p-q axis; the differences x = d_p - d_q, y = d_u - d_w act as two
  //      coordinates of an embedding. Candidate sweeps: x, y, x + y, x - y, d_p, d_q.
  //   2. per sweep: sort, and for every cut position inside the balance window count the
  //      boundary vertices of either side (difference arrays); keep the smallest.
  //   3. vertex Fiduccia-Mattheyses refinement of the winner (move a separator vertex to one
  //      side, pull its neighbours on the other side into the separator; hill-climbing with
  //      roll-back).
  bool separate(const std::vector<int>& s, int id, double bal, std::vector<int>* a,
                std::vector<int>* b, std::vector<int>* sep) {
    const int n = static_cast<int>(s.size());
    for (int f = 0; f < 4; ++f)
      for (int i = 0; i < n; ++i) d4_[f][s[i]] = -1;
    const int p = bfs_dist(s[0], id, d4_[0]);      // far from an arbitrary start
    for (int i = 0; i < n; ++i) d4_[0][s[i]] = -1;
    const int q = bfs_dist(p, id, d4_[0]);         // d4_[0] = hops from p; q is farthest from p
    if (d4_[0][q] < 2) return false;
    bfs_dist(q, id, d4_[1]);
    int u = s[0];
    long long best_sum = -1;
    for (int i = 0; i < n; ++i) {
      const int v = s[i];
      // far from the p-q axis, and not next to either end
      const long long off = static_cast<long long>(d4_[0][v]) + d4_[1][v] -
                            std::abs(d4_[0][v] - d4_[1][v]);
      if (off > best_sum) {
        best_sum = off;
        u = v;
      }
    }
    // Graphs of the keyframe loop (hundreds of poses, re-analysed whenever a vertex is added) take the
    // quick route: one axis only (x = d_p - d_q, d_p, d_q as sweeps). The second axis and its three
    // sweeps cost as much again and buy nothing measurable on graphs this small.
    const bool quick = g_.n <= kQuickOrderingVertices;
    if (quick) {
      for (int i = 0; i < n; ++i) d4_[2][s[i]] = d4_[3][s[i]] = 0;
    } else {
      const int w = bfs_dist(u, id, d4_[2]);
      bfs_dist(w, id, d4_[3]);
    }

    int lo = std::max(1, static_cast<int>(bal * n)), hi = std::min(n - 1, n - lo);
    // sort keys packed into one word: primary key, secondary key, position in s (all three are
    // small: hop-distance differences and a subset index)
    std::vector<uint64_t> keyed(n);
    std::vector<int> diff_a(n + 2), diff_b(n + 2), best_order;
    int best_cost = n + 1, best_k = -1;
    bool best_from_a = true;
    int best_cand = 0;
    const long long kBias = 1LL << 20;   // |keys| < 2^20: hop counts of a subset of < 2^19 vertices, doubled
    auto sweep = [&](int cand) {
      for (int i = 0; i < n; ++i) {
        const int v = s[i];
        const long long x = d4_[0][v] - d4_[1][v], y = d4_[2][v] - d4_[3][v];
        long long k1, k2;
        switch (cand) {
          case 0: k1 = x; k2 = y; break;
          case 1: k1 = y; k2 = x; break;
          case 2: k1 = x + y; k2 = x - y; break;
          case 3: k1 = x - y; k2 = x + y; break;
          case 4: k1 = d4_[0][v]; k2 = y; break;
          default: k1 = d4_[1][v]; k2 = y; break;
        }
        keyed[i] = (static_cast<uint64_t>(k1 + kBias) << 42) | (static_cast<uint64_t>(k2 + kBias) << 21) |
                   static_cast<uint64_t>(i);
      }
      std::sort(keyed.begin(), keyed.end());
      for (int i = 0; i < n; ++i) keyed[i] = static_cast<uint64_t>(s[keyed[i] & ((1u << 21) - 1)]);
    };
    // no admissible cut inside the balance window (small, dense subsets): widen the window
    for (int attempt = 0; attempt < 3 && best_k < 0;
         ++attempt, lo = attempt == 1 ? std::max(1, lo / 2) : 1, hi = n - lo)
      for (int cand = 0; cand < 6; ++cand) {
        if (quick && cand >= 1 && cand <= 3) continue;   // y == 0: these repeat sweep 0
        sweep(cand);
        for (int i = 0; i < n; ++i) pos_[keyed[i]] = i;
        std::fill(diff_a.begin(), diff_a.end(), 0);
        std::fill(diff_b.begin(), diff_b.end(), 0);
        for (int i = 0; i < n; ++i) {
          const int v = static_cast<int>(keyed[i]);
          int mx = i, mn = i;
          for (int t = g_.ptr[v]; t < g_.ptr[v + 1]; ++t) {
            const int x = g_.adj[t];
            if (mark_[x] != id) continue;
            mx = std::max(mx, pos_[x]);
            mn = std::min(mn, pos_[x]);
          }
          // with A = [0, k): v is on A's boundary for i < k <= mx, on B's boundary for mn < k <= i
          ++diff_a[i + 1];
          --diff_a[mx + 1];
          ++diff_b[mn + 1];
          --diff_b[i + 1];
        }
        int ca = 0, cb = 0;
        for (int k = 1; k <= hi; ++k) {
          ca += diff_a[k];
          cb += diff_b[k];
          if (k < lo) continue;
          // the separator is taken out of one side: that side must stay above the balance bound
          const int cost_a = k - ca >= lo ? ca : n + 1, cost_b = n - k - cb >= lo ? cb : n + 1;
          const int c = std::min(cost_a, cost_b);
          if (c < best_cost) {
            best_cost = c;
            best_k = k;
            best_from_a = cost_a <= cost_b;
            best_cand = cand;
          }
        }
      }
    if (best_k < 0) return false;
    sweep(best_cand);
    best_order.resize(n);
    for (int i = 0; i < n; ++i) best_order[i] = static_cast<int>(keyed[i]);
    // part: 0 = A, 1 = B, 2 = separator
    for (int i = 0; i < n; ++i) part_[best_order[i]] = i < best_k ? 0 : 1;
    for (int i = 0; i < n; ++i) {
      const int v = best_order[i];
      if ((i < best_k) != best_from_a) continue;
      const int other = best_from_a ? 1 : 0;
      for (int t = g_.ptr[v]; t < g_.ptr[v + 1]; ++t) {
        const int x = g_.adj[t];
        if (mark_[x] == id && part_[x] == other) {
          part_[v] = 2;
          break;
        }
      }
    }
    refine(s, id, lo);
    a->clear();
    b->clear();
    sep->clear();
    for (int i = 0; i < n; ++i) {
      const int v = best_order[i];
      (part_[v] == 0 ? a : part_[v] == 1 ? b : sep)->push_back(v);
    }
    return !a->empty() && !b->empty();
  }

  // Vertex-separator FM on part_ (0 / 1 / 2) of subset `id`; both sides keep >= lo vertices.
  void refine(const std::vector<int>& s, int id, int lo) {
    const int n = static_cast<int>(s.size());
    std::vector<int> sepv;
    int size[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
      ++size[part_[s[i]]];
      if (part_[s[i]] == 2) sepv.push_back(s[i]);
      lock_[s[i]] = 0;
    }
    struct Move { int v, to, n_pulled; };
    std::vector<Move> log;
    std::vector<int> pulled;
    // neighbours of a separator vertex on side 0 / 1, kept up to date move by move
    auto count_sides = [&](int v) {
      int nb[2] = {0, 0};
      for (int t = g_.ptr[v]; t < g_.ptr[v + 1]; ++t) {
        const int x = g_.adj[t];
        if (mark_[x] == id && part_[x] < 2) ++nb[part_[x]];
      }
      nbc_[0][v] = nb[0];
      nbc_[1][v] = nb[1];
    };
    for (int pass = 0; pass < 8; ++pass) {
      log.clear();
      pulled.clear();
      for (size_t i = 0; i < sepv.size(); ++i) count_sides(sepv[i]);
      const int start_sep = size[2];
      int best_sep = size[2], best_at = 0, best_imb = std::abs(size[0] - size[1]);
      const int patience = 40 + size[2] / 4;
      std::vector<int> touched;
      while (true) {
        int bv = -1, bto = 0, bgain = -(1 << 30), bslot = -1;
        for (size_t i = 0; i < sepv.size(); ++i) {
          const int v = sepv[i];
          if (part_[v] != 2 || lock_[v]) continue;
          for (int to = 0; to < 2; ++to) {
            const int pull = nbc_[1 - to][v];
            if (size[1 - to] - pull < lo) continue;
            const int gain = 1 - pull;
            // ties: towards the smaller side
            const bool better = gain > bgain || (gain == bgain && size[to] < size[bto]);
            if (better) {
              bgain = gain;
              bv = v;
              bto = to;
              bslot = static_cast<int>(i);
            }
          }
        }
        if (bv < 0) break;
        (void)bslot;
        // apply
        part_[bv] = bto;
        lock_[bv] = 1;
        touched.push_back(bv);
        ++size[bto];
        --size[2];
        int n_pulled = 0;
        for (int t = g_.ptr[bv]; t < g_.ptr[bv + 1]; ++t) {
          const int x = g_.adj[t];
          if (mark_[x] != id) continue;
          if (part_[x] == 2) {
            ++nbc_[bto][x];  // bv now sits on side bto
          } else if (part_[x] == 1 - bto) {
            part_[x] = 2;
            --size[1 - bto];
            ++size[2];
            sepv.push_back(x);
            pulled.push_back(x);
            ++n_pulled;
          }
        }
        // the pulled vertices left side 1 - bto: their separator neighbours lose one there, and
        // they get counts of their own
        for (int k = 0; k < n_pulled; ++k) {
          const int x = pulled[pulled.size() - 1 - k];
          for (int t = g_.ptr[x]; t < g_.ptr[x + 1]; ++t) {
            const int z = g_.adj[t];
            if (mark_[z] == id && part_[z] == 2) --nbc_[1 - bto][z];
          }
        }
        for (int k = 0; k < n_pulled; ++k) count_sides(pulled[pulled.size() - 1 - k]);
        const Move m = {bv, bto, n_pulled};
        log.push_back(m);
        const int imb = std::abs(size[0] - size[1]);
        if (size[2] < best_sep || (size[2] == best_sep && imb < best_imb)) {
          best_sep = size[2];
          best_imb = imb;
          best_at = static_cast<int>(log.size());
        }
        if (static_cast<int>(log.size()) - best_at > patience) break;
      }
      // roll back to the best prefix
      while (static_cast<int>(log.size()) > best_at) {
        const Move m = log.back();
        log.pop_back();
        for (int k = 0; k < m.n_pulled; ++k) {
          const int x = pulled.back();
          pulled.pop_back();
          part_[x] = 1 - m.to;
          ++size[1 - m.to];
          --size[2];
        }
        part_[m.v] = 2;
        --size[m.to];
        ++size[2];
      }
      for (size_t i = 0; i < touched.size(); ++i) lock_[touched[i]] = 0;
      // compact the separator list
      size_t wr = 0;
      for (size_t i = 0; i < sepv.size(); ++i)
        if (part_[sepv[i]] == 2 && !lock_[sepv[i]]) {
          lock_[sepv[i]] = 2;  // de-duplicate
          sepv[wr++] = sepv[i];
        }
      sepv.resize(wr);
      for (size_t i = 0; i < sepv.size(); ++i) lock_[sepv[i]] = 0;
      if (size[2] >= start_sep && pass > 0) break;
      if (best_at == 0) break;
    }
  }

  const Graph& g_;
  int leaf_;
  std::vector<int> mark_, dist_, owner_;
  std::vector<int> d4_[4], pos_, part_, lock_, queue_, nbc_[2];
  std::vector<std::vector<int> > groups_;
  std::vector<long long> load_;
  int cut_depth_ = 0;
  std::vector<int>* order_ = nullptr;
  int next_id_ = 0;
};

// Supernodes, panels, scatter tables and task lists of the single-GPU path (pgo_symbolic.h).
template <typename Key>
void bucket_tasks(std::vector<std::pair<Key, Task> >& in, int n_levels, std::vector<int>* ptr,
                  std::vector<Task>* out) {
  ptr->assign(n_levels + 1, 0);
  for (size_t i = 0; i < in.size(); ++i) ++(*ptr)[in[i].first + 1];
  for (int l = 0; l < n_levels; ++l) (*ptr)[l + 1] += (*ptr)[l];
  out->resize(in.size());
  std::vector<int> cur(ptr->begin(), ptr->end() - 1);
  for (size_t i = 0; i < in.size(); ++i) (*out)[cur[in[i].first]++] = in[i].second;
  std::vector<std::pair<Key, Task> >().swap(in);
}

bool build_supernodal(Symbolic& S, std::string* err) {
  Supernodal& N = S.sn;
  N = Supernodal();
  const int n = S.n;
  const std::vector<int>& cp = S.col_ptr;
  const std::vector<int>& ri = S.row_idx;
  // supernodes: maximal chains with nested structure and one owner
  for (int p = 0; p < n;) {
    int q = p;
    while (q + 1 < n && q + 1 - p < kMaxSuperWidth && S.parent[q] == q + 1 &&
           S.owner[q] == S.owner[q + 1] && cp[q + 1] - cp[q] == cp[q + 2] - cp[q + 1] + 1)
      ++q;
    N.sn_first.push_back(p);
    p = q + 1;
  }
  N.sn_first.push_back(n);
  N.n_super = static_cast<int>(N.sn_first.size()) - 1;
  // panels: balanced pieces of at most kPanelWidth columns
  std::vector<int> pn_of(n), sn_of(n);
  N.sn_pn_ptr.assign(1, 0);
  for (int s = 0; s < N.n_super; ++s) {
    const int c0 = N.sn_first[s], W = N.sn_first[s + 1] - c0;
    const int pieces = (W + kPanelWidth - 1) / kPanelWidth;
    for (int k = 0; k < pieces; ++k) {
      const int a = c0 + static_cast<int>(static_cast<int64_t>(W) * k / pieces);
      const int b = c0 + static_cast<int>(static_cast<int64_t>(W) * (k + 1) / pieces);
      for (int c = a; c < b; ++c) {
        pn_of[c] = static_cast<int>(N.pn_first.size());
        sn_of[c] = s;
      }
      N.pn_first.push_back(a);
      N.pn_sn.push_back(s);
    }
    N.sn_pn_ptr.push_back(static_cast<int>(N.pn_first.size()));
  }
  N.pn_first.push_back(n);
  N.n_panels = static_cast<int>(N.pn_first.size()) - 1;

  // scatter tables + panel levels (a source panel always precedes its targets)
  std::vector<int> plevel(N.n_panels, 0);
  N.pn_meta.assign(N.n_panels + 1, 0);
  N.pn_scratch.assign(N.n_panels, -1);
  std::vector<std::pair<int, Task> > ff, fa, fb;
  for (int K = 0; K < N.n_panels; ++K) {
    const int c0 = N.pn_first[K], w = N.pn_first[K + 1] - c0;
    const int base = cp[c0] + w, m = cp[c0 + 1] - base;  // below rows: ri[base .. base + m)
    N.pn_meta[K] = static_cast<int>(N.colbase.size());
    for (int b = 0; b < m;) {
      const int q = pn_of[ri[base + b]];
      const int q0 = N.pn_first[q], q1 = N.pn_first[q + 1];
      const int b_begin = b;
      while (b < m && ri[base + b] < q1) ++b;
      const int64_t off = static_cast<int64_t>(N.tbl.size());
      if (off + m > 0x7FFFFFF0LL) {
        if (err) *err = "scatter tables exceed 2^31 entries: graph too dense for this solver";
        return false;
      }
      const int* rows_q = ri.data() + cp[q0];
      const int len_q = cp[q0 + 1] - cp[q0];
      int at = 0;
      for (int a = b_begin; a < m; ++a) {
        const int r = ri[base + a];
        while (at < len_q && rows_q[at] != r) ++at;
        if (at == len_q) {
          if (err) *err = "internal: a source row is missing from its target panel";
          return false;
        }
        N.tbl.push_back(at);
      }
      for (int t = b_begin; t < b; ++t) {
        const int rb = ri[base + t];
        N.colbase.push_back(cp[rb] - (rb - q0));
        N.tbl_off.push_back(static_cast<int>(off) - b_begin);
      }
      plevel[q] = std::max(plevel[q], plevel[K] + 1);
    }
    N.update_blocks += static_cast<int64_t>(m) * (m + 1) / 2;
    const double Ws = 3.0 * w, Ms = 3.0 * m;
    N.flops += Ws * Ws * Ws / 3.0 + Ms * Ws * Ws + Ms * Ms * Ws;
    N.n_plevels = std::max(N.n_plevels, plevel[K] + 1);
    if (w <= kSmallWidth && m <= kSmallRows) {
      const Task t = {K, 0, m, 0};
      ff.push_back(std::make_pair(plevel[K], t));
    } else {
      int chunks = 0;
      for (int r0 = 0; r0 == 0 || r0 < m; r0 += kRowChunk, ++chunks) {
        const Task t = {K, r0, std::min(m, r0 + kRowChunk), 0};
        fa.push_back(std::make_pair(plevel[K], t));
      }
      if (chunks > 1) {
        N.pn_scratch[K] = static_cast<int>(N.scratch_blocks);
        N.scratch_blocks += static_cast<int64_t>(w) * w;
      }
      // Outer products. Inside a supernode every panel updates the supernode's later columns
      // (the next panel needs them); the update of the ANCESTORS (rows x rows below the
      // supernode) is applied once per supernode, by its last panel's level, with all the
      // supernode's columns as the inner dimension.
      const int sn_s = N.pn_sn[K];
      const int W = N.sn_first[sn_s + 1] - N.sn_first[sn_s], o = c0 - N.sn_first[sn_s];
      const int w_rest = W - o - w;                         // later columns of the supernode
      const int np = N.sn_pn_ptr[sn_s + 1] - N.sn_pn_ptr[sn_s];
      for (int pass = 0; pass < 2; ++pass) {
        // pass 0: inside the supernode (columns b < w_rest); pass 1 (last panel): the ancestors
        if (pass == 0 ? w_rest == 0 : w_rest != 0) continue;
        const int jlim = pass == 0 ? w_rest : m;
        const int wk = pass == 0 || np == 1 ? w : kPanelWidth;   // widest staged chunk
        const int tj = std::min((jlim + 7) / 8 * 8, 16);
        int ti = std::min((m + 7) / 8 * 8, 16);                  // two 8-row strips per warp (4 warps)
        while (ti > 8 && sn_tile_doubles(wk, ti, tj) > kTileSmemDoubles) ti -= 8;
        // the inner dimension of an ancestors' update is cut into groups of kUpdateGroup panels
        // (split-K: the tiles of different groups add into the same targets with atomics anyway)
        const int n_groups = pass == 0 ? 1 : (np + kUpdateGroup - 1) / kUpdateGroup;
        for (int gi = 0; gi < n_groups; ++gi) {
          const int first = pass == 0 ? 0 : np * gi / n_groups, end = pass == 0 ? 1 : np * (gi + 1) / n_groups;
          const int count = end - first, back = pass == 0 ? 0 : np - end;
          const int aux = ti | (tj << 8) | (count << 16) | (back << 22) | ((pass == 0 ? 1 : 0) << 28);
          for (int i0 = 0; i0 < m; i0 += ti)
            for (int j0 = 0; j0 < std::min(jlim, i0 + ti); j0 += tj) {
              const Task t = {K, i0, j0, aux};
              fb.push_back(std::make_pair(plevel[K], t));
            }
        }
      }
    }
    const int sn_id = N.pn_sn[K];
    PanelDesc pd = {c0, w, m, cp[c0], N.pn_meta[K], N.pn_scratch[K], c0 - N.sn_first[sn_id], sn_id};
    N.pn.push_back(pd);
    N.pn_owner.push_back(S.owner[c0]);
  }
  N.pn_meta[N.n_panels] = static_cast<int>(N.colbase.size());
  bucket_tasks(ff, N.n_plevels, &N.ff_ptr, &N.ff);
  bucket_tasks(fa, N.n_plevels, &N.fa_ptr, &N.fa);
  bucket_tasks(fb, N.n_plevels, &N.fb_ptr, &N.fb);

  // substitution: supernode levels and tasks
  std::vector<int> slevel(N.n_super, 0);
  std::vector<std::pair<int, Task> > ss, sa, sf, sb;
  for (int s = 0; s < N.n_super; ++s) {
    const int c0 = N.sn_first[s], W = N.sn_first[s + 1] - c0;
    const int base = cp[c0] + W, m = cp[c0 + 1] - base;
    for (int b = 0; b < m; ++b) {
      const int q = sn_of[ri[base + b]];
      slevel[q] = std::max(slevel[q], slevel[s] + 1);
    }
    N.n_slevels = std::max(N.n_slevels, slevel[s] + 1);
    SuperDesc sd = {c0, W, m, cp[c0], N.sn_pn_ptr[s], N.sn_pn_ptr[s + 1], 0, 0};
    N.sn.push_back(sd);
    N.sn_owner.push_back(S.owner[c0]);
    if (W <= kSubstWarpWidth) {
      const Task t = {s, 0, m, 0};
      ss.push_back(std::make_pair(slevel[s], t));
      continue;
    }
    const Task t = {s, 0, m, 0};
    sa.push_back(std::make_pair(slevel[s], t));
    for (int p = N.sn_pn_ptr[s]; p < N.sn_pn_ptr[s + 1]; ++p) {
      for (int r0 = 0; r0 < m; r0 += 32) {
        const Task x = {p, r0, std::min(m, r0 + 32), 0};
        sf.push_back(std::make_pair(slevel[s], x));
      }
      for (int r0 = 0; r0 < m; r0 += 64) {
        const Task x = {p, r0, std::min(m, r0 + 64), 0};
        sb.push_back(std::make_pair(slevel[s], x));
      }
    }
  }
  bucket_tasks(ss, N.n_slevels, &N.ss_ptr, &N.ss);
  bucket_tasks(sa, N.n_slevels, &N.sa_ptr, &N.sa);
  bucket_tasks(sf, N.n_slevels, &N.sf_ptr, &N.sf);
  bucket_tasks(sb, N.n_slevels, &N.sb_ptr, &N.sb);
  return true;
}

}  // namespace

Supernodal::Lists Supernodal::lists(int owner) const {
  Lists L;
  L.n_plevels = n_plevels;
  L.n_slevels = n_slevels;
  // filter one task list by the owner of its panel / supernode, keeping the level buckets
  auto filter = [&](const std::vector<int>& ptr, const std::vector<Task>& in, const std::vector<int>& own,
                    int n_levels, std::vector<int>* out_ptr, std::vector<Task>* out) {
    out_ptr->assign(n_levels + 1, 0);
    out->clear();
    for (int l = 0; l < n_levels; ++l) {
      for (int i = ptr[l]; i < ptr[l + 1]; ++i)
        if (owner == kAllOwners || own[in[i].id] == owner) out->push_back(in[i]);
      (*out_ptr)[l + 1] = static_cast<int>(out->size());
    }
  };
  filter(ff_ptr, ff, pn_owner, n_plevels, &L.ff_ptr, &L.ff);
  filter(fa_ptr, fa, pn_owner, n_plevels, &L.fa_ptr, &L.fa);
  filter(fb_ptr, fb, pn_owner, n_plevels, &L.fb_ptr, &L.fb);
  filter(ss_ptr, ss, sn_owner, n_slevels, &L.ss_ptr, &L.ss);
  filter(sa_ptr, sa, sn_owner, n_slevels, &L.sa_ptr, &L.sa);
  filter(sf_ptr, sf, pn_owner, n_slevels, &L.sf_ptr, &L.sf);
  filter(sb_ptr, sb, pn_owner, n_slevels, &L.sb_ptr, &L.sb);
  const int pair_doubles = (kPanelWidth * (kPanelWidth + 1) / 2 + 1) / 2;
  L.fa_smem.assign(n_plevels, 0);
  L.fa_smem_small.assign(n_plevels, 0);
  L.fa_large.assign(n_plevels, 0);
  L.fb_smem.assign(n_plevels, 0);
  L.ff_smem.assign(n_plevels, 0);
  L.ff_smem_small.assign(n_plevels, 0);
  L.ff_large.assign(n_plevels, 0);
  auto fused_need = [&](const Task& t) { return sn_fused_doubles(pn[t.id].w, pn[t.id].m); };
  for (int l = 0; l < n_plevels; ++l) {
    std::stable_partition(L.ff.begin() + L.ff_ptr[l], L.ff.begin() + L.ff_ptr[l + 1],
                          [&](const Task& t) { return fused_need(t) > kFusedSmallDoubles; });
    for (int i = L.ff_ptr[l]; i < L.ff_ptr[l + 1]; ++i) {
      const int need = fused_need(L.ff[i]);
      if (need > kFusedSmallDoubles) {
        ++L.ff_large[l];
        L.ff_smem[l] = std::max(L.ff_smem[l], need);
      } else {
        L.ff_smem_small[l] = std::max(L.ff_smem_small[l], need);
      }
    }
    auto factor_need = [&](const Task& t) {
      const int w = pn[t.id].w;
      return w * w * 9 + w * 9 + pair_doubles + 3 * w + 3 * w * (3 * (t.r1 - t.r0) + 1);
    };
    std::stable_partition(L.fa.begin() + L.fa_ptr[l], L.fa.begin() + L.fa_ptr[l + 1],
                          [&](const Task& t) { return factor_need(t) > kFactorSmallDoubles; });
    for (int i = L.fa_ptr[l]; i < L.fa_ptr[l + 1]; ++i) {
      const int need = factor_need(L.fa[i]);
      if (need > kFactorSmallDoubles) {
        ++L.fa_large[l];
        L.fa_smem[l] = std::max(L.fa_smem[l], need);
      } else {
        L.fa_smem_small[l] = std::max(L.fa_smem_small[l], need);
      }
    }
    for (int i = L.fb_ptr[l]; i < L.fb_ptr[l + 1]; ++i) {
      const Task& t = L.fb[i];
      const SuperDesc& sd = sn[pn[t.id].sn];
      const bool multi = sd.pn_end - sd.pn_begin > 1 && !((t.aux >> 28) & 1);
      const int w = multi ? kPanelWidth : pn[t.id].w;
      L.fb_smem[l] = std::max(L.fb_smem[l], sn_tile_doubles(w, t.aux & 0xFF, (t.aux >> 8) & 0xFF));
    }
  }
  auto subst_need = [&](const Task& t) { return sn_subst_doubles(sn[t.id].W); };
  L.ss_smem.assign(n_slevels, 0);
  L.ss_smem_small.assign(n_slevels, 0);
  L.ss_large.assign(n_slevels, 0);
  for (int l = 0; l < n_slevels; ++l) {
    std::stable_partition(L.ss.begin() + L.ss_ptr[l], L.ss.begin() + L.ss_ptr[l + 1],
                          [&](const Task& t) { return subst_need(t) > kSubstSmallDoubles; });
    for (int i = L.ss_ptr[l]; i < L.ss_ptr[l + 1]; ++i) {
      const int need = subst_need(L.ss[i]);
      if (need > kSubstSmallDoubles) {
        ++L.ss_large[l];
        L.ss_smem[l] = std::max(L.ss_smem[l], need);
      } else {
        L.ss_smem_small[l] = std::max(L.ss_smem_small[l], need);
      }
    }
  }
  L.sa_smem.assign(n_slevels, 0);
  for (int l = 0; l < n_slevels; ++l)
    for (int i = L.sa_ptr[l]; i < L.sa_ptr[l + 1]; ++i)
      L.sa_smem[l] = std::max(L.sa_smem[l], 6 * sn[L.sa[i].id].W + kPanelWidth * kPanelWidth * 9 +
                                                kPanelWidth * 9 + 3 * 256);
  return L;
}

bool analyse(int n, const std::vector<std::pair<int, int> >& edges, int ordering, int world,
             Symbolic* out, std::string* err) {
  const auto t0 = std::chrono::steady_clock::now();
  *out = Symbolic();
  Symbolic& S = *out;
  S.n = n;
  for (size_t e = 0; e < edges.size(); ++e)
    if (edges[e].first < 0 || edges[e].first >= n || edges[e].second < 0 || edges[e].second >= n) {
      if (err) *err = "edge endpoint outside the free-vertex range";
      return false;
    }
  Graph g = build_graph(n, edges);

  // ---- ordering -------------------------------------------------------------------------------
  if (world < 1 || (world & (world - 1)) != 0) {
    if (err) *err = "world size must be a power of two";
    return false;
  }
  S.world = world;
  std::vector<int> vertex_owner(n, 0);
  if (ordering == 1) {
    if (world != 1) {
      if (err) *err = "natural ordering cannot be partitioned";
      return false;
    }
    S.perm.resize(n);
    std::iota(S.perm.begin(), S.perm.end(), 0);
  } else {
    Dissector d(g, kLeafSize, world);
    d.run(&S.perm);
    vertex_owner = d.owner();
    // leaf groups become supernodes: clique inside the group, every member tied to the whole
    // boundary of the group (structure only; the numeric blocks of these pairs start at zero)
    std::vector<std::pair<int, int> > all(edges);
    std::vector<int> in_group(n, -1), boundary;
    const std::vector<std::vector<int> >& groups = d.groups();
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      const std::vector<int>& L = groups[gi];
      boundary.clear();
      for (size_t i = 0; i < L.size(); ++i) in_group[L[i]] = static_cast<int>(gi);
      for (size_t i = 0; i < L.size(); ++i)
        for (int t = g.ptr[L[i]]; t < g.ptr[L[i] + 1]; ++t) {
          const int x = g.adj[t];
          if (in_group[x] != static_cast<int>(gi)) boundary.push_back(x);
        }
      std::sort(boundary.begin(), boundary.end());
      boundary.erase(std::unique(boundary.begin(), boundary.end()), boundary.end());
      for (size_t i = 0; i < L.size(); ++i) {
        for (size_t j = i + 1; j < L.size(); ++j) all.push_back(std::make_pair(L[i], L[j]));
        for (size_t j = 0; j < boundary.size(); ++j) all.push_back(std::make_pair(L[i], boundary[j]));
      }
    }
    if (!groups.empty()) g = build_graph(n, all);
  }
  if (static_cast<int>(S.perm.size()) != n) {
    if (err) *err = "internal: ordering lost vertices";
    return false;
  }
  // shared separators last (a topological reordering of the elimination tree: same fill)
  {
    std::vector<int> owned, shared;
    for (int p = 0; p < n; ++p) (vertex_owner[S.perm[p]] < 0 ? shared : owned).push_back(S.perm[p]);
    S.first_shared = static_cast<int>(owned.size());
    owned.insert(owned.end(), shared.begin(), shared.end());
    S.perm.swap(owned);
    S.owner.resize(n);
    for (int p = 0; p < n; ++p) S.owner[p] = vertex_owner[S.perm[p]];
  }
  S.iperm.assign(n, -1);
  for (int p = 0; p < n; ++p) S.iperm[S.perm[p]] = p;

  // ---- elimination tree (Liu, with path compression) ------------------------------------------
  S.parent.assign(n, -1);
  {
    std::vector<int> ancestor(n, -1);
    for (int p = 0; p < n; ++p) {
      const int v = S.perm[p];
      for (int t = g.ptr[v]; t < g.ptr[v + 1]; ++t) {
        int r = S.iperm[g.adj[t]];
        if (r >= p) continue;
        while (ancestor[r] != -1 && ancestor[r] != p) {
          const int next = ancestor[r];
          ancestor[r] = p;
          r = next;
        }
        if (ancestor[r] == -1) {
          ancestor[r] = p;
          S.parent[r] = p;
        }
      }
    }
  }
  // children lists
  std::vector<int> child_ptr(n + 1, 0), child(n > 0 ? n : 0);
  for (int p = 0; p < n; ++p)
    if (S.parent[p] >= 0) ++child_ptr[S.parent[p] + 1];
  for (int p = 0; p < n; ++p) child_ptr[p + 1] += child_ptr[p];
  {
    std::vector<int> cur(child_ptr.begin(), child_ptr.end() - 1);
    for (int p = 0; p < n; ++p)
      if (S.parent[p] >= 0) child[cur[S.parent[p]]++] = p;
  }

  // ---- factor structure -------------------------------------------------------------------------
  std::vector<std::vector<int> > cols(n);
  {
    std::vector<int> flag(n, -1);
    for (int p = 0; p < n; ++p) {
      std::vector<int>& c = cols[p];
      flag[p] = p;
      const int v = S.perm[p];
      for (int t = g.ptr[v]; t < g.ptr[v + 1]; ++t) {
        const int q = S.iperm[g.adj[t]];
        if (q > p && flag[q] != p) {
          flag[q] = p;
          c.push_back(q);
        }
      }
      for (int t = child_ptr[p]; t < child_ptr[p + 1]; ++t) {
        const std::vector<int>& cc = cols[child[t]];
        for (size_t i = 0; i < cc.size(); ++i) {
          const int q = cc[i];
          if (q > p && flag[q] != p) {
            flag[q] = p;
            c.push_back(q);
          }
        }
      }
      std::sort(c.begin(), c.end());
    }
  }
  S.col_ptr.assign(n + 1, 0);
  for (int p = 0; p < n; ++p) S.col_ptr[p + 1] = S.col_ptr[p] + 1 + static_cast<int>(cols[p].size());
  S.nnzb = S.col_ptr[n];
  if (S.nnzb > 0x7FFFFFF0LL) {
    if (err) *err = "factor too large for 32-bit block positions";
    return false;
  }
  S.row_idx.resize(S.nnzb);
  S.n_ops = 0;
  for (int p = 0; p < n; ++p) {
    int w = S.col_ptr[p];
    S.row_idx[w++] = p;
    for (size_t i = 0; i < cols[p].size(); ++i) S.row_idx[w++] = cols[p][i];
    S.n_ops += static_cast<int64_t>(cols[p].size()) * (static_cast<int64_t>(cols[p].size()) + 1) / 2;
    std::vector<int>().swap(cols[p]);
  }
  // ---- levels -----------------------------------------------------------------------------------
  S.level.assign(n, 0);
  S.n_levels = n ? 1 : 0;
  for (int p = 0; p < n; ++p) {
    const int par = S.parent[p];
    if (par >= 0) S.level[par] = std::max(S.level[par], S.level[p] + 1);
    S.n_levels = std::max(S.n_levels, S.level[p] + 1);
  }
  // ---- partition bookkeeping -----------------------------------------------------------------------
  S.local_levels = 0;
  S.shared_min_level = S.n_levels;
  for (int p = 0; p < n; ++p) {
    if (S.owner[p] >= 0) S.local_levels = std::max(S.local_levels, S.level[p] + 1);
    else S.shared_min_level = std::min(S.shared_min_level, S.level[p]);
    const int par = S.parent[p];
    if (par >= 0 && S.owner[par] >= 0 && S.owner[par] != S.owner[p]) {
      if (err) *err = "internal: partition does not respect the elimination tree";
      return false;
    }
  }
  for (size_t e = 0; e < edges.size(); ++e) {
    const int oa = S.owner[S.iperm[edges[e].first]], ob = S.owner[S.iperm[edges[e].second]];
    if (oa >= 0 && ob >= 0 && oa != ob) {
      if (err) *err = "internal: an edge crosses two ranks' interiors";
      return false;
    }
  }
  if (!build_supernodal(S, err)) return false;
  S.analyse_seconds =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return true;
}

}  // namespace pgo
