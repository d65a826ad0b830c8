// GPU fabric planner: link matrix -> uniformity verdict, reduction trees, rings.
//
// Reference K5 (MX/src/kvstore/gpu_topology.h, comm_tree.h): build a link matrix from
// cudaDeviceGetP2PAttribute (:133-196), partition the GPUs with Kernighan-Lin bisection (:269-427) (or
// backtracking, :789-957) into a binary reduction tree per root, penalising links that are already used
// (MXNET_KVSTORE_TREE_LINK_USAGE_PENALTY = 0.7, comm_tree.h:56), so that a DGX-1 hybrid cube mesh reduces
// over NVLink instead of PCIe.
//
// On a B200 HGX box the matrix is uniform (every pair: 18 NVLink-5 links through NVSwitch), the planner
// detects that and emits flat "every peer is one hop" plans used by the one-/two-shot kernels; the general
// path (KL bisection with link-usage penalty) is kept so the same code plans rings/trees on asymmetric
// fabrics, and is unit-tested GPU-free with synthetic matrices exactly like the reference's
// tests/cpp/kvstore/gpu_topology_test.cc.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#define RT_API extern "C" __attribute__((visibility("default")))

namespace dlcfn {

using Mat = std::vector<std::vector<double>>;

static bool IsUniform(const Mat& w) {
  const int n = (int)w.size();
  double ref = -1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (i == j) continue;
      if (w[i][j] <= 0) return false;
      if (ref < 0) ref = w[i][j];
      else if (std::fabs(w[i][j] - ref) > 1e-9) return false;
    }
  return true;
}

// Kernighan-Lin: split `nodes` into two halves (the half containing `root` first) maximising the weight that
// stays INSIDE the halves (equivalently minimising the cut), equal sizes (+-1).
static void KLBisect(const Mat& w, const std::vector<int>& nodes, int root, std::vector<int>* a, std::vector<int>* b) {
  const int n = (int)nodes.size();
  std::vector<int> side(n);
  // initial split: root + its strongest neighbours on side 0
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  int ri = (int)(std::find(nodes.begin(), nodes.end(), root) - nodes.begin());
  std::sort(order.begin(), order.end(), [&](int x, int y) {
    if (x == ri) return true;
    if (y == ri) return false;
    return w[root][nodes[x]] > w[root][nodes[y]];
  });
  const int half = (n + 1) / 2;
  for (int k = 0; k < n; ++k) side[order[k]] = k < half ? 0 : 1;
  auto cut_gain = [&](int x, int y) {  // gain of swapping x (side 0) and y (side 1)
    double dx = 0, dy = 0;
    for (int k = 0; k < n; ++k) {
      if (k == x || k == y) continue;
      double wx = w[nodes[x]][nodes[k]], wy = w[nodes[y]][nodes[k]];
      dx += side[k] == 1 ? wx : -wx;  // external - internal for x
      dy += side[k] == 0 ? wy : -wy;
    }
    return dx + dy - 2 * w[nodes[x]][nodes[y]];
  };
  for (int pass = 0; pass < 8; ++pass) {
    double best = 1e-12;
    int bx = -1, by = -1;
    for (int x = 0; x < n; ++x) {
      if (side[x] != 0 || x == ri) continue;  // the root stays in its half
      for (int y = 0; y < n; ++y) {
        if (side[y] != 1) continue;
        double g = cut_gain(x, y);
        if (g > best) { best = g; bx = x; by = y; }
      }
    }
    if (bx < 0) break;
    std::swap(side[bx], side[by]);
  }
  a->clear(); b->clear();
  for (int k = 0; k < n; ++k) (side[k] == 0 ? a : b)->push_back(nodes[k]);
}

// Recursive tree construction: the root keeps one half, hands the other half to the best-linked node in it.
static void BuildTree(Mat& w, const std::vector<int>& nodes, int root, double penalty, std::vector<int>* parent) {
  if (nodes.size() <= 1) return;
  std::vector<int> a, b;
  KLBisect(w, nodes, root, &a, &b);
  if (b.empty()) return;
  int child = b[0];
  for (int c : b)
    if (w[root][c] > w[root][child]) child = c;
  (*parent)[child] = root;
  // a used link is worth less to the next tree that considers it (comm_tree.h:56)
  w[root][child] *= penalty;
  w[child][root] *= penalty;
  BuildTree(w, a, root, penalty, parent);
  BuildTree(w, b, child, penalty, parent);
}

// Backtracking planner (reference gpu_topology.h:789-957, MXNET_KVSTORE_TREE_BACKTRACK=1): the doubling tree - at every
// level each node already in the tree adopts one new child - searched exhaustively, depth first, for the assignment with the
// largest total link weight that uses only existing links (weight > 0).  KL bisection is a heuristic and can pair GPUs that
// have no direct link on sparse fabrics; this search never does, and for n <= 8 it is instantaneous.
struct Backtrack {
  const Mat& w;
  int n;
  std::vector<int> best_parent, parent;
  double best = -1;
  explicit Backtrack(const Mat& m) : w(m), n((int)m.size()) {}

  // members: nodes in the tree in adoption order.  Within a level the members [0, level_end) adopt in order; a member with
  // no free linked neighbour adopts nobody; a level in which nobody adopted while GPUs are still outside is a dead end.
  void Adopt(std::vector<int>& members, std::vector<char>& used, double total, int idx, int level_end, bool adopted) {
    if ((int)members.size() == n) {
      if (total > best) { best = total; best_parent = parent; }
      return;
    }
    if (idx == level_end) {
      if (adopted) Adopt(members, used, total, 0, (int)members.size(), false);
      return;
    }
    const int p = members[idx];
    bool any = false;
    for (int c = 0; c < n; ++c) {
      if (used[c] || w[p][c] <= 0) continue;
      any = true;
      used[c] = 1; parent[c] = p; members.push_back(c);
      Adopt(members, used, total + w[p][c], idx + 1, level_end, true);
      members.pop_back(); parent[c] = -2; used[c] = 0;
    }
    if (!any) Adopt(members, used, total, idx + 1, level_end, adopted);
  }
  bool Run(int root, std::vector<int>* out) {
    parent.assign(n, -2);
    parent[root] = -1;
    best = -1;
    std::vector<int> members{root};
    std::vector<char> used(n, 0);
    used[root] = 1;
    Adopt(members, used, 0.0, 0, 1, false);
    if (best < 0) return false;
    *out = best_parent;
    return true;
  }
};

}  // namespace dlcfn

using namespace dlcfn;

// One doubling tree per root by exhaustive backtracking over existing links (see Backtrack); the weight of every used link
// is multiplied by `penalty` before the next root is planned.  Returns 0, or -2 if some root has no spanning tree.
RT_API int dlcfn_topo_plan_trees_backtrack(const double* link, int n, double penalty, int* parents) {
  if (n <= 0 || n > 16) return -1;
  Mat w(n, std::vector<double>(n));
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) w[i][j] = link[i * n + j];
  for (int r = 0; r < n; ++r) {
    Backtrack bt(w);
    std::vector<int> parent;
    if (!bt.Run(r, &parent)) return -2;
    for (int i = 0; i < n; ++i) {
      parents[r * n + i] = parent[i];
      if (parent[i] >= 0) { w[i][parent[i]] *= penalty; w[parent[i]][i] *= penalty; }
    }
  }
  return 0;
}

// link: n*n row-major weights (0 = no direct link).  Returns 1 if all off-diagonal weights are equal and > 0.
RT_API int dlcfn_topo_is_uniform(const double* link, int n) {
  Mat w(n, std::vector<double>(n));
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) w[i][j] = link[i * n + j];
  return IsUniform(w) ? 1 : 0;
}

// One reduction tree per root.  parents: n*n ints, parents[r*n + i] = parent of GPU i in the tree rooted at r
// (-1 for the root).  Link usage is penalised across roots so the trees spread over the fabric.
RT_API int dlcfn_topo_plan_trees(const double* link, int n, double penalty, int* parents) {
  if (n <= 0) return -1;
  Mat w(n, std::vector<double>(n));
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) w[i][j] = link[i * n + j];
  std::vector<int> all(n);
  std::iota(all.begin(), all.end(), 0);
  for (int r = 0; r < n; ++r) {
    std::vector<int> parent(n, -2);
    parent[r] = -1;
    BuildTree(w, all, r, penalty, &parent);
    for (int i = 0; i < n; ++i) parents[r * n + i] = parent[i];
  }
  return 0;
}

// Greedy strongest-link ring starting at `start`; order receives the n GPU ids.
RT_API int dlcfn_topo_plan_ring(const double* link, int n, int start, int* order) {
  if (n <= 0 || start < 0 || start >= n) return -1;
  std::vector<char> used(n, 0);
  int cur = start;
  for (int k = 0; k < n; ++k) {
    order[k] = cur;
    used[cur] = 1;
    int best = -1;
    for (int j = 0; j < n; ++j)
      if (!used[j] && (best < 0 || link[cur * n + j] > link[cur * n + best])) best = j;
    cur = best;
  }
  return 0;
}
