// The reference's BVH, rebuilt on the host for the analytic primitives.
//
// rayColor gathers candidate shapes by a depth-first walk of the SAH tree built by
// generateBVH (helpers.h:381-472): children pushed (left, right) and popped from
// the back (render_final_project.cpp:492-512), every leaf whose padded box the ray
// LINE reaches with tmax > 0 contributes all its shapes, the FIRST shape wins ties
// of t (strict `<`, :531).  Two properties of that gather are visible in images
// and are therefore reproduced, not "fixed":
//   * ties: scenes with coplanar overlapping shapes (ceiling light panels lying in
//     the ceiling plane, stacked prisms) depend on the candidate order;
//   * culling is NOT conservative for shadow rays: the gather walks the
//     unnormalised light vector from isectP + sray*1e-3 (:814) while the occlusion
//     test walks the normalised one from isectP + s^*1e-3 (:838).  With the
//     sphere-light quirk (sampleRay returns a position, |sray| ~ tens of units)
//     occluders within the first centimetres are never gathered.
// This header replays the reference's top-down build on primitive centres
// (largest-extent axis, its quicksort, full SAH sweep with c_isect=1,
// c_trav=0.33, the n=2,3,4 special cases, leaf when extent < 1e-3 or the sweep
// does not pay) and emits the node array plus the primitive order in which the
// reference's stack visits the leaves.  The build is O(n^2) per node like the
// reference's; it is meant for the tens-to-hundreds of analytic primitives of the
// reference's scenes (triangle meshes go through the device LBVH instead).
#pragma once
#include <cfloat>
#include <functional>
#include <vector>

namespace drt {

struct RefNode {
  double lo[3], hi[3];   // BoundingVolume::lbound/ubound (geometry.cpp:2632-2655)
  int leaf;              // 1: holds primitives [first, first+count) of `order`
  int first, count;
  int left, right;       // children (interior nodes)
};

class ReferenceBVH {
 public:
  typedef std::function<void(int prim, double lo[3], double hi[3])> BoundsFn;

  std::vector<RefNode> nodes;   // nodes[0] is the root
  std::vector<int> order;       // primitives in reference candidate order

  // centers: 3 doubles per primitive (GeoPrimitive::center)
  void build(const double* centers, int n, const BoundsFn& bounds) {
    c_ = centers; bounds_ = bounds; nodes.clear(); order.clear();
    std::vector<int> all(n);
    for (int i = 0; i < n; i++) all[i] = i;
    gen(all);
  }

 private:
  const double* c_ = nullptr;
  BoundsFn bounds_;

  double ctr(int prim, int axis) const { return c_[3 * prim + axis]; }

  void centroidBounds(const std::vector<int>& idx, double lo[3], double hi[3]) const {   // helpers.h:330-362
    for (int a = 0; a < 3; a++) { lo[a] = FLT_MAX; hi[a] = FLT_MIN; }
    for (int i : idx)
      for (int a = 0; a < 3; a++) {
        if (ctr(i, a) < lo[a]) lo[a] = ctr(i, a);
        if (ctr(i, a) > hi[a]) hi[a] = ctr(i, a);
      }
  }
  float halfCost(const std::vector<int>& v, float base_area) const {                      // helpers.h:371-376
    double lo[3], hi[3];
    centroidBounds(v, lo, hi);
    double ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    return (float)((ex * ey * 2 + ex * ez * 2 + ey * ez * 2) / base_area * v.size());
  }
  void quicksort(std::vector<int>& idx, int axis, int low, int high) const {              // helpers.h:244-268
    if (low >= high) return;
    float pivot = (float)ctr(idx[high], axis);
    int i = low;
    for (int j = low; j < high; j++)
      if (ctr(idx[j], axis) < pivot) { std::swap(idx[j], idx[i]); i++; }
    std::swap(idx[i], idx[high]);
    quicksort(idx, axis, low, i - 1);
    quicksort(idx, axis, i + 1, high);
  }
  int makeNode(const std::vector<int>& idx) {                                              // geometry.cpp:2632-2655
    RefNode nd;
    for (int a = 0; a < 3; a++) { nd.lo[a] = FLT_MAX; nd.hi[a] = FLT_MIN; }   // ubound starts at FLT_MIN (tiny POSITIVE)
    for (int i : idx) {
      double lo[3], hi[3];
      bounds_(i, lo, hi);
      for (int a = 0; a < 3; a++) { if (lo[a] < nd.lo[a]) nd.lo[a] = lo[a]; if (nd.hi[a] < hi[a]) nd.hi[a] = hi[a]; }
    }
    for (int a = 0; a < 3; a++) { nd.lo[a] -= 1e-2; nd.hi[a] += 1e-2; }
    nd.leaf = 0; nd.first = nd.count = 0; nd.left = nd.right = -1;
    nodes.push_back(nd);
    return (int)nodes.size() - 1;
  }
  void makeLeaf(int node, const std::vector<int>& idx) {
    nodes[node].leaf = 1; nodes[node].first = (int)order.size(); nodes[node].count = (int)idx.size();
    order.insert(order.end(), idx.begin(), idx.end());
  }

  // Leaves are appended to `order` in the order the reference's stack pops them
  // (right subtree first), so a leaf's primitives are contiguous.
  int gen(std::vector<int> idx) {
    const int n = (int)idx.size();
    if (n == 1) { int me = makeNode(idx); makeLeaf(me, idx); return me; }
    double lo[3], hi[3];
    centroidBounds(idx, lo, hi);
    double ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    int axis = 0;
    if (ext[1] > ext[0]) axis = (ext[2] > ext[1]) ? 2 : 1;
    else if (ext[2] > ext[0]) axis = 2;
    if (ext[axis] < 1e-3) { int me = makeNode(idx); makeLeaf(me, idx); return me; }
    quicksort(idx, axis, 0, n - 1);
    const int me = makeNode(idx);
    int slice;
    if (n == 2) slice = 1;
    else if (n == 3) slice = 1;
    else if (n == 4) slice = 2;
    else {
      const float c_isect = 1, c_trav = 0.33f;                                             // render_final_project.cpp:77-78
      float base_area = (float)(ext[0] * ext[1] * 2 + ext[1] * ext[2] * 2 + ext[0] * ext[2] * 2);
      float best = FLT_MAX;
      slice = 1;
      for (int i = 1; i < n - 1; i++) {
        std::vector<int> a(idx.begin(), idx.begin() + i), b(idx.begin() + i, idx.end());
        float cost = c_trav + c_isect * (halfCost(a, base_area) + halfCost(b, base_area));
        if (cost < best) { best = cost; slice = i; }
      }
      if (c_isect * n <= best) { makeLeaf(me, idx); return me; }
    }
    // right child first: it is the one the reference's stack pops first
    int r = gen(std::vector<int>(idx.begin() + slice, idx.end()));
    int l = gen(std::vector<int>(idx.begin(), idx.begin() + slice));
    nodes[me].left = l; nodes[me].right = r;
    return me;
  }
};

}  // namespace drt
