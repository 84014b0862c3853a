// Item schedule of the 128->64 w-sweep kernel (convt_l4_sw.cu), shared between the device code and a host-only unit test
// (tests/test_schedule_cpu.py compiles this header with g++).  An item is (decode-block pair, depth) = 16 sweep steps.
#pragma once
#ifdef __CUDACC__
#define A3D_L4_HD __host__ __device__ __forceinline__
#else
#define A3D_L4_HD inline
#endif

namespace a3d {
namespace l4 {

constexpr int WIN = 16;   // input grid of the 128->64 layer (depths per decode-block pair)

// Output depth of item t for a class with depth parity pd.  pd = 1 classes run one plane behind (d = t - 1 mod 16):
// class (pd, .) reads the input planes d + pd - 1 and d + pd, so with this shift ALL FOUR classes working on item t read
// the same two planes (t - 1, t) at the same time (L2 reuse), and all four have their half-length item (one plane
// outside the grid) at t = 0 mod 16, which keeps the clusters of an item in step.
A3D_L4_HD int item_depth(int t, int pd) { return (t + (pd ? WIN - 1 : 0)) % WIN; }

// Work list of a cluster: up to two segments.  The regular clusters own one parity class for the whole launch and walk
// the items [0, limit) of that class in rounds of `per` items; the clusters left over after dividing the grid by four
// help two classes each with the contiguous tail [first, first + count).
//   Regular cluster cj takes item k * per + (cj + k) % per in round k: neighbouring clusters always work on neighbouring
// items (they share an activation plane through L2), and the rotation by k walks every cluster through all 16 depths --
// with the plain stride-`per` walk (per = 18) the half-length border items (one input plane outside the grid) all went to
// the even-numbered clusters, which then idled 6 % of the launch.
struct Seg { int q, first, count, per, cj, limit; };   // per > 0: regular (rounds); per == 0: contiguous tail
struct Sched { Seg s[2]; };

A3D_L4_HD int seg_item(const Seg& S, int k) {
  if (S.per == 0) return S.first + k;
  const int t = k * S.per + (S.cj + k) % S.per;
  return t < S.limit ? t : -1;             // only the last round can be incomplete
}

A3D_L4_HD Sched make_sched(int cluster_id, int n_clusters, int n_items) {
  Sched sc;
  const int per = n_clusters >> 2;          // regular clusters per class (the launcher passes 4k or 4k + 2 clusters)
  const int reg = per << 2;
  sc.s[1] = Seg{0, 0, 0, 0, 0, 0};
  // With two helper clusters, each helper finishes two classes: the regular clusters of a class walk items [0, n_reg),
  // the helper the tail [n_reg, n_items); n_reg / per = 2 (n_items - n_reg) balances both kinds of cluster.
  int n_reg = n_items;
  if (n_clusters > reg) n_reg = (int)(((long long)n_items * 2 * per + 2 * per) / (2 * per + 1));
  if (n_reg > n_items) n_reg = n_items;
  if (cluster_id < reg) {
    sc.s[0] = Seg{cluster_id & 3, 0, (n_reg + per - 1) / per, per, cluster_id >> 2, n_reg};
  } else {
    const int e = cluster_id - reg;         // 0 or 1
    sc.s[0] = Seg{2 * e, n_reg, n_items - n_reg, 0, 0, n_items};
    sc.s[1] = Seg{2 * e + 1, n_reg, n_items - n_reg, 0, 0, n_items};
  }
  return sc;
}

// Clusters (CTA pairs) to launch: 4k (k per output-parity class) or 4k + 2 -- the two left-over clusters help two
// classes each; tiny launches get no helpers and at most one cluster per (item, class).
A3D_L4_HD int num_clusters(int num_sms, int n_items) {
  int n_clusters = num_sms / 2;
  if ((n_clusters & 3) != 2 || n_items < 64) n_clusters &= ~3;
  if (n_clusters > 4 * n_items) n_clusters = 4 * n_items;
  if (n_clusters < 4) n_clusters = 4;
  return n_clusters;
}

}  // namespace l4
}  // namespace a3d
