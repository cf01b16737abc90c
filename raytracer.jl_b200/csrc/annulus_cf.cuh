// annulus_cf.cuh -- closed-form description of the mesh that init_annulus (src/GridAnnulus.jl:57-70) builds.
//
// The reference builds the annulus with Dict/Set based serial passes (primary_grid :72-142, edge_connectivity
// :515-595, secondary_nodes :607-698, constrain2layers! :296-321, discontinuous_boundaries :910-968,
// element_incidence :420-452).  Every array it produces is a pure function of (ntheta, nr, spacing), so here
// each entity (edge, node, element list, G column, halo row) is computed independently from its index:
// that is what lets one CUDA thread (or warp) produce one entity with no serial pass and no hash containers.
// The functions are __host__ __device__ so that the same code is exercised on the CPU by the unit tests
// (tests/cf_host_driver.cpp) and runs inside the kernels of annulus_build.cu.
//
// Notation: T = ntheta, M = nr + 7 rings, ring radii rc[1..M]; ring node (k, c) has id k + M (c - 1);
// centre node C = M T + 1; quad (k, c) = element k + (M-1)(c-1) with corners [(k,c), (k,c+1), (k+1,c+1), (k+1,c)];
// triangle c = element nq + c = [C, (1,c), (1,c+1)], nq = (M-1) T.  All ids 1-based in this header.
//
// Edge creation order of edge_connectivity (with its slot-1 marking quirk): quad 1 creates its
// bottom edge (id 1); EVERY quad e creates right = 3(e-1)+2, top = 3(e-1)+3, left = 3(e-1)+4 (so every radial
// edge exists twice, once as the right edge of the left quad and once as the left edge of the right quad);
// triangle c creates its ring edge Eq + 2(c-1) + 1 and the spoke Eq + 2(c-1) + 2 (Eq = 3 nq + 1).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CF_HD __host__ __device__ __forceinline__
#else
#define CF_HD inline
#endif

namespace cf {

typedef int64_t i64;

struct Params {
  i64 T, M;        // columns, rings
  i64 nq, nel;     // quads, elements
  i64 Eq, nE;      // last quad edge id, number of edges
  i64 nnods0;      // primary nodes (M T + 1)
  double spacing;
  double dth;      // 2 pi / T
  double eps;      // 2 pi - (1 - 1/T)   (GridAnnulus.jl:621)
  const double* rc;   // [M+1], rc[k] for k = 1..M (rc[0] unused)
  const int* lay;     // [M], layer id of quad row k = 1..M-1 (lay[0] unused); triangles are layer 1
  const int* disc;    // [M+2], disc[k] = 1 iff ring k is one of the 7 discontinuities
};

CF_HD double two_pi() { return 6.283185307179586; }  // 2 * Float64(pi)
CF_HD double pi_() { return 3.141592653589793; }

CF_HD i64 wrap_col(const Params& p, i64 c) { return c < 1 ? c + p.T : (c > p.T ? c - p.T : c); }
CF_HD i64 ring_node(const Params& p, i64 k, i64 c) { return k + p.M * (c - 1); }
CF_HD i64 quad_id(const Params& p, i64 k, i64 c) { return k + (p.M - 1) * (c - 1); }
CF_HD i64 tri_id(const Params& p, i64 c) { return p.nq + c; }
CF_HD void quad_kc(const Params& p, i64 e, i64& k, i64& c) {
  k = (e - 1) % (p.M - 1) + 1;
  c = (e - 1) / (p.M - 1) + 1;
}
CF_HD i64 edge_right(i64 e) { return 3 * (e - 1) + 2; }
CF_HD i64 edge_top(i64 e) { return 3 * (e - 1) + 3; }
CF_HD i64 edge_left(i64 e) { return 3 * (e - 1) + 4; }
CF_HD i64 edge_tri_ring(const Params& p, i64 c) { return p.Eq + 2 * (c - 1) + 1; }
CF_HD i64 edge_tri_spoke(const Params& p, i64 c) { return p.Eq + 2 * (c - 1) + 2; }

// Julia float div: round((x - rem(x, y)) / y), ties to even
CF_HD double julia_fdiv(double x, double y) { return rint((x - fmod(x, y)) / y); }

// Geometry of edge g: the (theta, r) pairs AFTER correct_theta (:710-725), ordered (lower id, higher id).
struct EdgeGeom {
  double t1, t2, r1, r2;
};

// kind: 0 azimuthal (ring k, columns c -> c+1), 1 radial (column c, rings k -> k+1), 2 spoke (column c)
CF_HD void edge_kind(const Params& p, i64 g, int& kind, i64& k, i64& c) {
  if (g == 1) {
    kind = 0;
    k = 1;
    c = 1;
  } else if (g <= p.Eq) {
    const i64 e = (g - 2) / 3 + 1;
    const int s = (int)((g - 2) % 3);
    i64 qk, qc;
    quad_kc(p, e, qk, qc);
    if (s == 0) {  // right
      kind = 1;
      k = qk;
      c = wrap_col(p, qc + 1);
    } else if (s == 1) {  // top
      kind = 0;
      k = qk + 1;
      c = qc;
    } else {  // left
      kind = 1;
      k = qk;
      c = qc;
    }
  } else {
    const i64 t = g - p.Eq - 1;
    const i64 tc = t / 2 + 1;
    if (t % 2 == 0) {
      kind = 0;
      k = 1;
      c = tc;
    } else {
      kind = 2;
      k = 1;
      c = wrap_col(p, tc + 1);
    }
  }
}

CF_HD EdgeGeom edge_geom(const Params& p, i64 g) {
  int kind;
  i64 k, c;
  edge_kind(p, g, kind, k, c);
  EdgeGeom q;
  if (kind == 0) {
    // nodes (k, c) and (k, c+1); sorted by id: for the wrap column (c == T) the lower id is (k, 1)
    if (c < p.T) {
      q.t1 = p.dth * (double)(c - 1);
      q.t2 = p.dth * (double)c;
    } else {
      q.t1 = p.dth * (double)0;
      q.t2 = p.dth * (double)(p.T - 1);
    }
    q.r1 = p.rc[k];
    q.r2 = p.rc[k];
    if (fabs(q.t1 - q.t2) >= p.eps) {  // correct_theta, non-centre branch
      if (q.t1 < pi_())
        q.t1 = q.t1 + two_pi();
      else if (q.t2 < pi_())
        q.t2 = q.t2 + two_pi();
    }
  } else if (kind == 1) {
    q.t1 = q.t2 = p.dth * (double)(c - 1);
    q.r1 = p.rc[k];
    q.r2 = p.rc[k + 1];
  } else {
    // spoke: lower id is the ring node (1, c), higher id the centre (theta = 0, r = 0); centre branch: max
    const double tc = p.dth * (double)(c - 1);
    const double tm = tc > 0.0 ? tc : 0.0;
    q.t1 = q.t2 = tm;
    q.r1 = p.rc[1];
    q.r2 = 0.0;
  }
  return q;
}

// edge_length :700-708 and npoints = Int(L div spacing) :642
CF_HD i64 edge_npoints(const Params& p, i64 g) {
  const EdgeGeom q = edge_geom(p, g);
  double L;
  if (q.t1 == q.t2)
    L = sqrt(q.r1 * q.r1 + q.r2 * q.r2 - 2 * q.r1 * q.r2 * cos(q.t1 - q.t2));
  else
    L = q.r1 * fabs(q.t2 - q.t1);
  const double np = julia_fdiv(L, p.spacing);
  return np > 0.0 ? (i64)np : 0;
}

// (theta, r) of the j-th (1-based) secondary node of edge g with np points (:650-657)
CF_HD void secondary_coord(const Params& p, i64 g, i64 np, i64 j, double& th, double& r) {
  const EdgeGeom q = edge_geom(p, g);
  const double Lt = q.t2 - q.t1, Lr = q.r2 - q.r1;
  th = q.t1 + Lt * (double)j / (double)(np + 1);
  r = q.r1 + Lr * (double)j / (double)(np + 1);
}

// edge2el: the (at most two) elements that receive the secondary nodes of edge g
CF_HD void edge_elements(const Params& p, i64 g, i64& a, i64& b) {
  if (g == 1) {
    a = 1;
    b = tri_id(p, 1);
    return;
  }
  if (g <= p.Eq) {
    const i64 e = (g - 2) / 3 + 1;
    const int s = (int)((g - 2) % 3);
    i64 k, c;
    quad_kc(p, e, k, c);
    a = e;
    if (s == 0)
      b = quad_id(p, k, wrap_col(p, c + 1));
    else if (s == 1)
      b = (k + 1 <= p.M - 1) ? quad_id(p, k + 1, c) : 0;
    else
      b = quad_id(p, k, wrap_col(p, c - 1));
    return;
  }
  const i64 t = g - p.Eq - 1;
  const i64 tc = t / 2 + 1;
  a = tri_id(p, tc);
  b = (t % 2 == 0) ? quad_id(p, 1, tc) : tri_id(p, wrap_col(p, tc + 1));
}

// Edges whose secondary nodes appear in the e2n list of element e, ascending edge id (== list order).
// Returns the count (<= 7).
CF_HD int element_edges(const Params& p, i64 e, i64 out[8]) {
  int n = 0;
  if (e <= p.nq) {
    i64 k, c;
    quad_kc(p, e, k, c);
    if (e == 1) out[n++] = 1;
    out[n++] = edge_right(e);
    out[n++] = edge_top(e);
    out[n++] = edge_left(e);
    out[n++] = edge_right(quad_id(p, k, wrap_col(p, c - 1)));
    out[n++] = edge_left(quad_id(p, k, wrap_col(p, c + 1)));
    if (k >= 2)
      out[n++] = edge_top(quad_id(p, k - 1, c));
    else
      out[n++] = edge_tri_ring(p, c);
  } else {
    const i64 c = e - p.nq;
    if (c == 1) out[n++] = 1;
    out[n++] = edge_tri_ring(p, c);
    out[n++] = edge_tri_spoke(p, c);
    out[n++] = edge_tri_spoke(p, wrap_col(p, c - 1));
  }
  for (int i = 1; i < n; ++i) {  // insertion sort (n <= 7)
    const i64 v = out[i];
    int j = i - 1;
    while (j >= 0 && out[j] > v) {
      out[j + 1] = out[j];
      --j;
    }
    out[j + 1] = v;
  }
  return n;
}

// Is quad e directly BELOW a discontinuity ring (its 3rd corner lies on one)?  (:920)
CF_HD bool is_below_quad(const Params& p, i64 e) {
  if (e > p.nq) return false;
  i64 k, c;
  quad_kc(p, e, k, c);
  return p.disc[k + 1] != 0;
}

// layer-constrained neighbours of element e (element_neighbours :473-507 + constrain2layers! :296-321),
// ascending ids.  Returns the count (<= 11).
CF_HD int element_neighbours(const Params& p, i64 e, i64 out[12]) {
  int n = 0;
  if (e <= p.nq) {
    i64 k, c;
    quad_kc(p, e, k, c);
    for (int dc = -1; dc <= 1; ++dc)
      for (int dk = -1; dk <= 1; ++dk) {
        if (dc == 0 && dk == 0) continue;
        const i64 kk = k + dk;
        if (kk < 1 || kk > p.M - 1) continue;
        if (p.lay[kk] != p.lay[k]) continue;
        out[n++] = quad_id(p, kk, wrap_col(p, c + dc));
      }
    if (k == 1 && p.lay[1] == 1)
      for (int dc = -1; dc <= 1; ++dc) out[n++] = tri_id(p, wrap_col(p, c + dc));
  } else {
    const i64 c = e - p.nq;
    if (p.lay[1] == 1)
      for (int dc = -1; dc <= 1; ++dc) out[n++] = quad_id(p, 1, wrap_col(p, c + dc));
    out[n++] = tri_id(p, wrap_col(p, c - 1));
    out[n++] = tri_id(p, wrap_col(p, c + 1));
  }
  for (int i = 1; i < n; ++i) {
    const i64 v = out[i];
    int j = i - 1;
    while (j >= 0 && out[j] > v) {
      out[j + 1] = out[j];
      --j;
    }
    out[j + 1] = v;
  }
  // unique (tiny T can alias columns)
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (m == 0 || out[m - 1] != out[i]) out[m++] = out[i];
  return m;
}

// Insert element c and its neighbours into a sorted unique set (<= 48 entries)
CF_HD void set_insert(i64* set, int& n, i64 v) {
  int lo = 0;
  while (lo < n && set[lo] < v) ++lo;
  if (lo < n && set[lo] == v) return;
  for (int j = n; j > lo; --j) set[j] = set[j - 1];
  set[lo] = v;
  ++n;
}
CF_HD void set_add_patch(const Params& p, i64* set, int& n, i64 c) {
  set_insert(set, n, c);
  i64 nb[12];
  const int m = element_neighbours(p, c, nb);
  for (int i = 0; i < m; ++i) set_insert(set, n, nb[i]);
}

// G column (element_incidence :420-452) of a ring node (k, c) that is NOT replaced by twins in the cells below
// a discontinuity: containing quads (k-1 | k, c-1 | c) -- only (k, .) when ring k is a discontinuity -- plus
// the two triangles for k == 1.
CF_HD int g_column_ring_node(const Params& p, i64 k, i64 c, i64* set) {
  int n = 0;
  const i64 cl = wrap_col(p, c - 1);
  if (k - 1 >= 1 && !p.disc[k]) {
    set_add_patch(p, set, n, quad_id(p, k - 1, cl));
    set_add_patch(p, set, n, quad_id(p, k - 1, c));
  }
  if (k <= p.M - 1) {
    set_add_patch(p, set, n, quad_id(p, k, cl));
    set_add_patch(p, set, n, quad_id(p, k, c));
  }
  if (k == 1) {
    set_add_patch(p, set, n, tri_id(p, cl));
    set_add_patch(p, set, n, tri_id(p, c));
  }
  return n;
}

// G column of the secondary nodes of edge g (original ids)
CF_HD int g_column_edge(const Params& p, i64 g, i64* set) {
  int n = 0;
  i64 a, b;
  edge_elements(p, g, a, b);
  // the top edge of a below-quad keeps its ORIGINAL nodes only in the quad above (the below-quad gets twins)
  const bool a_is_twinned = (g >= 2 && g <= p.Eq && (g - 2) % 3 == 1 && is_below_quad(p, a));
  if (!a_is_twinned) set_add_patch(p, set, n, a);
  if (b) set_add_patch(p, set, n, b);
  return n;
}

}  // namespace cf

// -----------------------------------------------------------------------------------------------------------
// Entities that need the prefix sums over edges / below-quads
namespace cf {

struct Tables {
  const i64* eoff;      // [nE + 1] exclusive prefix sum of the per-edge point counts (edge g -> eoff[g-1])
  const i64* twin_off;  // [7 T + 1] exclusive prefix sum of (2 + np(top edge)) over the below-quads
  const i64* kd;        // [7] ring index k of the seven quad rows directly below a discontinuity, ascending
  i64 nnods1;           // nodes before the twins (primary + secondary)
  i64 nnods;            // all nodes
};

CF_HD i64 edge_np_from_off(const Tables& tb, i64 g) { return tb.eoff[g] - tb.eoff[g - 1]; }

// index d (0..6) of the below-row holding ring k, or -1
CF_HD int below_row(const Tables& tb, i64 k) {
  for (int d = 0; d < 7; ++d)
    if (tb.kd[d] == k) return d;
  return -1;
}
CF_HD i64 below_index(const Params& p, const Tables& tb, i64 e) {  // -1 if e is not a below-quad
  if (e > p.nq) return -1;
  i64 k, c;
  quad_kc(p, e, k, c);
  const int d = below_row(tb, k);
  return d < 0 ? -1 : (c - 1) * 7 + d;
}

CF_HD i64 elem_list_len(const Params& p, const Tables& tb, i64 e) {
  i64 ed[8];
  const int ne = element_edges(p, e, ed);
  i64 len = e <= p.nq ? 4 : 3;
  for (int i = 0; i < ne; ++i) len += edge_np_from_off(tb, ed[i]);
  return len;
}

// e2n list of element e after discontinuous_boundaries (twins substituted in the below-quads).
// Writes id + bias (bias = 0: 1-based int64 for the ABI; bias = -1: 0-based int32 for the device mesh).
template <typename OutT>
CF_HD void elem_list_fill(const Params& p, const Tables& tb, i64 e, OutT* out, i64 bias = 0) {
  i64 ed[8];
  const int ne = element_edges(p, e, ed);
  i64 o = 0;
  const i64 b = below_index(p, tb, e);
  if (e <= p.nq) {
    i64 k, c;
    quad_kc(p, e, k, c);
    const i64 c1 = wrap_col(p, c + 1);
    out[o++] = (OutT)(ring_node(p, k, c) + bias);
    out[o++] = (OutT)(ring_node(p, k, c1) + bias);
    if (b >= 0) {
      out[o++] = (OutT)(tb.nnods1 + tb.twin_off[b] + 1 + bias);
      out[o++] = (OutT)(tb.nnods1 + tb.twin_off[b] + 2 + bias);
    } else {
      out[o++] = (OutT)(ring_node(p, k + 1, c1) + bias);
      out[o++] = (OutT)(ring_node(p, k + 1, c) + bias);
    }
  } else {
    const i64 c = e - p.nq;
    out[o++] = (OutT)(p.nnods0 + bias);  // centre
    out[o++] = (OutT)(ring_node(p, 1, c) + bias);
    out[o++] = (OutT)(ring_node(p, 1, wrap_col(p, c + 1)) + bias);
  }
  for (int i = 0; i < ne; ++i) {
    const i64 g = ed[i];
    const i64 np = edge_np_from_off(tb, g);
    const bool twinned = b >= 0 && g == edge_top(e);
    const i64 base = (twinned ? tb.nnods1 + tb.twin_off[b] + 2 : p.nnods0 + tb.eoff[g - 1]) + bias;
    for (i64 j = 1; j <= np; ++j) out[o++] = (OutT)(base + j);
  }
}

// (orig, twin) ids of halo row h (0 <= h < H): below-quad b, position pos inside it
CF_HD void halo_pair(const Params& p, const Tables& tb, i64 b, i64 pos, i64& orig, i64& twin) {
  const i64 c = b / 7 + 1;
  const i64 k = tb.kd[b % 7];
  twin = tb.nnods1 + tb.twin_off[b] + pos + 1;
  if (pos == 0)
    orig = ring_node(p, k + 1, wrap_col(p, c + 1));
  else if (pos == 1)
    orig = ring_node(p, k + 1, c);
  else
    orig = p.nnods0 + tb.eoff[edge_top(quad_id(p, k, c)) - 1] + (pos - 1);
}

// upper-bound style search: largest i in [0, n) with off[i] <= x   (off ascending, off[0] == 0)
CF_HD i64 find_segment(const i64* off, i64 n, i64 x) {
  i64 lo = 0, hi = n;
  while (hi - lo > 1) {
    const i64 mid = (lo + hi) >> 1;
    if (off[mid] <= x)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// G column of node v (1-based) except the centre node (handled separately: 2T entries).  <= 48 entries.
CF_HD int g_column(const Params& p, const Tables& tb, i64 v, i64* set) {
  if (v < p.nnods0) {
    const i64 k = (v - 1) % p.M + 1, c = (v - 1) / p.M + 1;
    return g_column_ring_node(p, k, c, set);
  }
  if (v <= tb.nnods1) {
    const i64 s = v - p.nnods0 - 1;  // 0-based secondary index
    // edge g (1-based) with eoff[g-1] <= s < eoff[g]; skip empty edges
    i64 lo = 0, hi = p.nE;  // search over g-1 in [0, nE)
    while (hi - lo > 1) {
      const i64 mid = (lo + hi) >> 1;
      if (tb.eoff[mid] <= s)
        lo = mid;
      else
        hi = mid;
    }
    return g_column_edge(p, lo + 1, set);
  }
  const i64 tw = v - tb.nnods1 - 1;
  const i64 b = find_segment(tb.twin_off, 7 * p.T, tw);
  int n = 0;
  set_add_patch(p, set, n, quad_id(p, tb.kd[b % 7], b / 7 + 1));
  return n;
}

// G column of the centre node: every triangle and (same layer) every first-row quad, ascending
CF_HD i64 g_column_centre_len(const Params& p) { return p.lay[1] == 1 ? 2 * p.T : p.T; }
CF_HD i64 g_column_centre_entry(const Params& p, i64 i) {  // i-th (0-based) entry
  if (p.lay[1] == 1) return i < p.T ? quad_id(p, 1, i + 1) : tri_id(p, i - p.T + 1);
  return tri_id(p, i + 1);
}

}  // namespace cf

// -----------------------------------------------------------------------------------------------------------
// Host-side O(nr) set-up of the parameter tables (ring radii, quad-row layers, discontinuity flags)
#include <algorithm>
#include <vector>
namespace cf {

struct HostParams {
  std::vector<double> rc;
  std::vector<int> lay, disc;
  std::vector<i64> kd;
  Params p;
};

// primary_grid :72-95 (ring radii), constrain2layers! :296-313 + find_boundary :374-381 (quad-row layers)
inline void make_params(i64 ntheta, i64 nr, double spacing, HostParams& hp) {
  const double R = 6371.0;
  const double rl[7] = {R - 20.0, R - 35.0, R - 210.0, R - 410.0, R - 660.0, R - 2740.0, R - 2891.5};
  const i64 M = nr + 7, T = ntheta;
  std::vector<double> col(rl, rl + 7);
  const i64 nlin = nr;
  for (i64 j = 0; j < nlin; ++j) {  // LinRange(0.1, R, nr)[j+1] = (1 - t) * a + t * b, t = j / (nr - 1)
    const double t = nlin > 1 ? (double)j / (double)(nlin - 1) : 0.0;
    col.push_back(nlin > 1 ? (1.0 - t) * 0.1 + t * R : 0.1);
  }
  std::sort(col.begin(), col.end());
  hp.rc.assign(M + 2, 0.0);
  for (i64 k = 1; k <= M; ++k) hp.rc[k] = col[k - 1];
  hp.disc.assign(M + 2, 0);
  for (i64 k = 1; k <= M; ++k)
    for (int q = 0; q < 7; ++q)
      if (hp.rc[k] == rl[q]) hp.disc[k] = 1;
  const double rlayer[8] = {R, R - 20, R - 35, R - 210, R - 410, R - 660, R - 2740, R - 2891.5};
  hp.lay.assign(M + 1, 0);
  for (i64 k = 1; k <= M - 1; ++k) {
    const double c = (hp.rc[k] + hp.rc[k] + hp.rc[k + 1] + hp.rc[k + 1]) * 0.25;
    int L = 0;
    if (c < rlayer[7])
      L = 1;
    else
      for (int i = 0; i < 7; ++i)
        if (rlayer[i] > c && c > rlayer[i + 1]) L = i + 2;
    hp.lay[k] = L;
  }
  hp.kd.clear();
  for (i64 k = 1; k <= M - 1; ++k)
    if (hp.disc[k + 1]) hp.kd.push_back(k);
  Params& p = hp.p;
  p.T = T;
  p.M = M;
  p.nq = (M - 1) * T;
  p.nel = p.nq + T;
  p.Eq = 3 * p.nq + 1;
  p.nE = p.Eq + 2 * T;
  p.nnods0 = M * T + 1;
  p.spacing = spacing;
  p.dth = 2 * 3.141592653589793 / (double)T;
  p.eps = 2 * 3.141592653589793 - (1.0 - 1.0 / (double)T);
  p.rc = hp.rc.data();
  p.lay = hp.lay.data();
  p.disc = hp.disc.data();
}

}  // namespace cf
