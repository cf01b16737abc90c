// rt_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never measured as product).
//
// A plain C++ restatement of the reference algorithm of albert-de-montserrat/RayTracer.jl for the
// shortest-path-method hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library.  The product (raytracer.jl_b200/csrc) never links it.
//
// PARITY STATUS: "parity unpinned" at the level of golden vectors -- the reference is pure Julia, ships no
// tests / fixtures / known-answer vectors, and Julia is not installed here, so the reference cannot be run.
// The oracle is pinned instead by (1) a line-by-line restatement with file:line citations below,
// (2) the closed-form node/element/halo counts of SURVEY.md section 8 (tests/test_oracle_builder.py),
// (3) algebraic invariants of SURVEY.md section 8c and (4) an independent binary-heap Dijkstra that must give
// bit-identical travel times (ora_dijkstra*).
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -fopenmp; no FMA contraction anywhere, because
// plain Julia never contracts a*b+c).
//
// All node / element ids crossing this API are 1-based int64 exactly as in Julia.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <set>
#include <string>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

namespace {

const double R_EARTH = 6371.0;  // src/utils.jl:2
// src/GridAnnulus.jl:73 (Float32 literals, all exactly representable in binary)
const double RL[7] = {6371.0 - 20.0, 6371.0 - 35.0, 6371.0 - 210.0, 6371.0 - 410.0,
                      6371.0 - 660.0, 6371.0 - 2740.0, 6371.0 - 2891.5};
const double PI = 3.141592653589793;  // Float64(pi)

// Julia Base: lerpi(j, d, a, b) = (t = j/d; (1-t)*a + t*b)   (LinRange element j+1 of d+1)
inline double lerpi(i64 j, i64 d, double a, double b) {
  double t = (double)j / (double)d;
  return (1.0 - t) * a + t * b;
}

// Julia float `div(x, y)` = round((x - rem(x, y)) / y), rem == C fmod (exact), round == ties-to-even.
inline double julia_fdiv(double x, double y) { return std::nearbyint((x - std::fmod(x, y)) / y); }

struct Mesh {
  std::vector<double> x, z, theta, r;
  std::vector<std::vector<i64>> e2n;  // [nel], 1-based node ids
  std::vector<std::vector<i64>> nbr;  // [nel], 1-based element ids
  std::vector<int8_t> el_type;        // 0 = :Quad, 1 = :Tri
  i64 ntheta = 0, nr = 0, nel = 0, nnods = 0;
  std::vector<i64> G_colptr, G_rowval;  // SparseMatrixCSC{Bool,Int64}(nel x nnods), 1-based
  std::vector<i64> halo;                // (2H x 2) column-major, 1-based
  i64 halo_rows = 0;
  std::string err;
};

// ---------------------------------------------------------------------------------------------------------
// src/GridAnnulus.jl:473-507  element_neighbours
void element_neighbours(Mesh& m) {
  const i64 nel = m.nel;
  // incidence_matrix = sparse(J=element, I=node): column(node) = ascending element ids containing node
  std::vector<std::vector<i64>> node2el(m.nnods + 1);
  for (i64 e = 1; e <= nel; ++e)
    for (i64 nd : m.e2n[e - 1]) node2el[nd].push_back(e);  // ascending because e ascends
  m.nbr.assign(nel, {});
  // NOTE reference quirk :490 -- `for node in 1:nel` (nel == number of ring nodes; the centre node is skipped)
  for (i64 node = 1; node <= nel && node <= m.nnods; ++node) {
    const auto& els = node2el[node];
    for (i64 e1 : els) {
      auto& cur = m.nbr[e1 - 1];
      for (i64 e2 : els)
        if (e1 != e2 && std::find(cur.begin(), cur.end(), e2) == cur.end()) cur.push_back(e2);
    }
  }
}

// src/GridAnnulus.jl:72-142  primary_grid
void primary_grid(Mesh& m, i64 ntheta, i64 nr_in) {
  i64 nr = nr_in + 7;  // :75
  i64 nn = nr * ntheta;
  i64 nels = (nr - 1) * ntheta;
  m.ntheta = ntheta;
  m.nr = nr;
  m.r.assign(nn + 1, 0.0);
  m.theta.assign(nn + 1, 0.0);
  double dth = 2 * PI / (double)ntheta;  // :81
  double r_in = 0.1, r_out = R_EARTH;    // :84
  std::vector<double> rcol;
  for (int k = 0; k < 7; ++k) rcol.push_back(RL[k]);
  i64 nlin = nr - 7;
  for (i64 j = 0; j < nlin; ++j) rcol.push_back(nlin > 1 ? lerpi(j, nlin - 1, r_in, r_out) : r_in);
  std::sort(rcol.begin(), rcol.end());  // :87
  for (i64 ii = 1; ii <= ntheta; ++ii)
    for (i64 k = 1; k <= nr; ++k) {
      i64 id = k + nr * (ii - 1);
      m.r[id - 1] = rcol[k - 1];
      m.theta[id - 1] = dth * (double)(ii - 1);  // :92
    }
  // centre :95 already 0,0
  m.e2n.clear();
  m.e2n.resize(nels + ntheta);
  m.el_type.assign(nels + ntheta, 0);
  for (i64 ii = 1; ii <= ntheta; ++ii)  // :99-111
    for (i64 k = 1; k <= nr - 1; ++k) {
      i64 idx = k + (nr - 1) * (ii - 1);
      i64 idx1 = k + nr * (ii - 1);
      i64 idx2 = (ii < ntheta) ? k + nr * ii : k;
      m.e2n[idx - 1] = {idx1, idx2, idx2 + 1, idx1 + 1};
    }
  for (i64 ii = 1; ii <= ntheta; ++ii) {  // :114-121
    i64 idx = 1 + nr * (ii - 1);
    i64 third = (ii == ntheta) ? 1 : idx + nr;
    m.e2n[nels + ii - 1] = {nn + 1, idx, third};
    m.el_type[nels + ii - 1] = 1;
  }
  m.nel = nels + ntheta;
  m.nnods = nn + 1;
  element_neighbours(m);
  m.x.resize(m.nnods);
  m.z.resize(m.nnods);
  for (i64 i = 0; i < m.nnods; ++i) {  // @cartesian :27-29
    m.x[i] = m.r[i] * std::sin(m.theta[i]);
    m.z[i] = m.r[i] * std::cos(m.theta[i]);
  }
}

struct Edges {
  std::vector<std::pair<i64, i64>> nodes;  // sorted (lo, hi)
  std::vector<std::vector<i64>> els;       // edge2el (set semantics; insertion order irrelevant)
};

// src/GridAnnulus.jl:515-595  edge_connectivity (with the slot-1 quirk at :566-571)
void edge_connectivity(const Mesh& m, Edges& E) {
  const i64 nel = m.nel;
  std::vector<i64> el2edge(4 * nel, 0);
  auto has_node = [&](i64 el, i64 nd) {
    const auto& v = m.e2n[el - 1];
    return std::find(v.begin(), v.end(), nd) != v.end();
  };
  for (i64 iel = 1; iel <= nel; ++iel) {
    const auto& el = m.e2n[iel - 1];
    int nedge = (int)el.size();  // only primary nodes exist at this point
    for (int ie = 0; ie < nedge; ++ie) {
      if (el2edge[ie + 4 * (iel - 1)] != 0) continue;
      i64 a = el[ie], b = el[(ie + 1) % nedge];  // map_rectangle / map_triangle :519-526
      if (a > b) std::swap(a, b);                // sort!(edge; dims=1)
      E.nodes.push_back({a, b});
      E.els.push_back({iel});
      i64 gid = (i64)E.nodes.size();
      el2edge[ie + 4 * (iel - 1)] = gid;
      for (i64 ieln : m.nbr[iel - 1]) {
        // issubset(edge[:, iedge], edge_neighbour): both endpoints among the neighbour's nodes;
        // the inner `for i in 1:nedge ... break` always lands on i = 1  (:566-571)
        if (has_node(ieln, a) && has_node(ieln, b)) {
          el2edge[0 + 4 * (ieln - 1)] = gid;
          auto& s = E.els.back();
          if (std::find(s.begin(), s.end(), ieln) == s.end()) s.push_back(ieln);
        }
      }
    }
  }
}

// src/GridAnnulus.jl:607-698 secondary_nodes, :700-708 edge_length, :710-725 correct_theta
void secondary_nodes(Mesh& m, double spacing) {
  Edges E;
  edge_connectivity(m, E);
  const i64 nedges = (i64)E.nodes.size();
  const double eps = 2 * PI - (1.0 - 1.0 / (double)m.ntheta);  // :621
  const i64 icenter = m.nr * m.ntheta + 1;                      // :622
  const i64 nnods0 = (i64)m.r.size();
  std::vector<double> thmid, rmid;
  i64 gidx = 0;
  for (i64 i = 0; i < nedges; ++i) {
    i64 n1 = E.nodes[i].first, n2 = E.nodes[i].second;
    double t1 = m.theta[n1 - 1], t2 = m.theta[n2 - 1];
    double r1 = m.r[n1 - 1], r2 = m.r[n2 - 1];
    if (n1 != icenter && n2 != icenter) {  // correct_theta
      if (std::fabs(t1 - t2) >= eps) {
        if (t1 < PI)
          t1 = t1 + 2 * PI;
        else if (t2 < PI)
          t2 = t2 + 2 * PI;
      }
    } else {
      double tm = std::max(t1, t2);
      t1 = tm;
      t2 = tm;
    }
    double L;
    if (t1 == t2)
      L = std::sqrt(r1 * r1 + r2 * r2 - 2 * r1 * r2 * std::cos(t1 - t2));  // polardistance :706
    else
      L = r1 * std::fabs(t2 - t1);  // arclength :708
    i64 np = (i64)julia_fdiv(L, spacing);  // :642
    if (np > 0) {
      for (i64 j = 1; j <= np; ++j) {
        ++gidx;
        double Lt = t2 - t1, Lr = r2 - r1;
        double dt = Lt * (double)j / (double)(np + 1);  // :653
        double dr = Lr * (double)j / (double)(np + 1);
        thmid.push_back(t1 + dt);
        rmid.push_back(r1 + dr);
        for (i64 iel : E.els[i]) m.e2n[iel - 1].push_back(gidx + nnods0);  // :661-663
      }
    }
  }
  m.theta.insert(m.theta.end(), thmid.begin(), thmid.end());
  m.r.insert(m.r.end(), rmid.begin(), rmid.end());
  m.nnods = (i64)m.r.size();
  m.x.resize(m.nnods);
  m.z.resize(m.nnods);
  for (i64 i = 0; i < m.nnods; ++i) {  // :671
    m.x[i] = m.r[i] * std::sin(m.theta[i]);
    m.z[i] = m.r[i] * std::cos(m.theta[i]);
  }
}

// src/GridAnnulus.jl:374-381
int find_boundary(double ri, const double* rlayer /*8*/) {
  if (ri < rlayer[7]) return 1;
  for (int i = 0; i < 7; ++i)
    if (rlayer[i] > ri && ri > rlayer[i + 1]) return i + 2;
  return 0;  // Julia would return `nothing`
}

// src/GridAnnulus.jl:296-321
void constrain2layers(Mesh& m) {
  const double rlayer[8] = {R_EARTH,        R_EARTH - 20,  R_EARTH - 35,   R_EARTH - 210,
                            R_EARTH - 410,  R_EARTH - 660, R_EARTH - 2740, R_EARTH - 2891.5};
  std::vector<int> lay(m.nel);
  for (i64 i = 0; i < m.nel; ++i) {
    const auto& e = m.e2n[i];
    double c;
    if (m.el_type[i] == 0)
      c = (m.r[e[0] - 1] + m.r[e[1] - 1] + m.r[e[2] - 1] + m.r[e[3] - 1]) * 0.25;
    else
      c = (m.r[e[0] - 1] + m.r[e[1] - 1] + m.r[e[2] - 1]) * 0.33;
    lay[i] = find_boundary(c, rlayer);
  }
  for (i64 i = 0; i < m.nel; ++i) {
    std::vector<i64> keep;
    for (i64 nb : m.nbr[i])
      if (lay[nb - 1] == lay[i]) keep.push_back(nb);
    m.nbr[i].swap(keep);
  }
}

// src/GridAnnulus.jl:910-968
void discontinuous_boundaries(Mesh& m) {
  std::vector<i64> idx;
  i64 counter = m.nnods;
  const i64 nnods = m.nnods;
  for (i64 i = 0; i < m.nel; ++i) {
    auto& e = m.e2n[i];
    if (e.size() < 3) continue;
    double r3 = m.r[e[2] - 1];
    int ib = -1;
    for (int k = 0; k < 7; ++k)
      if (r3 == RL[k]) {
        ib = k;
        break;
      }
    if (ib < 0) continue;
    for (size_t j = 0; j < e.size(); ++j) {
      i64 node = e[j];
      if (m.r[node - 1] == RL[ib]) {
        ++counter;
        e[j] = counter;
        idx.push_back(node);
      }
    }
  }
  const i64 H = (i64)idx.size();
  for (i64 k = 0; k < H; ++k) {
    double th = m.theta[idx[k] - 1];
    double rr = m.r[idx[k] - 1] - 0.05;  // :938
    m.theta.push_back(th);
    m.r.push_back(rr);
    m.x.push_back(rr * std::sin(th));  // polar2cartesian :55
    m.z.push_back(rr * std::cos(th));
  }
  m.halo_rows = 2 * H;
  m.halo.assign(4 * H, 0);
  for (i64 k = 0; k < H; ++k) {  // :945-950  (column-major 2H x 2)
    m.halo[k] = idx[k];
    m.halo[k + 2 * H] = k + 1 + nnods;
    m.halo[k + H + 2 * H] = idx[k];
    m.halo[k + H] = k + 1 + nnods;
  }
  m.nnods = (i64)m.r.size();
}

// src/GridAnnulus.jl:420-452  element_incidence; sparse(I,J,V) sorts rows within a column and ORs duplicates
void element_incidence(Mesh& m) {
  const i64 n = m.nnods;
  std::vector<i64> cnt(n + 1, 0);
  for (i64 e = 0; e < m.nel; ++e) cnt[0] += 0;
  for (i64 e = 0; e < m.nel; ++e)
    for (i64 nd : m.e2n[e]) cnt[nd] += 1 + (i64)m.nbr[e].size();
  std::vector<i64> off(n + 2, 0);
  for (i64 v = 1; v <= n; ++v) off[v + 1] = off[v] + cnt[v];
  std::vector<i64> tmp(off[n + 1]);
  std::vector<i64> fill(off.begin(), off.end());
  for (i64 e = 0; e < m.nel; ++e)
    for (i64 nd : m.e2n[e]) {
      tmp[fill[nd]++] = e + 1;
      for (i64 nb : m.nbr[e]) tmp[fill[nd]++] = nb;
    }
  m.G_colptr.assign(n + 1, 1);
  m.G_rowval.clear();
  for (i64 v = 1; v <= n; ++v) {
    auto b = tmp.begin() + off[v], e = tmp.begin() + off[v + 1];
    std::sort(b, e);
    e = std::unique(b, e);
    m.G_rowval.insert(m.G_rowval.end(), b, e);
    m.G_colptr[v] = (i64)m.G_rowval.size() + 1;
  }
}

// flat views used by the solvers -------------------------------------------------------------------------
template <typename T>
struct Graph2DT {
  i64 n, nel;
  const i64 *e2n_off, *e2n_idx;  // e2n_off[nel+1] 0-based offsets; e2n_idx 1-based node ids
  const i64 *colptr, *rowval;    // Julia CSC, 1-based
  const i64* halo;
  i64 halo_rows;
  const T *x, *z, *U;
};
using Graph2D = Graph2DT<double>;

// src/SSSP/bfm.jl:186 + src/GridAnnulus.jl:808-815: dGi + 2.0 * sqrt(0 + dx^2 + dz^2) / (Ui + Uj)
// T = float restates the Float32 path of src/SSSP/bfm_gpu.jl:487-526 (every array cast to Float32 :170-205):
// dist0[Gi] + 2 * distance(xi, zi, x[Gi], z[Gi]) / (Ui + U[Gi]) with every operation rounded to Float32.
template <typename T>
inline T cand2d(const Graph2DT<T>& g, T dj, i64 i0, i64 j0) {
  T dx = g.x[i0] - g.x[j0], dz = g.z[i0] - g.z[j0];
  T d = T(0);
  d += dx * dx;
  d += dz * dz;
  T len2 = T(2) * std::sqrt(d);
  T w = len2 / (g.U[i0] + g.U[j0]);
  return dj + w;
}

const double INF = std::numeric_limits<double>::infinity();

}  // namespace

extern "C" {

// ------------------------------------------------------------------ annulus builder (init_annulus :57-70)
void* ora_annulus_build(i64 ntheta, i64 nr, double spacing) {
  Mesh* m = new Mesh();
  primary_grid(*m, ntheta, nr);
  secondary_nodes(*m, spacing);
  constrain2layers(*m);
  discontinuous_boundaries(*m);
  element_incidence(*m);
  return m;
}

void ora_mesh_sizes(void* h, i64* out /*[8]: n, nel, sum_e2n, nnzG, halo_rows, sum_nbr, ntheta, nr*/) {
  Mesh* m = (Mesh*)h;
  i64 se = 0, sn = 0;
  for (auto& e : m->e2n) se += (i64)e.size();
  for (auto& e : m->nbr) sn += (i64)e.size();
  out[0] = m->nnods;
  out[1] = m->nel;
  out[2] = se;
  out[3] = (i64)m->G_rowval.size();
  out[4] = m->halo_rows;
  out[5] = sn;
  out[6] = m->ntheta;
  out[7] = m->nr;
}

void ora_mesh_export(void* h, double* x, double* z, double* th, double* r, i64* e2n_off, i64* e2n_idx,
                     i64* colptr, i64* rowval, i64* halo, i64* nbr_off, i64* nbr_idx, int8_t* el_type) {
  Mesh* m = (Mesh*)h;
  std::memcpy(x, m->x.data(), sizeof(double) * m->nnods);
  std::memcpy(z, m->z.data(), sizeof(double) * m->nnods);
  std::memcpy(th, m->theta.data(), sizeof(double) * m->nnods);
  std::memcpy(r, m->r.data(), sizeof(double) * m->nnods);
  i64 o = 0, on = 0;
  for (i64 e = 0; e < m->nel; ++e) {
    e2n_off[e] = o;
    for (i64 nd : m->e2n[e]) e2n_idx[o++] = nd;
    nbr_off[e] = on;
    for (i64 nb : m->nbr[e]) nbr_idx[on++] = nb;
    el_type[e] = m->el_type[e];
  }
  e2n_off[m->nel] = o;
  nbr_off[m->nel] = on;
  std::memcpy(colptr, m->G_colptr.data(), sizeof(i64) * (m->nnods + 1));
  std::memcpy(rowval, m->G_rowval.data(), sizeof(i64) * m->G_rowval.size());
  if (m->halo_rows) std::memcpy(halo, m->halo.data(), sizeof(i64) * 2 * m->halo_rows);
}

void ora_mesh_free(void* h) { delete (Mesh*)h; }

// --------------------------------------------------------------------------------- velocity interpolation
// src/utils.jl:38-44 (buffer < 0) and src/ShortestPath.jl:74-90 (buffer >= 0).  Interpolations.jl
// gridded linear (version unpinned -> "parity unpinned" at the ulp level): i = clamp(searchsortedlast(knots,x),
// 1, nk-1); f = (x-k[i])/(k[i+1]-k[i]); v = (1-f)*y[i] + f*y[i+1]; outside the knots -> error (Throw()).
int ora_interp_velocity(const double* kr, const double* kv, i64 nk, const double* r, i64 n, double buffer,
                        double* out) {
  for (i64 i = 0; i < n; ++i) {
    double xq = r[i];
    if (buffer >= 0.0) {
      for (int k = 0; k < 7; ++k)
        if (xq == RL[k]) {
          xq = xq + buffer;
          break;
        }
    }
    if (!(xq >= kr[0] && xq <= kr[nk - 1])) return 1;  // BoundsError in the reference
    i64 idx = (i64)(std::upper_bound(kr, kr + nk, xq) - kr);  // searchsortedlast (1-based)
    if (idx < 1) idx = 1;
    if (idx > nk - 1) idx = nk - 1;
    double f = (xq - kr[idx - 1]) / (kr[idx] - kr[idx - 1]);
    out[i] = (1.0 - f) * kv[idx - 1] + f * kv[idx];
  }
  return 0;
}

// src/GridAnnulus.jl:823-840 closest_point: first index minimising sqrt((a-pa)^2 + (b-pb)^2)
i64 ora_closest_point(const double* a, const double* b, i64 n, double pa, double pb) {
  double best = INF;
  i64 index = -1;
  for (i64 i = 0; i < n; ++i) {
    double da = a[i] - pa, db = b[i] - pb;
    double di = std::sqrt(da * da + db * db);
    if (di < best) {
      index = i + 1;
      best = di;
    }
  }
  return index;
}

// ------------------------------------------------------------------------------------ bfm (src/SSSP/bfm.jl)
// stats[0] = sweeps, stats[1] = candidate evaluations (E_relaxed), stats[2] = active-vertex updates,
// stats[3] = E_graph (sum over vertices of |scan list|)
extern "C++" {
template <typename T>
static int bfm_impl(i64 n, i64 nel, const i64* e2n_off, const i64* e2n_idx, const i64* colptr, const i64* rowval,
                    const i64* halo, i64 halo_rows, const T* x, const T* z, const T* U, const T* U2,
                    const T* r, i64 source, int nthreads, i64 max_sweeps, T* dist, i64* prev, i64* stats) {
  const T INF = std::numeric_limits<T>::infinity();
  // U2 != null: dual-velocity relax, _relax!(..., U::Matrix) bfm.jl:113-159 (U = U[:,1], U2 = U[:,2])
  Graph2DT<T> g{n, nel, e2n_off, e2n_idx, colptr, rowval, halo, halo_rows, x, z, U};
  if (source < 1 || source > n) return 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  std::vector<T> dist0(n);
  std::vector<uint8_t> Q(n, 0);
  for (i64 i = 0; i < n; ++i) prev[i] = 0;  // reference leaves p undefined (bfm.jl:12)
  // init_halo_path! :64-70 (n = length(halo) / 2 == all 2H rows)
  for (i64 k = 0; k < halo_rows; ++k) {
    i64 h1 = halo[k], h2 = halo[k + halo_rows];
    prev[h2 - 1] = h1;
    prev[h1 - 1] = h2;
  }
  // init_Q! :74-80
  for (i64 p = colptr[source - 1]; p < colptr[source]; ++p) {
    i64 el = rowval[p - 1];
    for (i64 q = e2n_off[el - 1]; q < e2n_off[el]; ++q) Q[e2n_idx[q] - 1] = 1;
  }
  for (i64 i = 0; i < n; ++i) dist[i] = INF;
  dist[source - 1] = T(0);
  for (i64 i = 0; i < n; ++i) dist0[i] = dist[i];
  i64 sweeps = 0, evals = 0, updates = 0;
  std::vector<i64> active;
  active.reserve(n);
  while (true) {
    active.clear();
    for (i64 i = 0; i < n; ++i)
      if (Q[i]) active.push_back(i);  // findall(Q) :107
    if (active.empty()) break;        // sum(Q) != 0 :29
    if (max_sweeps > 0 && sweeps >= max_sweeps) break;
    i64 ev = 0;
    const i64 na = (i64)active.size();
    // relax! :100-111 -> _relax! :161-210
#pragma omp parallel for schedule(static) reduction(+ : ev)
    for (i64 a = 0; a < na; ++a) {
      i64 i0 = active[a];
      T di = dist0[i0];
      for (i64 p = colptr[i0]; p < colptr[i0 + 1]; ++p) {
        i64 el = rowval[p - 1];
        for (i64 q = e2n_off[el - 1]; q < e2n_off[el]; ++q) {
          i64 j0 = e2n_idx[q] - 1;
          T dj = dist0[j0];
          T delta;
          if (dj == INF) {
            delta = INF;
          } else if (!U2) {
            delta = cand2d(g, dj, i0, j0);
          } else {
            // head_idx = (ri > r[Gi]) + 1; tail_idx = (head_idx == 1) + 1; muladd(2, len / (Ui[tail] + U[Gi, head]), dGi)
            const bool down = r[i0] > r[j0];
            const T ut = down ? U[i0] : U2[i0];
            const T uh = down ? U2[j0] : U[j0];
            const T dx = x[i0] - x[j0], dz = z[i0] - z[j0];
            const T q = std::sqrt(dx * dx + dz * dz) / (ut + uh);
            delta = 2 * q + dj;  // muladd: the product by 2 is exact, fused or not
          }
          if (di > delta) {
            di = delta;
            prev[i0] = j0 + 1;
          }
          ++ev;
        }
      }
      dist[i0] = di;
    }
    evals += ev;
    updates += na;
    // update_halo! :54-62 (serial order == single-thread reference)
    for (i64 k = 0; k < halo_rows; ++k) {
      i64 h1 = halo[k] - 1, h2 = halo[k + halo_rows] - 1;
      if (dist[h1] < dist0[h1] && dist[h2] > dist[h1]) {
        dist[h2] = dist[h1];
        prev[h2] = prev[h1];
      }
    }
    std::fill(Q.begin(), Q.end(), 0);  // :37
    // update_Q! :82-98 (byte flags: benign races only)
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
      if (dist[i] < INF && dist[i] < dist0[i]) {
        for (i64 p = colptr[i]; p < colptr[i + 1]; ++p) {
          i64 el = rowval[p - 1];
          for (i64 q = e2n_off[el - 1]; q < e2n_off[el]; ++q) {
            i64 j0 = e2n_idx[q] - 1;
            if (!Q[j0]) Q[j0] = 1;
          }
        }
      }
    }
    std::memcpy(dist0.data(), dist, sizeof(T) * n);  // :43
    ++sweeps;
  }
  if (stats) {
    i64 eg = 0;
    for (i64 i = 0; i < n; ++i)
      for (i64 p = colptr[i]; p < colptr[i + 1]; ++p) {
        i64 el = rowval[p - 1];
        eg += e2n_off[el] - e2n_off[el - 1];
      }
    stats[0] = sweeps;
    stats[1] = evals;
    stats[2] = updates;
    stats[3] = eg;
  }
  return 0;
}
}  // extern "C++"

int ora_bfm(i64 n, i64 nel, const i64* e2n_off, const i64* e2n_idx, const i64* colptr, const i64* rowval,
            const i64* halo, i64 halo_rows, const double* x, const double* z, const double* U, i64 source,
            int nthreads, i64 max_sweeps, double* dist, i64* prev, i64* stats) {
  return bfm_impl<double>(n, nel, e2n_off, e2n_idx, colptr, rowval, halo, halo_rows, x, z, U, nullptr, nullptr, source,
                          nthreads, max_sweeps, dist, prev, stats);
}

// Float32 comparison path (src/SSSP/bfm_gpu.jl:170-205): x, z, U cast to Float32 (round to nearest), travel times
// kept and relaxed in Float32 (:487-526); halo rows applied in serial order like the CPU path.  dist_out holds the
// Float32 values widened to double.
int ora_bfm_f32(i64 n, i64 nel, const i64* e2n_off, const i64* e2n_idx, const i64* colptr, const i64* rowval,
                const i64* halo, i64 halo_rows, const double* x, const double* z, const double* U, i64 source,
                int nthreads, double* dist_out, i64* prev, i64* stats) {
  std::vector<float> xf(n), zf(n), Uf(n), df(n);
  for (i64 i = 0; i < n; ++i) {
    xf[i] = (float)x[i];
    zf[i] = (float)z[i];
    Uf[i] = (float)U[i];
  }
  int rc = bfm_impl<float>(n, nel, e2n_off, e2n_idx, colptr, rowval, halo, halo_rows, xf.data(), zf.data(), Uf.data(),
                           nullptr, nullptr, source, nthreads, 0, df.data(), prev, stats);
  for (i64 i = 0; i < n; ++i) dist_out[i] = (double)df[i];
  return rc;
}

// bfm with U::Matrix (dual velocity): U1 = U[:,1], U2 = U[:,2], r = gr.r
int ora_bfm_dual(i64 n, i64 nel, const i64* e2n_off, const i64* e2n_idx, const i64* colptr, const i64* rowval,
                 const i64* halo, i64 halo_rows, const double* x, const double* z, const double* r, const double* U1,
                 const double* U2, i64 source, int nthreads, double* dist, i64* prev, i64* stats) {
  return bfm_impl<double>(n, nel, e2n_off, e2n_idx, colptr, rowval, halo, halo_rows, x, z, U1, U2, r, source, nthreads,
                          0, dist, prev, stats);
}

// dual_velocity(r, interpolant; buffer) src/utils.jl:51-66 -> V[n x 2] column-major
int ora_dual_velocity(const double* kr, const double* kv, i64 nk, const double* r, i64 n, double buffer, double* out) {
  auto itp = [&](double xq, double& v) {
    if (!(xq >= kr[0] && xq <= kr[nk - 1])) return 1;
    i64 idx = (i64)(std::upper_bound(kr, kr + nk, xq) - kr);
    if (idx < 1) idx = 1;
    if (idx > nk - 1) idx = nk - 1;
    double f = (xq - kr[idx - 1]) / (kr[idx] - kr[idx - 1]);
    v = (1.0 - f) * kv[idx - 1] + f * kv[idx];
    return 0;
  };
  for (i64 i = 0; i < n; ++i) {
    bool on = false;
    for (int k = 0; k < 7; ++k) on = on || r[i] == RL[k];
    if (on) {
      if (itp(r[i] - buffer, out[i]) || itp(r[i] + buffer, out[n + i])) return 1;
    } else {
      if (itp(r[i], out[i])) return 1;
      out[n + i] = out[i];
    }
  }
  return 0;
}

// Independent check: binary-heap Dijkstra on the same (symmetric) scan lists + halo rows as 0-weight edges.
// Must give bit-identical dist (fp `+` is monotone => least fixed point is schedule independent).
int ora_dijkstra(i64 n, i64 nel, const i64* e2n_off, const i64* e2n_idx, const i64* colptr, const i64* rowval,
                 const i64* halo, i64 halo_rows, const double* x, const double* z, const double* U, i64 source,
                 double* dist) {
  Graph2D g{n, nel, e2n_off, e2n_idx, colptr, rowval, halo, halo_rows, x, z, U};
  std::vector<std::vector<i64>> twins(n);
  for (i64 k = 0; k < halo_rows; ++k) twins[halo[k] - 1].push_back(halo[k + halo_rows] - 1);
  for (i64 i = 0; i < n; ++i) dist[i] = INF;
  dist[source - 1] = 0.0;
  typedef std::pair<double, i64> PQE;
  std::priority_queue<PQE, std::vector<PQE>, std::greater<PQE>> pq;
  pq.push({0.0, source - 1});
  std::vector<uint8_t> done(n, 0);
  while (!pq.empty()) {
    PQE t = pq.top();
    pq.pop();
    i64 u = t.second;
    if (done[u]) continue;
    done[u] = 1;
    double du = dist[u];
    for (i64 v : twins[u])
      if (du < dist[v]) {
        dist[v] = du;
        pq.push({du, v});
      }
    for (i64 p = colptr[u]; p < colptr[u + 1]; ++p) {
      i64 el = rowval[p - 1];
      for (i64 q = e2n_off[el - 1]; q < e2n_off[el]; ++q) {
        i64 v = e2n_idx[q] - 1;
        if (done[v]) continue;
        double c = cand2d(g, du, v, u);  // evaluated from v's side exactly as _relax! does
        if (c < dist[v]) {
          dist[v] = c;
          pq.push({c, v});
        }
      }
    }
  }
  return 0;
}

// src/SSSP/ssspm.jl:30-40  recontruct_path(prev::Vector, source, receiver) -> [receiver ... source]
// returns path length, or -1 if the chase does not reach `source` within n steps (the reference would
// loop forever / throw), or -(needed) if cap is too small.
i64 ora_reconstruct_path(const i64* prev, i64 n, i64 source, i64 receiver, i64* out, i64 cap) {
  i64 len = 0;
  if (receiver < 1 || receiver > n) return -1;
  if (len < cap) out[len] = receiver;
  ++len;
  i64 ip = (receiver == source) ? source : prev[receiver - 1];
  i64 steps = 0;
  while (ip != source) {
    if (ip < 1 || ip > n || ++steps > n) return -1;
    if (len < cap) out[len] = ip;
    ++len;
    ip = prev[ip - 1];
  }
  if (len < cap) out[len] = source;
  ++len;
  return len <= cap ? len : -len;
}

// ------------------------------------------------------------------------------------------ 3-D grid
// src/StructuredGrid.jl:35-45 grid (LinRange axes), :90-96 CartesianIndex (x fastest), :225-235 spherical2cart
// coord_system 0: Cartesian axes as they are; 1: axes are (theta, phi, r) -> spherical2cart.
void ora_grid3d_coords(const double* c0, const double* c1, const i64* nn, int coord_system, double* X,
                       double* Y, double* Z) {
  std::vector<double> ax[3];
  for (int d = 0; d < 3; ++d) {
    ax[d].resize(nn[d]);
    for (i64 i = 0; i < nn[d]; ++i) ax[d][i] = nn[d] > 1 ? lerpi(i, nn[d] - 1, c0[d], c1[d]) : c0[d];
  }
  i64 I = 0;
  for (i64 k = 0; k < nn[2]; ++k)
    for (i64 j = 0; j < nn[1]; ++j)
      for (i64 i = 0; i < nn[0]; ++i, ++I) {
        double a = ax[0][i], b = ax[1][j], c = ax[2][k];
        if (coord_system == 0) {
          X[I] = a;
          Y[I] = b;
          Z[I] = c;
        } else {
          X[I] = c * std::cos(b) * std::sin(a);
          Y[I] = c * std::sin(b) * std::sin(a);
          Z[I] = c * std::cos(a);
        }
      }
}

// 3-D edge weight, two reference definitions (selected with ora_set_weight3d):
//   mode 0 (default)  src/SSSP/weights.jl:20   edge_weight = distance3D(p1,p2) * (1/abs(U1+U2)) * 2
//   mode 1            src/Dijsktra.jl:388 (BFM/foo!) and :44 (dijsktra)   fw(a,b) / abs(U[a] + U[b]) * 0.5, fw = distance3D
// distance3D: StructuredGrid.jl:239-241.
static int g_weight3d = 0;
void ora_set_weight3d(int mode) { g_weight3d = mode ? 1 : 0; }
// window half width of nodal_incidence(gr; neighbour_levels = L) (StructuredGrid.jl:177-212): level 0 is the
// 26-neighbourhood (radius 1); every expansion round deep-copies Q and unions Q0[n] for n in Q0[i], so the radius
// DOUBLES per round: 2^L (5^3 window at L = 1, 9^3 at L = 2, 17^3 at L = 3), self included once L >= 1.
int ora_window3d(int star_levels) { return 1 << star_levels; }
extern "C++" {
template <typename T>
static inline T cand3d(const T* X, const T* Y, const T* Z, const T* U, T dj, i64 i, i64 j) {
  T dx = X[i] - X[j], dy = Y[i] - Y[j], dz = Z[i] - Z[j];
  T d = std::sqrt(dx * dx + dy * dy + dz * dz);
  T w = g_weight3d ? d / std::fabs(U[i] + U[j]) * T(0.5) : d * (T(1) / std::fabs(U[i] + U[j])) * T(2);
  return dj + w;
}
}  // extern "C++"

// nodal_incidence(gr; neighbour_levels) transliterated with the reference's container semantics (Dict of Sets built
// from the 8-node hexes of connectivity(gr), StructuredGrid.jl:121-168, then `neighbour_levels` rounds of
// Q0 = deepcopy(Q); union!(Q[i], Q0[n]) for n in Q0[i]).  Only for small grids: pins the implicit window used by the
// solvers.  Output CSR: off[n+1], list ascending per node (1-based); with list == null only off is written.
void ora_nodal_incidence3d(const i64* nn, int star_levels, i64* off, i64* list) {
  const i64 nx = nn[0], ny = nn[1], nz = nn[2], n = nx * ny * nz;
  std::vector<std::set<i64>> Q(n);
  for (i64 k = 0; k + 1 < nz; ++k)
    for (i64 j = 0; j + 1 < ny; ++j)
      for (i64 i = 0; i + 1 < nx; ++i) {
        i64 el[8];
        int c = 0;
        for (i64 dk = 0; dk < 2; ++dk)
          for (i64 dj = 0; dj < 2; ++dj)
            for (i64 di = 0; di < 2; ++di) el[c++] = (i + di) + nx * ((j + dj) + ny * (k + dk));
        for (int a = 0; a < 8; ++a)
          for (int b = 0; b < 8; ++b)
            if (a != b) Q[el[a]].insert(el[b]);
      }
  for (int lev = 0; lev < star_levels; ++lev) {
    const std::vector<std::set<i64>> Q0 = Q;
    for (i64 i = 0; i < n; ++i)
      for (i64 m : Q0[i]) Q[i].insert(Q0[m].begin(), Q0[m].end());
  }
  i64 o = 0;
  for (i64 i = 0; i < n; ++i) {
    off[i] = o;
    if (list)
      for (i64 m : Q[i]) list[o++] = m + 1;
    else
      o += (i64)Q[i].size();
  }
  off[n] = o;
}

// star-L adjacency of nodal_incidence (StructuredGrid.jl:177-223): L = 0 -> 26-neighbourhood without self;
// L >= 1 -> clipped (2*2^L+1)^3 window INCLUDING self (checked against ora_nodal_incidence3d, the literal Dict/Set build).  Canonical scan order = ascending linear id (the
// reference iterates a Julia Set, whose order is not reproducible).  Control flow = BFM/foo!/goo!
// (src/Dijsktra.jl:294-343, 376-403).
extern "C++" {
template <typename T>
static int bfm3d_impl(const i64* nn, int star_levels, const T* X, const T* Y, const T* Z, const T* U, i64 source,
                      int nthreads, i64 max_sweeps, T* dist, i64* prev, i64* stats) {
  const T INF = std::numeric_limits<T>::infinity();
  const i64 nx = nn[0], ny = nn[1], nz = nn[2], n = nx * ny * nz;
  const i64 w = (i64)1 << star_levels;
  const bool self = star_levels >= 1;
  if (source < 1 || source > n) return 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  std::vector<T> dist0(n);
  std::vector<uint8_t> act(n, 0), imp(n, 0);
  for (i64 i = 0; i < n; ++i) {
    dist[i] = INF;
    prev[i] = 0;
  }
  dist[source - 1] = T(0);
  dist0.assign(dist, dist + n);
  auto mark_window = [&](i64 I, std::vector<uint8_t>& flags) {
    i64 i = I % nx, j = (I / nx) % ny, k = I / (nx * ny);
    for (i64 kk = std::max<i64>(0, k - w); kk <= std::min(nz - 1, k + w); ++kk)
      for (i64 jj = std::max<i64>(0, j - w); jj <= std::min(ny - 1, j + w); ++jj)
        for (i64 ii = std::max<i64>(0, i - w); ii <= std::min(nx - 1, i + w); ++ii) {
          i64 J = ii + nx * (jj + ny * kk);
          if (J == I && !self) continue;
          flags[J] = 1;
        }
  };
  mark_window(source - 1, act);
  i64 sweeps = 0, evals = 0, updates = 0;
  std::vector<i64> active;
  while (true) {
    active.clear();
    for (i64 i = 0; i < n; ++i)
      if (act[i]) active.push_back(i);
    if (active.empty()) break;
    if (max_sweeps > 0 && sweeps >= max_sweeps) break;
    i64 ev = 0;
    const i64 na = (i64)active.size();
#pragma omp parallel for schedule(static) reduction(+ : ev)
    for (i64 a = 0; a < na; ++a) {
      i64 I = active[a];
      i64 i = I % nx, j = (I / nx) % ny, k = I / (nx * ny);
      T di = dist0[I];
      for (i64 kk = std::max<i64>(0, k - w); kk <= std::min(nz - 1, k + w); ++kk)
        for (i64 jj = std::max<i64>(0, j - w); jj <= std::min(ny - 1, j + w); ++jj)
          for (i64 ii = std::max<i64>(0, i - w); ii <= std::min(nx - 1, i + w); ++ii) {
            i64 J = ii + nx * (jj + ny * kk);
            if (J == I && !self) continue;
            T dj = dist0[J];
            T t = (dj == INF) ? INF : cand3d<T>(X, Y, Z, U, dj, I, J);
            if (di > t) {
              di = t;
              prev[I] = J + 1;
            }
            ++ev;
          }
      dist[I] = di;
    }
    evals += ev;
    updates += na;
    std::fill(act.begin(), act.end(), 0);
    for (i64 I = 0; I < n; ++I)
      if (dist[I] < dist0[I]) mark_window(I, act);  // goo!
    std::memcpy(dist0.data(), dist, sizeof(T) * n);
    ++sweeps;
  }
  if (stats) {
    i64 eg = 0;
    for (i64 k = 0; k < nz; ++k)
      for (i64 j = 0; j < ny; ++j)
        for (i64 i = 0; i < nx; ++i) {
          i64 cx = std::min(nx - 1, i + w) - std::max<i64>(0, i - w) + 1;
          i64 cy = std::min(ny - 1, j + w) - std::max<i64>(0, j - w) + 1;
          i64 cz = std::min(nz - 1, k + w) - std::max<i64>(0, k - w) + 1;
          eg += cx * cy * cz - (self ? 0 : 1);
        }
    stats[0] = sweeps;
    stats[1] = evals;
    stats[2] = updates;
    stats[3] = eg;
  }
  return 0;
}
}  // extern "C++"

int ora_bfm3d(const i64* nn, int star_levels, const double* X, const double* Y, const double* Z,
              const double* U, i64 source, int nthreads, i64 max_sweeps, double* dist, i64* prev, i64* stats) {
  return bfm3d_impl<double>(nn, star_levels, X, Y, Z, U, source, nthreads, max_sweeps, dist, prev, stats);
}

// Float32 variant (the 3-D benchmarks of the reference run the grid in Float32, benchmarks/cpu.jl:9-13): coordinates
// and U cast to Float32, weights.jl:20 evaluated in Float32; dist_out = the Float32 values widened to double.
int ora_bfm3d_f32(const i64* nn, int star_levels, const double* X, const double* Y, const double* Z,
                  const double* U, i64 source, int nthreads, double* dist_out, i64* prev, i64* stats) {
  const i64 n = nn[0] * nn[1] * nn[2];
  std::vector<float> Xf(n), Yf(n), Zf(n), Uf(n), df(n);
  for (i64 i = 0; i < n; ++i) {
    Xf[i] = (float)X[i];
    Yf[i] = (float)Y[i];
    Zf[i] = (float)Z[i];
    Uf[i] = (float)U[i];
  }
  int rc = bfm3d_impl<float>(nn, star_levels, Xf.data(), Yf.data(), Zf.data(), Uf.data(), source, nthreads, 0,
                             df.data(), prev, stats);
  for (i64 i = 0; i < n; ++i) dist_out[i] = (double)df[i];
  return rc;
}

int ora_dijkstra3d(const i64* nn, int star_levels, const double* X, const double* Y, const double* Z,
                   const double* U, i64 source, double* dist) {
  const i64 nx = nn[0], ny = nn[1], nz = nn[2], n = nx * ny * nz;
  const i64 w = (i64)1 << star_levels;
  for (i64 i = 0; i < n; ++i) dist[i] = INF;
  dist[source - 1] = 0.0;
  typedef std::pair<double, i64> PQE;
  std::priority_queue<PQE, std::vector<PQE>, std::greater<PQE>> pq;
  pq.push({0.0, source - 1});
  std::vector<uint8_t> done(n, 0);
  while (!pq.empty()) {
    PQE t = pq.top();
    pq.pop();
    i64 I = t.second;
    if (done[I]) continue;
    done[I] = 1;
    i64 i = I % nx, j = (I / nx) % ny, k = I / (nx * ny);
    for (i64 kk = std::max<i64>(0, k - w); kk <= std::min(nz - 1, k + w); ++kk)
      for (i64 jj = std::max<i64>(0, j - w); jj <= std::min(ny - 1, j + w); ++jj)
        for (i64 ii = std::max<i64>(0, i - w); ii <= std::min(nx - 1, i + w); ++ii) {
          i64 J = ii + nx * (jj + ny * kk);
          if (J == I || done[J]) continue;
          double c = cand3d(X, Y, Z, U, dist[I], J, I);
          if (c < dist[J]) {
            dist[J] = c;
            pq.push({c, J});
          }
        }
  }
  return 0;
}

// src/Interpolations/interpolation.jl:5-18 interpolate!(V, gr); bilinear.jl:1-30; barycentric.jl:1-32.
// Points are Point2D{Polar}(theta, r) (GridAnnulus.jl:25): .x = theta, .z = r.  The reference wraps the sums in
// @muladd (fusing is up to LLVM -> unpinned at the ulp level); here: plain left-to-right arithmetic.
void ora_interpolate_cells(i64 nel, const i64* e2n_off, const i64* e2n_idx, const int8_t* el_type,
                           const double* theta, const double* r, double* V) {
  for (i64 e = 0; e < nel; ++e) {
    const i64* el = e2n_idx + e2n_off[e];
    const i64 len = e2n_off[e + 1] - e2n_off[e];
    if (el_type[e] == 0) {
      if (len < 4) continue;
      double z1 = r[el[0] - 1], z2 = r[el[3] - 1];
      double x1 = theta[el[0] - 1], x2 = theta[el[1] - 1];
      if (x2 - x1 > PI) x1 += 2 * PI;
      const double vu1 = V[el[0] - 1], vu2 = V[el[1] - 1], vu3 = V[el[2] - 1], vu4 = V[el[3] - 1];
      const double dx21 = x2 - x1, dz21 = z2 - z1;
      for (i64 q = 4; q < len; ++q) {
        const i64 nd = el[q] - 1;
        const double dx2 = x2 - theta[nd], dx1 = theta[nd] - x1, dz2 = z2 - r[nd], dz1 = r[nd] - z1;
        V[nd] = 1 / (dx21 * dz21) * (vu1 * dx2 * dz2 + vu2 * dx1 * dz2 + vu4 * dx2 * dz1 + vu3 * dx1 * dz1);
      }
    } else {
      if (len < 3) continue;
      const double x1 = theta[el[0] - 1], x2 = theta[el[1] - 1], x3 = theta[el[2] - 1];
      const double z1 = r[el[0] - 1], z2 = r[el[1] - 1], z3 = r[el[2] - 1];
      const double vu1 = V[el[0] - 1], vu2 = V[el[1] - 1], vu3 = V[el[2] - 1];
      for (i64 q = 3; q < len; ++q) {
        const i64 nd = el[q] - 1;
        const double x = theta[nd], z = r[nd];
        const double den = (z2 - z3) * (x1 - x3) + (x3 - x2) * (z1 - z3);
        const double N1 = ((z2 - z3) * (x - x3) + (x3 - x2) * (z - z3)) / den;
        const double N2 = ((z3 - z1) * (x - x3) + (x1 - x3) * (z - z3)) / den;
        const double N3 = 1 - N1 - N2;
        V[nd] = N1 * vu1 + N2 * vu2 + N3 * vu3;
      }
    }
  }
}

// src/GridAnnulus.jl:763-804 nodal_incidence(gr::Grid2D): Q[nJ] = nodes of every cell containing nJ, self excluded,
// each once.  Output: deg[n] and (if list != null) the neighbours ASCENDING per node (set semantics).
void ora_nodal_adjacency(i64 n, i64 nel, const i64* e2n_off, const i64* e2n_idx, i64* deg, i64* off, i64* list) {
  std::vector<std::vector<i64>> Q(n);
  for (i64 e = 0; e < nel; ++e)
    for (i64 a = e2n_off[e]; a < e2n_off[e + 1]; ++a)
      for (i64 b = e2n_off[e]; b < e2n_off[e + 1]; ++b)
        if (e2n_idx[a] != e2n_idx[b]) Q[e2n_idx[a] - 1].push_back(e2n_idx[b]);
  i64 o = 0;
  for (i64 v = 0; v < n; ++v) {
    std::sort(Q[v].begin(), Q[v].end());
    Q[v].erase(std::unique(Q[v].begin(), Q[v].end()), Q[v].end());
    deg[v] = (i64)Q[v].size();
    off[v] = o;
    if (list) std::copy(Q[v].begin(), Q[v].end(), list + o);
    o += deg[v];
  }
  off[n] = o;
}

// ---------------------------------------------------------------------------------------------------------
// Multiphase / layer-restricted propagation.
// partition_grid(gr) src/topology/topology.jl:183-206 with find_layer_number :137-147: id > 0 = "Layer_id",
// id < 0 = "Boundary_(-id)"; the radius is rounded to 2 digits first (round(r; digits = 2) = round(r * 100) / 100).
void ora_partition_grid(const double* r, i64 n, int32_t* id) {
  for (i64 i = 0; i < n; ++i) {
    const double ri = std::nearbyint(r[i] * 100.0) / 100.0;
    int b = 0;
    for (int k = 0; k < 7; ++k)
      if (ri == RL[k]) {
        b = k + 1;
        break;
      }
    if (b) {
      id[i] = -b;
      continue;
    }
    int layer = 0;
    if (ri > RL[0]) {
      layer = 1;
    } else if (ri < RL[6]) {
      layer = 8;
    } else {
      for (int k = 0; k < 6; ++k)
        if (RL[k] > ri && ri > RL[k + 1]) {
          layer = k + 2;
          break;
        }
    }
    id[i] = layer;  // 0 = find_layer_number returned nothing (cannot happen for rounded radii off the boundaries)
  }
}

// The inner loop of bfm_multiphase (src/SSSP/bfm_multiphase.jl:118-150) restated on the two-level graph of bfm: Jacobi
// sweeps that CONTINUE from a given (dist, prev) state and only ever relax / activate nodes with allowed[i] != 0
// (`ID[Gi] in current_level`, :121, :196); the frontier starts as the allowed nodes of the star patches of the seed
// nodes; update_halo! acts on allowed targets only.  The reference routine itself is unfinished (it calls the undefined
// _relax_bfm! and fillfalse!), so the relax step is _relax!(U::Vector) of bfm.jl:161-210.
int ora_bfm_continue(i64 n, i64 nel, const i64* e2n_off, const i64* e2n_idx, const i64* colptr, const i64* rowval,
                     const i64* halo, i64 halo_rows, const double* x, const double* z, const double* U,
                     const uint8_t* allowed, const i64* seeds, i64 nseeds, int nthreads, double* dist, i64* prev,
                     i64* stats) {
  Graph2D g{n, nel, e2n_off, e2n_idx, colptr, rowval, halo, halo_rows, x, z, U};
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  // A restart treats the seeds as nodes that have just improved: their travel times first cross their halo rows
  // (update_halo! with "is a seed" in place of "improved in this sweep", serial row order, allowed targets only; a
  // node set this way counts as a seed for the rows after it), then the frontier is the allowed part of the star
  // patches of all of them (update_Q! for the seeds).  Without this a phase that restarts on a discontinuity could
  // never enter the layer below it: the twins of the boundary nodes only receive values through the halo rule.
  std::vector<uint8_t> seeded(n, 0);
  for (i64 k = 0; k < nseeds; ++k) {
    if (seeds[k] < 1 || seeds[k] > n) return 1;
    seeded[seeds[k] - 1] = 1;
  }
  for (i64 k = 0; k < halo_rows; ++k) {
    const i64 h1 = halo[k] - 1, h2 = halo[k + halo_rows] - 1;
    if (!seeded[h1] || (allowed && !allowed[h2])) continue;
    if (dist[h2] > dist[h1]) {
      dist[h2] = dist[h1];
      prev[h2] = prev[h1];
      seeded[h2] = 1;
    }
  }
  std::vector<double> dist0(dist, dist + n);
  std::vector<uint8_t> Q(n, 0);
  for (i64 s = 0; s < n; ++s) {
    if (!seeded[s]) continue;
    for (i64 p = colptr[s]; p < colptr[s + 1]; ++p) {
      const i64 el = rowval[p - 1];
      for (i64 q = e2n_off[el - 1]; q < e2n_off[el]; ++q)
        if (!allowed || allowed[e2n_idx[q] - 1]) Q[e2n_idx[q] - 1] = 1;
    }
  }
  i64 sweeps = 0, evals = 0, updates = 0;
  std::vector<i64> active;
  while (true) {
    active.clear();
    for (i64 i = 0; i < n; ++i)
      if (Q[i]) active.push_back(i);
    if (active.empty()) break;
    i64 ev = 0;
    const i64 na = (i64)active.size();
#pragma omp parallel for schedule(static) reduction(+ : ev)
    for (i64 a = 0; a < na; ++a) {
      const i64 i0 = active[a];
      double di = dist0[i0];
      for (i64 p = colptr[i0]; p < colptr[i0 + 1]; ++p) {
        const i64 el = rowval[p - 1];
        for (i64 q = e2n_off[el - 1]; q < e2n_off[el]; ++q) {
          const i64 j0 = e2n_idx[q] - 1;
          const double dj = dist0[j0];
          const double delta = dj == INF ? INF : cand2d(g, dj, i0, j0);
          if (di > delta) {
            di = delta;
            prev[i0] = j0 + 1;
          }
          ++ev;
        }
      }
      dist[i0] = di;
    }
    evals += ev;
    updates += na;
    for (i64 k = 0; k < halo_rows; ++k) {
      const i64 h1 = halo[k] - 1, h2 = halo[k + halo_rows] - 1;
      if (allowed && !allowed[h2]) continue;
      if (dist[h1] < dist0[h1] && dist[h2] > dist[h1]) {
        dist[h2] = dist[h1];
        prev[h2] = prev[h1];
      }
    }
    std::fill(Q.begin(), Q.end(), 0);
    for (i64 i = 0; i < n; ++i)
      if (dist[i] < INF && dist[i] < dist0[i])
        for (i64 p = colptr[i]; p < colptr[i + 1]; ++p) {
          const i64 el = rowval[p - 1];
          for (i64 q = e2n_off[el - 1]; q < e2n_off[el]; ++q) {
            const i64 j0 = e2n_idx[q] - 1;
            if (!allowed || allowed[j0]) Q[j0] = 1;
          }
        }
    std::memcpy(dist0.data(), dist, sizeof(double) * n);
    ++sweeps;
  }
  if (stats) {
    stats[0] = sweeps;
    stats[1] = evals;
    stats[2] = updates;
    stats[3] = 0;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Alternative solvers of the reference on the star-0 node graph nodal_incidence(gr) (CSR: off[n+1] 0-based, list 1-based
// neighbour ids).  Weight (both): 2 * distance(xi, zi, x[i], z[i]) / abs(U[i] + Ui) with the scalar distance
// sqrt((ax-bx)^2 + (az-bz)^2) (src/GridAnnulus.jl:806).
static inline double w_nodal(const double* x, const double* z, const double* U, i64 a, i64 b) {
  const double dx = x[a] - x[b], dz = z[a] - z[b];
  return 2 * std::sqrt(dx * dx + dz * dz) / std::fabs(U[b] + U[a]);
}

// dijkstra(G::Dict, source, gr, U) src/SSSP/dijkstra.jl:68-136, _relax_dijkstra! :138-162, min_distance :164-178,
// literally: Q is a set scanned linearly for its minimum (O(|Q|) per settled node).  The reference's Q is a Julia Set,
// whose iteration order decides between equal distances and is not reproducible: an ordered set stands in for it
// (first minimum in ascending id).  prev: 0 = never set.
int ora_dijkstra_nodal(i64 n, const i64* off, const i64* list, const double* x, const double* z, const double* U,
                       i64 source, double* dist, i64* prev) {
  if (source < 1 || source > n) return 1;
  for (i64 i = 0; i < n; ++i) {
    dist[i] = INF;
    prev[i] = 0;
  }
  dist[source - 1] = 0.0;
  std::set<i64> Q;
  Q.insert(source - 1);
  std::vector<uint8_t> settled(n, 0);
  while (!Q.empty()) {
    double d = INF;
    i64 Qi = -1;
    for (i64 q : Q)
      if (dist[q] < d) {
        d = dist[q];
        Qi = q;
      }
    if (Qi < 0) break;  // only unreachable (Inf) entries cannot occur: nodes enter Q with a finite value
    const double di = dist[Qi];
    for (i64 e = off[Qi]; e < off[Qi + 1]; ++e) {
      const i64 i = list[e] - 1;
      if (settled[i]) continue;
      const double delta = di + w_nodal(x, z, U, Qi, i);
      if (delta < dist[i]) {
        prev[i] = Qi + 1;
        dist[i] = delta;
        Q.insert(i);
      }
    }
    Q.erase(Qi);
    settled[Qi] = 1;
  }
  return 0;
}

// radius_stepping(Gsp, source, gr, U) src/SSSP/radius_stepping.jl:7-46 with relaxation! :58-71 (single thread: the
// threaded loop races on dist / p), min_distance :73-84 and update! :48-56, literally.  Returns the iteration count.
i64 ora_radius_stepping_nodal(i64 n, const i64* off, const i64* list, const double* x, const double* z,
                              const double* U, i64 source, double* dist, i64* prev) {
  if (source < 1 || source > n) return -1;
  std::vector<uint8_t> Q(n, 1), F(n, 0);
  Q[source - 1] = 0;
  F[source - 1] = 1;
  for (i64 i = 0; i < n; ++i) {
    dist[i] = INF;
    prev[i] = 0;
  }
  dist[source - 1] = 0.0;
  i64 it = 1;
  auto left = [&]() {
    i64 c = 0;
    for (i64 i = 0; i < n; ++i) c += Q[i];
    return c;
  };
  while (left() != 0) {
    for (i64 i = 0; i < n; ++i)  // relaxation!
      if (F[i])
        for (i64 e = off[i]; e < off[i + 1]; ++e) {
          const i64 j = list[e] - 1;
          if (!Q[j]) continue;
          const double delta = dist[i] + w_nodal(x, z, U, i, j);
          if (dist[j] > delta) {
            dist[j] = delta;
            prev[j] = i + 1;
          }
        }
    double D = INF;  // min_distance(Q, dist)
    for (i64 i = 0; i < n; ++i)
      if (Q[i] && dist[i] < D) D = dist[i];
    for (i64 i = 0; i < n; ++i) {  // update!
      F[i] = 0;
      if (Q[i] && dist[i] <= D) {
        Q[i] = 0;
        F[i] = 1;
      }
    }
    ++it;
  }
  return it;
}

int ora_num_threads() {
#ifdef _OPENMP
  return omp_get_num_procs();  // not omp_get_max_threads(): an earlier solve's omp_set_num_threads(1) would stick
#else
  return 1;
#endif
}

}  // extern "C"
