// cf_host_driver.cpp -- TEST INFRASTRUCTURE: runs the closed-form mesh description of
// raytracer.jl_b200/csrc/annulus_cf.cuh sequentially on the CPU so that its integer topology can be compared
// with the oracle's literal restatement without a GPU (tests/test_annulus_cf.py).  The product runs the very
// same inline functions inside CUDA kernels (annulus_build.cu); nothing here is linked into librt_sssp.so.
#include <cstring>
#include <vector>

#include "../raytracer.jl_b200/csrc/annulus_cf.cuh"

using cf::i64;

struct CfMesh {
  cf::HostParams hp;
  std::vector<i64> eoff, twin_off;
  cf::Tables tb;
  std::vector<double> theta, r;
  std::vector<i64> e2n_off, e2n_idx, g_off, g_idx, halo, nbr_off, nbr_idx;
};

extern "C" {

void* cfh_build(i64 ntheta, i64 nr, double spacing) {
  CfMesh* m = new CfMesh();
  cf::make_params(ntheta, nr, spacing, m->hp);
  const cf::Params& p = m->hp.p;
  m->eoff.assign(p.nE + 1, 0);
  for (i64 g = 1; g <= p.nE; ++g) m->eoff[g] = m->eoff[g - 1] + cf::edge_npoints(p, g);
  cf::Tables& tb = m->tb;
  tb.eoff = m->eoff.data();
  tb.kd = m->hp.kd.data();
  tb.nnods1 = p.nnods0 + m->eoff[p.nE];
  const i64 nb = 7 * p.T;
  m->twin_off.assign(nb + 1, 0);
  for (i64 b = 0; b < nb; ++b) {
    const i64 e = cf::quad_id(p, m->hp.kd[b % 7], b / 7 + 1);
    m->twin_off[b + 1] = m->twin_off[b] + 2 + cf::edge_np_from_off(tb, cf::edge_top(e));
  }
  tb.twin_off = m->twin_off.data();
  const i64 H = m->twin_off[nb];
  tb.nnods = tb.nnods1 + H;
  const i64 n = tb.nnods;
  // coordinates (theta, r)
  m->theta.assign(n, 0.0);
  m->r.assign(n, 0.0);
  for (i64 c = 1; c <= p.T; ++c)
    for (i64 k = 1; k <= p.M; ++k) {
      m->theta[cf::ring_node(p, k, c) - 1] = p.dth * (double)(c - 1);
      m->r[cf::ring_node(p, k, c) - 1] = p.rc[k];
    }
  for (i64 g = 1; g <= p.nE; ++g) {
    const i64 np = cf::edge_np_from_off(tb, g);
    for (i64 j = 1; j <= np; ++j)
      cf::secondary_coord(p, g, np, j, m->theta[p.nnods0 + m->eoff[g - 1] + j - 1], m->r[p.nnods0 + m->eoff[g - 1] + j - 1]);
  }
  m->halo.assign(4 * H, 0);
  for (i64 b = 0; b < nb; ++b)
    for (i64 pos = 0; pos < m->twin_off[b + 1] - m->twin_off[b]; ++pos) {
      i64 o, t;
      cf::halo_pair(p, tb, b, pos, o, t);
      const i64 h = m->twin_off[b] + pos;
      m->halo[h] = o;
      m->halo[h + 2 * H] = t;
      m->halo[h + H] = t;
      m->halo[h + H + 2 * H] = o;
      m->theta[t - 1] = m->theta[o - 1];
      m->r[t - 1] = m->r[o - 1] - 0.05;
    }
  // e2n
  m->e2n_off.assign(p.nel + 1, 0);
  for (i64 e = 1; e <= p.nel; ++e) m->e2n_off[e] = m->e2n_off[e - 1] + cf::elem_list_len(p, tb, e);
  m->e2n_idx.assign(m->e2n_off[p.nel], 0);
  for (i64 e = 1; e <= p.nel; ++e) cf::elem_list_fill(p, tb, e, m->e2n_idx.data() + m->e2n_off[e - 1]);
  // G
  m->g_off.assign(n + 1, 0);
  i64 set[64];
  for (i64 v = 1; v <= n; ++v) {
    const i64 len = (v == p.nnods0) ? cf::g_column_centre_len(p) : cf::g_column(p, tb, v, set);
    m->g_off[v] = m->g_off[v - 1] + len;
  }
  m->g_idx.assign(m->g_off[n], 0);
  for (i64 v = 1; v <= n; ++v) {
    i64* out = m->g_idx.data() + m->g_off[v - 1];
    if (v == p.nnods0) {
      for (i64 i = 0; i < cf::g_column_centre_len(p); ++i) out[i] = cf::g_column_centre_entry(p, i);
    } else {
      const int len = cf::g_column(p, tb, v, set);
      for (int i = 0; i < len; ++i) out[i] = set[i];
    }
  }
  // neighbours (ascending)
  m->nbr_off.assign(p.nel + 1, 0);
  i64 nb12[12];
  for (i64 e = 1; e <= p.nel; ++e) {
    const int c = cf::element_neighbours(p, e, nb12);
    m->nbr_off[e] = m->nbr_off[e - 1] + c;
    for (int i = 0; i < c; ++i) m->nbr_idx.push_back(nb12[i]);
  }
  return m;
}

void cfh_sizes(void* h, i64* out) {
  CfMesh* m = (CfMesh*)h;
  out[0] = m->tb.nnods;
  out[1] = m->hp.p.nel;
  out[2] = (i64)m->e2n_idx.size();
  out[3] = (i64)m->g_idx.size();
  out[4] = (i64)m->halo.size() / 2;
  out[5] = (i64)m->nbr_idx.size();
}

void cfh_export(void* h, double* theta, double* r, i64* e2n_off, i64* e2n_idx, i64* g_off, i64* g_idx, i64* halo,
                i64* nbr_off, i64* nbr_idx) {
  CfMesh* m = (CfMesh*)h;
  std::memcpy(theta, m->theta.data(), m->theta.size() * 8);
  std::memcpy(r, m->r.data(), m->r.size() * 8);
  std::memcpy(e2n_off, m->e2n_off.data(), m->e2n_off.size() * 8);
  std::memcpy(e2n_idx, m->e2n_idx.data(), m->e2n_idx.size() * 8);
  std::memcpy(g_off, m->g_off.data(), m->g_off.size() * 8);
  std::memcpy(g_idx, m->g_idx.data(), m->g_idx.size() * 8);
  std::memcpy(halo, m->halo.data(), m->halo.size() * 8);
  std::memcpy(nbr_off, m->nbr_off.data(), m->nbr_off.size() * 8);
  std::memcpy(nbr_idx, m->nbr_idx.data(), m->nbr_idx.size() * 8);
}

void cfh_free(void* h) { delete (CfMesh*)h; }
}
