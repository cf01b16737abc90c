// annulus_build.cu -- init_annulus(ntheta, nr; spacing) (src/GridAnnulus.jl:57-70) as CUDA kernels.
//
// The reference runs six serial Dict/Set passes (primary_grid, edge_connectivity, secondary_nodes,
// constrain2layers!, discontinuous_boundaries, element_incidence) and cannot build the large configs at all
// (its secondary-node scratch is nedges*1776 doubles, :618).  Here the mesh is generated entity by entity from
// the closed-form description in annulus_cf.cuh: one thread per edge / node / element / halo row, three prefix
// sums (points per edge, twins per below-quad, list lengths), no hash containers, no host pass over O(n) data.
// The host only prepares the O(nr) parameter tables (ring radii, row layers) and, for rt_mesh_export, the
// O(nel) neighbour lists.
#include <cub/device/device_scan.cuh>

#include "annulus_cf.cuh"
#include "mesh2d.cuh"

namespace {

using cf::Params;
using cf::Tables;

__global__ void edge_np_kernel(Params p, i64* __restrict__ cnt) {
  const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (g <= p.nE) cnt[g - 1] = cf::edge_npoints(p, g);
}

__global__ void twin_cnt_kernel(Params p, const i64* __restrict__ eoff, const i64* __restrict__ kd,
                                i64* __restrict__ cnt) {
  const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= 7 * p.T) return;
  const i64 e = cf::quad_id(p, kd[b % 7], b / 7 + 1);
  const i64 g = cf::edge_top(e);
  cnt[b] = 2 + (eoff[g] - eoff[g - 1]);
}

// (theta, r) of primary and secondary nodes
__global__ void node_polar_kernel(Params p, Tables tb, double* __restrict__ theta, double* __restrict__ r) {
  const i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (v > tb.nnods1) return;
  if (v < p.nnods0) {
    const i64 k = (v - 1) % p.M + 1, c = (v - 1) / p.M + 1;
    theta[v - 1] = p.dth * (double)(c - 1);
    r[v - 1] = p.rc[k];
  } else if (v == p.nnods0) {
    theta[v - 1] = 0.0;
    r[v - 1] = 0.0;
  } else {
    const i64 s = v - p.nnods0 - 1;
    const i64 g = cf::find_segment(tb.eoff, p.nE, s) + 1;
    const i64 np = tb.eoff[g] - tb.eoff[g - 1];
    double th, rr;
    cf::secondary_coord(p, g, np, s - tb.eoff[g - 1] + 1, th, rr);
    theta[v - 1] = th;
    r[v - 1] = rr;
  }
}

// twins (discontinuous_boundaries :936-950): coordinates and both halves of the halo matrix
__global__ void twin_kernel(Params p, Tables tb, i64 H, double* __restrict__ theta, double* __restrict__ r,
                            i64* __restrict__ halo /* (2H x 2) column-major, 1-based */) {
  const i64 h = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const i64 b = cf::find_segment(tb.twin_off, 7 * p.T, h);
  i64 o, t;
  cf::halo_pair(p, tb, b, h - tb.twin_off[b], o, t);
  theta[t - 1] = theta[o - 1];
  r[t - 1] = r[o - 1] - 0.05;
  halo[h] = o;
  halo[h + 2 * H] = t;
  halo[h + H] = t;
  halo[h + H + 2 * H] = o;
}

// x = r sin(theta), z = r cos(theta)  (@cartesian :27-29, polar2cartesian :55)
__global__ void cartesian_kernel(const double* __restrict__ theta, const double* __restrict__ r, i64 n,
                                 double* __restrict__ x, double* __restrict__ z) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = theta[i], rr = r[i];
  x[i] = __dmul_rn(rr, sin(t));
  z[i] = __dmul_rn(rr, cos(t));
}

__global__ void elem_len_kernel(Params p, Tables tb, i64* __restrict__ len) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (e <= p.nel) len[e - 1] = cf::elem_list_len(p, tb, e);
}
__global__ void elem_fill_kernel(Params p, Tables tb, const i64* __restrict__ off, i32* __restrict__ e2n_off,
                                 i32* __restrict__ e2n_idx) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (e > p.nel + 1) return;
  e2n_off[e - 1] = (i32)off[e - 1];
  if (e <= p.nel) cf::elem_list_fill<i32>(p, tb, e, e2n_idx + off[e - 1], -1);
}

__global__ void g_len_kernel(Params p, Tables tb, i64* __restrict__ len) {
  const i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (v > tb.nnods) return;
  if (v == p.nnods0) {
    len[v - 1] = cf::g_column_centre_len(p);
    return;
  }
  i64 set[64];
  len[v - 1] = cf::g_column(p, tb, v, set);
}
__global__ void g_fill_kernel(Params p, Tables tb, const i64* __restrict__ g_off, i32* __restrict__ g_idx) {
  const i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (v > tb.nnods) return;
  i32* out = g_idx + g_off[v - 1];
  if (v == p.nnods0) {
    const i64 len = cf::g_column_centre_len(p);
    for (i64 i = 0; i < len; ++i) out[i] = (i32)(cf::g_column_centre_entry(p, i) - 1);
    return;
  }
  i64 set[64];
  const int len = cf::g_column(p, tb, v, set);
  for (int i = 0; i < len; ++i) out[i] = (i32)(set[i] - 1);
}

int scan64(i64* in_out_plus1, i64 count, cudaStream_t s) {
  // exclusive prefix sum in place over count + 1 entries (the last input entry must be 0)
  size_t bytes = 0;
  RT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in_out_plus1, in_out_plus1, (int)(count + 1), s));
  DevBuf<char> tmp;
  RT_TRY(tmp.alloc(bytes));
  RT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in_out_plus1, in_out_plus1, (int)(count + 1), s));
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

int last_of(const i64* dev, i64 index, i64* out, cudaStream_t s) {
  RT_CUDA(cudaMemcpyAsync(out, dev + index, sizeof(i64), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

}  // namespace

int annulus_build_device(rt_mesh* h, i64 ntheta, i64 nr, double spacing) {
  // correct_theta's wrap test needs dtheta < 1 - 1/ntheta (ntheta >= 7); smaller rings alias their columns
  RT_ARG(ntheta >= 8, "rt_annulus_build supports ntheta >= 8");
  RT_ARG(nr >= 2 && spacing > 0.0, "rt_annulus_build needs nr >= 2 and spacing > 0");
  cudaStream_t s = h->stream;
  cf::HostParams hp;
  cf::make_params(ntheta, nr, spacing, hp);
  RT_ARG(hp.kd.size() == 7, "ring radii do not contain the seven discontinuities exactly once");
  Params p = hp.p;
  RT_ARG(p.nE < (i64)2000000000, "mesh too large");

  DevBuf<double> d_rc;
  DevBuf<int> d_lay, d_disc;
  DevBuf<i64> d_kd;
  RT_TRY(d_rc.upload(hp.rc.data(), hp.rc.size(), s));
  RT_TRY(d_lay.upload(hp.lay.data(), hp.lay.size(), s));
  RT_TRY(d_disc.upload(hp.disc.data(), hp.disc.size(), s));
  RT_TRY(d_kd.upload(hp.kd.data(), hp.kd.size(), s));
  p.rc = d_rc.p;
  p.lay = d_lay.p;
  p.disc = d_disc.p;

  // ---- points per edge -> eoff
  DevBuf<i64> eoff, twin_off;
  RT_TRY(eoff.alloc(p.nE + 1));
  RT_TRY(eoff.zero(s));
  edge_np_kernel<<<grid_for(p.nE, 256), 256, 0, s>>>(p, eoff.p);
  RT_TRY(scan64(eoff.p, p.nE, s));
  i64 nsec = 0;
  RT_TRY(last_of(eoff.p, p.nE, &nsec, s));
  // ---- twins per below-quad -> twin_off
  const i64 nb = 7 * p.T;
  RT_TRY(twin_off.alloc(nb + 1));
  RT_TRY(twin_off.zero(s));
  twin_cnt_kernel<<<grid_for(nb, 256), 256, 0, s>>>(p, eoff.p, d_kd.p, twin_off.p);
  RT_TRY(scan64(twin_off.p, nb, s));
  i64 H = 0;
  RT_TRY(last_of(twin_off.p, nb, &H, s));

  Tables tb;
  tb.eoff = eoff.p;
  tb.twin_off = twin_off.p;
  tb.kd = d_kd.p;
  tb.nnods1 = p.nnods0 + nsec;
  tb.nnods = tb.nnods1 + H;
  const i64 n = tb.nnods;
  RT_ARG(n < (i64)2000000000, "mesh too large for int32 node ids");

  Mesh2D* mp = new Mesh2D();
  h->m2 = mp;
  h->kind = 2;
  Mesh2D& m = *mp;
  m.n = n;
  m.nel = p.nel;
  m.ntheta = p.T;
  m.nr = p.M;
  m.halo_rows = 2 * H;
  m.has_polar = true;
  RT_TRY(m.theta.alloc(n));
  RT_TRY(m.r.alloc(n));
  RT_TRY(m.x.alloc(n));
  RT_TRY(m.z.alloc(n));
  node_polar_kernel<<<grid_for(tb.nnods1, 256), 256, 0, s>>>(p, tb, m.theta.p, m.r.p);
  DevBuf<i64> halo_d;
  RT_TRY(halo_d.alloc(4 * H));
  if (H) twin_kernel<<<grid_for(H, 256), 256, 0, s>>>(p, tb, H, m.theta.p, m.r.p, halo_d.p);
  cartesian_kernel<<<grid_for(n, 256), 256, 0, s>>>(m.theta.p, m.r.p, n, m.x.p, m.z.p);
  RT_CUDA(cudaGetLastError());

  // ---- e2n
  {
    DevBuf<i64> off;
    RT_TRY(off.alloc(p.nel + 1));
    RT_TRY(off.zero(s));
    elem_len_kernel<<<grid_for(p.nel, 256), 256, 0, s>>>(p, tb, off.p);
    RT_TRY(scan64(off.p, p.nel, s));
    RT_TRY(last_of(off.p, p.nel, &m.sum_e2n, s));
    RT_ARG(m.sum_e2n < (i64)2147483000, "sum|e2n| exceeds int32");
    RT_TRY(m.e2n_off.alloc(p.nel + 1));
    RT_TRY(m.e2n_idx.alloc(m.sum_e2n));
    elem_fill_kernel<<<grid_for(p.nel + 1, 128), 128, 0, s>>>(p, tb, off.p, m.e2n_off.p, m.e2n_idx.p);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaStreamSynchronize(s));
  }
  // ---- G
  {
    RT_TRY(m.g_off.alloc(n + 1));
    RT_TRY(m.g_off.zero(s));
    g_len_kernel<<<grid_for(n, 128), 128, 0, s>>>(p, tb, m.g_off.p);
    RT_TRY(scan64(m.g_off.p, n, s));
    RT_TRY(last_of(m.g_off.p, n, &m.nnzG, s));
    RT_TRY(m.g_idx.alloc(m.nnzG));
    g_fill_kernel<<<grid_for(n, 128), 128, 0, s>>>(p, tb, m.g_off.p, m.g_idx.p);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaStreamSynchronize(s));
  }
  // ---- O(nel) host tables that only rt_mesh_export hands out (gr.neighbours, gr.element_type)
  {
    const Params& hpp = hp.p;
    m.nbr_off_h.assign(hpp.nel + 1, 0);
    m.el_type_h.assign(hpp.nel, 0);
    i64 nb12[12];
    for (i64 e = 1; e <= hpp.nel; ++e) {
      const int c = cf::element_neighbours(hpp, e, nb12);
      m.nbr_off_h[e] = m.nbr_off_h[e - 1] + c;
      for (int i = 0; i < c; ++i) m.nbr_idx_h.push_back(nb12[i]);
      m.el_type_h[e - 1] = e <= hpp.nq ? 0 : 1;
    }
    m.sum_nbr = (i64)m.nbr_idx_h.size();
  }
  std::vector<i64> halo_h(4 * H);
  if (H) RT_CUDA(cudaMemcpyAsync(halo_h.data(), halo_d.p, 4 * H * sizeof(i64), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return mesh2d_finalize(h, halo_h.data());
}
