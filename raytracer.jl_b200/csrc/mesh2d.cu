// mesh2d.cu -- adopt / finalize / export the two-level annulus graph on the device.
// Replaces the host-side containers of the reference: Grid2D (src/GridAnnulus.jl:9-21), the
// SparseMatrixCSC G of element_incidence (:420-452), halo::Matrix{Int64} (:943-950) and the adjacency
// containers of src/topology/topology.jl:1-111.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <map>

#include "mesh2d.cuh"

namespace {

// narrowing happens only after the 64-bit value was range checked: an id such as 2^32 + 5 must not wrap into a valid one
__global__ void cvt_i64_i32_kernel(const i64* __restrict__ src, i32* __restrict__ dst, i64 count, i64 bias, i64 lo,
                                   i64 hi, int* bad) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const i64 v = src[i] + bias;
  if (v < lo || v >= hi) {
    *bad = 1;
    dst[i] = 0;
  } else {
    dst[i] = (i32)v;
  }
}
__global__ void cvt_i64_i64_kernel(const i64* __restrict__ src, i64* __restrict__ dst, i64 count, i64 bias) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = src[i] + bias;
}
__global__ void cvt_i32_i64_kernel(const i32* __restrict__ src, i64* __restrict__ dst, i64 count, i64 bias) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = (i64)src[i] + bias;
}

// ---- range checks (the library must never index out of bounds because of a bad caller array) -----------
__global__ void check_range_kernel(const i32* __restrict__ v, i64 count, i32 lo, i32 hi, int* bad) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count && (v[i] < lo || v[i] >= hi)) *bad = 1;
}
__global__ void check_mono32_kernel(const i32* __restrict__ off, i64 count /*entries*/, i32 last, int* bad) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < count && off[i] > off[i + 1]) *bad = 1;
  if (i == 0 && (off[0] != 0 || off[count - 1] != last)) *bad = 1;
}
__global__ void check_mono64_kernel(const i64* __restrict__ off, i64 count, i64 last, int* bad) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < count && off[i] > off[i + 1]) *bad = 1;
  if (i == 0 && (off[0] != 0 || off[count - 1] != last)) *bad = 1;
}

// ---- work items: start flag of node v = (v % 32 == 0) or its G column differs from the one of v-1 -------
__global__ void item_flags_kernel(const i64* __restrict__ g_off, const i32* __restrict__ g_idx, i64 n,
                                  i32* __restrict__ flags) {
  i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  int start = 0;
  if ((v & 31) == 0) {
    start = 1;
  } else {
    i64 a0 = g_off[v - 1], a1 = g_off[v], b1 = g_off[v + 1];
    if (a1 - a0 != b1 - a1) {
      start = 1;
    } else {
      for (i64 k = 0; k < a1 - a0; ++k)
        if (g_idx[a0 + k] != g_idx[a1 + k]) {
          start = 1;
          break;
        }
    }
  }
  flags[v] = start;
}
__global__ void item_scatter_kernel(const i32* __restrict__ flags, const i32* __restrict__ idx, i64 n,
                                    i32* __restrict__ item_first) {
  i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n && flags[v]) item_first[idx[v]] = (i32)v;
}

// ---- n2e: transpose of e2n (one warp per element) ------------------------------------------------------
__global__ void n2e_count_kernel(const i32* __restrict__ e2n_off, const i32* __restrict__ e2n_idx, i64 nel,
                                 i32* __restrict__ cnt) {
  i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= nel) return;
  for (i32 q = e2n_off[w] + lane; q < e2n_off[w + 1]; q += 32) atomicAdd(&cnt[e2n_idx[q]], 1);
}
__global__ void n2e_fill_kernel(const i32* __restrict__ e2n_off, const i32* __restrict__ e2n_idx, i64 nel,
                                const i32* __restrict__ n2e_off, i32* __restrict__ cursor,
                                i32* __restrict__ n2e_idx) {
  i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= nel) return;
  for (i32 q = e2n_off[w] + lane; q < e2n_off[w + 1]; q += 32) {
    i32 nd = e2n_idx[q];
    i32 slot = atomicAdd(&cursor[nd], 1);
    n2e_idx[n2e_off[nd] + slot] = (i32)w;
  }
}

// ---- E_graph = sum_v sum_{el in G[:,v]} |e2n[el]| ---------------------------------------------------------
__global__ void graph_edges_kernel(const i64* __restrict__ g_off, const i32* __restrict__ g_idx,
                                   const i32* __restrict__ e2n_off, i64 n, u64* __restrict__ out) {
  i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 s = 0;
  if (v < n)
    for (i64 c = g_off[v]; c < g_off[v + 1]; ++c) {
      i32 el = g_idx[c];
      s += (u64)(e2n_off[el + 1] - e2n_off[el]);
    }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// ---- interpolate!(V, gr) (src/Interpolations/interpolation.jl:5-18): the reference loops over the elements in
// order and overwrites the secondary nodes of each, so a node shared by two cells keeps the value computed in the
// cell with the LARGER id.  Pass 1 finds that owner per node, pass 2 evaluates bilinear (bilinear.jl:1-17, in
// (theta, r) space, including the wrap-column quirk) / barycentric (barycentric.jl:1-15) weights there.
__global__ void interp_owner_kernel(const i32* __restrict__ e2n_off, const i32* __restrict__ e2n_idx, i64 nel,
                                    const int8_t* __restrict__ el_type, i32* __restrict__ owner) {
  const i64 e = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= nel) return;
  const int nc = el_type[e] == 0 ? 4 : 3;
  for (i32 q = e2n_off[e] + nc + lane; q < e2n_off[e + 1]; q += 32) atomicMax(&owner[e2n_idx[q]], (i32)e);
}
__global__ void interp_cells_kernel(const i32* __restrict__ e2n_off, const i32* __restrict__ e2n_idx, i64 nel,
                                    const int8_t* __restrict__ el_type, const i32* __restrict__ owner,
                                    const double* __restrict__ theta, const double* __restrict__ r,
                                    const double* __restrict__ V0, double* __restrict__ V) {
  const i64 e = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= nel) return;
  const i32* el = e2n_idx + e2n_off[e];
  const int len = e2n_off[e + 1] - e2n_off[e];
  const double PI = 3.141592653589793;
  if (el_type[e] == 0) {
    if (len < 4) return;
    const double z1 = r[el[0]], z2 = r[el[3]];
    double x1 = theta[el[0]];
    const double x2 = theta[el[1]];
    if (x2 - x1 > PI) x1 += 2 * PI;
    const double vu1 = V0[el[0]], vu2 = V0[el[1]], vu3 = V0[el[2]], vu4 = V0[el[3]];
    const double dx21 = x2 - x1, dz21 = z2 - z1;
    for (int q = 4 + lane; q < len; q += 32) {
      const int nd = el[q];
      if (owner[nd] != (i32)e) continue;
      const double dx2 = x2 - theta[nd], dx1 = theta[nd] - x1, dz2 = z2 - r[nd], dz1 = r[nd] - z1;
      V[nd] = 1 / (dx21 * dz21) * (vu1 * dx2 * dz2 + vu2 * dx1 * dz2 + vu4 * dx2 * dz1 + vu3 * dx1 * dz1);
    }
  } else {
    if (len < 3) return;
    const double x1 = theta[el[0]], x2 = theta[el[1]], x3 = theta[el[2]];
    const double z1 = r[el[0]], z2 = r[el[1]], z3 = r[el[2]];
    const double vu1 = V0[el[0]], vu2 = V0[el[1]], vu3 = V0[el[2]];
    const double den = (z2 - z3) * (x1 - x3) + (x3 - x2) * (z1 - z3);
    for (int q = 3 + lane; q < len; q += 32) {
      const int nd = el[q];
      if (owner[nd] != (i32)e) continue;
      const double x = theta[nd], z = r[nd];
      const double N1 = ((z2 - z3) * (x - x3) + (x3 - x2) * (z - z3)) / den;
      const double N2 = ((z3 - z1) * (x - x3) + (x1 - x3) * (z - z3)) / den;
      const double N3 = 1 - N1 - N2;
      V[nd] = N1 * vu1 + N2 * vu2 + N3 * vu3;
    }
  }
}

template <typename T>
int scan_exclusive(const T* in, T* out, i64 count, cudaStream_t s) {
  size_t bytes = 0;
  RT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)count, s));
  DevBuf<char> tmp;
  RT_TRY(tmp.alloc(bytes));
  RT_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)count, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

}  // namespace

int mesh2d_finalize(rt_mesh* h, const i64* halo_host) {
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  const i64 n = m.n, nel = m.nel;
  RT_ARG(n > 0 && n < (i64)2000000000 && nel > 0 && nel < (i64)2000000000, "mesh too large for int32 ids");
  RT_ARG(m.sum_e2n < (i64)2147483000, "sum|e2n| exceeds int32");

  // validate caller arrays once
  {
    DevBuf<int> bad;
    RT_TRY(bad.alloc(1));
    RT_TRY(bad.zero(s));
    check_mono32_kernel<<<grid_for(nel + 1, 256), 256, 0, s>>>(m.e2n_off.p, nel + 1, (i32)m.sum_e2n, bad.p);
    check_mono64_kernel<<<grid_for(n + 1, 256), 256, 0, s>>>(m.g_off.p, n + 1, m.nnzG, bad.p);
    if (m.sum_e2n) check_range_kernel<<<grid_for(m.sum_e2n, 256), 256, 0, s>>>(m.e2n_idx.p, m.sum_e2n, 0, (i32)n, bad.p);
    if (m.nnzG) check_range_kernel<<<grid_for(m.nnzG, 256), 256, 0, s>>>(m.g_idx.p, m.nnzG, 0, (i32)nel, bad.p);
    int hb = 0;
    RT_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    RT_ARG(hb == 0, "graph arrays are inconsistent (offsets not monotone or ids out of range)");
  }

  // ---- work items
  {
    DevBuf<i32> flags, idx;
    RT_TRY(flags.alloc(n + 1));
    RT_TRY(idx.alloc(n + 1));
    RT_TRY(flags.zero(s));
    item_flags_kernel<<<grid_for(n, 256), 256, 0, s>>>(m.g_off.p, m.g_idx.p, n, flags.p);
    RT_TRY(scan_exclusive<i32>(flags.p, idx.p, n + 1, s));
    i32 ni = 0;
    RT_CUDA(cudaMemcpyAsync(&ni, idx.p + n, sizeof(i32), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    m.n_items = ni;
    RT_TRY(m.item_first.alloc(ni + 1));
    item_scatter_kernel<<<grid_for(n, 256), 256, 0, s>>>(flags.p, idx.p, n, m.item_first.p);
    i32 n32 = (i32)n;
    RT_CUDA(cudaMemcpyAsync(m.item_first.p + ni, &n32, sizeof(i32), cudaMemcpyHostToDevice, s));
    RT_CUDA(cudaStreamSynchronize(s));
  }

  // ---- n2e
  {
    DevBuf<i32> cnt;
    RT_TRY(cnt.alloc(n + 1));
    RT_TRY(cnt.zero(s));
    n2e_count_kernel<<<grid_for(nel * 32, 256), 256, 0, s>>>(m.e2n_off.p, m.e2n_idx.p, nel, cnt.p);
    RT_TRY(m.n2e_off.alloc(n + 1));
    RT_TRY(scan_exclusive<i32>(cnt.p, m.n2e_off.p, n + 1, s));
    RT_TRY(cnt.zero(s));
    RT_TRY(m.n2e_idx.alloc(m.sum_e2n));
    n2e_fill_kernel<<<grid_for(nel * 32, 256), 256, 0, s>>>(m.e2n_off.p, m.e2n_idx.p, nel, m.n2e_off.p, cnt.p,
                                                           m.n2e_idx.p);
    RT_CUDA(cudaStreamSynchronize(s));
  }

  // ---- E_graph
  {
    DevBuf<u64> acc;
    RT_TRY(acc.alloc(1));
    RT_TRY(acc.zero(s));
    graph_edges_kernel<<<grid_for(n, 256), 256, 0, s>>>(m.g_off.p, m.g_idx.p, m.e2n_off.p, n, acc.p);
    u64 v = 0;
    RT_CUDA(cudaMemcpyAsync(&v, acc.p, sizeof(u64), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    m.graph_edges = (i64)v;
  }

  // ---- halo tables (host side: 2H rows, tiny)
  const i64 R2 = m.halo_rows;
  m.H = R2 / 2;
  m.halo_structured = (R2 % 2 == 0);
  std::vector<i32> h1(R2), h2(R2);
  for (i64 k = 0; k < R2; ++k) {
    i64 a = halo_host[k], b = halo_host[k + R2];
    RT_ARG(a >= 1 && a <= n && b >= 1 && b <= n, "halo id out of range");
    h1[k] = (i32)(a - 1);
    h2[k] = (i32)(b - 1);
  }
  if (m.halo_structured) {
    // expected shape (src/GridAnnulus.jl:943-950): rows [0,H) = (orig_k, twin_k), rows [H,2H) = (twin_k, orig_k),
    // every twin distinct and never an orig.
    const i64 H = m.H;
    std::vector<uint8_t> is_orig(n, 0), is_twin(n, 0);
    for (i64 k = 0; k < H && m.halo_structured; ++k) {
      if (h1[k + H] != h2[k] || h2[k + H] != h1[k]) m.halo_structured = false;
      is_orig[h1[k]] = 1;
    }
    for (i64 k = 0; k < H && m.halo_structured; ++k) {
      if (is_twin[h2[k]] || is_orig[h2[k]]) m.halo_structured = false;
      is_twin[h2[k]] = 1;
    }
  }
  RT_TRY(m.halo_h1.upload(h1.data(), R2, s));
  RT_TRY(m.halo_h2.upload(h2.data(), R2, s));
  if (m.halo_structured && m.H > 0) {
    // second-half rows grouped by orig, ascending row order inside a group == serial order of update_halo!
    const i64 H = m.H;
    std::vector<i32> order(H);
    for (i64 k = 0; k < H; ++k) order[k] = (i32)k;
    std::stable_sort(order.begin(), order.end(), [&](i32 a, i32 b) { return h1[a] < h1[b]; });
    std::vector<i32> uo, off, tw(H);
    for (i64 q = 0; q < H; ++q) {
      i32 k = order[q];
      if (uo.empty() || uo.back() != h1[k]) {
        uo.push_back(h1[k]);
        off.push_back((i32)q);
      }
      tw[q] = h2[k];
    }
    off.push_back((i32)H);
    m.n_h2_orig = (i64)uo.size();
    RT_TRY(m.h2_orig.upload(uo.data(), uo.size(), s));
    RT_TRY(m.h2_off.upload(off.data(), off.size(), s));
    RT_TRY(m.h2_twin.upload(tw.data(), tw.size(), s));
  }
  {
    // zero-weight coupling partners of every halo node (row (a, b) lets a improve b), sorted by node, for the
    // near-far schedule
    std::vector<std::pair<i32, i32>> pr;
    pr.reserve(R2);
    for (i64 k = 0; k < R2; ++k) pr.push_back({h1[k], h2[k]});
    std::sort(pr.begin(), pr.end());
    pr.erase(std::unique(pr.begin(), pr.end()), pr.end());
    std::vector<i32> nd, off, part;
    for (size_t q = 0; q < pr.size(); ++q) {
      if (nd.empty() || nd.back() != pr[q].first) {
        nd.push_back(pr[q].first);
        off.push_back((i32)q);
      }
      part.push_back(pr[q].second);
    }
    off.push_back((i32)pr.size());
    m.n_hn = (i64)nd.size();
    RT_TRY(m.hn_node.upload(nd.data(), nd.size(), s));
    RT_TRY(m.hn_off.upload(off.data(), off.size(), s));
    RT_TRY(m.hn_part.upload(part.data(), part.size(), s));
  }
  {
    // init_halo_path! (src/SSSP/bfm.jl:64-70) in serial row order; last writer wins
    std::map<i32, i32> init;
    for (i64 k = 0; k < R2; ++k) {
      init[h2[k]] = h1[k];
      init[h1[k]] = h2[k];
    }
    std::vector<i32> nd, vl;
    for (auto& kv : init) {
      nd.push_back(kv.first);
      vl.push_back(kv.second);
    }
    m.n_hinit = (i64)nd.size();
    RT_TRY(m.hinit_node.upload(nd.data(), nd.size(), s));
    RT_TRY(m.hinit_val.upload(vl.data(), vl.size(), s));
  }
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

int mesh2d_from_host(rt_mesh* h, i64 n, i64 nel, const i64* e2n_off, const i64* e2n_idx, const i64* colptr,
                     const i64* rowval, const i64* halo, i64 halo_rows, const double* x, const double* z,
                     const double* theta, const double* r) {
  RT_ARG(n > 0 && nel > 0 && e2n_off && e2n_idx && colptr && rowval && x && z, "null graph array");
  RT_ARG(halo_rows >= 0 && (halo_rows == 0 || halo), "null halo");
  RT_ARG(n < (i64)2000000000 && nel < (i64)2000000000, "mesh too large");
  Mesh2D* mp = new Mesh2D();
  h->m2 = mp;
  h->kind = 2;
  Mesh2D& m = *mp;
  cudaStream_t s = h->stream;
  m.n = n;
  m.nel = nel;
  m.sum_e2n = e2n_off[nel];
  m.nnzG = colptr[n] - 1;
  m.halo_rows = halo_rows;
  RT_ARG(e2n_off[0] == 0 && m.sum_e2n >= 0 && colptr[0] == 1 && m.nnzG >= 0, "bad offsets");
  RT_ARG(m.sum_e2n < (i64)2147483000, "sum|e2n| exceeds int32");
  RT_TRY(m.x.upload(x, n, s));
  RT_TRY(m.z.upload(z, n, s));
  m.has_polar = theta && r;
  if (m.has_polar) {
    RT_TRY(m.theta.upload(theta, n, s));
    RT_TRY(m.r.upload(r, n, s));
  }
  {
    DevBuf<i64> tmp;
    DevBuf<int> bad;
    RT_TRY(bad.alloc(1));
    RT_TRY(bad.zero(s));
    RT_TRY(tmp.upload(e2n_off, nel + 1, s));
    RT_TRY(m.e2n_off.alloc(nel + 1));
    cvt_i64_i32_kernel<<<grid_for(nel + 1, 256), 256, 0, s>>>(tmp.p, m.e2n_off.p, nel + 1, 0, 0, m.sum_e2n + 1, bad.p);
    RT_TRY(tmp.upload(e2n_idx, m.sum_e2n, s));
    RT_TRY(m.e2n_idx.alloc(m.sum_e2n));
    if (m.sum_e2n)
      cvt_i64_i32_kernel<<<grid_for(m.sum_e2n, 256), 256, 0, s>>>(tmp.p, m.e2n_idx.p, m.sum_e2n, -1, 0, n, bad.p);
    RT_TRY(tmp.upload(colptr, n + 1, s));
    RT_TRY(m.g_off.alloc(n + 1));
    cvt_i64_i64_kernel<<<grid_for(n + 1, 256), 256, 0, s>>>(tmp.p, m.g_off.p, n + 1, -1);
    RT_TRY(tmp.upload(rowval, m.nnzG, s));
    RT_TRY(m.g_idx.alloc(m.nnzG));
    if (m.nnzG) cvt_i64_i32_kernel<<<grid_for(m.nnzG, 256), 256, 0, s>>>(tmp.p, m.g_idx.p, m.nnzG, -1, 0, nel, bad.p);
    int hb = 0;
    RT_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    RT_ARG(hb == 0, "graph arrays hold ids outside 1..n / 1..nel (checked as 64-bit values before narrowing)");
  }
  return mesh2d_finalize(h, halo);
}

int mesh2d_sizes(const rt_mesh* h, i64 sizes[8]) {
  const Mesh2D& m = *h->m2;
  sizes[0] = m.n;
  sizes[1] = m.nel;
  sizes[2] = m.sum_e2n;
  sizes[3] = m.nnzG;
  sizes[4] = m.halo_rows;
  sizes[5] = m.sum_nbr;
  sizes[6] = m.ntheta;
  sizes[7] = m.nr;
  return RT_OK;
}

int mesh2d_export(const rt_mesh* h, double* x, double* z, double* theta, double* r, i64* e2n_off, i64* e2n_idx,
                  i64* colptr, i64* rowval, i64* halo, i64* nbr_off, i64* nbr_idx, int8_t* el_type) {
  const Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  const i64 n = m.n, nel = m.nel;
  if (x) RT_CUDA(cudaMemcpyAsync(x, m.x.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (z) RT_CUDA(cudaMemcpyAsync(z, m.z.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (theta || r) RT_ARG(m.has_polar, "mesh has no (theta, r) arrays");
  if (theta) RT_CUDA(cudaMemcpyAsync(theta, m.theta.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (r) RT_CUDA(cudaMemcpyAsync(r, m.r.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  DevBuf<i64> tmp;
  auto out32 = [&](const i32* src, i64 count, i64 bias, i64* dst) -> int {
    if (!dst || count == 0) return RT_OK;
    RT_TRY(tmp.alloc(count));
    cvt_i32_i64_kernel<<<grid_for(count, 256), 256, 0, s>>>(src, tmp.p, count, bias);
    RT_CUDA(cudaMemcpyAsync(dst, tmp.p, count * sizeof(i64), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    return RT_OK;
  };
  RT_TRY(out32(m.e2n_off.p, nel + 1, 0, e2n_off));
  RT_TRY(out32(m.e2n_idx.p, m.sum_e2n, 1, e2n_idx));
  RT_TRY(out32(m.g_idx.p, m.nnzG, 1, rowval));
  if (colptr) {
    RT_TRY(tmp.alloc(n + 1));
    cvt_i64_i64_kernel<<<grid_for(n + 1, 256), 256, 0, s>>>(m.g_off.p, tmp.p, n + 1, 1);
    RT_CUDA(cudaMemcpyAsync(colptr, tmp.p, (n + 1) * sizeof(i64), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
  }
  if (halo && m.halo_rows) {
    const i64 R2 = m.halo_rows;
    std::vector<i32> h1(R2), h2(R2);
    RT_CUDA(cudaMemcpyAsync(h1.data(), m.halo_h1.p, R2 * sizeof(i32), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaMemcpyAsync(h2.data(), m.halo_h2.p, R2 * sizeof(i32), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    for (i64 k = 0; k < R2; ++k) {
      halo[k] = (i64)h1[k] + 1;
      halo[k + R2] = (i64)h2[k] + 1;
    }
  }
  if (nbr_off || nbr_idx || el_type) {
    RT_ARG((i64)m.nbr_off_h.size() == nel + 1, "mesh was adopted from arrays: no neighbours / element types");
    if (nbr_off) std::copy(m.nbr_off_h.begin(), m.nbr_off_h.end(), nbr_off);
    if (nbr_idx) std::copy(m.nbr_idx_h.begin(), m.nbr_idx_h.end(), nbr_idx);
    if (el_type) std::copy(m.el_type_h.begin(), m.el_type_h.end(), el_type);
  }
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

void mesh2d_free(rt_mesh* h) {
  if (h->m2) {
    if (h->m2->counters_host) cudaFreeHost(h->m2->counters_host);
    if (h->m2->canon) canon_ws_free(h->m2->canon);
    if (h->m2->round_graph) cudaGraphExecDestroy((cudaGraphExec_t)h->m2->round_graph);
    delete h->m2;
  }
  h->m2 = nullptr;
}

int mesh2d_closest(const rt_mesh* h, const double* pa, const double* pb, i64 npts, int system, i64* out) {
  const Mesh2D& m = *h->m2;
  if (system == 1) RT_ARG(m.has_polar, "mesh has no (theta, r) arrays");
  const double* a = system == 1 ? m.theta.p : m.x.p;
  const double* b = system == 1 ? m.r.p : m.z.p;
  return closest_point_device(a, b, m.n, pa, pb, npts, out, h->stream);
}

int mesh2d_coords(const rt_mesh* h, const double** x, const double** z, const double** theta, const double** r) {
  const Mesh2D& m = *h->m2;
  *x = m.x.p;
  *z = m.z.p;
  *theta = m.has_polar ? m.theta.p : nullptr;
  *r = m.has_polar ? m.r.p : nullptr;
  return RT_OK;
}

// interpolate!(V, gr): V_dev in/out (device, n doubles).  el_type: host array (0 = :Quad, 1 = :Tri) or null for a
// mesh built by rt_annulus_build.
int mesh2d_interpolate_cells(rt_mesh* h, const int8_t* el_type_host, double* V_dev) {
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  RT_ARG(m.has_polar, "mesh has no (theta, r) arrays");
  const int8_t* et = el_type_host;
  if (!et) {
    RT_ARG((i64)m.el_type_h.size() == m.nel, "element types are needed for a mesh adopted from arrays");
    et = m.el_type_h.data();
  }
  DevBuf<int8_t> d_et;
  DevBuf<i32> owner;
  DevBuf<double> V0;
  RT_TRY(d_et.upload(et, m.nel, s));
  RT_TRY(owner.alloc(m.n));
  RT_CUDA(cudaMemsetAsync(owner.p, 0xff, m.n * sizeof(i32), s));
  RT_TRY(V0.alloc(m.n));
  RT_CUDA(cudaMemcpyAsync(V0.p, V_dev, m.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  interp_owner_kernel<<<grid_for(m.nel * 32, 256), 256, 0, s>>>(m.e2n_off.p, m.e2n_idx.p, m.nel, d_et.p, owner.p);
  interp_cells_kernel<<<grid_for(m.nel * 32, 256), 256, 0, s>>>(m.e2n_off.p, m.e2n_idx.p, m.nel, d_et.p, owner.p,
                                                               m.theta.p, m.r.p, V0.p, V_dev);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}
