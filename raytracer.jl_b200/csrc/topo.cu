// topo.cu -- node-to-node topology of the annulus mesh on the device:
//   * nodal_incidence(gr::Grid2D) src/GridAnnulus.jl:763-804 (star-0: every node that shares a cell, self
//     excluded, each neighbour once), nodal_degree src/topology/topology.jl:70-77 and the CSR container
//     SparseAdjencyList{list,deg,idx} of sparse_adjacency_list :94-111;
//   * symrcm(adjgr, degrees) src/SSSP/rcm.jl:2-46 ((reverse) Cuthill-McKee).
//
// The adjacency is never materialised for RCM: neighbours of v are enumerated on the fly as
// "for e in elements containing v, for u in e2n[e]", and a duplicate (u also sits in another cell of v) is
// detected in O(1) through the transpose: u is emitted for cell e only if no cell e' < e of v contains u, i.e. no
// e' in n2e[u] is also in n2e[v].  The reference iterates Julia Sets (irreproducible order); here neighbours come
// in (ascending cell id, list order) and BFS children are ordered by (position of the earliest parent, node id).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "mesh2d.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct TP {
  const i32* __restrict__ e2n_off;
  const i32* __restrict__ e2n_idx;
  const i32* __restrict__ n2e_off;
  const i32* __restrict__ n2e_idx;
  i64 n;
};

// is neighbour u of v (found in cell e) a first occurrence?  (no smaller cell of v contains u)
__device__ __forceinline__ bool first_occurrence(const TP& t, int v, int u, int e) {
  for (int a = t.n2e_off[u]; a < t.n2e_off[u + 1]; ++a) {
    const int eu = t.n2e_idx[a];
    if (eu >= e) continue;
    for (int b = t.n2e_off[v]; b < t.n2e_off[v + 1]; ++b)
      if (t.n2e_idx[b] == eu) return false;
  }
  return true;
}

// warp per node: visit every unique neighbour (cells ascending, list order); f(u) is called by ONE lane per u
template <typename F>
__device__ __forceinline__ void for_each_neighbour(const TP& t, int v, int lane, F f) {
  // cells of v in ascending id: selection over the (<= 4, centre: ntheta) entries of n2e[v]
  const int b0 = t.n2e_off[v], b1 = t.n2e_off[v + 1];
  int last = -1;
  for (int r = b0; r < b1; ++r) {
    int e = 0x7fffffff;
    for (int b = b0; b < b1; ++b) {
      const int c = t.n2e_idx[b];
      if (c > last && c < e) e = c;
    }
    if (e == 0x7fffffff) break;
    last = e;
    for (int q = t.e2n_off[e] + lane; q < t.e2n_off[e + 1]; q += 32) {
      const int u = t.e2n_idx[q];
      if (u != v && first_occurrence(t, v, u, e)) f(u, q - t.e2n_off[e]);
    }
  }
}

__global__ void degree_kernel(TP t, i32* __restrict__ deg) {
  const i64 v = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= t.n) return;
  int cnt = 0;
  for_each_neighbour(t, (int)v, lane, [&](int, int) { ++cnt; });
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  if (lane == 0) deg[v] = cnt;
}

// ordered fill: lanes handle consecutive list positions, so a ballot prefix keeps (cell, list order)
__global__ void adjacency_fill_kernel(TP t, const i64* __restrict__ off, i64* __restrict__ list) {
  const i64 v = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= t.n) return;
  const int b0 = t.n2e_off[v], b1 = t.n2e_off[v + 1];
  i64 o = off[v];
  int last = -1;
  for (int r = b0; r < b1; ++r) {
    int e = 0x7fffffff;
    for (int b = b0; b < b1; ++b) {
      const int c = t.n2e_idx[b];
      if (c > last && c < e) e = c;
    }
    if (e == 0x7fffffff) break;
    last = e;
    const int s = t.e2n_off[e], m = t.e2n_off[e + 1] - s;
    for (int k0 = 0; k0 < m; k0 += 32) {
      const int k = k0 + lane;
      int u = -1;
      bool keep = false;
      if (k < m) {
        u = t.e2n_idx[s + k];
        keep = u != (int)v && first_occurrence(t, (int)v, u, e);
      }
      const unsigned ball = __ballot_sync(FULL, keep);
      if (keep) list[o + __popc(ball & ((1u << lane) - 1u))] = (i64)u + 1;
      o += __popc(ball);
    }
  }
}

// ------------------------------------------------------------------------------------------------- RCM
__global__ void seed_keys_kernel(const i32* __restrict__ deg, i64 n, u64* __restrict__ keys) {
  const i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) keys[v] = ((u64)(unsigned)deg[v] << 32) | (u64)v;
}
// first node in (degree, id) order that is not placed yet (rcm.jl:13-22)
__global__ void next_seed_kernel(const u64* __restrict__ sorted, const i32* __restrict__ level, i64 n, i64 from,
                                 u64* __restrict__ best) {
  const i64 k = from + (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && level[(int)(sorted[k] & 0xffffffffull)] < 0) atomicMin(best, (u64)k);
}
__global__ void place_seed_kernel(const u64* __restrict__ sorted, u64 k, i32* __restrict__ level, i32* __restrict__ order,
                                  i64 pos) {
  const int v = (int)(sorted[k] & 0xffffffffull);
  level[v] = 0;
  order[pos] = v;
}
// expand one BFS level: frontier = order[base, base + nf); children get level L + 1 and the position of their
// earliest parent
__global__ void rcm_expand_kernel(TP t, const i32* __restrict__ order, i64 base, i64 nf, int L, i32* __restrict__ level,
                                  unsigned* __restrict__ ppos, i32* __restrict__ next, u64* __restrict__ n_next) {
  const i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nf) return;
  const int v = order[base + w];
  const unsigned p = (unsigned)(base + w);
  for_each_neighbour(t, v, lane, [&](int u, int) {
    const int lu = atomicCAS(&level[u], -1, L + 1);
    if (lu == -1) next[atomicAdd(n_next, 1ull)] = u;
    if (lu == -1 || lu == L + 1) atomicMin(&ppos[u], p);
  });
}
__global__ void rcm_keys_kernel(const i32* __restrict__ next, i64 nn, const unsigned* __restrict__ ppos,
                                u64* __restrict__ keys) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nn) keys[k] = ((u64)ppos[next[k]] << 32) | (u64)(unsigned)next[k];
}
__global__ void rcm_place_kernel(const u64* __restrict__ keys, i64 nn, i32* __restrict__ order, i64 base) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nn) order[base + k] = (i32)(keys[k] & 0xffffffffull);
}
__global__ void rcm_reverse_kernel(const i32* __restrict__ order, i64 n, i64* __restrict__ perm) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) perm[k] = (i64)order[n - 1 - k] + 1;  // reverse(F), 1-based
}

int sort_keys(u64* keys_in, u64* keys_out, i64 count, cudaStream_t s) {
  size_t bytes = 0;
  RT_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys_in, keys_out, (int)count, 0, 64, s));
  DevBuf<char> tmp;
  RT_TRY(tmp.alloc(bytes));
  RT_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, keys_in, keys_out, (int)count, 0, 64, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

TP make_tp(const Mesh2D& m) {
  TP t;
  t.e2n_off = m.e2n_off.p;
  t.e2n_idx = m.e2n_idx.p;
  t.n2e_off = m.n2e_off.p;
  t.n2e_idx = m.n2e_idx.p;
  t.n = m.n;
  return t;
}

}  // namespace

// deg_out[n] (may be null); list_off[n+1] 0-based offsets (may be null); list_idx: 1-based neighbour ids, capacity cap
// (null -> only degrees / offsets).  idx of the reference's SparseAdjencyList = list_off + 1.
int mesh2d_nodal_adjacency(rt_mesh* h, i64* deg_out, i64* list_off, i64* list_idx, i64 cap) {
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  const i64 n = m.n;
  const TP t = make_tp(m);
  DevBuf<i32> deg;
  RT_TRY(deg.alloc(n));
  degree_kernel<<<grid_for(n * 32, 256), 256, 0, s>>>(t, deg.p);
  RT_CUDA(cudaGetLastError());
  std::vector<i32> hd(n);
  RT_CUDA(cudaMemcpyAsync(hd.data(), deg.p, n * sizeof(i32), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  std::vector<i64> off(n + 1, 0);
  for (i64 v = 0; v < n; ++v) off[v + 1] = off[v] + hd[v];
  if (deg_out)
    for (i64 v = 0; v < n; ++v) deg_out[v] = hd[v];
  if (list_off) std::copy(off.begin(), off.end(), list_off);
  if (!list_idx) return RT_OK;
  RT_ARG(cap >= off[n], "adjacency list capacity too small");
  DevBuf<i64> doff, dlist;
  RT_TRY(doff.upload(off.data(), n + 1, s));
  RT_TRY(dlist.alloc(off[n]));
  adjacency_fill_kernel<<<grid_for(n * 32, 256), 256, 0, s>>>(t, doff.p, dlist.p);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaMemcpyAsync(list_idx, dlist.p, off[n] * sizeof(i64), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

// perm_out[n]: 1-based; position k of the reordered mesh holds old node perm_out[k]  (gr.x .= gr.x[prm])
int mesh2d_rcm(rt_mesh* h, i64* perm_out) {
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  const i64 n = m.n;
  const TP t = make_tp(m);
  DevBuf<i32> deg, level, order, next;
  DevBuf<unsigned> ppos;
  DevBuf<u64> keys, sorted, lkeys, lsorted, scal;
  RT_TRY(deg.alloc(n));
  RT_TRY(level.alloc(n));
  RT_TRY(order.alloc(n));
  RT_TRY(next.alloc(n));
  RT_TRY(ppos.alloc(n));
  RT_TRY(keys.alloc(n));
  RT_TRY(sorted.alloc(n));
  RT_TRY(lkeys.alloc(n));
  RT_TRY(lsorted.alloc(n));
  RT_TRY(scal.alloc(2));
  degree_kernel<<<grid_for(n * 32, 256), 256, 0, s>>>(t, deg.p);
  seed_keys_kernel<<<grid_for(n, 256), 256, 0, s>>>(deg.p, n, keys.p);
  RT_CUDA(cudaGetLastError());
  RT_TRY(sort_keys(keys.p, sorted.p, n, s));  // sortperm(degrees), ties by id
  RT_CUDA(cudaMemsetAsync(level.p, 0xff, n * sizeof(i32), s));
  RT_CUDA(cudaMemsetAsync(ppos.p, 0xff, n * sizeof(unsigned), s));
  i64 placed = 0, seed_from = 0;
  while (placed < n) {
    // next component: first unplaced node in (degree, id) order
    u64 hbest = ~0ull;
    RT_CUDA(cudaMemsetAsync(scal.p, 0xff, sizeof(u64), s));
    next_seed_kernel<<<grid_for(n - seed_from, 256), 256, 0, s>>>(sorted.p, level.p, n, seed_from, scal.p);
    RT_CUDA(cudaMemcpyAsync(&hbest, scal.p, sizeof(u64), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    if (hbest == ~0ull) break;
    seed_from = (i64)hbest + 1;
    place_seed_kernel<<<1, 1, 0, s>>>(sorted.p, hbest, level.p, order.p, placed);
    i64 base = placed, nf = 1;
    placed += 1;
    int L = 0;
    while (nf > 0) {
      RT_CUDA(cudaMemsetAsync(scal.p + 1, 0, sizeof(u64), s));
      rcm_expand_kernel<<<grid_for(nf * 32, 256), 256, 0, s>>>(t, order.p, base, nf, L, level.p, ppos.p, next.p,
                                                               scal.p + 1);
      u64 nn = 0;
      RT_CUDA(cudaMemcpyAsync(&nn, scal.p + 1, sizeof(u64), cudaMemcpyDeviceToHost, s));
      RT_CUDA(cudaStreamSynchronize(s));
      if (nn == 0) break;
      rcm_keys_kernel<<<grid_for((i64)nn, 256), 256, 0, s>>>(next.p, (i64)nn, ppos.p, lkeys.p);
      RT_TRY(sort_keys(lkeys.p, lsorted.p, (i64)nn, s));
      rcm_place_kernel<<<grid_for((i64)nn, 256), 256, 0, s>>>(lsorted.p, (i64)nn, order.p, placed);
      base = placed;
      nf = (i64)nn;
      placed += nf;
      ++L;
    }
  }
  RT_ARG(placed == n, "internal error: RCM did not place every node");
  DevBuf<i64> perm;
  RT_TRY(perm.alloc(n));
  rcm_reverse_kernel<<<grid_for(n, 256), 256, 0, s>>>(order.p, n, perm.p);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaMemcpyAsync(perm_out, perm.p, n * sizeof(i64), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}
