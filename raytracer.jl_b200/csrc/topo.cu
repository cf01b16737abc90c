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

// =========================================================================================================
// Alternative solvers of the reference behind the Dijkstra / RadiusStepping result structs:
//   dijkstra(G::Dict, source, gr, U)          src/SSSP/dijkstra.jl:68-136  (_relax_dijkstra! :138-162)
//   radius_stepping(Gsp, source, gr, U)       src/SSSP/radius_stepping.jl:7-46 (relaxation! :58-71, update! :48-56)
// Both run on the star-0 node graph nodal_incidence(gr) (no halo coupling) with the weight
// 2 * distance(xi, zi, xj, zj) / abs(U[j] + U[i]) and settle nodes in order of travel time; their travel times are the
// least fixed point of that graph, which a label-correcting frontier relaxation reaches bit for bit (fp `+` is monotone),
// so the device never serialises on a priority queue.  Their predecessors follow from the settle order: a node keeps
// the FIRST settled neighbour that gave it its final value (strict `<`), i.e. the tight predecessor with the smallest
// (travel time, id); nodes whose only tight predecessors are coincident duplicates at EQUAL travel time (zero-weight
// edges) take the duplicate that was resolved first.  (Ties in the reference's dijkstra are decided by the iteration order
// of a Julia Set, which is not reproducible: ascending id stands in for it, as in radius_stepping's index loop.)
namespace {

struct SN {
  TP t;
  const double* __restrict__ x;
  const double* __restrict__ z;
  const double* __restrict__ U;
  double* dist;
  i32* prev;
  i32* zl;  // pass in which the predecessor was fixed (0 = regular tight predecessor), -1 = open
  i32* fr0;
  i32* fr1;
  unsigned* inq;
  u64* cnt;  // [0],[1] frontier sizes (ping-pong) [2] open nodes [3] fixed in this pass [4] relaxations
  int source;
};

__device__ __forceinline__ double w_nodal(const SN& p, int a, int b) {
  const double dx = __dsub_rn(p.x[a], p.x[b]), dz = __dsub_rn(p.z[a], p.z[b]);
  const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz)));
  return __ddiv_rn(__dmul_rn(2.0, d), fabs(__dadd_rn(p.U[b], p.U[a])));
}

__global__ void sn_init_kernel(SN p) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.t.n) {
    p.dist[i] = i == p.source ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
    p.prev[i] = -1;
    p.zl[i] = -1;
    p.inq[i] = 0u;
  }
  if (i == 0) {
    p.fr0[0] = p.source;
    p.cnt[0] = 1ull;
    for (int k = 1; k < 8; ++k) p.cnt[k] = 0ull;
  }
}
// one warp per frontier node: push dist + w to every star-0 neighbour
__global__ void sn_relax_kernel(SN p, int cur) {
  const i32* fr = cur ? p.fr1 : p.fr0;
  i32* nx = cur ? p.fr0 : p.fr1;
  const i64 nf = (i64)p.cnt[cur];
  const int lane = threadIdx.x & 31;
  for (i64 q = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nf; q += ((i64)gridDim.x * blockDim.x) >> 5) {
    const int i = fr[q];
    if (lane == 0) {
      atomicExch(&p.inq[i], 0u);  // a later improvement of i queues it again ...
      __threadfence();            // ... and the flag is down before the travel time is read
    }
    __syncwarp();
    const double di = __ldcg(&p.dist[i]);
    u64 nrel = 0;
    for_each_neighbour(p.t, i, lane, [&](int u, int) {
      ++nrel;
      const double delta = __dadd_rn(di, w_nodal(p, i, u));
      const u64 bits = (u64)__double_as_longlong(delta);
      if (bits < (u64)__double_as_longlong(__ldcg(&p.dist[u]))) {
        const u64 old = atomicMin((u64*)&p.dist[u], bits);
        if (bits < old && atomicExch(&p.inq[u], 1u) == 0u) nx[atomicAdd(&p.cnt[cur ^ 1], 1ull)] = u;
      }
    });
    for (int o = 16; o; o >>= 1) nrel += __shfl_xor_sync(FULL, nrel, o);
    if (lane == 0) atomicAdd(&p.cnt[4], nrel);
  }
}
// predecessor, regular case: tight neighbour with a strictly smaller travel time, smallest (time, id)
__global__ void sn_prev_kernel(SN p, i32* __restrict__ open_list) {
  const i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= p.t.n) return;
  const double di = p.dist[i];
  if (!(di < __longlong_as_double(0x7ff0000000000000LL)) || (int)i == p.source) return;  // warp-uniform
  u64 bd = ~0ull;
  int bj = 0x7fffffff;
  for_each_neighbour(p.t, (int)i, lane, [&](int j, int) {
    const double dj = p.dist[j];
    if (!(dj < di)) return;
    if (__dadd_rn(dj, w_nodal(p, j, (int)i)) != di) return;
    const u64 b = (u64)__double_as_longlong(dj);
    if (b < bd || (b == bd && j < bj)) {
      bd = b;
      bj = j;
    }
  });
  for (int o = 16; o; o >>= 1) {
    const u64 od = __shfl_xor_sync(FULL, bd, o);
    const int oj = __shfl_xor_sync(FULL, bj, o);
    if (od < bd || (od == bd && oj < bj)) {
      bd = od;
      bj = oj;
    }
  }
  if (lane == 0) {
    if (bd != ~0ull) {
      p.prev[i] = bj;
      p.zl[i] = 0;
    } else {
      open_list[atomicAdd(&p.cnt[2], 1ull)] = (i32)i;
    }
  }
}
// zero-weight case, pass k: among the equal-time tight neighbours fixed in an earlier pass, the one fixed first, then
// the smallest id; decisions are applied after the kernel so that a pass sees a consistent state
__global__ void sn_zero_kernel(SN p, const i32* __restrict__ open_list, i64 n_open, int pass, i32* __restrict__ pend) {
  const i64 q = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= n_open) return;
  const int i = open_list[q];
  if (p.zl[i] >= 0) return;  // fixed in an earlier pass (warp-uniform)
  const double di = p.dist[i];
  int bl = 0x7fffffff, bj = 0x7fffffff;
  for_each_neighbour(p.t, i, lane, [&](int j, int) {
    if (p.dist[j] != di) return;
    const int l = j == p.source ? 0 : p.zl[j];
    if (l < 0 || l >= pass) return;
    if (__dadd_rn(di, w_nodal(p, j, i)) != di) return;
    if (l < bl || (l == bl && j < bj)) {
      bl = l;
      bj = j;
    }
  });
  for (int o = 16; o; o >>= 1) {
    const int ol = __shfl_xor_sync(FULL, bl, o), oj = __shfl_xor_sync(FULL, bj, o);
    if (ol < bl || (ol == bl && oj < bj)) {
      bl = ol;
      bj = oj;
    }
  }
  if (lane == 0) pend[q] = bl != 0x7fffffff ? bj : -1;
}
__global__ void sn_zero_apply_kernel(SN p, const i32* __restrict__ open_list, i64 n_open, int pass,
                                     const i32* __restrict__ pend) {
  const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_open) return;
  const int i = open_list[q];
  if (p.zl[i] >= 0 || pend[q] < 0) return;
  p.prev[i] = pend[q];
  p.zl[i] = pass;
  atomicAdd(&p.cnt[3], 1ull);
}

}  // namespace

// algorithm: 0 = dijkstra, 1 = radius_stepping (same tables; the caller wraps them in the matching result struct).
int mesh2d_sssp_nodal(rt_mesh* h, const double* U_dev, i64 source1, int algorithm, double* dist_out, i64* prev_out,
                      rt_stats* stats) {
  (void)algorithm;
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  const i64 n = m.n;
  RT_ARG(source1 >= 1 && source1 <= n, "source out of range");
  DevBuf<double> dist;
  DevBuf<i32> prev, zl, fr0, fr1, open_list, pend;
  DevBuf<unsigned> inq;
  DevBuf<u64> cnt;
  RT_TRY(dist.alloc(n));
  RT_TRY(prev.alloc(n));
  RT_TRY(zl.alloc(n));
  RT_TRY(fr0.alloc(n));
  RT_TRY(fr1.alloc(n));
  RT_TRY(open_list.alloc(n));
  RT_TRY(pend.alloc(n));
  RT_TRY(inq.alloc(n));
  RT_TRY(cnt.alloc(8));
  SN p;
  p.t = make_tp(m);
  p.x = m.x.p;
  p.z = m.z.p;
  p.U = U_dev;
  p.dist = dist.p;
  p.prev = prev.p;
  p.zl = zl.p;
  p.fr0 = fr0.p;
  p.fr1 = fr1.p;
  p.inq = inq.p;
  p.cnt = cnt.p;
  p.source = (int)(source1 - 1);
  cudaEvent_t ev0, ev1;
  RT_CUDA(cudaEventCreate(&ev0));
  RT_CUDA(cudaEventCreate(&ev1));
  cudaEventRecord(ev0, s);
  sn_init_kernel<<<grid_for(n, 256), 256, 0, s>>>(p);
  u64 hc[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  int cur = 0;
  i64 rounds = 0, launches = 1;
  int rc = RT_OK;
  while (hc[cur] > 0) {
    cudaMemsetAsync(cnt.p + (cur ^ 1), 0, sizeof(u64), s);
    sn_relax_kernel<<<(unsigned)std::min<i64>(((i64)hc[cur] * 32 + 255) / 256, 148 * 16), 256, 0, s>>>(p, cur);
    cudaMemcpyAsync(hc, cnt.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      rc = RT_ERR_CUDA;
      break;
    }
    cur ^= 1;
    ++rounds;
    ++launches;
    if (rounds > 4 * n + 64) {
      rc = RT_ERR_CUDA;
      break;
    }
  }
  if (rc == RT_OK) {
    sn_prev_kernel<<<grid_for(n * 32, 256), 256, 0, s>>>(p, open_list.p);
    cudaMemcpyAsync(hc, cnt.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = RT_ERR_CUDA;
    ++launches;
    const i64 n_open = (i64)hc[2];
    for (int pass = 1; rc == RT_OK && n_open > 0 && pass < 64; ++pass) {
      cudaMemsetAsync(cnt.p + 3, 0, sizeof(u64), s);
      sn_zero_kernel<<<grid_for(n_open * 32, 256), 256, 0, s>>>(p, open_list.p, n_open, pass, pend.p);
      sn_zero_apply_kernel<<<grid_for(n_open, 256), 256, 0, s>>>(p, open_list.p, n_open, pass, pend.p);
      cudaMemcpyAsync(hc, cnt.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) rc = RT_ERR_CUDA;
      launches += 2;
      if (hc[3] == 0) break;  // nothing left that a fixed neighbour explains
    }
  }
  cudaEventRecord(ev1, s);
  if (rc == RT_OK && dist_out) {
    if (cudaMemcpyAsync(dist_out, dist.p, n * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = RT_ERR_CUDA;
  }
  if (rc == RT_OK && prev_out) rc = prev_to_host_i64(prev.p, n, prev_out, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) rc = RT_ERR_CUDA;
  if (stats) {
    *stats = rt_stats{};
    stats->sweeps = rounds;
    stats->relaxed_edges = (i64)hc[4];
    stats->total_launches = launches;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    stats->kernel_ms = ms;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  if (rc != RT_OK) rt_set_error("CUDA failure in the nodal-graph solver: %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}
