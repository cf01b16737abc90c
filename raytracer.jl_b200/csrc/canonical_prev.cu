// canonical_prev.cu -- the reference's predecessor table, exact ties included, from converged travel times.
//
// The reference relaxes in Jacobi sweeps with a strict `>` (src/SSSP/bfm.jl:161-210): prev[i] is the FIRST candidate in
// scan order that gives i its final travel time in the EARLIEST sweep in which that happens, and update_halo! (:54-62)
// copies (dist, prev) across a halo row in the sweep in which the row's first node improved.  A work-efficient schedule
// reaches the same travel times (least fixed point) but visits candidates in another order, so its predecessors differ
// on exact ties -- which are systematic on these meshes (every radial edge exists twice, collinear nodes in
// constant-velocity layers).
//
// Given the converged travel times T this file REPLAYS the reference's sweeps on the only part of the graph that can
// matter for the final predecessors:
//   1. near-tight edges j -> i: fl(T[j] + w_ji) <= T[i] + slack, kept per node in the reference's scan order (first
//      occurrence of a node).  slack = 8 ulp(max T) is absolute: whether a candidate gives i its final value in some
//      sweep depends on j's value in that sweep only up to the rounding of the sum, and a value of j that close to T[j]
//      can only have arrived over near-tight edges itself (fp `+` is monotone), recursively back to the source.
//   2. the reference's sweep structure -- double-buffered relax with strict `>` over the kept candidates, update_halo!
//      in its serial row semantics, frontier = successors of the nodes that improved -- restricted to those edges.
//      Values far above the final ones differ from the reference's (fewer paths), values within the slack are the
//      reference's bit for bit, and so is the predecessor written when a node reaches its final value.
// (A plain breadth-first levelling of the exactly-tight edges -- SURVEY.md A.5 -- is NOT enough: with dozens of exact
// ties per node a final value is regularly first reached through a NON-final value of a neighbour that rounds to the
// same sum; measured on constant-velocity meshes.)
// Work: one tightness scan of the graph (about one sweep) plus sweeps over a graph with a handful of edges per node.
// Checked bit for bit against the predecessors of the reference schedule (tests/test_canonical_prev.py).
#include <cub/device/device_scan.cuh>

#include <cstring>

#include "mesh2d.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int KT = 8;  // near-tight candidates kept in the fixed slots of a node; longer lists go to the overflow CSR
constexpr int MODE_F64 = 0, MODE_DUAL = 1, MODE_F32 = 2, MODE_3D = 3, MODE_3D_F32 = 4;

struct CP {
  // ---- 2-D two-level graph
  const double* __restrict__ x;
  const double* __restrict__ z;
  const double* __restrict__ U1;
  const double* __restrict__ U2;
  const double* __restrict__ r;
  const i32* __restrict__ e2n_off;
  const i32* __restrict__ e2n_idx;
  const i64* __restrict__ g_off;
  const i32* __restrict__ g_idx;
  const i32* __restrict__ item_first;
  // ---- 3-D structured grid (implicit window adjacency, canonical scan order = ascending linear id)
  const double* __restrict__ X3;
  const double* __restrict__ Y3;
  const double* __restrict__ Z3;
  int nx, ny, nz, w3, self3, wmode3;
  // ---- common
  const double* __restrict__ fin;  // converged travel times [n]
  double slack;                    // absolute near-tightness slack
  i64 n, n_items;
  int source;
  i32* nt_fix;     // [n x KT] near-tight candidates in scan order (overflow nodes: slot 0 = row in the overflow CSR)
  i32* nt_cnt;     // [n] list length (<= KT), or -1 = the list lives in the overflow CSR
  i32* ovf_node;   // overflow nodes
  i32* ovf_len;    // [n_ovf + 1] lengths -> offsets after the scan
  i32* ovf_idx;
  i32* succ_off;   // [n + 1]
  i32* succ_cur;   // [n]
  i32* succ_idx;
  double* rd;      // replay travel times (this sweep) ...
  double* rd0;     // ... and of the previous sweep
  i32* prev;       // [n] output (entries of unreached nodes are left alone)
  i32* stamp;      // [n] last sweep in which the node was queued
  i32* act0;
  i32* act1;
  // halo (structured: rows [0,H) orig -> twin, rows [H,2H) twin -> orig, twins grouped by orig in row order)
  const i32* __restrict__ h1;
  const i32* __restrict__ h2;
  i64 H;
  const i32* __restrict__ g_orig;
  const i32* __restrict__ g_toff;
  const i32* __restrict__ g_twin;
  i64 n_groups;
  // control: [0] cur parity [1] sweep [2] done [3] n overflow nodes; cnt[0], cnt[1] = list sizes
  int* ctl;
  unsigned long long* cnt;
};

// travel time through candidate j with value dj, for target i -- the exact expression of the solve's relax mode
template <int MODE>
__device__ __forceinline__ double cand_value(const CP& p, int i, int j, double dj) {
  if (MODE == MODE_3D || MODE == MODE_3D_F32)
    return exact_cand3<MODE == MODE_3D_F32>(dj, p.X3[i], p.Y3[i], p.Z3[i], p.U1[i], p.X3[j], p.Y3[j], p.Z3[j], p.U1[j],
                                            p.wmode3);
  double ut = p.U1[i], us = p.U1[j];
  if (MODE == MODE_DUAL) {  // bfm.jl:113-159: head = candidate, tail = target; head_idx = (r_i > r_Gi) + 1
    const bool down = p.r[i] > p.r[j];
    ut = down ? p.U1[i] : p.U2[i];
    us = down ? p.U2[j] : p.U1[j];
  }
  return exact_cand2<MODE == MODE_F32>(dj, p.x[j], p.z[j], us, p.x[i], p.z[i], ut);
}

// can fl(dj + w) <= di + slack hold?  w = 2 sqrt(d2) / ssum >= (di + slack - dj) is certain when d2 is clearly larger
__device__ __forceinline__ bool maybe_near_tight(double di, double dj, double d2, double ssum, double slack, bool f32) {
  const double t = (di - dj) + slack;
  if (!(ssum > 0.0)) return true;
  const double hi = (t + di * (f32 ? 1.3e-7 : 4e-15)) * ssum * 0.5;
  return !(d2 > hi * hi * (f32 ? 1.0 + 2e-6 : 1.0 + 1e-9));
}

template <int MODE>
__device__ __forceinline__ bool is_near_tight2(const CP& p, double di, double xi, double zi, double u1i, double u2i,
                                               double ri, double dj, double xj, double zj, double u1j, double u2j,
                                               double rj) {
  constexpr bool F32 = MODE == MODE_F32;
  double ut = u1i, us = u1j;
  if (MODE == MODE_DUAL) {
    const bool down = ri > rj;
    ut = down ? u1i : u2i;
    us = down ? u2j : u1j;
  }
  const double dx = __dsub_rn(xi, xj), dz = __dsub_rn(zi, zj);
  const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz));
  if (!maybe_near_tight(di, dj, d2, __dadd_rn(ut, us), p.slack, F32)) return false;
  return exact_cand2<F32>(dj, xj, zj, us, xi, zi, ut) <= di + p.slack;
}

// ---- step 1 (2-D): warp per work item (<= 32 targets sharing one G column); the candidates of the column pass through
// a warp-private shared-memory slab 32 at a time, every lane tests them in scan order against its own target
template <int MODE>
__global__ void __launch_bounds__(128) near_tight2_kernel(CP p) {
  __shared__ double2 s_xz[4][32], s_ud[4][32], s_u2r[MODE == MODE_DUAL ? 4 : 1][32];
  __shared__ int s_id[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  for (i64 it = (i64)blockIdx.x * 4 + warp; it < p.n_items; it += (i64)gridDim.x * 4) {
    const int v0 = p.item_first[it];
    const int t = p.item_first[it + 1] - v0;
    const int i = v0 + min(lane, t - 1);
    const double di = p.fin[i];
    const bool want = lane < t && di < INF && i != p.source;
    if (!__any_sync(FULL, want)) {
      if (lane < t) p.nt_cnt[i] = 0;
      continue;
    }
    const double xi = p.x[i], zi = p.z[i], u1i = p.U1[i];
    const double u2i = MODE == MODE_DUAL ? p.U2[i] : 0.0, ri = MODE == MODE_DUAL ? p.r[i] : 0.0;
    const double lim = di + p.slack;
    int b[KT];
#pragma unroll
    for (int q = 0; q < KT; ++q) b[q] = -1;
    int cnt = 0;
    const i64 c0 = p.g_off[v0], c1 = p.g_off[v0 + 1];
    for (i64 c = c0; c < c1; ++c) {
      const int el = p.g_idx[c];
      const int s = p.e2n_off[el];
      const int m = p.e2n_off[el + 1] - s;
      for (int k0 = 0; k0 < m; k0 += 32) {
        __syncwarp();
        if (k0 + lane < m) {
          const int j = p.e2n_idx[s + k0 + lane];
          s_id[warp][lane] = j;
          s_xz[warp][lane] = make_double2(p.x[j], p.z[j]);
          s_ud[warp][lane] = make_double2(p.U1[j], p.fin[j]);
          if (MODE == MODE_DUAL) s_u2r[warp][lane] = make_double2(p.U2[j], p.r[j]);
        }
        __syncwarp();
        const int mm = min(32, m - k0);
        for (int q = 0; q < mm; ++q) {
          const double2 ud = s_ud[warp][q];
          if (!want || cnt > KT || !(ud.y <= lim)) continue;
          const int j = s_id[warp][q];
          if (j == i) continue;
          const double2 xz = s_xz[warp][q];
          double u2j = 0.0, rj = 0.0;
          if (MODE == MODE_DUAL) {
            u2j = s_u2r[warp][q].x;
            rj = s_u2r[warp][q].y;
          }
          if (!is_near_tight2<MODE>(p, di, xi, zi, u1i, u2i, ri, ud.y, xz.x, xz.y, ud.x, u2j, rj)) continue;
          bool seen = false;
#pragma unroll
          for (int e = 0; e < KT; ++e) seen = seen || b[e] == j;
          if (seen) continue;
          if (cnt < KT) {
#pragma unroll
            for (int e = 0; e < KT; ++e)
              if (e == cnt) b[e] = j;
          }
          ++cnt;  // KT + 1 = overflow: the list is rebuilt by the overflow kernels
        }
      }
    }
    if (lane < t) {
      if (cnt > KT) {
        const int row = atomicAdd(&p.ctl[3], 1);
        p.ovf_node[row] = i;
        p.nt_cnt[i] = -1;
        p.nt_fix[(i64)i * KT] = row;
      } else {
        p.nt_cnt[i] = cnt;
#pragma unroll
        for (int e = 0; e < KT; ++e) p.nt_fix[(i64)i * KT + e] = b[e];
      }
    }
  }
}
// overflow nodes (2-D): one warp per node walks its whole scan list; FILL = 0 counts, 1 writes (scan order, duplicates
// of a node are kept: a later duplicate can never win against the first, strict `>`)
template <int MODE, int FILL>
__global__ void near_tight2_overflow_kernel(CP p) {
  const int lane = threadIdx.x & 31;
  const int no = p.ctl[3];
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < no; w += (gridDim.x * blockDim.x) >> 5) {
    const int i = p.ovf_node[w];
    const double di = p.fin[i], xi = p.x[i], zi = p.z[i], u1i = p.U1[i];
    const double u2i = MODE == MODE_DUAL ? p.U2[i] : 0.0, ri = MODE == MODE_DUAL ? p.r[i] : 0.0;
    const double lim = di + p.slack;
    int total = 0;
    const int base = FILL ? p.ovf_len[w] : 0;
    for (i64 c = p.g_off[i]; c < p.g_off[i + 1]; ++c) {
      const int el = p.g_idx[c];
      const int s = p.e2n_off[el], m = p.e2n_off[el + 1] - s;
      for (int k0 = 0; k0 < m; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        int j = -1;
        if (k < m) {
          j = p.e2n_idx[s + k];
          const double dj = p.fin[j];
          if (j != i && dj <= lim)
            hit = is_near_tight2<MODE>(p, di, xi, zi, u1i, u2i, ri, dj, p.x[j], p.z[j], p.U1[j],
                                       MODE == MODE_DUAL ? p.U2[j] : 0.0, MODE == MODE_DUAL ? p.r[j] : 0.0);
        }
        const unsigned ball = __ballot_sync(FULL, hit);
        if (FILL && hit) p.ovf_idx[base + total + __popc(ball & ((1u << lane) - 1u))] = j;
        total += __popc(ball);
      }
    }
    if (!FILL && lane == 0) p.ovf_len[w] = total;
  }
}

// ---- step 1 (3-D): thread per node over its clipped window in ascending linear id
template <bool F32>
__device__ __forceinline__ bool is_near_tight3(const CP& p, double di, double xi, double yi, double zi, double ui, i64 J) {
  const double dj = p.fin[J];
  if (!(dj <= di + p.slack)) return false;
  const double xj = p.X3[J], yj = p.Y3[J], zj = p.Z3[J], uj = p.U1[J];
  const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj), dz = __dsub_rn(zi, zj);
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  if (!maybe_near_tight(di, dj, d2, screen_ssum3(fabs(__dadd_rn(ui, uj)), p.wmode3), p.slack, F32)) return false;
  return exact_cand3<F32>(dj, xi, yi, zi, ui, xj, yj, zj, uj, p.wmode3) <= di + p.slack;
}
template <bool F32>
__global__ void __launch_bounds__(128) near_tight3_kernel(CP p) {
  const i64 I = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= p.n) return;
  const double di = p.fin[I];
  int cnt = 0;
  int b[KT];
#pragma unroll
  for (int q = 0; q < KT; ++q) b[q] = -1;
  if (di < __longlong_as_double(0x7ff0000000000000LL) && (int)I != p.source) {
    const int i = (int)(I % p.nx), j = (int)((I / p.nx) % p.ny), k = (int)(I / ((i64)p.nx * p.ny));
    const double xi = p.X3[I], yi = p.Y3[I], zi = p.Z3[I], ui = p.U1[I];
    const int w = p.w3;
    for (int zz = max(0, k - w); zz <= min(p.nz - 1, k + w) && cnt <= KT; ++zz)
      for (int yy = max(0, j - w); yy <= min(p.ny - 1, j + w) && cnt <= KT; ++yy)
        for (int xx = max(0, i - w); xx <= min(p.nx - 1, i + w); ++xx) {
          const i64 J = (i64)xx + (i64)p.nx * ((i64)yy + (i64)p.ny * zz);
          if (J == I) continue;
          if (!is_near_tight3<F32>(p, di, xi, yi, zi, ui, J)) continue;
          if (cnt < KT) {
#pragma unroll
            for (int e = 0; e < KT; ++e)
              if (e == cnt) b[e] = (int)J;
          }
          ++cnt;
        }
  }
  if (cnt > KT) {
    const int row = atomicAdd(&p.ctl[3], 1);
    p.ovf_node[row] = (i32)I;
    p.nt_cnt[I] = -1;
    p.nt_fix[I * KT] = row;
  } else {
    p.nt_cnt[I] = cnt;
#pragma unroll
    for (int e = 0; e < KT; ++e) p.nt_fix[I * KT + e] = b[e];
  }
}
template <bool F32, int FILL>
__global__ void near_tight3_overflow_kernel(CP p) {
  const int lane = threadIdx.x & 31;
  const int no = p.ctl[3];
  for (int wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wq < no; wq += (gridDim.x * blockDim.x) >> 5) {
    const i64 I = p.ovf_node[wq];
    const int i = (int)(I % p.nx), j = (int)((I / p.nx) % p.ny), k = (int)(I / ((i64)p.nx * p.ny));
    const double di = p.fin[I], xi = p.X3[I], yi = p.Y3[I], zi = p.Z3[I], ui = p.U1[I];
    const int w = p.w3;
    const int x0 = max(0, i - w), x1 = min(p.nx - 1, i + w), y0 = max(0, j - w), y1 = min(p.ny - 1, j + w);
    const int z0 = max(0, k - w), z1 = min(p.nz - 1, k + w);
    const int cx = x1 - x0 + 1, cy = y1 - y0 + 1, cells = cx * cy * (z1 - z0 + 1);
    int total = 0;
    const int base = FILL ? p.ovf_len[wq] : 0;
    for (int t0 = 0; t0 < cells; t0 += 32) {  // window cells in ascending linear id
      const int t = t0 + lane;
      bool hit = false;
      i64 J = -1;
      if (t < cells) {
        const int xx = x0 + t % cx, yy = y0 + (t / cx) % cy, zz = z0 + t / (cx * cy);
        J = (i64)xx + (i64)p.nx * ((i64)yy + (i64)p.ny * zz);
        hit = J != I && is_near_tight3<F32>(p, di, xi, yi, zi, ui, J);
      }
      const unsigned ball = __ballot_sync(FULL, hit);
      if (FILL && hit) p.ovf_idx[base + total + __popc(ball & ((1u << lane) - 1u))] = (i32)J;
      total += __popc(ball);
    }
    if (!FILL && lane == 0) p.ovf_len[wq] = total;
  }
}

// ---- candidate list of node i
__device__ __forceinline__ void list_of(const CP& p, i64 i, const i32*& lst, int& len) {
  const int c = p.nt_cnt[i];
  if (c >= 0) {
    lst = p.nt_fix + i * KT;
    len = c;
  } else {
    const int row = p.nt_fix[i * KT];
    lst = p.ovf_idx + p.ovf_len[row];
    len = p.ovf_len[row + 1] - p.ovf_len[row];
  }
}

// ---- successors (transpose of the kept edges)
__global__ void succ_count_kernel(CP p) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const i32* lst;
  int len;
  list_of(p, i, lst, len);
  for (int e = 0; e < len; ++e) atomicAdd(&p.succ_off[lst[e]], 1);
}
__global__ void succ_fill_kernel(CP p) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const i32* lst;
  int len;
  list_of(p, i, lst, len);
  for (int e = 0; e < len; ++e) {
    const int j = lst[e];
    p.succ_idx[p.succ_off[j] + atomicAdd(&p.succ_cur[j], 1)] = (i32)i;
  }
}

// ---- step 2: the replay
__global__ void rp_init_kernel(CP p) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.n) {
    const double v = i == p.source ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
    p.rd[i] = v;
    p.rd0[i] = v;
    p.stamp[i] = -1;
  }
}
// init_Q!: the frontier of the first sweep = the successors of the source
__global__ void rp_seed_kernel(CP p) {
  if (blockIdx.x || threadIdx.x) return;
  unsigned long long c = 0;
  for (int e = p.succ_off[p.source]; e < p.succ_off[p.source + 1]; ++e) {
    const int s = p.succ_idx[e];
    if (p.stamp[s] != 0) {
      p.stamp[s] = 0;
      p.act0[c++] = s;
    }
  }
  p.cnt[0] = c;
  p.cnt[1] = 0ull;
  p.ctl[0] = 0;
  p.ctl[1] = 0;
  p.ctl[2] = c == 0ull;
}
__global__ void rp_begin_kernel(CP p) {
  int* c = p.ctl;
  if (c[2]) return;
  if (c[1] > 0) c[0] ^= 1;  // the list filled during the previous sweep becomes the frontier
  if (p.cnt[c[0]] == 0ull) {
    c[2] = 1;
    return;
  }
  c[1] += 1;
  p.cnt[c[0] ^ 1] = 0ull;
}
// _relax! over the kept candidates: double-buffered, strict `>`, scan order
template <int MODE>
__global__ void rp_relax_kernel(CP p) {
  if (p.ctl[2]) return;
  const int cur = p.ctl[0];
  const i32* act = cur ? p.act1 : p.act0;
  const i64 na = (i64)p.cnt[cur];
  for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < na; q += (i64)gridDim.x * blockDim.x) {
    const int i = act[q];
    const i32* lst;
    int len;
    list_of(p, i, lst, len);
    double di = p.rd0[i];
    int bp = -1;
    for (int e = 0; e < len; ++e) {
      const int j = lst[e];
      const double dj = p.rd0[j];
      if (!(dj < di)) continue;  // dj + w >= dj >= di (also dj == Inf): `di > delta` cannot hold
      const double delta = cand_value<MODE>(p, i, j, dj);
      if (di > delta) {
        di = delta;
        bp = j;
      }
    }
    p.rd[i] = di;
    if (bp >= 0) p.prev[i] = bp;
  }
}
// update_halo! first half of the rows (orig_k -> twin_k) ...
__global__ void rp_halo1_kernel(CP p) {
  if (p.ctl[2]) return;
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.H) return;
  const int a = p.h1[k], b = p.h2[k];
  const double da = p.rd[a];
  if (da < p.rd0[a] && p.rd[b] > da) {
    p.rd[b] = da;
    p.prev[b] = p.prev[a];
  }
}
// ... and second half (twin -> orig) per orig in ascending row order
__global__ void rp_halo2_kernel(CP p) {
  if (p.ctl[2]) return;
  const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.n_groups) return;
  const int o = p.g_orig[g];
  double d_o = p.rd[o];
  int p_o = -2;
  for (int q = p.g_toff[g]; q < p.g_toff[g + 1]; ++q) {
    const int b = p.g_twin[q];
    const double db = p.rd[b];
    if (db < p.rd0[b] && d_o > db) {
      d_o = db;
      p_o = p.prev[b];
    }
  }
  if (p_o != -2) {
    p.rd[o] = d_o;
    p.prev[o] = p_o;
  }
}
// update_Q! + copyto!(dist0, dist): an improved node queues its successors for the next sweep
__device__ __forceinline__ void rp_commit_node(const CP& p, int i, int sweep, i32* nx, int cur) {
  const double d = p.rd[i];
  if (!(d < p.rd0[i])) return;
  p.rd0[i] = d;
  for (int e = p.succ_off[i]; e < p.succ_off[i + 1]; ++e) {
    const int s = p.succ_idx[e];
    if (p.stamp[s] != sweep && atomicExch(&p.stamp[s], sweep) != sweep) nx[atomicAdd(&p.cnt[cur ^ 1], 1ull)] = s;
  }
}
__global__ void rp_commit_kernel(CP p) {
  if (p.ctl[2]) return;
  const int cur = p.ctl[0], sweep = p.ctl[1];
  const i32* act = cur ? p.act1 : p.act0;
  i32* nx = cur ? p.act0 : p.act1;
  const i64 na = (i64)p.cnt[cur];
  for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < na; q += (i64)gridDim.x * blockDim.x)
    rp_commit_node(p, act[q], sweep, nx, cur);
}
__global__ void rp_commit_halo_kernel(CP p) {
  if (p.ctl[2]) return;
  const int cur = p.ctl[0], sweep = p.ctl[1];
  i32* nx = cur ? p.act0 : p.act1;
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 2 * p.H) return;
  rp_commit_node(p, p.h2[k], sweep, nx, cur);
}
__global__ void max_finite_kernel(const double* __restrict__ d, i64 n, unsigned long long* out) {
  unsigned long long m = 0ull;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const double v = d[i];
    if (v < __longlong_as_double(0x7ff0000000000000LL)) m = max(m, (unsigned long long)__double_as_longlong(v));
  }
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

}  // namespace

// Workspace of the pass (allocated on first use, kept with the mesh / grid).
struct CanonWs {
  DevBuf<i32> nt_fix, nt_cnt, ovf_node, ovf_len, ovf_idx, succ_off, succ_cur, succ_idx, stamp, act0, act1;
  DevBuf<double> rd, rd0;
  DevBuf<int> ctl;
  DevBuf<unsigned long long> cnt;
  DevBuf<uint8_t> scan_tmp;
  size_t scan_bytes = 0;
  bool ready = false;
};

void canon_ws_free(CanonWs* w) { delete w; }

namespace {

int ensure_canon_ws(CanonWs& w, i64 n, cudaStream_t s) {
  if (w.ready) return RT_OK;
  RT_TRY(w.nt_fix.alloc((size_t)n * KT));
  RT_TRY(w.nt_cnt.alloc(n));
  RT_TRY(w.ovf_node.alloc(n));
  RT_TRY(w.ovf_len.alloc(n + 1));
  RT_TRY(w.succ_off.alloc(n + 1));
  RT_TRY(w.succ_cur.alloc(n));
  RT_TRY(w.stamp.alloc(n));
  RT_TRY(w.act0.alloc(n));
  RT_TRY(w.act1.alloc(n));
  RT_TRY(w.rd.alloc(n));
  RT_TRY(w.rd0.alloc(n));
  RT_TRY(w.ctl.alloc(8));
  RT_TRY(w.cnt.alloc(4));
  cub::DeviceScan::ExclusiveSum(nullptr, w.scan_bytes, w.succ_off.p, w.succ_off.p, n + 1, s);
  RT_TRY(w.scan_tmp.alloc(w.scan_bytes));
  w.ready = true;
  return RT_OK;
}

template <int MODE>
void launch_relax(const CP& p, unsigned grid, cudaStream_t s) {
  rp_relax_kernel<MODE><<<grid, 128, 0, s>>>(p);
}

int run_canonical(rt_mesh* h, CP& p, CanonWs& w, int mode, i64* sweeps_out, i64* launches_out) {
  cudaStream_t s = h->stream;
  const i64 n = p.n;
  const bool f32 = mode == MODE_F32 || mode == MODE_3D_F32;
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  p.nt_fix = w.nt_fix.p;
  p.nt_cnt = w.nt_cnt.p;
  p.ovf_node = w.ovf_node.p;
  p.ovf_len = w.ovf_len.p;
  p.succ_off = w.succ_off.p;
  p.succ_cur = w.succ_cur.p;
  p.rd = w.rd.p;
  p.rd0 = w.rd0.p;
  p.stamp = w.stamp.p;
  p.act0 = w.act0.p;
  p.act1 = w.act1.p;
  p.ctl = w.ctl.p;
  p.cnt = w.cnt.p;
  // slack: 8 ulp of the largest travel time, in the arithmetic of the solve
  RT_CUDA(cudaMemsetAsync(w.cnt.p, 0, 4 * sizeof(unsigned long long), s));
  max_finite_kernel<<<(unsigned)(sm_count * 4), 256, 0, s>>>(p.fin, n, w.cnt.p + 2);
  unsigned long long mb = 0;
  RT_CUDA(cudaMemcpyAsync(&mb, w.cnt.p + 2, sizeof(mb), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  double tmax;
  std::memcpy(&tmax, &mb, sizeof(double));
  p.slack = 8.0 * tmax * (f32 ? 1.1920929e-7 : 2.220446049250313e-16);
  RT_CUDA(cudaMemsetAsync(w.ctl.p, 0, 8 * sizeof(int), s));
  RT_CUDA(cudaMemsetAsync(w.succ_off.p, 0, (n + 1) * sizeof(i32), s));
  RT_CUDA(cudaMemsetAsync(w.succ_cur.p, 0, n * sizeof(i32), s));
  i64 launches = 1;
  // ---- near-tight lists
  const unsigned gt = (unsigned)std::max<i64>(1, std::min<i64>((p.n_items + 3) / 4, (i64)sm_count * 32));
  switch (mode) {
    case MODE_F64: near_tight2_kernel<MODE_F64><<<gt, 128, 0, s>>>(p); break;
    case MODE_DUAL: near_tight2_kernel<MODE_DUAL><<<gt, 128, 0, s>>>(p); break;
    case MODE_F32: near_tight2_kernel<MODE_F32><<<gt, 128, 0, s>>>(p); break;
    case MODE_3D: near_tight3_kernel<false><<<grid_for(n, 128), 128, 0, s>>>(p); break;
    default: near_tight3_kernel<true><<<grid_for(n, 128), 128, 0, s>>>(p); break;
  }
  int hctl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  RT_CUDA(cudaMemcpyAsync(hctl, w.ctl.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  const int n_ovf = hctl[3];
  ++launches;
  if (n_ovf > 0) {
    const unsigned go = (unsigned)std::min<i64>(((i64)n_ovf * 32 + 127) / 128, (i64)sm_count * 16);
    switch (mode) {
      case MODE_F64: near_tight2_overflow_kernel<MODE_F64, 0><<<go, 128, 0, s>>>(p); break;
      case MODE_DUAL: near_tight2_overflow_kernel<MODE_DUAL, 0><<<go, 128, 0, s>>>(p); break;
      case MODE_F32: near_tight2_overflow_kernel<MODE_F32, 0><<<go, 128, 0, s>>>(p); break;
      case MODE_3D: near_tight3_overflow_kernel<false, 0><<<go, 128, 0, s>>>(p); break;
      default: near_tight3_overflow_kernel<true, 0><<<go, 128, 0, s>>>(p); break;
    }
    RT_CUDA(cudaMemsetAsync(w.ovf_len.p + n_ovf, 0, sizeof(i32), s));
    size_t sb = w.scan_bytes;
    cub::DeviceScan::ExclusiveSum(w.scan_tmp.p, sb, w.ovf_len.p, w.ovf_len.p, n_ovf + 1, s);
    i32 total = 0;
    RT_CUDA(cudaMemcpyAsync(&total, w.ovf_len.p + n_ovf, sizeof(i32), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    if (w.ovf_idx.n < (size_t)total) RT_TRY(w.ovf_idx.alloc((size_t)total + (size_t)total / 4 + 1024));
    p.ovf_idx = w.ovf_idx.p;
    switch (mode) {
      case MODE_F64: near_tight2_overflow_kernel<MODE_F64, 1><<<go, 128, 0, s>>>(p); break;
      case MODE_DUAL: near_tight2_overflow_kernel<MODE_DUAL, 1><<<go, 128, 0, s>>>(p); break;
      case MODE_F32: near_tight2_overflow_kernel<MODE_F32, 1><<<go, 128, 0, s>>>(p); break;
      case MODE_3D: near_tight3_overflow_kernel<false, 1><<<go, 128, 0, s>>>(p); break;
      default: near_tight3_overflow_kernel<true, 1><<<go, 128, 0, s>>>(p); break;
    }
    launches += 3;
  } else {
    p.ovf_idx = w.ovf_idx.p;
  }
  // ---- successors
  succ_count_kernel<<<grid_for(n, 256), 256, 0, s>>>(p);
  {
    size_t sb = w.scan_bytes;
    cub::DeviceScan::ExclusiveSum(w.scan_tmp.p, sb, w.succ_off.p, w.succ_off.p, n + 1, s);
  }
  i32 n_edges = 0;
  RT_CUDA(cudaMemcpyAsync(&n_edges, w.succ_off.p + n, sizeof(i32), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  if (w.succ_idx.n < (size_t)n_edges) RT_TRY(w.succ_idx.alloc((size_t)n_edges + (size_t)n_edges / 4 + 1024));
  p.succ_idx = w.succ_idx.p;
  succ_fill_kernel<<<grid_for(n, 256), 256, 0, s>>>(p);
  rp_init_kernel<<<grid_for(n, 256), 256, 0, s>>>(p);
  rp_seed_kernel<<<1, 32, 0, s>>>(p);
  launches += 5;
  // ---- replay of the reference's sweeps
  const unsigned gsm = (unsigned)(sm_count * 4);
  const bool halo = p.H > 0;
  hctl[2] = 0;
  i64 enq = 0;
  while (!hctl[2]) {
    for (int r = 0; r < 32; ++r) {
      rp_begin_kernel<<<1, 1, 0, s>>>(p);
      switch (mode) {
        case MODE_F64: launch_relax<MODE_F64>(p, gsm, s); break;
        case MODE_DUAL: launch_relax<MODE_DUAL>(p, gsm, s); break;
        case MODE_F32: launch_relax<MODE_F32>(p, gsm, s); break;
        case MODE_3D: launch_relax<MODE_3D>(p, gsm, s); break;
        default: launch_relax<MODE_3D_F32>(p, gsm, s); break;
      }
      if (halo) {
        rp_halo1_kernel<<<grid_for(p.H, 256), 256, 0, s>>>(p);
        rp_halo2_kernel<<<grid_for(p.n_groups, 256), 256, 0, s>>>(p);
      }
      rp_commit_kernel<<<gsm, 256, 0, s>>>(p);
      if (halo) rp_commit_halo_kernel<<<grid_for(2 * p.H, 256), 256, 0, s>>>(p);
      launches += halo ? 6 : 3;
    }
    enq += 32;
    RT_CUDA(cudaMemcpyAsync(hctl, w.ctl.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    if (enq > 4 * n + 64) {
      rt_set_error("canonical_prev did not terminate");
      return RT_ERR_CUDA;
    }
  }
  RT_CUDA(cudaGetLastError());
  if (sweeps_out) *sweeps_out = hctl[1];
  if (launches_out) *launches_out = launches;
  return RT_OK;
}

}  // namespace

// fin: converged travel times [n] (plain doubles); prev: [n] int32, rewritten for every reached node but the source.
// mode: 0 fp64 (U1 = U), 1 dual velocity (U1, U2 = the two columns, needs gr.r), 2 Float32 arithmetic (x, z, U already rounded)
int canonical_prev_2d(rt_mesh* h, const double* x, const double* z, const double* U1, const double* U2, int mode,
                      const double* fin, int source, i32* prev, i64* sweeps_out, i64* launches_out) {
  Mesh2D& m = *h->m2;
  if (m.halo_rows > 0 && !m.halo_structured) {
    rt_set_error("canonical_prev needs the halo matrix of init_annulus ((orig, twin) rows, then (twin, orig) rows)");
    return RT_ERR_UNSUPPORTED;
  }
  if (!m.canon) m.canon = new CanonWs();
  CanonWs& w = *m.canon;
  RT_TRY(ensure_canon_ws(w, m.n, h->stream));
  CP p = {};
  p.x = x;
  p.z = z;
  p.U1 = U1;
  p.U2 = U2 ? U2 : U1;
  p.r = m.r.p;
  p.e2n_off = m.e2n_off.p;
  p.e2n_idx = m.e2n_idx.p;
  p.g_off = m.g_off.p;
  p.g_idx = m.g_idx.p;
  p.item_first = m.item_first.p;
  p.fin = fin;
  p.n = m.n;
  p.n_items = m.n_items;
  p.source = source;
  p.prev = prev;
  p.h1 = m.halo_h1.p;
  p.h2 = m.halo_h2.p;
  p.H = m.halo_rows > 0 ? m.H : 0;
  p.g_orig = m.h2_orig.p;
  p.g_toff = m.h2_off.p;
  p.g_twin = m.h2_twin.p;
  p.n_groups = m.halo_rows > 0 ? m.n_h2_orig : 0;
  return run_canonical(h, p, w, mode, sweeps_out, launches_out);
}

// 3-D structured grid: the same pass on the implicit window adjacency (no halo).  X, Y, Z, U are the arrays the solve
// used (Float32-rounded copies when f32).
int canonical_prev_3d(rt_mesh* h, CanonWs** ws, const Grid3Desc& g, const double* U, bool f32, const double* fin,
                      i64 source, i32* prev, i64* launches_out) {
  const i64 n = (i64)g.nx * g.ny * g.nz;
  if (!*ws) *ws = new CanonWs();
  CanonWs& w = **ws;
  RT_TRY(ensure_canon_ws(w, n, h->stream));
  CP p = {};
  p.U1 = U;
  p.U2 = U;
  p.fin = fin;
  p.n = n;
  p.n_items = 0;
  p.source = (int)source;
  p.prev = prev;
  p.X3 = g.X;
  p.Y3 = g.Y;
  p.Z3 = g.Z;
  p.nx = g.nx;
  p.ny = g.ny;
  p.nz = g.nz;
  p.w3 = g.w;
  p.self3 = g.self;
  p.wmode3 = g.wmode;
  p.H = 0;
  p.n_groups = 0;
  return run_canonical(h, p, w, f32 ? MODE_3D_F32 : MODE_3D, nullptr, launches_out);
}
