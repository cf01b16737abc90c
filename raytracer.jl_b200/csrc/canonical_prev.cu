// canonical_prev.cu -- the reference's predecessor table, exact ties included, from converged travel times.
//
// The reference relaxes in Jacobi sweeps with a strict `>` (src/SSSP/bfm.jl:161-210): prev[i] is the FIRST candidate in
// scan order among those that attain the final dist[i] in the EARLIEST sweep in which that value is attainable, and
// update_halo! (:54-62) copies (dist, prev) across a halo row in the sweep in which the row's first node improved.  A
// work-efficient schedule reaches the same travel times (least fixed point) but visits candidates in another order, so
// its predecessors differ on exact ties -- which are systematic on these meshes (every radial edge exists twice).
//
// Given the converged dist this file rebuilds the sweep structure without sweeping:
//   1. tight edges  j -> i  :  fl(dist[j] + w_ji) == dist[i] bitwise, j != i, in the reference's scan order of i;
//   2. level(source) = 0, level(i) = 1 + min level over tight j  == the sweep in which i reaches its final value
//      (breadth-first search over the tight edges; halo rows pass the level on inside a sweep, first-half rows before
//      second-half rows, rows of one orig in row order -- the serial semantics of update_halo!);
//   3. prev[i] = first tight j in scan order with level(j) = level(i) - 1; halo-set nodes inherit their partner's prev.
// Checked bit for bit against the predecessors of the reference schedule (tests/test_canonical_prev.py).
#include <cub/device/device_scan.cuh>

#include "mesh2d.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int KT = 8;  // distinct tight predecessors kept per node; more -> the node is rescanned on demand (overflow list)
constexpr int MODE_F64 = 0, MODE_DUAL = 1, MODE_F32 = 2;

struct CP {
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
  const double* __restrict__ dist;  // converged travel times [n]
  i64 n, n_items;
  int source;
  i32* tight;      // [n x KT] tight predecessors in scan order (first occurrence of each node)
  i32* tcnt;       // [n] how many (KT + 1 = overflow: rescan)
  i32* ovf;        // overflow node list
  i32* succ_cnt;   // [n + 1] -> succ_off after the scan
  i32* succ_cur;   // [n]
  i32* succ_idx;
  i32* level;      // [n]
  i32* prev;       // [n] output (entries of unreached nodes are left alone)
  i32* fr0;
  i32* fr1;
  // halo (structured: rows [0,H) orig -> twin, rows [H,2H) twin -> orig): groups = origs with their twins in row order
  const i32* __restrict__ hn_index;   // node -> row of the halo-node table, -1 if the node is on no halo row
  const i32* __restrict__ hn_group;   // halo-node row -> group
  const i32* __restrict__ g_orig;
  const i32* __restrict__ g_toff;
  const i32* __restrict__ g_twin;
  i32* g_stamp;    // [n_groups] last level at which the group was queued
  i32* g_list;
  int n_groups;
  // control block: [0] cur parity [1] level L [2] done [3] n overflow [4] queued groups; counts: cnt[0], cnt[1]
  int* ctl;
  unsigned long long* cnt;
  // 3-D structured grid (implicit window adjacency, canonical scan order = ascending linear id); x, z unused there
  const double* __restrict__ X3;
  const double* __restrict__ Y3;
  const double* __restrict__ Z3;
  int nx, ny, nz, w3, self3, wmode3;
};

template <int MODE>
__device__ __forceinline__ bool is_tight(const CP& p, double di, double xi, double zi, double u1i, double u2i,
                                         double ri, double dj, double xj, double zj, double u1j, double u2j, double rj) {
  constexpr bool F32 = MODE == MODE_F32;
  double ut = u1i, us = u1j;
  if (MODE == MODE_DUAL) {  // bfm.jl:113-159: head = candidate, tail = target; head_idx = (r_i > r_Gi) + 1
    const bool down = ri > rj;
    ut = down ? u1i : u2i;
    us = down ? u2j : u1j;
  }
  const double dx = __dsub_rn(xi, xj), dz = __dsub_rn(zi, zj);
  const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz));
  if (!screen_maybe_tight_t<F32>(di, dj, d2, __dadd_rn(ut, us))) return false;
  return exact_cand2<F32>(dj, xj, zj, us, xi, zi, ut) == di;
}

// step 1: warp per work item (<= 32 targets sharing one G column); the candidates of the column pass through a
// warp-private shared-memory slab 32 at a time, every lane tests them in scan order against its own target
template <int MODE>
__global__ void __launch_bounds__(128) tight_build_kernel(CP p) {
  __shared__ double2 s_xz[4][32], s_ud[4][32], s_u2r[MODE == MODE_DUAL ? 4 : 1][32];
  __shared__ int s_id[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  for (i64 it = (i64)blockIdx.x * 4 + warp; it < p.n_items; it += (i64)gridDim.x * 4) {
    const int v0 = p.item_first[it];
    const int t = p.item_first[it + 1] - v0;
    const int i = v0 + min(lane, t - 1);
    const double di = p.dist[i];
    const bool want = lane < t && di < INF && i != p.source;
    if (!__any_sync(FULL, want)) {
      if (lane < t) p.tcnt[i] = 0;
      continue;
    }
    const double xi = p.x[i], zi = p.z[i], u1i = p.U1[i];
    const double u2i = MODE == MODE_DUAL ? p.U2[i] : 0.0, ri = MODE == MODE_DUAL ? p.r[i] : 0.0;
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
          s_ud[warp][lane] = make_double2(p.U1[j], p.dist[j]);
          if (MODE == MODE_DUAL) s_u2r[warp][lane] = make_double2(p.U2[j], p.r[j]);
        }
        __syncwarp();
        const int mm = min(32, m - k0);
        for (int q = 0; q < mm; ++q) {
          const double2 ud = s_ud[warp][q];
          if (!want || !(ud.y <= di)) continue;
          const int j = s_id[warp][q];
          if (j == i) continue;
          const double2 xz = s_xz[warp][q];
          double u2j = 0.0, rj = 0.0;
          if (MODE == MODE_DUAL) {
            u2j = s_u2r[warp][q].x;
            rj = s_u2r[warp][q].y;
          }
          if (!is_tight<MODE>(p, di, xi, zi, u1i, u2i, ri, ud.y, xz.x, xz.y, ud.x, u2j, rj)) continue;
          bool seen = false;
#pragma unroll
          for (int e = 0; e < KT; ++e) seen = seen || b[e] == j;
          if (seen) continue;
          if (cnt < KT) {
#pragma unroll
            for (int e = 0; e < KT; ++e)
              if (e == cnt) b[e] = j;
          }
          ++cnt;
        }
      }
    }
    if (lane < t) {
      p.tcnt[i] = cnt > KT ? KT + 1 : cnt;
#pragma unroll
      for (int e = 0; e < KT; ++e) p.tight[(i64)i * KT + e] = b[e];
      if (cnt > KT) p.ovf[atomicAdd(&p.ctl[3], 1)] = i;
    }
  }
}

__global__ void succ_count_kernel(CP p) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int c = min(p.tcnt[i], KT);
  for (int e = 0; e < c; ++e) atomicAdd(&p.succ_cnt[p.tight[i * KT + e]], 1);
}
__global__ void succ_fill_kernel(CP p) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int c = min(p.tcnt[i], KT);
  for (int e = 0; e < c; ++e) {
    const int j = p.tight[i * KT + e];
    p.succ_idx[p.succ_cnt[j] + atomicAdd(&p.succ_cur[j], 1)] = (i32)i;  // succ_cnt holds the offsets by now
  }
}
__global__ void cp_init_kernel(CP p) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.n) p.level[i] = i == p.source ? 0 : -1;
  if (i < p.n_groups) p.g_stamp[i] = -1;
  if (i == 0) {
    p.fr0[0] = p.source;
    p.cnt[0] = 1ull;
    p.cnt[1] = 0ull;
    p.ctl[0] = 0;
    p.ctl[1] = 0;
    p.ctl[2] = 0;
    p.ctl[4] = 0;
  }
}

// ---- one level = begin, expand, (overflow pull), select, halo
__global__ void cp_begin_kernel(CP p) {
  int* c = p.ctl;
  if (c[2]) return;
  if (c[1] > 0) c[0] ^= 1;  // the list filled during the previous level becomes the frontier
  if (p.cnt[c[0]] == 0ull) {
    c[2] = 1;
    return;
  }
  c[1] += 1;
  c[4] = 0;
  p.cnt[c[0] ^ 1] = 0ull;
}
__global__ void cp_expand_kernel(CP p) {
  if (p.ctl[2]) return;
  const int cur = p.ctl[0], L = p.ctl[1];
  const i32* fr = cur ? p.fr1 : p.fr0;
  i32* nx = cur ? p.fr0 : p.fr1;
  const i64 nf = (i64)p.cnt[cur];
  for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < nf; q += (i64)gridDim.x * blockDim.x) {
    const int j = fr[q];
    for (int e = p.succ_cnt[j]; e < p.succ_cnt[j + 1]; ++e) {
      const int i = p.succ_idx[e];
      if (p.level[i] == -1 && atomicCAS(&p.level[i], -1, L) == -1) nx[atomicAdd(&p.cnt[cur ^ 1], 1ull)] = i;
    }
  }
}
// nodes with more than KT distinct tight predecessors: rescan the whole list (one warp per node, lanes split it)
template <int MODE>
__global__ void cp_overflow_kernel(CP p) {
  if (p.ctl[2]) return;
  const int cur = p.ctl[0], L = p.ctl[1];
  i32* nx = cur ? p.fr0 : p.fr1;
  const int lane = threadIdx.x & 31;
  const int no = p.ctl[3];
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < no; w += (gridDim.x * blockDim.x) >> 5) {
    const int i = p.ovf[w];
    if (p.level[i] != -1) continue;  // warp-uniform
    const double di = p.dist[i], xi = p.x[i], zi = p.z[i], u1i = p.U1[i];
    const double u2i = MODE == MODE_DUAL ? p.U2[i] : 0.0, ri = MODE == MODE_DUAL ? p.r[i] : 0.0;
    int bpos = 0x7fffffff, bid = -1, pos0 = 0;
    for (i64 c = p.g_off[i]; c < p.g_off[i + 1]; ++c) {
      const int el = p.g_idx[c];
      const int s = p.e2n_off[el], m = p.e2n_off[el + 1] - s;
      for (int k = lane; k < m; k += 32) {
        const int j = p.e2n_idx[s + k];
        if (j == i || p.level[j] != L - 1 || pos0 + k > bpos) continue;
        const double dj = p.dist[j];
        if (!(dj <= di)) continue;
        if (is_tight<MODE>(p, di, xi, zi, u1i, u2i, ri, dj, p.x[j], p.z[j], p.U1[j], MODE == MODE_DUAL ? p.U2[j] : 0.0,
                           MODE == MODE_DUAL ? p.r[j] : 0.0)) {
          bpos = pos0 + k;
          bid = j;
        }
      }
      pos0 += m;
    }
    for (int o = 16; o; o >>= 1) {
      const int op = __shfl_xor_sync(FULL, bpos, o), oi = __shfl_xor_sync(FULL, bid, o);
      if (op < bpos) {
        bpos = op;
        bid = oi;
      }
    }
    if (lane == 0 && bid >= 0) {
      p.level[i] = L;
      p.prev[i] = bid;
      nx[atomicAdd(&p.cnt[cur ^ 1], 1ull)] = i;
    }
  }
}
__global__ void cp_select_kernel(CP p) {
  if (p.ctl[2]) return;
  const int cur = p.ctl[0], L = p.ctl[1];
  const i32* nx = cur ? p.fr0 : p.fr1;
  const i64 nn = (i64)p.cnt[cur ^ 1];
  for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < nn; q += (i64)gridDim.x * blockDim.x) {
    const int i = nx[q];
    const int c = p.tcnt[i];
    if (c <= KT) {  // (overflow nodes got their predecessor in cp_overflow_kernel)
      for (int e = 0; e < c; ++e) {
        const int j = p.tight[(i64)i * KT + e];
        if (p.level[j] == L - 1) {
          p.prev[i] = j;
          break;
        }
      }
    }
    if (p.n_groups > 0) {
      const int hq = p.hn_index[i];
      if (hq >= 0) {
        const int g = p.hn_group[hq];
        if (g >= 0 && atomicExch(&p.g_stamp[g], L) != L) p.g_list[atomicAdd(&p.ctl[4], 1)] = g;
      }
    }
  }
}
// update_halo! inside sweep L, per orig with its twins in row order: first half (orig -> twins) if the orig reached its
// final value in this sweep, else second half (first twin in row order that did passes value and predecessor to the orig)
__global__ void cp_halo_kernel(CP p) {
  if (p.ctl[2]) return;
  const int cur = p.ctl[0], L = p.ctl[1];
  i32* nx = cur ? p.fr0 : p.fr1;
  const int ng = p.ctl[4];
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ng; q += gridDim.x * blockDim.x) {
    const int g = p.g_list[q];
    const int o = p.g_orig[g];
    const int lo = p.level[o];
    if (lo == L) {
      if (o == p.source) continue;  // the source never improves: its rows never fire (bfm.jl:56)
      for (int e = p.g_toff[g]; e < p.g_toff[g + 1]; ++e) {
        const int b = p.g_twin[e];
        if (p.level[b] == -1 && p.dist[b] == p.dist[o]) {
          p.level[b] = L;
          p.prev[b] = p.prev[o];
          nx[atomicAdd(&p.cnt[cur ^ 1], 1ull)] = b;
        }
      }
    } else if (lo == -1) {
      for (int e = p.g_toff[g]; e < p.g_toff[g + 1]; ++e) {
        const int b = p.g_twin[e];
        if (p.level[b] == L && b != p.source && p.dist[b] == p.dist[o]) {
          p.level[o] = L;
          p.prev[o] = p.prev[b];
          nx[atomicAdd(&p.cnt[cur ^ 1], 1ull)] = o;
          break;
        }
      }
    }
  }
}

// ---- 3-D: tight predecessors of node I inside its clipped window, ascending linear id (thread per node)
template <bool F32>
__device__ __forceinline__ bool is_tight3(const CP& p, double di, double xi, double yi, double zi, double ui, i64 J) {
  const double dj = p.dist[J];
  if (!(dj <= di)) return false;
  const double xj = p.X3[J], yj = p.Y3[J], zj = p.Z3[J], uj = p.U1[J];
  const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj), dz = __dsub_rn(zi, zj);
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  if (!screen_maybe_tight_t<F32>(di, dj, d2, screen_ssum3(fabs(__dadd_rn(ui, uj)), p.wmode3))) return false;
  return exact_cand3<F32>(dj, xi, yi, zi, ui, xj, yj, zj, uj, p.wmode3) == di;
}
template <bool F32>
__global__ void __launch_bounds__(128) tight_build3_kernel(CP p) {
  const i64 I = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= p.n) return;
  const double di = p.dist[I];
  int cnt = 0;
  int b[KT];
#pragma unroll
  for (int q = 0; q < KT; ++q) b[q] = -1;
  if (di < __longlong_as_double(0x7ff0000000000000LL) && (int)I != p.source) {
    const int i = (int)(I % p.nx), j = (int)((I / p.nx) % p.ny), k = (int)(I / ((i64)p.nx * p.ny));
    const double xi = p.X3[I], yi = p.Y3[I], zi = p.Z3[I], ui = p.U1[I];
    const int w = p.w3;
    for (int zz = max(0, k - w); zz <= min(p.nz - 1, k + w); ++zz)
      for (int yy = max(0, j - w); yy <= min(p.ny - 1, j + w); ++yy)
        for (int xx = max(0, i - w); xx <= min(p.nx - 1, i + w); ++xx) {
          const i64 J = (i64)xx + (i64)p.nx * ((i64)yy + (i64)p.ny * zz);
          if (J == I) continue;
          if (!is_tight3<F32>(p, di, xi, yi, zi, ui, J)) continue;
          if (cnt < KT) {
#pragma unroll
            for (int e = 0; e < KT; ++e)
              if (e == cnt) b[e] = (int)J;
          }
          ++cnt;
        }
  }
  p.tcnt[I] = cnt > KT ? KT + 1 : cnt;
#pragma unroll
  for (int e = 0; e < KT; ++e) p.tight[I * KT + e] = b[e];
  if (cnt > KT) p.ovf[atomicAdd(&p.ctl[3], 1)] = (i32)I;
}
template <bool F32>
__global__ void cp_overflow3_kernel(CP p) {
  if (p.ctl[2]) return;
  const int cur = p.ctl[0], L = p.ctl[1];
  i32* nx = cur ? p.fr0 : p.fr1;
  const int lane = threadIdx.x & 31;
  const int no = p.ctl[3];
  for (int wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wq < no; wq += (gridDim.x * blockDim.x) >> 5) {
    const i64 I = p.ovf[wq];
    if (p.level[I] != -1) continue;  // warp-uniform
    const int i = (int)(I % p.nx), j = (int)((I / p.nx) % p.ny), k = (int)(I / ((i64)p.nx * p.ny));
    const double di = p.dist[I], xi = p.X3[I], yi = p.Y3[I], zi = p.Z3[I], ui = p.U1[I];
    const int w = p.w3;
    const int x0 = max(0, i - w), x1 = min(p.nx - 1, i + w), y0 = max(0, j - w), y1 = min(p.ny - 1, j + w);
    const int z0 = max(0, k - w), z1 = min(p.nz - 1, k + w);
    const int cx = x1 - x0 + 1, cy = y1 - y0 + 1, total = cx * cy * (z1 - z0 + 1);
    i64 best = 0x7fffffffffffffffLL;
    for (int t = lane; t < total; t += 32) {  // window cells in ascending linear id
      const int xx = x0 + t % cx, yy = y0 + (t / cx) % cy, zz = z0 + t / (cx * cy);
      const i64 J = (i64)xx + (i64)p.nx * ((i64)yy + (i64)p.ny * zz);
      if (J == I || J > best || p.level[J] != L - 1) continue;
      if (is_tight3<F32>(p, di, xi, yi, zi, ui, J)) best = J;
    }
    for (int o = 16; o; o >>= 1) {
      const i64 ob = __shfl_xor_sync(FULL, best, o);
      best = ob < best ? ob : best;
    }
    if (lane == 0 && best != 0x7fffffffffffffffLL) {
      p.level[I] = L;
      p.prev[I] = (i32)best;
      nx[atomicAdd(&p.cnt[cur ^ 1], 1ull)] = (i32)I;
    }
  }
}

template <int MODE>
void launch_tight(const CP& p, unsigned grid, cudaStream_t s) {
  tight_build_kernel<MODE><<<grid, 128, 0, s>>>(p);
}
template <int MODE>
void launch_overflow(const CP& p, unsigned grid, cudaStream_t s) {
  cp_overflow_kernel<MODE><<<grid, 128, 0, s>>>(p);
}

}  // namespace

// Workspace of the pass (allocated on first use, kept with the mesh).
struct CanonWs {
  DevBuf<i32> tight, tcnt, ovf, succ_cnt, succ_cur, succ_idx, level, fr0, fr1, hn_group, g_stamp, g_list;
  DevBuf<int> ctl;
  DevBuf<unsigned long long> cnt;
  DevBuf<uint8_t> scan_tmp;
  size_t scan_bytes = 0;
  bool ready = false;
};

void canon_ws_free(CanonWs* w) { delete w; }

namespace {

int ensure_canon_ws(CanonWs& w, i64 n, int n_groups, cudaStream_t s) {
  if (w.ready) return RT_OK;
  RT_TRY(w.tight.alloc((size_t)n * KT));
  RT_TRY(w.tcnt.alloc(n));
  RT_TRY(w.ovf.alloc(n));
  RT_TRY(w.succ_cnt.alloc(n + 1));
  RT_TRY(w.succ_cur.alloc(n));
  RT_TRY(w.succ_idx.alloc((size_t)n * KT));
  RT_TRY(w.level.alloc(n));
  RT_TRY(w.fr0.alloc(n));
  RT_TRY(w.fr1.alloc(n));
  RT_TRY(w.ctl.alloc(8));
  RT_TRY(w.cnt.alloc(2));
  RT_TRY(w.g_stamp.alloc(std::max(n_groups, 1)));
  RT_TRY(w.g_list.alloc(std::max(n_groups, 1)));
  cub::DeviceScan::ExclusiveSum(nullptr, w.scan_bytes, w.succ_cnt.p, w.succ_cnt.p, n + 1, s);
  RT_TRY(w.scan_tmp.alloc(w.scan_bytes));
  w.ready = true;
  return RT_OK;
}

void bind_ws(CP& p, CanonWs& w) {
  p.tight = w.tight.p;
  p.tcnt = w.tcnt.p;
  p.ovf = w.ovf.p;
  p.succ_cnt = w.succ_cnt.p;
  p.succ_cur = w.succ_cur.p;
  p.succ_idx = w.succ_idx.p;
  p.level = w.level.p;
  p.fr0 = w.fr0.p;
  p.fr1 = w.fr1.p;
  p.hn_group = w.hn_group.p;
  p.g_stamp = w.g_stamp.p;
  p.g_list = w.g_list.p;
  p.ctl = w.ctl.p;
  p.cnt = w.cnt.p;
}

// kind: 0..2 = 2-D relax modes (MODE_*), 3 = 3-D fp64, 4 = 3-D Float32
int run_canonical(rt_mesh* h, CP& p, CanonWs& w, int kind, i64 n_items, i64* levels_out, i64* launches_out) {
  cudaStream_t s = h->stream;
  const i64 n = p.n;
  const int n_groups = p.n_groups;
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  const unsigned gfull = grid_for(std::max<i64>(n, n_groups), 256);
  RT_CUDA(cudaMemsetAsync(w.ctl.p, 0, 8 * sizeof(int), s));
  RT_CUDA(cudaMemsetAsync(w.succ_cnt.p, 0, (n + 1) * sizeof(i32), s));
  RT_CUDA(cudaMemsetAsync(w.succ_cur.p, 0, n * sizeof(i32), s));
  const unsigned gt = (unsigned)std::min<i64>((n_items + 3) / 4, (i64)sm_count * 32);
  if (kind == MODE_DUAL)
    launch_tight<MODE_DUAL>(p, gt, s);
  else if (kind == MODE_F32)
    launch_tight<MODE_F32>(p, gt, s);
  else if (kind == MODE_F64)
    launch_tight<MODE_F64>(p, gt, s);
  else if (kind == 3)
    tight_build3_kernel<false><<<grid_for(n, 128), 128, 0, s>>>(p);
  else
    tight_build3_kernel<true><<<grid_for(n, 128), 128, 0, s>>>(p);
  succ_count_kernel<<<grid_for(n, 256), 256, 0, s>>>(p);
  cub::DeviceScan::ExclusiveSum(w.scan_tmp.p, w.scan_bytes, w.succ_cnt.p, w.succ_cnt.p, n + 1, s);
  succ_fill_kernel<<<grid_for(n, 256), 256, 0, s>>>(p);
  cp_init_kernel<<<gfull, 256, 0, s>>>(p);
  i64 launches = 6;
  int hctl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  RT_CUDA(cudaMemcpyAsync(hctl, w.ctl.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  const bool has_ovf = hctl[3] > 0;
  const unsigned gsm = (unsigned)(sm_count * 2);
  i64 enq = 0;
  while (!hctl[2]) {
    for (int r = 0; r < 32; ++r) {
      cp_begin_kernel<<<1, 1, 0, s>>>(p);
      cp_expand_kernel<<<gsm, 256, 0, s>>>(p);
      if (has_ovf) {
        if (kind == MODE_DUAL)
          launch_overflow<MODE_DUAL>(p, gsm, s);
        else if (kind == MODE_F32)
          launch_overflow<MODE_F32>(p, gsm, s);
        else if (kind == MODE_F64)
          launch_overflow<MODE_F64>(p, gsm, s);
        else if (kind == 3)
          cp_overflow3_kernel<false><<<gsm, 128, 0, s>>>(p);
        else
          cp_overflow3_kernel<true><<<gsm, 128, 0, s>>>(p);
      }
      cp_select_kernel<<<gsm, 256, 0, s>>>(p);
      if (n_groups > 0) cp_halo_kernel<<<gsm, 256, 0, s>>>(p);
      launches += 3 + (has_ovf ? 1 : 0) + (n_groups > 0 ? 1 : 0);
    }
    enq += 32;
    RT_CUDA(cudaMemcpyAsync(hctl, w.ctl.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    if (enq > n + 64) {
      rt_set_error("canonical_prev did not terminate");
      return RT_ERR_CUDA;
    }
  }
  RT_CUDA(cudaGetLastError());
  if (levels_out) *levels_out = hctl[1];  // == the reference's sweep count
  if (launches_out) *launches_out = launches;
  return RT_OK;
}

}  // namespace

// dist: converged travel times [n] (plain doubles); prev: [n] int32, overwritten for every reached node but the source.
// mode: 0 fp64 (U1 = U), 1 dual velocity (U1, U2 = the two columns, needs gr.r), 2 Float32 arithmetic (x, z, U already rounded)
int canonical_prev_2d(rt_mesh* h, const double* x, const double* z, const double* U1, const double* U2, int mode,
                      const double* dist, int source, i32* prev, i64* levels_out, i64* launches_out) {
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  const i64 n = m.n;
  if (m.halo_rows > 0 && !m.halo_structured) {
    rt_set_error("canonical_prev needs the halo matrix of init_annulus ((orig, twin) rows, then (twin, orig) rows)");
    return RT_ERR_UNSUPPORTED;
  }
  if (!m.canon) m.canon = new CanonWs();
  CanonWs& w = *m.canon;
  const int n_groups = (int)m.n_h2_orig;
  const bool first = !w.ready;
  RT_TRY(ensure_canon_ws(w, n, n_groups, s));
  if (first && m.n_hn > 0) {  // halo-node row -> group (orig and twins of one orig share a group)
    std::vector<i32> hn(m.n_hn), go(std::max(n_groups, 1)), gt(std::max<i64>(m.H, 1)), goff(n_groups + 1);
    RT_CUDA(cudaMemcpy(hn.data(), m.hn_node.p, m.n_hn * sizeof(i32), cudaMemcpyDeviceToHost));
    std::vector<i32> grp(m.n_hn, -1);
    if (n_groups > 0) {
      RT_CUDA(cudaMemcpy(go.data(), m.h2_orig.p, n_groups * sizeof(i32), cudaMemcpyDeviceToHost));
      RT_CUDA(cudaMemcpy(goff.data(), m.h2_off.p, (n_groups + 1) * sizeof(i32), cudaMemcpyDeviceToHost));
      RT_CUDA(cudaMemcpy(gt.data(), m.h2_twin.p, m.H * sizeof(i32), cudaMemcpyDeviceToHost));
      auto row_of = [&](i32 node) { return (i64)(std::lower_bound(hn.begin(), hn.end(), node) - hn.begin()); };
      for (int g = 0; g < n_groups; ++g) {
        grp[row_of(go[g])] = g;
        for (int e = goff[g]; e < goff[g + 1]; ++e) grp[row_of(gt[e])] = g;
      }
    }
    RT_TRY(w.hn_group.upload(grp.data(), grp.size()));
  }
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
  p.dist = dist;
  p.n = n;
  p.n_items = m.n_items;
  p.source = source;
  p.prev = prev;
  bind_ws(p, w);
  p.hn_index = m.hn_index.p;
  p.g_orig = m.h2_orig.p;
  p.g_toff = m.h2_off.p;
  p.g_twin = m.h2_twin.p;
  p.n_groups = n_groups;
  return run_canonical(h, p, w, mode, m.n_items, levels_out, launches_out);
}

// 3-D structured grid: the same pass on the implicit window adjacency (canonical scan order = ascending linear id; no
// halo).  X, Y, Z, U are the arrays the solve used (Float32-rounded copies when f32).
int canonical_prev_3d(rt_mesh* h, CanonWs** ws, const Grid3Desc& g, const double* U, bool f32, const double* dist,
                      i64 source, i32* prev, i64* launches_out) {
  const i64 n = (i64)g.nx * g.ny * g.nz;
  if (!*ws) *ws = new CanonWs();
  CanonWs& w = **ws;
  RT_TRY(ensure_canon_ws(w, n, 0, h->stream));
  CP p = {};
  p.U1 = U;
  p.U2 = U;
  p.dist = dist;
  p.n = n;
  p.n_items = 0;
  p.source = (int)source;
  p.prev = prev;
  bind_ws(p, w);
  p.n_groups = 0;
  p.X3 = g.X;
  p.Y3 = g.Y;
  p.Z3 = g.Z;
  p.nx = g.nx;
  p.ny = g.ny;
  p.nz = g.nz;
  p.w3 = g.w;
  p.self3 = g.self;
  p.wmode3 = g.wmode;
  return run_canonical(h, p, w, f32 ? 4 : 3, 0, nullptr, launches_out);
}
