// bfm2d_push.cu -- work-efficient schedule for bfm on the two-level annulus graph (schedule = 1, "near-far").
//
// The reference's Jacobi sweeps (src/SSSP/bfm.jl:29-47) re-relax every vertex 40-70 times on these meshes
// (measured: E_relaxed / E_graph = 39.6 ... 67).  Because fp `+` is monotone and weights are >= 0 the converged
// travel times are the least fixed point and do not depend on the relaxation order, so this file
// uses a label-correcting PUSH schedule ordered by travel time instead:
//   * dist is updated with a 64-bit atomicMin on the bit pattern of the (non-negative) double;
//   * a node whose value improved carries a flag; it is released (pushes dist + w to its whole star patch) only
//     once its value is below the moving threshold tau ("near"); improvements above tau wait in a far list;
//     when no near work is left tau moves to (smallest waiting value + delta);
//   * a released work item (<= 32 nodes sharing one G column) is expanded by one CTA: its column's elements are
//     spread over the warps, lanes hold TARGET nodes (coalesced id loads, gathers of x, z, U), the released
//     sources are broadcast from shared memory, and the edge weight 2*len/(U_i+U_j) is computed in registers
//     with the reference's operation order (it is bitwise symmetric in i and j);
//   * the halo rule (update_halo!, bfm.jl:54-62) is a zero-weight coupling between twins and is pushed too.
// Travel times are therefore bit-identical to the Jacobi schedule / the reference.  Predecessors are assigned
// afterwards by a deterministic pass: prev[i] = first candidate in the reference's scan order that is
// bit-exactly tight (dist[j] + w == dist[i]) with dist[j] < dist[i]; nodes that owe their value to a
// zero-weight coupling (twins, coincident duplicates) inherit along that coupling.  This equals the
// reference's prev except on exact ties (which are systematic on these meshes: every radial edge exists twice).
#include <cooperative_groups.h>

#include <cstring>
#include <cub/device/device_scan.cuh>

#include "mesh2d.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int PUSH_BLOCK = 128;
constexpr int PUSH_GY = 8;  // element groups per released item (x 4 warps = 32 elements per pass)

struct PP {
  const double* __restrict__ x;
  const double* __restrict__ z;
  const double* __restrict__ U;
  const double* __restrict__ U1;  // dual-velocity relax (bfm.jl:113-159): U[:,1] / U[:,2] and gr.r
  const double* __restrict__ U2;
  const double* __restrict__ r;
  const i32* __restrict__ e2n_off;
  const i32* __restrict__ e2n_idx;
  const i64* __restrict__ g_off;
  const i32* __restrict__ g_idx;
  const i32* __restrict__ item_first;
  const i32* __restrict__ node_item;
  const i32* __restrict__ hn_node;
  const i32* __restrict__ hn_off;
  const i32* __restrict__ hn_part;
  const i32* __restrict__ hn_index;  // node -> row of hn_off (or -1)
  int n_hn;
  int source;  // the source never 'improves', so update_halo! never fires from it (bfm.jl:56)
  double* dist;   // travel times; element stride ds (1 = plain array, 2 = first word of the packed (dist, key) pairs)
  int ds;
  u64* keys;      // packed mode: keys[2*j + 1] = predecessor key of node j (same allocation as dist)
  i32* prev;
  unsigned* pend_mask;
  unsigned* far_mask;
  unsigned* infar;
  unsigned* cur_mask;
  u64* counters;  // [0],[1] near counts (ping-pong) [2] evals [3] releases [4],[5] far counts (ping-pong)
                  // [6] unresolved count [7] scratch
  double* tau;    // [0] tau [1] delta [2] min far (bits)
  i32 *nearq0, *nearq1;  // selected with nq() / fq(): no dynamic indexing, so the struct stays out of local memory
  i32 *farq0, *farq1;
  int* ctl;       // device-side round control: [0] cur [1] fcur [2] mode (1 push, 2 advance) [3] done [4] rounds
                  // [5] push rounds
  // batch of sources solved in lock step (state arrays hold nb slices; pp_view() selects one)
  int nb;
  int cta_units;   // long columns: 1 = CTA per (item, element group) with block barriers, 0 = warp per (item, element)
  int compact;       // 1: long-column units park the targets that survive the group screen and evaluate full warps
  int group_screen;  // 1: per (item, target) disc bound before the source loop (screen.h: group_cannot_improve_t)
  int warp_units;  // 1: short columns -> warp-per-item push (push2d_warp_body), 0: CTA per (item, element group)
  i64 n, n_items;
  const int* sources;  // [nb] 0-based
  // flattened work of a batch round (written by round_begin): flat[0 .. nb] = exclusive prefix over the sources of
  // the near-list slots of the sources that push this round, flat[FLAT_STRIDE + (0 .. nb)] = the same for the far-list
  // slots of the sources whose threshold advances.  One launch then serves all sources with perfect load balance.
  i64* flat;
  unsigned* ticket;  // finished CTAs of the kernel that carries the round control in its tail (round_tail)
  // short-column meshes: de-duplicated target list per work item (first occurrences of the column's scan list, scan
  // order kept): a node shared by several elements of the column is visited once per released item.  tgt_off == null:
  // not built; an empty list = this item walks its elements (column too long for the builder).
  const i64* __restrict__ tgt_off;
  const i32* __restrict__ tgt_idx;
  i32* flat_b;   // owner source of the first flat_cap global near-list slots of this round (written by prep)
  i64 flat_cap;
};
constexpr int MAX_NB = 1024;          // sources advancing in lock step (one thread each in round_begin)
constexpr int FLAT_STRIDE = MAX_NB + 1;

// source that owns global slot g: largest b with base[b] <= g (sources without slots are skipped)
__device__ __forceinline__ int flat_owner(const i64* __restrict__ base, int nb, i64 g) {
  int lo = 0, hi = nb;  // invariant: base[lo] <= g < base[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (base[mid] <= g)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__device__ __forceinline__ i32* nq(const PP& p, int k) { return k ? p.nearq1 : p.nearq0; }
__device__ __forceinline__ i32* fq(const PP& p, int k) { return k ? p.farq1 : p.farq0; }

// slice b of the per-source state (mesh arrays are shared)
__device__ __forceinline__ PP pp_view(const PP& p, int b) {
  PP v = p;
  v.dist = p.dist + (i64)b * p.n * p.ds;
  v.keys = p.keys ? p.keys + (i64)b * p.n * 2 : nullptr;
  v.prev = p.prev + (i64)b * p.n;
  v.pend_mask = p.pend_mask + (i64)b * p.n_items;
  v.far_mask = p.far_mask + (i64)b * p.n_items;
  v.infar = p.infar + (i64)b * p.n_items;
  v.cur_mask = p.cur_mask + (i64)b * p.n_items;
  v.counters = p.counters + (i64)b * 8;
  v.tau = p.tau + (i64)b * 4;
  v.nearq0 = p.nearq0 + (i64)b * p.n_items;
  v.nearq1 = p.nearq1 + (i64)b * p.n_items;
  v.farq0 = p.farq0 + (i64)b * p.n_items;
  v.farq1 = p.farq1 + (i64)b * p.n_items;
  v.ctl = p.ctl + (i64)b * 8;
  v.source = p.sources ? p.sources[b] : p.source;
  v.nb = 1;
  return v;
}

// exact_cand2<F32> (exact.h); F32: every operation rounded to Float32 (bfm_gpu.jl:487-526)
template <bool F32 = false>
__device__ __forceinline__ double edge_delta(double di, double xi, double zi, double Ui, double xj, double zj,
                                             double Uj) {
  return exact_cand2<F32>(di, xi, zi, Ui, xj, zj, Uj);
}
// relax modes of the push kernels: plain fp64, dual velocity (bfm.jl:113-159), Float32 arithmetic
constexpr int MODE_F64 = 0, MODE_DUAL = 1, MODE_F32 = 2;

__global__ void hn_index_kernel(const i32* __restrict__ hn_node, i64 n_hn, i32* __restrict__ hn_index) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_hn) hn_index[hn_node[k]] = (i32)k;
}
__global__ void node_item_kernel(const i32* __restrict__ item_first, i64 n_items, i32* __restrict__ node_item) {
  const i64 it = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (it >= n_items) return;
  const int v0 = item_first[it], t = item_first[it + 1] - v0;
  if (lane < t) node_item[v0 + lane] = (i32)it;
}

// mean travel time across a cell (first to third node of every element = the diagonal of a quad): the scale
// of the long edges of a star patch, used to size the near-far bucket automatically
__global__ void wdiag_kernel(PP p, i64 nel, double* __restrict__ sum) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  double wt = 0.0;
  if (e < nel && p.e2n_off[e + 1] - p.e2n_off[e] >= 3) {
    const int a = p.e2n_idx[p.e2n_off[e]], b = p.e2n_idx[p.e2n_off[e] + 2];
    wt = edge_delta(0.0, p.x[a], p.z[a], p.U[a], p.x[b], p.z[b], p.U[b]);
    if (!(wt == wt) || wt > 1e300) wt = 0.0;
  }
  for (int o = 16; o; o >>= 1) wt += __shfl_xor_sync(FULL, wt, o);
  if ((threadIdx.x & 31) == 0 && wt > 0.0) atomicAdd(sum, wt);
}

// disc around the released sources of an item and their largest velocity (screen.h: group_cannot_improve_t)
struct GroupDisc {
  double cx, cz, rho, umax;
};
__device__ __forceinline__ GroupDisc group_disc(bool on, double x, double z, double u) {
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  double xmn = on ? x : INF, xmx = on ? x : -INF, zmn = on ? z : INF, zmx = on ? z : -INF, um = on ? u : 0.0;
  for (int o = 16; o; o >>= 1) {
    xmn = fmin(xmn, __shfl_xor_sync(FULL, xmn, o));
    xmx = fmax(xmx, __shfl_xor_sync(FULL, xmx, o));
    zmn = fmin(zmn, __shfl_xor_sync(FULL, zmn, o));
    zmx = fmax(zmx, __shfl_xor_sync(FULL, zmx, o));
    um = fmax(um, __shfl_xor_sync(FULL, um, o));
  }
  GroupDisc g;
  g.cx = 0.5 * (xmn + xmx);
  g.cz = 0.5 * (zmn + zmx);
  const double hx = xmx - xmn, hz = zmx - zmn;
  g.rho = 0.5 * sqrt(hx * hx + hz * hz) * (1.0 + 1e-12) + 1e-300;
  g.umax = um;
  return g;
}

// de-duplicated target lists (see PP::tgt_off): one warp per item gathers the ids of its column into shared memory and
// keeps the first occurrence of every node; FILL = 0 counts, 1 writes.  Columns above TGT_CAP entries get no list.
constexpr int TGT_CAP = 2048;
template <int FILL>
__global__ void __launch_bounds__(128) item_targets_kernel(PP p, i64* __restrict__ len_or_off, i32* __restrict__ out) {
  __shared__ int buf[4][TGT_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (i64 it = (i64)blockIdx.x * 4 + warp; it < p.n_items; it += (i64)gridDim.x * 4) {
    const int v0 = p.item_first[it];
    const i64 c0 = p.g_off[v0], c1 = p.g_off[v0 + 1];
    int m = 0;
    bool fits = true;
    __syncwarp();
    for (i64 c = c0; c < c1 && fits; ++c) {
      const int el = p.g_idx[c];
      const int s = p.e2n_off[el], len = p.e2n_off[el + 1] - s;
      if (m + len > TGT_CAP) {
        fits = false;
        break;
      }
      for (int k = lane; k < len; k += 32) buf[warp][m + k] = p.e2n_idx[s + k];
      m += len;
    }
    __syncwarp();
    if (!fits) {
      if (!FILL && lane == 0) len_or_off[it] = 0;
      continue;
    }
    int total = 0;
    const i64 base = FILL ? len_or_off[it] : 0;
    for (int k0 = 0; k0 < m; k0 += 32) {
      const int k = k0 + lane;
      bool uniq = false;
      int id = -1;
      if (k < m) {
        id = buf[warp][k];
        uniq = true;
        for (int q = 0; q < k; ++q)
          if (buf[warp][q] == id) {
            uniq = false;
            break;
          }
      }
      const unsigned ball = __ballot_sync(FULL, uniq);
      if (FILL && uniq) out[base + total + __popc(ball & ((1u << lane) - 1u))] = id;
      total += __popc(ball);
    }
    if (!FILL && lane == 0) len_or_off[it] = total;
  }
}

// flag node j (whose value just improved to d) for propagation: near list if d < tau, else far list
__device__ __forceinline__ void enqueue(const PP& p, int j, double d, double tau, i32* __restrict__ near_next,
                                        int nxt, i32* __restrict__ far_list, int fcur) {
  const int it = p.node_item[j];
  const unsigned bit = 1u << (j - p.item_first[it]);
  if (d < tau) {
    const unsigned old = atomicOr(&p.pend_mask[it], bit);
    if (old == 0u) near_next[atomicAdd(&p.counters[nxt], 1ull)] = it;
    if (__ldcg(&p.far_mask[it]) & bit) atomicAnd(&p.far_mask[it], ~bit);
  } else {
    atomicOr(&p.far_mask[it], bit);
    if (atomicExch(&p.infar[it], 1u) == 0u) far_list[atomicAdd(&p.counters[4 + fcur], 1ull)] = it;
  }
}

// try dist[j] = min(dist[j], d); returns true if it improved
__device__ __forceinline__ bool relax_to(const PP& p, int j, double d) {
  const u64 bits = (u64)__double_as_longlong(d);
  const u64 old = atomicMin((u64*)&p.dist[(i64)j * p.ds], bits);
  return bits < old;
}


// ---- packed (travel time, predecessor key) pairs: one 128-bit atomic compare-and-swap keeps them consistent, so
// no separate predecessor pass is needed.  Order: smaller time wins; on an exact tie a REGULAR predecessor
// (positive-weight edge) with a smaller node id wins; zero-weight couplings (coincident nodes, halo twins) only
// win by strict improvement (no cycles among equal-time nodes).  The result is schedule independent: every tight
// predecessor eventually pushes its final value, so the key ends as the smallest id among them.
constexpr u64 KEY_NONE = ~0ull;
constexpr u64 KEY_ZW = 1ull << 40;    // zero-weight edge between coincident nodes
constexpr u64 KEY_HALO = 1ull << 41;  // zero-weight halo coupling (resolved to the twin's predecessor afterwards)
constexpr u64 KEY_ZMASK = KEY_ZW | KEY_HALO;

struct alignas(16) DP {
  u64 d;  // bit pattern of the (non-negative) travel time
  u64 k;
};
__device__ __forceinline__ DP dp_cas(DP* addr, DP expected, DP desired) {
  DP old;
  asm volatile(
      "{\n"
      ".reg .b128 e, d, o;\n"
      "mov.b128 e, {%2, %3};\n"
      "mov.b128 d, {%4, %5};\n"
      "atom.global.relaxed.gpu.cas.b128 o, [%6], e, d;\n"
      "mov.b128 {%0, %1}, o;\n"
      "}\n"
      : "=l"(old.d), "=l"(old.k)
      : "l"(expected.d), "l"(expected.k), "l"(desired.d), "l"(desired.k), "l"(addr)
      : "memory");
  return old;
}
__device__ __forceinline__ DP dp_load(const PP& p, int j) {
  DP v;
  const u64* a = (const u64*)p.dist + 2 * (i64)j;
  v.d = __ldcg(a);
  v.k = __ldcg(a + 1);
  return v;  // a torn read only costs one failed CAS
}
// returns true iff the travel time strictly improved
__device__ __forceinline__ bool dp_update(const PP& p, int j, DP cur, double delta, u64 key) {
  const u64 db = (u64)__double_as_longlong(delta);
  DP* addr = (DP*)p.dist + j;
  while (true) {
    const bool strictly = db < cur.d;
    const bool tie = db == cur.d && !(key & KEY_ZMASK) && key < cur.k;
    if (!strictly && !tie) return false;
    DP des;
    des.d = db;
    des.k = key;
    const DP old = dp_cas(addr, cur, des);
    if (old.d == cur.d && old.k == cur.k) return strictly;
    cur = old;
  }
}

// round step 1: take the released bits of every item of the near list
__global__ void prep_kernel(PP p, const i32* __restrict__ near_cur, int cur) {
  const i64 slot = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= (i64)p.counters[cur]) return;
  const int it = near_cur[slot];
  p.cur_mask[slot] = atomicExch(&p.pend_mask[it], 0u);
}

// round step 2: one CTA per released item; warps split the elements of its G column; lanes = targets.
// All state that another SM may have written in an earlier phase of the SAME launch (persistent kernel) is read
// with ld.global.cg (L2) so that a stale L1 line can never hide an update; mesh arrays are read-only.
template <bool PACKED>
__device__ __forceinline__ void push2d_body_t(const PP& p, const i32* near_cur, int cur, i32* near_next,
                                              i32* far_list, int fcur) {
  __shared__ double2 sxz[32], sUd[32];  // (x, z) and (U, dist) of the released sources: one LDS.128 each
  __shared__ int s_id[32];
  __shared__ int s_ns;
  __shared__ double s_dmin;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const i64 n_near = (i64)__ldcg(&p.counters[cur]);
  const double tau = __ldcg(&p.tau[0]);
  u64 evals = 0;
  // work unit = (near-list slot, element group gy): the column of one released item is spread over PUSH_GY
  // CTAs x 4 warps (warp per element), so a round with few released items still fills the machine
  for (i64 unit = blockIdx.x; unit < n_near * PUSH_GY; unit += gridDim.x) {
    const i64 slot = unit / PUSH_GY;
    const int gy = (int)(unit - slot * PUSH_GY);
    const int it = __ldcg(&near_cur[slot]);
    const int v0 = p.item_first[it];
    const i64 c0 = p.g_off[v0], c1 = p.g_off[v0 + 1];
    if (c0 + (i64)gy * nwarp >= c1 && gy != 0) continue;  // block-uniform: this group has no element
    const unsigned mask = __ldcg(&p.cur_mask[slot]);
    __syncthreads();  // smem reuse across units
    if (warp == 0) {
      // compact the released sources into shared memory
      const bool on = (mask >> lane) & 1u;
      const int pos = __popc(mask & ((1u << lane) - 1u));
      if (on) {
        const int i = v0 + lane;
        sxz[pos] = make_double2(p.x[i], p.z[i]);
        sUd[pos] = make_double2(p.U[i], __ldcg(&p.dist[(i64)i * p.ds]));
        s_id[pos] = i;
      }
      double dm = on ? __ldcg(&p.dist[(i64)(v0 + lane) * p.ds]) : __longlong_as_double(0x7ff0000000000000LL);
      for (int o = 16; o; o >>= 1) dm = fmin(dm, __shfl_xor_sync(FULL, dm, o));
      if (lane == 0) {
        s_ns = __popc(mask);
        s_dmin = dm;
      }
    }
    __syncthreads();
    const int ns = s_ns;
    if (ns == 0) continue;
    const double dmin = s_dmin;
    // zero-weight halo coupling (update_halo!): lane s of the last warp of group 0 serves source s
    if (gy == 0 && warp == nwarp - 1 && p.n_hn > 0 && lane < ns && s_id[lane] != p.source) {
      const int lo = p.hn_index[s_id[lane]];
      if (lo >= 0) {
        const double d = sUd[lane].y;
        for (int q = p.hn_off[lo]; q < p.hn_off[lo + 1]; ++q) {
          const int b = p.hn_part[q];
          if (PACKED) {
            const DP cb = dp_load(p, b);
            if ((u64)__double_as_longlong(d) < cb.d && dp_update(p, b, cb, d, KEY_HALO | (u64)s_id[lane]))
              enqueue(p, b, d, tau, near_next, cur ^ 1, far_list, fcur);
          } else if (d < __ldcg(&p.dist[(i64)b * p.ds]) && relax_to(p, b, d)) {
            enqueue(p, b, d, tau, near_next, cur ^ 1, far_list, fcur);
          }
        }
      }
    }
    for (i64 c = c0 + (i64)gy * nwarp + warp; c < c1; c += (i64)PUSH_GY * nwarp) {
      const int el = p.g_idx[c];
      const int s = p.e2n_off[el];
      const int m = p.e2n_off[el + 1] - s;
      // software-pipelined target loop: the gathers of the next target are in flight while this one is evaluated
      int k = lane;
      int j = k < m ? p.e2n_idx[s + k] : -1;
      double dj = 0.0, xj = 0.0, zj = 0.0, Uj = 0.0;
      u64 kj = KEY_NONE, kjn = KEY_NONE;
      if (j >= 0) {
        if (PACKED) kj = __ldcg(p.keys + 2 * (i64)j + 1);
        dj = __ldcg(&p.dist[(i64)j * p.ds]);
        xj = p.x[j];
        zj = p.z[j];
        Uj = p.U[j];
      }
      while (k < m) {
        const int kn = k + 32;
        const int jn = kn < m ? p.e2n_idx[s + kn] : -1;
        double djn = 0.0, xjn = 0.0, zjn = 0.0, Ujn = 0.0;
        if (jn >= 0) {
          if (PACKED) kjn = __ldcg(p.keys + 2 * (i64)jn + 1);
          djn = __ldcg(&p.dist[(i64)jn * p.ds]);
          xjn = p.x[jn];
          zjn = p.z[jn];
          Ujn = p.U[jn];
        }
        if (dmin < dj) {  // else every released source is at or behind this target: nothing can improve
          double best = dj;
          u64 bkey = kj;
          bool changed = false;
          for (int q = 0; q < ns; ++q) {
            const double2 ud = sUd[q];  // (U, dist) of source q
            const double di = ud.y;
            if (!(di < best)) continue;  // di + w >= di >= best: cannot improve
            const double2 xz = sxz[q];
            {
              const double dx = __dsub_rn(xz.x, xj), dz = __dsub_rn(xz.y, zj);
              const double d2 = __fma_rn(dx, dx, dz * dz);
              if (screen_cannot_improve(best, di, d2, __dadd_rn(ud.x, Uj))) continue;
            }
            const double delta = edge_delta(di, xz.x, xz.y, ud.x, xj, zj, Uj);
            if (PACKED) {
              // delta == di means the weight was absorbed (coincident nodes): a zero-weight coupling
              const u64 key = (delta == di ? KEY_ZW : 0ull) | (u64)s_id[q];
              if (delta < best) {
                best = delta;
                bkey = key;
                changed = true;
              } else if (delta == best && !(key & KEY_ZMASK) && key < bkey) {
                bkey = key;
                changed = true;
              }
            } else {
              best = delta < best ? delta : best;
            }
          }
          if (PACKED) {
            if (changed) {
              DP cur_dp;
              cur_dp.d = (u64)__double_as_longlong(dj);
              cur_dp.k = kj;
              if (dp_update(p, j, cur_dp, best, bkey)) enqueue(p, j, best, tau, near_next, cur ^ 1, far_list, fcur);
            }
          } else if (best < dj && relax_to(p, j, best)) {
            enqueue(p, j, best, tau, near_next, cur ^ 1, far_list, fcur);
          }
        }
        k = kn;
        j = jn;
        kj = kjn;
        dj = djn;
        xj = xjn;
        zj = zjn;
        Uj = Ujn;
      }
      if (lane == 0) evals += (u64)m * (u64)ns;
    }
    if (gy == 0 && threadIdx.x == 0) atomicAdd(&p.counters[3], (u64)ns);
  }
  if (lane == 0 && evals) atomicAdd(&p.counters[2], evals);
}
// ---------------------------------------------------------------------------------------------------------
// Warp-level push for meshes whose G columns are short (coarse spacing: a column holds a few hundred candidates):
// one warp owns one released item, no block barrier.  The element offsets of the column are fetched by the lanes
// in parallel and prefix-summed, so the targets of ALL its elements form one flat index space that the lanes walk
// with full utilisation (an element holds ~10 nodes there); sources sit in a warp-private shared-memory slab.
template <bool PACKED, int MODE, bool COUNT = false>
__device__ __forceinline__ void push2d_warp_unit(const PP& p, int it, unsigned mask, int cur, i32* near_next,
                                                 i32* far_list, int fcur, double tau, double2* sxz, double2* sUd, double2* sU2r,
                                                 int* s_id, int* s_pre, int* s_start) {
  constexpr bool DUAL = MODE == MODE_DUAL, F32 = MODE == MODE_F32;
  const int lane = threadIdx.x & 31;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const int v0 = p.item_first[it];
  __syncwarp();
  const bool on = (mask >> lane) & 1u;
  const int pos = __popc(mask & ((1u << lane) - 1u));
  double dmy = INF;
  if (on) {
    const int i = v0 + lane;
    dmy = __ldcg(&p.dist[(i64)i * p.ds]);
    sxz[pos] = make_double2(p.x[i], p.z[i]);
    sUd[pos] = make_double2(DUAL ? p.U1[i] : p.U[i], dmy);
    if (DUAL) sU2r[pos] = make_double2(p.U2[i], p.r[i]);
    s_id[pos] = i;
  }
  double dmin = dmy;
  for (int o = 16; o; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(FULL, dmin, o));
  const int ns = __popc(mask);
  __syncwarp();
  // zero-weight halo coupling
  if (p.n_hn > 0 && lane < ns && s_id[lane] != p.source) {
    const int lo = p.hn_index[s_id[lane]];
    if (lo >= 0) {
      const double d = sUd[lane].y;
      for (int q = p.hn_off[lo]; q < p.hn_off[lo + 1]; ++q) {
        const int b = p.hn_part[q];
        if (PACKED) {
          const DP cb = dp_load(p, b);
          if ((u64)__double_as_longlong(d) < cb.d && dp_update(p, b, cb, d, KEY_HALO | (u64)s_id[lane]))
            enqueue(p, b, d, tau, near_next, cur ^ 1, far_list, fcur);
        } else if (d < __ldcg(&p.dist[(i64)b * p.ds]) && relax_to(p, b, d)) {
          enqueue(p, b, d, tau, near_next, cur ^ 1, far_list, fcur);
        }
      }
    }
  }
  u64 evals = 0;
  unsigned n_scr = 0, n_ex = 0;  // COUNT: candidates that reached the screen / the exact evaluation (this lane)
  const i64 c0 = p.g_off[v0], c1 = p.g_off[v0 + 1];
  for (i64 cb0 = c0; cb0 < c1; cb0 += 32) {  // element batches of 32 (a column has <= 16 elements, the centre 2T)
    const int ne = (int)min((i64)32, c1 - cb0);
    int s_l = 0, m_l = 0;
    if (lane < ne) {
      const int el = p.g_idx[cb0 + lane];
      s_l = p.e2n_off[el];
      m_l = p.e2n_off[el + 1] - s_l;
    }
    int incl = m_l;  // inclusive prefix sum of the element lengths
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += up;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    __syncwarp();
    s_pre[lane] = incl;
    s_start[lane] = s_l;
    __syncwarp();
    evals += (u64)total * (u64)ns;
    for (int t = lane; t < total; t += 32) {
      // element e with pre[e-1] <= t < pre[e]
      int lo = 0, hi = ne - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_pre[mid] <= t)
          lo = mid + 1;
        else
          hi = mid;
      }
      const int k = t - (lo ? s_pre[lo - 1] : 0);
      const int j = p.e2n_idx[s_start[lo] + k];
      const double dj = __ldcg(&p.dist[(i64)j * p.ds]);
      if (!(dmin < dj)) continue;
      u64 kj = KEY_NONE;
      if (PACKED) kj = __ldcg(p.keys + 2 * (i64)j + 1);
      const double xj = p.x[j], zj = p.z[j];
      const double Uj = DUAL ? p.U1[j] : p.U[j];
      const double U2j = DUAL ? p.U2[j] : 0.0, rj = DUAL ? p.r[j] : 0.0;
      double best = dj;
      u64 bkey = kj;
      bool changed = false;
      for (int q = 0; q < ns; ++q) {
        const double2 ud = sUd[q];
        const double di = ud.y;
        if (!(di < best)) continue;
        if (COUNT) ++n_scr;
        const double2 xz = sxz[q];
        double us = ud.x, ut = Uj;  // velocities of the source (head) and of the target (tail)
        if (DUAL) {
          const double2 u2r = sU2r[q];
          const bool down = rj > u2r.y;  // head_idx = (r_i > r_Gi) + 1 with i = target, Gi = source
          ut = down ? Uj : U2j;
          us = down ? u2r.x : ud.x;
        }
        {
          const double dx = __dsub_rn(xz.x, xj), dz = __dsub_rn(xz.y, zj);
          const double d2 = __fma_rn(dx, dx, dz * dz);
          if (screen_cannot_improve_t<F32>(best, di, d2, __dadd_rn(ut, us))) continue;
        }
        if (COUNT) ++n_ex;
        const double delta = edge_delta<F32>(di, xz.x, xz.y, us, xj, zj, ut);
        if (PACKED) {
          const u64 key = (delta == di ? KEY_ZW : 0ull) | (u64)s_id[q];
          if (delta < best) {
            best = delta;
            bkey = key;
            changed = true;
          } else if (delta == best && !(key & KEY_ZMASK) && key < bkey) {
            bkey = key;
            changed = true;
          }
        } else {
          best = delta < best ? delta : best;
        }
      }
      if (PACKED) {
        if (changed) {
          DP cur_dp;
          cur_dp.d = (u64)__double_as_longlong(dj);
          cur_dp.k = kj;
          if (dp_update(p, j, cur_dp, best, bkey)) enqueue(p, j, best, tau, near_next, cur ^ 1, far_list, fcur);
        }
      } else if (best < dj && relax_to(p, j, best)) {
        enqueue(p, j, best, tau, near_next, cur ^ 1, far_list, fcur);
      }
    }
    __syncwarp();
  }
  if (COUNT) {
    for (int o = 16; o; o >>= 1) {
      n_scr += __shfl_xor_sync(FULL, n_scr, o);
      n_ex += __shfl_xor_sync(FULL, n_ex, o);
    }
    if (lane == 0) {
      atomicAdd(&p.counters[6], (u64)n_scr);
      atomicAdd(&p.counters[7], (u64)n_ex);
    }
  }
  if (lane == 0) {
    atomicAdd(&p.counters[2], evals);
    atomicAdd(&p.counters[3], (u64)ns);
  }
}

// The same unit over the item's DE-DUPLICATED target list (PP::tgt_off): no element offsets, no prefix sums, every
// node of the patch is visited once; the id of the next round and the gathers that hang on it are in flight while this
// round's targets are evaluated.
template <bool PACKED, int MODE>
__device__ __forceinline__ void push2d_tgt_unit(const PP& p, int it, unsigned mask, i64 t0, i64 t1, int cur, i32* near_next,
                                                i32* far_list, int fcur, double tau, double2* sxz, double2* sUd,
                                                double2* sU2r, int* s_id) {
  constexpr bool DUAL = MODE == MODE_DUAL, F32 = MODE == MODE_F32;
  constexpr bool COUNT = false;
  const int lane = threadIdx.x & 31;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const int v0 = p.item_first[it];
  __syncwarp();
  const bool on = (mask >> lane) & 1u;
  const int pos = __popc(mask & ((1u << lane) - 1u));
  double dmy = INF;
  if (on) {
    const int i = v0 + lane;
    dmy = __ldcg(&p.dist[(i64)i * p.ds]);
    sxz[pos] = make_double2(p.x[i], p.z[i]);
    sUd[pos] = make_double2(DUAL ? p.U1[i] : p.U[i], dmy);
    if (DUAL) sU2r[pos] = make_double2(p.U2[i], p.r[i]);
    s_id[pos] = i;
  }
  double dmin = dmy;
  for (int o = 16; o; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(FULL, dmin, o));
  const int ns = __popc(mask);
  __syncwarp();
  if (p.n_hn > 0 && lane < ns && s_id[lane] != p.source) {  // zero-weight halo coupling
    const int lo = p.hn_index[s_id[lane]];
    if (lo >= 0) {
      const double d = sUd[lane].y;
      for (int q = p.hn_off[lo]; q < p.hn_off[lo + 1]; ++q) {
        const int b = p.hn_part[q];
        if (PACKED) {
          const DP cb = dp_load(p, b);
          if ((u64)__double_as_longlong(d) < cb.d && dp_update(p, b, cb, d, KEY_HALO | (u64)s_id[lane]))
            enqueue(p, b, d, tau, near_next, cur ^ 1, far_list, fcur);
        } else if (d < __ldcg(&p.dist[(i64)b * p.ds]) && relax_to(p, b, d)) {
          enqueue(p, b, d, tau, near_next, cur ^ 1, far_list, fcur);
        }
      }
    }
  }
  u64 evals = 0;
  unsigned n_scr = 0, n_ex = 0;
  {
    // de-duplicated targets of the item, one coalesced id load per 32 targets; the id of the next round and the
    // gathers that hang on it are issued before this round's targets are evaluated
    const int len = (int)(t1 - t0);
    const i32* __restrict__ tg = p.tgt_idx + t0;
    evals = (u64)len * (u64)ns;
    int j = lane < len ? tg[lane] : -1;
    double dj = 0.0, xj = 0.0, zj = 0.0, Uj = 0.0, U2j = 0.0, rj = 0.0;
    u64 kj = KEY_NONE;
    if (j >= 0) {
      if (PACKED) kj = __ldcg(p.keys + 2 * (i64)j + 1);
      dj = __ldcg(&p.dist[(i64)j * p.ds]);
      xj = p.x[j];
      zj = p.z[j];
      Uj = DUAL ? p.U1[j] : p.U[j];
      if (DUAL) {
        U2j = p.U2[j];
        rj = p.r[j];
      }
    }
    for (int t = lane; t < len; t += 32) {
      const int tn = t + 32;
      const int jn = tn < len ? tg[tn] : -1;
      double djn = 0.0, xjn = 0.0, zjn = 0.0, Ujn = 0.0, U2jn = 0.0, rjn = 0.0;
      u64 kjn = KEY_NONE;
      if (jn >= 0) {
        if (PACKED) kjn = __ldcg(p.keys + 2 * (i64)jn + 1);
        djn = __ldcg(&p.dist[(i64)jn * p.ds]);
        xjn = p.x[jn];
        zjn = p.z[jn];
        Ujn = DUAL ? p.U1[jn] : p.U[jn];
        if (DUAL) {
          U2jn = p.U2[jn];
          rjn = p.r[jn];
        }
      }
      if (dmin < dj) {  // (items of a coarse mesh release one or two sources: no group screen here)
        double best = dj;
        u64 bkey = kj;
        bool changed = false;
        for (int q = 0; q < ns; ++q) {
          const double2 ud = sUd[q];
          const double di = ud.y;
          if (!(di < best)) continue;
          if (COUNT) ++n_scr;
          const double2 xz = sxz[q];
          double us = ud.x, ut = Uj;
          if (DUAL) {
            const double2 u2r = sU2r[q];
            const bool down = rj > u2r.y;
            ut = down ? Uj : U2j;
            us = down ? u2r.x : ud.x;
          }
          {
            const double dx = __dsub_rn(xz.x, xj), dz = __dsub_rn(xz.y, zj);
            const double d2 = __fma_rn(dx, dx, dz * dz);
            if (screen_cannot_improve_t<F32>(best, di, d2, __dadd_rn(ut, us))) continue;
          }
          if (COUNT) ++n_ex;
          const double delta = edge_delta<F32>(di, xz.x, xz.y, us, xj, zj, ut);
          if (PACKED) {
            const u64 key = (delta == di ? KEY_ZW : 0ull) | (u64)s_id[q];
            if (delta < best) {
              best = delta;
              bkey = key;
              changed = true;
            } else if (delta == best && !(key & KEY_ZMASK) && key < bkey) {
              bkey = key;
              changed = true;
            }
          } else {
            best = delta < best ? delta : best;
          }
        }
        if (PACKED) {
          if (changed) {
            DP cur_dp;
            cur_dp.d = (u64)__double_as_longlong(dj);
            cur_dp.k = kj;
            if (dp_update(p, j, cur_dp, best, bkey)) enqueue(p, j, best, tau, near_next, cur ^ 1, far_list, fcur);
          }
        } else if (best < dj && relax_to(p, j, best)) {
          enqueue(p, j, best, tau, near_next, cur ^ 1, far_list, fcur);
        }
      }
      j = jn;
      kj = kjn;
      dj = djn;
      xj = xjn;
      zj = zjn;
      Uj = Ujn;
      U2j = U2jn;
      rj = rjn;
    }
  }
  (void)n_scr;
  (void)n_ex;
  if (lane == 0) {
    atomicAdd(&p.counters[2], evals);
    atomicAdd(&p.counters[3], (u64)ns);
  }
}

// warp-level units of ONE source: slots of its near list
template <int MODE, bool COUNT = false>
__device__ __forceinline__ void push2d_warp_body(const PP& p, const i32* near_cur, int cur, i32* near_next,
                                                 i32* far_list, int fcur, i64 first_warp, i64 n_warps) {
  constexpr bool DUAL = MODE == MODE_DUAL;
  __shared__ double2 w_sxz[PUSH_BLOCK / 32][32], w_sUd[PUSH_BLOCK / 32][32], w_sU2r[DUAL ? PUSH_BLOCK / 32 : 1][32];
  __shared__ int w_id[PUSH_BLOCK / 32][32], w_pre[PUSH_BLOCK / 32][32], w_start[PUSH_BLOCK / 32][32];
  const int warp = threadIdx.x >> 5;
  const i64 n_near = (i64)__ldcg(&p.counters[cur]);
  const double tau = __ldcg(&p.tau[0]);
  for (i64 slot = first_warp; slot < n_near; slot += n_warps) {
    const int it = __ldcg(&near_cur[slot]);
    const unsigned mask = __ldcg(&p.cur_mask[slot]);
    if (mask == 0u) continue;
    if (p.ds == 2)
      push2d_warp_unit<true, MODE, COUNT>(p, it, mask, cur, near_next, far_list, fcur, tau, w_sxz[warp], w_sUd[warp],
                                          w_sU2r[DUAL ? warp : 0], w_id[warp], w_pre[warp], w_start[warp]);
    else
      push2d_warp_unit<false, MODE, COUNT>(p, it, mask, cur, near_next, far_list, fcur, tau, w_sxz[warp], w_sUd[warp],
                                           w_sU2r[DUAL ? warp : 0], w_id[warp], w_pre[warp], w_start[warp]);
  }
}

// One target against all released sources of an item (sources in the warp-private shared-memory slab).
template <bool PACKED, int MODE, bool COUNT>
__device__ __forceinline__ void push2d_eval_target(const PP& p, int ns, const double2* sxz, const double2* sUd,
                                                   const double2* sU2r, const int* s_id, int j, double dj, u64 kj, double xj,
                                                   double zj, double Uj, double U2j, double rj, double tau, i32* near_next,
                                                   int cur, i32* far_list, int fcur, unsigned& n_scr, unsigned& n_ex) {
  constexpr bool DUAL = MODE == MODE_DUAL, F32 = MODE == MODE_F32;
  double best = dj;
  u64 bkey = kj;
  bool changed = false;
  for (int q = 0; q < ns; ++q) {
    const double2 ud = sUd[q];
    const double di = ud.y;
    if (!(di < best)) continue;
    if (COUNT) ++n_scr;
    const double2 xz = sxz[q];
    double us = ud.x, ut = Uj;  // velocities of the source (head) and of the target (tail)
    if (DUAL) {
      const double2 u2r = sU2r[q];
      const bool down = rj > u2r.y;  // head_idx = (r_i > r_Gi) + 1 with i = target, Gi = source
      ut = down ? Uj : U2j;
      us = down ? u2r.x : ud.x;
    }
    {
      const double dx = __dsub_rn(xz.x, xj), dz = __dsub_rn(xz.y, zj);
      const double d2 = __fma_rn(dx, dx, dz * dz);
      if (screen_cannot_improve_t<F32>(best, di, d2, __dadd_rn(ut, us))) continue;
    }
    if (COUNT) ++n_ex;
    const double delta = edge_delta<F32>(di, xz.x, xz.y, us, xj, zj, ut);
    if (PACKED) {
      const u64 key = (delta == di ? KEY_ZW : 0ull) | (u64)s_id[q];
      if (delta < best) {
        best = delta;
        bkey = key;
        changed = true;
      } else if (delta == best && !(key & KEY_ZMASK) && key < bkey) {
        bkey = key;
        changed = true;
      }
    } else {
      best = delta < best ? delta : best;
    }
  }
  if (PACKED) {
    if (changed) {
      DP cur_dp;
      cur_dp.d = (u64)__double_as_longlong(dj);
      cur_dp.k = kj;
      if (dp_update(p, j, cur_dp, best, bkey)) enqueue(p, j, best, tau, near_next, cur ^ 1, far_list, fcur);
    }
  } else if (best < dj && relax_to(p, j, best)) {
    enqueue(p, j, best, tau, near_next, cur ^ 1, far_list, fcur);
  }
}

// Warp-private buffer in which the lanes park the targets that survive the group screen; the source loop only runs on
// full warps of survivors (lanes whose target was screened out would otherwise idle through the whole loop).
struct LiveBuf {
  double2 a[64];  // (dist, x) of the target
  double2 b[64];  // (z, U)
  double2 c[64];  // dual velocity: (U2, r)
  u64 k[64];      // predecessor key
  int j[64];
};

// ---------------------------------------------------------------------------------------------------------
// Barrier-free push for long columns: warp-level unit = (released item, element lane e): the warp keeps its own copy
// of the released sources (warp-private smem slab) and walks the elements e, e + PUSH_GE, ... of the column with the
// software-pipelined target loop.  (The CTA-level variant above spends a third of its stall time in __syncthreads.)
constexpr int PUSH_GE = 16;
#ifndef RT_FAR_EVERY
#define RT_FAR_EVERY 3
#endif
constexpr int FAR_EVERY = RT_FAR_EVERY;  // the threshold-advance kernels are enqueued every FAR_EVERY-th round
#ifndef RT_PUSH_SPLIT
#define RT_PUSH_SPLIT 32
#endif
#ifndef RT_PUSH_FILL
#define RT_PUSH_FILL (148 * 96)
#endif
constexpr int PUSH_SPLIT = RT_PUSH_SPLIT;     // max warps per (item, element) in sparse rounds
constexpr int PUSH_FILL = RT_PUSH_FILL;  // warps the split aims to create (4 waves of resident warps; measured optimum is flat)
template <bool PACKED, int MODE, bool COUNT = false>
__device__ __forceinline__ void push2d_elem_unit(const PP& p, int it, unsigned mask, int e0, int sub, int kst, int cur,
                                                 i32* near_next,
                                                 i32* far_list, int fcur, double tau, double2* sxz, double2* sUd, double2* sU2r,
                                                 int* s_id, LiveBuf* lb) {
  constexpr bool DUAL = MODE == MODE_DUAL, F32 = MODE == MODE_F32;
  const int lane = threadIdx.x & 31;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const int v0 = p.item_first[it];
  const i64 c0 = p.g_off[v0], c1 = p.g_off[v0 + 1];
  if (c0 + e0 >= c1 && e0 != 0) return;  // warp-uniform: no element for this lane of the column
  __syncwarp();
  const bool on = (mask >> lane) & 1u;
  const int pos = __popc(mask & ((1u << lane) - 1u));
  double dmy = INF, gx = 0.0, gz = 0.0, gu = 0.0;
  if (on) {
    const int i = v0 + lane;
    dmy = __ldcg(&p.dist[(i64)i * p.ds]);
    gx = p.x[i];
    gz = p.z[i];
    gu = DUAL ? p.U1[i] : p.U[i];
    sxz[pos] = make_double2(gx, gz);
    sUd[pos] = make_double2(gu, dmy);
    if (DUAL) {
      const double u2 = p.U2[i];
      sU2r[pos] = make_double2(u2, p.r[i]);
      gu = fmax(gu, u2);
    }
    s_id[pos] = i;
  }
  double dmin = dmy;
  for (int o = 16; o; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(FULL, dmin, o));
  const GroupDisc gd = group_disc(on, gx, gz, gu);
  const int ns = __popc(mask);
  __syncwarp();
  if (e0 == 0 && sub == 0 && p.n_hn > 0 && lane < ns && s_id[lane] != p.source) {  // zero-weight halo coupling
    const int lo = p.hn_index[s_id[lane]];
    if (lo >= 0) {
      const double d = sUd[lane].y;
      for (int q = p.hn_off[lo]; q < p.hn_off[lo + 1]; ++q) {
        const int b = p.hn_part[q];
        if (PACKED) {
          const DP cb = dp_load(p, b);
          if ((u64)__double_as_longlong(d) < cb.d && dp_update(p, b, cb, d, KEY_HALO | (u64)s_id[lane]))
            enqueue(p, b, d, tau, near_next, cur ^ 1, far_list, fcur);
        } else if (d < __ldcg(&p.dist[(i64)b * p.ds]) && relax_to(p, b, d)) {
          enqueue(p, b, d, tau, near_next, cur ^ 1, far_list, fcur);
        }
      }
    }
  }
  u64 evals = 0;
  unsigned n_scr = 0, n_ex = 0, n_grp = 0;  // COUNT: candidates at the screen / exact evaluation; targets cut by the group screen
  const bool compact = p.compact != 0;
  int nlive = 0;  // survivors parked in the warp's buffer (warp-uniform)
  for (i64 c = c0 + e0; c < c1; c += PUSH_GE) {
    const int el = p.g_idx[c];
    const int s = p.e2n_off[el];
    const int m = p.e2n_off[el + 1] - s;
    // the target ids run two iterations ahead of the evaluation and the gathers one iteration ahead, so neither the
    // id load nor the gathers that depend on it are waited for in the iteration that issues them
    // (sub, kst): when a round releases few items the targets of an element are split over kst/32 warps
    int k = lane + 32 * sub;
    int j = k < m ? p.e2n_idx[s + k] : -1;
    int jn = k + kst < m ? p.e2n_idx[s + k + kst] : -1;
    double dj = 0.0, xj = 0.0, zj = 0.0, Uj = 0.0, U2j = 0.0, rj = 0.0;
    u64 kj = KEY_NONE, kjn = KEY_NONE;
    if (j >= 0) {
      if (PACKED) kj = __ldcg(p.keys + 2 * (i64)j + 1);
      dj = __ldcg(&p.dist[(i64)j * p.ds]);
      xj = p.x[j];
      zj = p.z[j];
      Uj = DUAL ? p.U1[j] : p.U[j];
      if (DUAL) {
        U2j = p.U2[j];
        rj = p.r[j];
      }
    }
    while (k - lane < m) {  // warp-uniform trip count (the survivors are exchanged with warp-wide primitives)
      const int kn = k + kst;
      const int jnn = kn + kst < m ? p.e2n_idx[s + kn + kst] : -1;
      double djn = 0.0, xjn = 0.0, zjn = 0.0, Ujn = 0.0, U2jn = 0.0, rjn = 0.0;
      if (jn >= 0) {
        if (PACKED) kjn = __ldcg(p.keys + 2 * (i64)jn + 1);
        djn = __ldcg(&p.dist[(i64)jn * p.ds]);
        xjn = p.x[jn];
        zjn = p.z[jn];
        Ujn = DUAL ? p.U1[jn] : p.U[jn];
        if (DUAL) {
          U2jn = p.U2[jn];
          rjn = p.r[jn];
        }
      }
      bool live = j >= 0 && dmin < dj;
      if (live && p.group_screen) {  // no source of the item can reach this target in time: skip the whole source loop
        const double dxc = __dsub_rn(xj, gd.cx), dzc = __dsub_rn(zj, gd.cz);
        live = !group_cannot_improve_t<F32>(dj, dmin, __fma_rn(dxc, dxc, dzc * dzc), gd.rho,
                                            __dadd_rn(DUAL ? fmax(Uj, U2j) : Uj, gd.umax));
        if (COUNT && !live) ++n_grp;
      }
      if (compact) {
        const unsigned ball = __ballot_sync(FULL, live);
        if (ball) {
          if (live) {
            const int w = nlive + __popc(ball & ((1u << lane) - 1u));
            lb->a[w] = make_double2(dj, xj);
            lb->b[w] = make_double2(zj, Uj);
            if (DUAL) lb->c[w] = make_double2(U2j, rj);
            lb->k[w] = kj;
            lb->j[w] = j;
          }
          nlive += __popc(ball);
          __syncwarp();
          if (nlive >= 32) {  // a full warp of survivors: evaluate the last 32
            const int w = nlive - 32 + lane;
            const double2 ta = lb->a[w], tb = lb->b[w];
            const double2 tc = DUAL ? lb->c[w] : make_double2(0.0, 0.0);
            const u64 tk = lb->k[w];
            const int tj = lb->j[w];
            __syncwarp();
            nlive -= 32;
            push2d_eval_target<PACKED, MODE, COUNT>(p, ns, sxz, sUd, sU2r, s_id, tj, ta.x, tk, ta.y, tb.x, tb.y, tc.x, tc.y, tau,
                                                    near_next, cur, far_list, fcur, n_scr, n_ex);
          }
        }
      } else if (live) {
        push2d_eval_target<PACKED, MODE, COUNT>(p, ns, sxz, sUd, sU2r, s_id, j, dj, kj, xj, zj, Uj, U2j, rj, tau, near_next, cur,
                                                far_list, fcur, n_scr, n_ex);
      }
      k = kn;
      j = jn;
      jn = jnn;
      kj = kjn;
      dj = djn;
      xj = xjn;
      zj = zjn;
      Uj = Ujn;
      U2j = U2jn;
      rj = rjn;
    }
    if (sub == 0) evals += (u64)m * (u64)ns;
  }
  if (nlive > 0) {  // the last, partial warp of survivors
    if (lane < nlive) {
      const double2 ta = lb->a[lane], tb = lb->b[lane];
      const double2 tc = DUAL ? lb->c[lane] : make_double2(0.0, 0.0);
      push2d_eval_target<PACKED, MODE, COUNT>(p, ns, sxz, sUd, sU2r, s_id, lb->j[lane], ta.x, lb->k[lane], ta.y, tb.x, tb.y, tc.x,
                                              tc.y, tau, near_next, cur, far_list, fcur, n_scr, n_ex);
    }
    __syncwarp();
  }
  if (COUNT) {
    for (int o = 16; o; o >>= 1) {
      n_scr += __shfl_xor_sync(FULL, n_scr, o);
      n_ex += __shfl_xor_sync(FULL, n_ex, o);
    }
    if (lane == 0) {
      atomicAdd(&p.counters[6], (u64)n_scr);
      atomicAdd(&p.counters[7], (u64)n_ex);
    }
  }
  if (lane == 0) {
    if (evals) atomicAdd(&p.counters[2], evals);
    if (e0 == 0 && sub == 0) atomicAdd(&p.counters[3], (u64)ns);
  }
}
template <int MODE, bool COUNT = false>
__device__ __forceinline__ void push2d_elem_body(const PP& p, const i32* near_cur, int cur, i32* near_next,
                                                 i32* far_list, int fcur, i64 first_warp, i64 n_warps) {
  constexpr bool DUAL = MODE == MODE_DUAL;
  __shared__ double2 e_sxz[PUSH_BLOCK / 32][32], e_sUd[PUSH_BLOCK / 32][32], e_sU2r[DUAL ? PUSH_BLOCK / 32 : 1][32];
  __shared__ int e_id[PUSH_BLOCK / 32][32];
  __shared__ LiveBuf e_lb[PUSH_BLOCK / 32];
  const int warp = threadIdx.x >> 5;
  const i64 n_near = (i64)__ldcg(&p.counters[cur]);
  const double tau = __ldcg(&p.tau[0]);
  // a round that releases few items would leave most of the machine idle while single warps walk whole elements
  // (tens of dependent gather rounds each): split every element over up to PUSH_SPLIT warps then
  const i64 base = n_near * PUSH_GE;
  const int split = (int)max((i64)1, min((i64)PUSH_SPLIT, (i64)PUSH_FILL / max(base, (i64)1)));
  const int kst = 32 * split;
  for (i64 unit = first_warp; unit < base * split; unit += n_warps) {
    const i64 eu = unit / split;
    const int sub = (int)(unit - eu * split);
    const i64 slot = eu / PUSH_GE;
    const int e0 = (int)(eu - slot * PUSH_GE);
    const int it = __ldcg(&near_cur[slot]);
    const unsigned mask = __ldcg(&p.cur_mask[slot]);
    if (mask == 0u) continue;
    if (p.ds == 2)
      push2d_elem_unit<true, MODE, COUNT>(p, it, mask, e0, sub, kst, cur, near_next, far_list, fcur, tau, e_sxz[warp],
                                          e_sUd[warp], e_sU2r[DUAL ? warp : 0], e_id[warp], &e_lb[warp]);
    else
      push2d_elem_unit<false, MODE, COUNT>(p, it, mask, e0, sub, kst, cur, near_next, far_list, fcur, tau, e_sxz[warp],
                                           e_sUd[warp], e_sU2r[DUAL ? warp : 0], e_id[warp], &e_lb[warp]);
  }
}

template <bool WARP, int MODE, bool COUNT = false>
__device__ __forceinline__ void push2d_body(const PP& p, const i32* near_cur, int cur, i32* near_next,
                                            i32* far_list, int fcur) {
  if (WARP) {
    const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
    push2d_warp_body<MODE, COUNT>(p, near_cur, cur, near_next, far_list, fcur, gw, nw);
  } else if (p.cta_units) {
    if (p.ds == 2)
      push2d_body_t<true>(p, near_cur, cur, near_next, far_list, fcur);
    else
      push2d_body_t<false>(p, near_cur, cur, near_next, far_list, fcur);
  } else {
    const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
    push2d_elem_body<MODE, COUNT>(p, near_cur, cur, near_next, far_list, fcur, gw, nw);
  }
}
// host-driven rounds with profiling timers (single source): also counts the candidates that reach the screen and the
// exact evaluation (rt_stats.screened_edges / exact_edges) -- the counting costs a few percent, so only here
template <bool WARP, int MODE>
__global__ void __launch_bounds__(PUSH_BLOCK, 6) push2d_kernel(PP p, const i32* __restrict__ near_cur, int cur,
                                                           i32* __restrict__ near_next, i32* __restrict__ far_list,
                                                           int fcur) {
  push2d_body<WARP, MODE, true>(p, near_cur, cur, near_next, far_list, fcur);
}

// threshold advance, step 1: smallest waiting value
__global__ void far_min_kernel(PP p, const i32* __restrict__ far_cur, int fcur) {
  const i64 slot = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  u64 best = ~0ull;
  if (slot < (i64)p.counters[4 + fcur]) {
    const int it = far_cur[slot];
    const unsigned m = p.far_mask[it];
    if ((m >> lane) & 1u) best = (u64)__double_as_longlong(p.dist[(i64)(p.item_first[it] + lane) * p.ds]);
  }
  for (int o = 16; o; o >>= 1) {
    const u64 other = __shfl_xor_sync(FULL, best, o);
    best = other < best ? other : best;
  }
  if (lane == 0 && best != ~0ull) atomicMin((u64*)&p.tau[2], best);
}
// step 2: tau = min + delta; release the waiting nodes below it, keep the others in the new far list
__global__ void far_release_kernel(PP p, const i32* __restrict__ far_cur, int fcur, i32* __restrict__ far_next,
                                   i32* __restrict__ near_next, int nxt) {
  const i64 slot = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slot >= (i64)p.counters[4 + fcur]) return;
  const double tau = __dadd_rn(p.tau[2], p.tau[1]);
  const int it = far_cur[slot];
  const unsigned m = p.far_mask[it];
  const bool mine = (m >> lane) & 1u;
  const bool rel = mine && p.dist[(i64)(p.item_first[it] + lane) * p.ds] < tau;
  const unsigned relm = __ballot_sync(FULL, rel);
  if (lane == 0) {
    const unsigned keep = m & ~relm;
    p.far_mask[it] = keep;
    if (keep)
      far_next[atomicAdd(&p.counters[4 + (fcur ^ 1)], 1ull)] = it;
    else
      p.infar[it] = 0u;
    if (relm) {
      const unsigned old = atomicOr(&p.pend_mask[it], relm);
      if (old == 0u) near_next[atomicAdd(&p.counters[nxt], 1ull)] = it;
    }
    if (slot == 0) p.tau[0] = tau;
  }
}

__global__ void push_init_kernel(PP pb, i64 n, double delta) {
  const PP p = pp_view(pb, blockIdx.y);
  const int source = p.source;
  i32* near0 = nq(p, 0);
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    p.dist[i * p.ds] = (i == source) ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
    if (p.ds == 2) p.keys[2 * i + 1] = KEY_NONE;
    p.prev[i] = -1;
  }
  if (i == 0) {
    const int it = p.node_item[source];
    p.pend_mask[it] = 1u << (source - p.item_first[it]);
    near0[0] = it;
    for (int k = 0; k < 8; ++k) {
      p.counters[k] = k == 0 ? 1ull : 0ull;
      p.ctl[k] = 0;
    }
    p.tau[0] = delta;
    p.tau[1] = delta;
  }
}

// ------------------------------------------------------------------------------------------- predecessors
// pass 1: prev[i] = first candidate in scan order with dist[j] < dist[i] and dist[j] + w == dist[i] (bitwise).
// Same (target, part) lane layout as the Jacobi relax kernel; parts are merged on the scan position.
__global__ void __launch_bounds__(256) prev_tight_kernel(PP p, i64 n_items, int source, i32* __restrict__ unres) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  for (i64 it = (i64)blockIdx.x * wpb + (threadIdx.x >> 5); it < n_items; it += (i64)gridDim.x * wpb) {
    const int v0 = p.item_first[it];
    const int t = p.item_first[it + 1] - v0;
    int lg = 0;
    while ((1 << lg) < t) ++lg;
    const int tl = 1 << lg, parts = 32 >> lg;
    const int ti = lane & (tl - 1), part = lane >> lg;
    const int i = v0 + min(ti, t - 1);
    const double xi = p.x[i], zi = p.z[i], Ui = p.U[i], di = p.dist[i];
    const bool want = di < INF && i != source;
    int bpos = 0x7fffffff, bid = -1;
    if (__any_sync(FULL, want)) {
      int pos_base = 0;
      const i64 c0 = p.g_off[v0], c1 = p.g_off[v0 + 1];
      for (i64 c = c0; c < c1; ++c) {
        const int el = p.g_idx[c];
        const int s = p.e2n_off[el];
        const int m = p.e2n_off[el + 1] - s;
        for (int k = part; k < m; k += parts) {
          const int j = p.e2n_idx[s + k];
          const double dj = p.dist[j];
          if (want && bid < 0 && dj < di) {
            const double xj = p.x[j], zj = p.z[j], Uj = p.U[j];
            const double dx = __dsub_rn(xi, xj), dz = __dsub_rn(zi, zj);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dz, dz));
            if (!screen_maybe_tight(di, dj, d2, __dadd_rn(Ui, Uj))) continue;
            const double delta = edge_delta(dj, xi, zi, Ui, xj, zj, Uj);
            if (delta == di) {
              bpos = pos_base + k;
              bid = j;
            }
          }
        }
        pos_base += m;
        if (__all_sync(FULL, !want || bid >= 0)) break;  // every lane has its first tight candidate
      }
    }
    for (int off = 16; off >= tl; off >>= 1) {
      const int op = __shfl_xor_sync(FULL, bpos, off);
      const int oi = __shfl_xor_sync(FULL, bid, off);
      if (op < bpos) {
        bpos = op;
        bid = oi;
      }
    }
    if (part == 0 && ti < t && want) {
      if (bid >= 0) {
        p.prev[i] = bid;
      } else {
        p.prev[i] = -2;  // reached but unresolved (overrides the halo-init value until pass 2 settles it)
        unres[atomicAdd(&p.counters[6], 1ull)] = i;
      }
    }
  }
}

// pass 2 (iterated): nodes without a strictly-earlier tight predecessor owe their value to a zero-weight
// coupling: a halo twin (inherit ITS predecessor, as update_halo! does) or a coincident duplicate at equal
// travel time (first one in scan order that is already resolved).
__global__ void prev_resolve_kernel(PP p, const i32* __restrict__ unres, i64 n_unres, int source,
                                    i32* __restrict__ unres_next, int* __restrict__ pending_prev) {
  const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_unres) return;
  const int i = unres[q];
  const double di = p.dist[i];
  int found = -1;
  // (a) halo partners
  if (p.n_hn > 0) {
    int lo = 0, hi = p.n_hn;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (p.hn_node[mid] < i)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo < p.n_hn && p.hn_node[lo] == i)
      for (int r = p.hn_off[lo]; r < p.hn_off[lo + 1] && found < 0; ++r) {
        const int b = p.hn_part[r];
        if (p.dist[b] == di) {
          if (b == source)
            found = source;
          else if (p.prev[b] >= 0)
            found = p.prev[b];
        }
      }
  }
  // (b) equal-time tight candidates (zero-weight edges between coincident nodes), scan order
  if (found < 0) {
    const double xi = p.x[i], zi = p.z[i], Ui = p.U[i];
    for (i64 c = p.g_off[i]; c < p.g_off[i + 1] && found < 0; ++c) {
      const int el = p.g_idx[c];
      for (int k = p.e2n_off[el]; k < p.e2n_off[el + 1]; ++k) {
        const int j = p.e2n_idx[k];
        if (j == i || p.dist[j] != di) continue;
        if (j != source && p.prev[j] < 0) continue;
        if (edge_delta(di, xi, zi, Ui, p.x[j], p.z[j], p.U[j]) == di) {
          found = j;
          break;
        }
      }
    }
  }
  if (found >= 0)
    pending_prev[q] = found;  // applied after the kernel: all decisions of one iteration see the same state
  else
    pending_prev[q] = -1;
  (void)unres_next;
}
__global__ void prev_apply_kernel(PP p, const i32* __restrict__ unres, i64 n_unres,
                                  const int* __restrict__ pending_prev, i32* __restrict__ unres_next) {
  const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_unres) return;
  const int i = unres[q];
  if (pending_prev[q] >= 0)
    p.prev[i] = pending_prev[q];
  else
    unres_next[atomicAdd(&p.counters[7], 1ull)] = i;
}
__global__ void prev_giveup_kernel(PP p, const i32* __restrict__ unres, i64 n_unres) {
  const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n_unres) p.prev[unres[q]] = -1;
}
__global__ void prev_halo_init_kernel(PP pb, const i32* __restrict__ hnode, const i32* __restrict__ hval, i64 nh) {
  const PP p = pp_view(pb, blockIdx.y);
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nh) p.prev[hnode[k]] = hval[k];
}

// ---------------------------------------------------------------------------------------------------------
// Device-controlled rounds: the host enqueues a fixed sequence of launches per round and only synchronises every
// `check_every` rounds; which phase runs (push / threshold advance / nothing) is decided on the device.
// after_far: the far kernels were enqueued between the previous round_begin and this one.  They are only enqueued
// every FAR_EVERY-th round (in most rounds they have nothing to do), so a source whose threshold advance has been
// requested (mode 2) stays put -- prep / push return at once -- until a round_begin that follows them.
// The round control for nb <= 32 sources, executed by ONE full warp (lane = source).  Used by round_begin_kernel<32> and by
// round_tail below; counters / tau are read past the L1 (the tail runs at the end of a kernel whose other CTAs updated
// them with atomics).
__device__ __forceinline__ void round_begin_warp(const PP& pb, int after_far) {
  const int b = threadIdx.x & 31;
  i64 cn = 0, cf = 0;
  if (b < pb.nb) {
    const PP p = pp_view(pb, b);
    int* c = p.ctl;
    if (!c[3] && !(c[2] == 2 && !after_far)) {
      if (c[2] == 1)
        c[0] ^= 1;
      else if (c[2] == 2)
        c[1] ^= 1;
      const int cur = c[0], fcur = c[1];
      const u64 n_near = __ldcg(&p.counters[cur]), n_far = __ldcg(&p.counters[4 + fcur]);
      if (n_near == 0 && n_far == 0) {
        c[2] = 0;
        c[3] = 1;
      } else {
        c[4] += 1;
        if (n_near > 0) {
          c[2] = 1;
          c[5] += 1;
          p.counters[cur ^ 1] = 0;
        } else {
          c[2] = 2;
          p.tau[2] = __longlong_as_double(-1LL);
          p.counters[4 + (fcur ^ 1)] = 0;
        }
      }
    }
    if (!c[3]) {
      if (c[2] == 1) cn = (i64)__ldcg(&p.counters[c[0]]);
      if (c[2] == 2) cf = (i64)__ldcg(&p.counters[4 + c[1]]);
    }
  }
  i64 in = cn, fi = cf;
  for (int o = 1; o < 32; o <<= 1) {
    const i64 a = __shfl_up_sync(FULL, in, o), f = __shfl_up_sync(FULL, fi, o);
    if (b >= o) {
      in += a;
      fi += f;
    }
  }
  if (b < pb.nb) {
    pb.flat[b + 1] = in;
    pb.flat[FLAT_STRIDE + b + 1] = fi;
  }
  if (b == 0) {
    pb.flat[0] = 0;
    pb.flat[FLAT_STRIDE] = 0;
  }
}
// Round control folded into the tail of the kernel that precedes it (option fuse_begin, nb <= 32): the last CTA to finish
// runs round_begin for the coming round, which saves one launch per round.  tail < 0: nothing; otherwise tail = the
// after_far argument.  Every thread of every CTA must reach this call.  Measured (B200, graph replay): SLOWER than the
// one-warp kernel of its own -- config[1] 1.159 s against 1.136 s per source, annulus 180x50 @1 km 91.5 against 88.8 ms:
// the ticket atomics of ~2400 CTAs and the serial control at the very end of the push cost more than a launch inside
// a graph.  Kept behind the option (default off).
__device__ __forceinline__ void round_tail(const PP& pb, int tail) {
  if (tail < 0) return;
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(pb.ticket, 1u) == gridDim.x * gridDim.y - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32) round_begin_warp(pb, tail);
  if (threadIdx.x == 0) *pb.ticket = 0u;
}
template <int BS>
__global__ void __launch_bounds__(BS) round_begin_kernel(PP pb, int after_far) {
  const int b = threadIdx.x;
  i64 cn = 0, cf = 0;  // this source's near slots to push / far slots to advance in the coming round
  if (b < pb.nb) {
    const PP p = pp_view(pb, b);
    int* c = p.ctl;
    if (!c[3] && !(c[2] == 2 && !after_far)) {
      if (c[2] == 1)
        c[0] ^= 1;
      else if (c[2] == 2)
        c[1] ^= 1;
      const int cur = c[0], fcur = c[1];
      const u64 n_near = p.counters[cur], n_far = p.counters[4 + fcur];
      if (n_near == 0 && n_far == 0) {
        c[2] = 0;
        c[3] = 1;
      } else {
        c[4] += 1;
        if (n_near > 0) {
          c[2] = 1;
          c[5] += 1;
          p.counters[cur ^ 1] = 0;
        } else {
          c[2] = 2;
          p.tau[2] = __longlong_as_double(-1LL);
          p.counters[4 + (fcur ^ 1)] = 0;
        }
      }
    }
    if (!c[3]) {
      if (c[2] == 1) cn = (i64)p.counters[c[0]];
      if (c[2] == 2) cf = (i64)p.counters[4 + c[1]];
    }
  }
  if (BS == 32) {  // nb <= 32: warp scan
    i64 in = cn, fi = cf;
    for (int o = 1; o < 32; o <<= 1) {
      const i64 a = __shfl_up_sync(FULL, in, o), f = __shfl_up_sync(FULL, fi, o);
      if (b >= o) {
        in += a;
        fi += f;
      }
    }
    if (b < pb.nb) {
      pb.flat[b + 1] = in;
      pb.flat[FLAT_STRIDE + b + 1] = fi;
    }
    if (b == 0) {
      pb.flat[0] = 0;
      pb.flat[FLAT_STRIDE] = 0;
    }
  } else {
    __shared__ i64 sn[BS], sf[BS];
    sn[b] = cn;
    sf[b] = cf;
    __syncthreads();
    for (int o = 1; o < BS; o <<= 1) {  // Hillis-Steele inclusive scan
      const i64 a = b >= o ? sn[b - o] : 0, f = b >= o ? sf[b - o] : 0;
      __syncthreads();
      sn[b] += a;
      sf[b] += f;
      __syncthreads();
    }
    if (b < pb.nb) {
      pb.flat[b + 1] = sn[b];
      pb.flat[FLAT_STRIDE + b + 1] = sf[b];
    }
    if (b == 0) {
      pb.flat[0] = 0;
      pb.flat[FLAT_STRIDE] = 0;
    }
  }
}
__global__ void prep_dc_kernel(PP pb) {
  const int nb = pb.nb;
  const i64* base = pb.flat;
  const i64 total = base[nb];
  for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (i64)gridDim.x * blockDim.x) {
    const int b = flat_owner(base, nb, g);
    if (g < pb.flat_cap) pb.flat_b[g] = b;
    const i64 slot = g - base[b];
    const i64 o = (i64)b * pb.n_items;
    const i32* near_cur = (pb.ctl[b * 8] ? pb.nearq1 : pb.nearq0) + o;
    pb.cur_mask[o + slot] = atomicExch(&pb.pend_mask[o + near_cur[slot]], 0u);
  }
}
// resident CTAs per SM the register allocation aims at: 5 -> 96 registers (40 B of spills), 6 -> 80 (132 B), 4 -> 124
// (none).  Measured on config[1] / annulus 180x50 @1 km (B200, r02p): 5: 1.137 s / 89 ms, 6: 1.178 s / 97 ms, 4: 1.187 s /
// 91 ms, 7: 1.164 s, 8: 1.305 s, 3: 1.350 s / 99 ms.
#ifndef RT_PUSH_MINB
#define RT_PUSH_MINB 5
#endif
template <bool WARP, int MODE>
__global__ void __launch_bounds__(PUSH_BLOCK, RT_PUSH_MINB) push2d_dc_kernel(PP pb, int tail) {
  if (WARP) {
    // warp-level units of ALL sources in one flat index space (slot g of the concatenated near lists)
    constexpr bool DUAL = MODE == MODE_DUAL;
    __shared__ double2 w_sxz[PUSH_BLOCK / 32][32], w_sUd[PUSH_BLOCK / 32][32], w_sU2r[DUAL ? PUSH_BLOCK / 32 : 1][32];
    __shared__ int w_id[PUSH_BLOCK / 32][32], w_pre[PUSH_BLOCK / 32][32], w_start[PUSH_BLOCK / 32][32];
    const int warp = threadIdx.x >> 5;
    const int nb = pb.nb;
    const i64* base = pb.flat;
    const i64 total = base[nb];
    const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 g = gw; g < total; g += nw) {
      const int b = nb == 1 ? 0 : flat_owner(base, nb, g);
      const i64 slot = g - base[b];
      const PP p = pp_view(pb, b);
      const int cur = p.ctl[0], fcur = p.ctl[1];
      const int it = __ldcg(&nq(p, cur)[slot]);
      const unsigned mask = __ldcg(&p.cur_mask[slot]);
      if (mask == 0u) continue;
      if (p.tgt_off && p.tgt_off[it + 1] > p.tgt_off[it]) continue;  // served by push2d_tgt_dc_kernel
      const double tau = __ldcg(&p.tau[0]);
      if (p.ds == 2)
        push2d_warp_unit<true, MODE>(p, it, mask, cur, nq(p, cur ^ 1), fq(p, fcur), fcur, tau, w_sxz[warp], w_sUd[warp],
                                     w_sU2r[DUAL ? warp : 0], w_id[warp], w_pre[warp], w_start[warp]);
      else
        push2d_warp_unit<false, MODE>(p, it, mask, cur, nq(p, cur ^ 1), fq(p, fcur), fcur, tau, w_sxz[warp], w_sUd[warp],
                                      w_sU2r[DUAL ? warp : 0], w_id[warp], w_pre[warp], w_start[warp]);
    }
  } else {
    const PP p = pp_view(pb, blockIdx.y);
    if (p.ctl[2] == 1) {
      const int cur = p.ctl[0], fcur = p.ctl[1];
      push2d_body<WARP, MODE>(p, nq(p, cur), cur, nq(p, cur ^ 1), fq(p, fcur), fcur);
    }
  }
  round_tail(pb, tail);
}
// rare path of the kernel below, kept out of line so that it does not cost the common path registers
template <int MODE>
__device__ __noinline__ void push2d_walk_fallback(const PP& pb, int b, int it, unsigned mask, int cur, int fcur, double tau,
                                                  double2* sxz, double2* sUd, double2* sU2r, int* s_id, int* s_pre,
                                                  int* s_start) {
  const PP p = pp_view(pb, b);
  if (p.ds == 2)
    push2d_warp_unit<true, MODE>(p, it, mask, cur, nq(p, cur ^ 1), fq(p, fcur), fcur, tau, sxz, sUd, sU2r, s_id, s_pre, s_start);
  else
    push2d_warp_unit<false, MODE>(p, it, mask, cur, nq(p, cur ^ 1), fq(p, fcur), fcur, tau, sxz, sUd, sU2r, s_id, s_pre, s_start);
}

// warp-level units of ALL sources over the de-duplicated target lists (short-column meshes)
#ifndef RT_TGT_MINB
#define RT_TGT_MINB 8
#endif
template <int MODE>
__global__ void __launch_bounds__(PUSH_BLOCK, RT_TGT_MINB) push2d_tgt_dc_kernel(PP pb, int tail) {
  constexpr bool DUAL = MODE == MODE_DUAL;
  __shared__ double2 w_sxz[PUSH_BLOCK / 32][32], w_sUd[PUSH_BLOCK / 32][32], w_sU2r[DUAL ? PUSH_BLOCK / 32 : 1][32];
  __shared__ int w_id[PUSH_BLOCK / 32][32], w_pre[PUSH_BLOCK / 32][32], w_start[PUSH_BLOCK / 32][32];
  const int warp = threadIdx.x >> 5;
  const int nb = pb.nb;
  const i64* base = pb.flat;
  const i64 total = base[nb];
  const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  for (i64 g = gw; g < total; g += nw) {
    const int b = nb == 1 ? 0 : (g < pb.flat_cap ? __ldcg(&pb.flat_b[g]) : flat_owner(base, nb, g));
    const i64 slot = g - base[b];
    const PP p = pp_view(pb, b);
    const int cur = p.ctl[0], fcur = p.ctl[1];
    const int it = __ldcg(&nq(p, cur)[slot]);
    const unsigned mask = __ldcg(&p.cur_mask[slot]);
    if (mask == 0u) continue;
    const i64 t0 = p.tgt_off[it], t1 = p.tgt_off[it + 1];
    const double tau = __ldcg(&p.tau[0]);
    if (t1 <= t0) {  // no list (a column longer than the builder's cap, e.g. the centre node): walk its elements
      push2d_walk_fallback<MODE>(pb, b, it, mask, cur, fcur, tau, w_sxz[warp], w_sUd[warp], w_sU2r[DUAL ? warp : 0], w_id[warp],
                                 w_pre[warp], w_start[warp]);
      continue;
    }
    if (p.ds == 2)
      push2d_tgt_unit<true, MODE>(p, it, mask, t0, t1, cur, nq(p, cur ^ 1), fq(p, fcur), fcur, tau, w_sxz[warp], w_sUd[warp],
                                  w_sU2r[DUAL ? warp : 0], w_id[warp]);
    else
      push2d_tgt_unit<false, MODE>(p, it, mask, t0, t1, cur, nq(p, cur ^ 1), fq(p, fcur), fcur, tau, w_sxz[warp], w_sUd[warp],
                                   w_sU2r[DUAL ? warp : 0], w_id[warp]);
  }
  round_tail(pb, tail);
}

// threshold advance over the concatenated far lists of the advancing sources: THREAD per far-list slot (an item of a
// coarse mesh holds one or two nodes: a warp per slot would idle 30 lanes); warps that lie inside one source -- nearly
// all of them -- reduce / append once per warp
__global__ void far_min_dc_kernel(PP pb) {
  const int nb = pb.nb;
  const i64* base = pb.flat + FLAT_STRIDE;
  const i64 total = base[nb];
  const int lane = threadIdx.x & 31;
  const i64 gsize = (i64)gridDim.x * blockDim.x;
  for (i64 g0 = (i64)blockIdx.x * blockDim.x + (threadIdx.x & ~31); g0 < total; g0 += gsize) {  // warp-uniform trip count
    const i64 g = g0 + lane;
    const bool valid = g < total;
    int b = -1;
    u64 best = ~0ull;
    if (valid) {
      b = nb == 1 ? 0 : flat_owner(base, nb, g);
      const i64 o = (i64)b * pb.n_items;
      const i32* far_cur = (pb.ctl[b * 8 + 1] ? pb.farq1 : pb.farq0) + o;
      const int it = far_cur[g - base[b]];
      unsigned m = pb.far_mask[o + it];
      const i64 v0 = (i64)b * pb.n + pb.item_first[it];
      while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1u;
        const u64 v = (u64)__double_as_longlong(pb.dist[(v0 + k) * pb.ds]);
        best = v < best ? v : best;
      }
    }
    const int b0 = __shfl_sync(FULL, b, 0);
    if (__all_sync(FULL, !valid || b == b0)) {
      for (int o = 16; o; o >>= 1) {
        const u64 other = __shfl_xor_sync(FULL, best, o);
        best = other < best ? other : best;
      }
      if (lane == 0 && best != ~0ull) atomicMin((u64*)&pb.tau[(i64)b0 * 4 + 2], best);
    } else if (valid && best != ~0ull) {
      atomicMin((u64*)&pb.tau[(i64)b * 4 + 2], best);
    }
  }
}
__global__ void far_release_dc_kernel(PP pb, int tail) {
  const int nb = pb.nb;
  const i64* base = pb.flat + FLAT_STRIDE;
  const i64 total = base[nb];
  const int lane = threadIdx.x & 31;
  const i64 gsize = (i64)gridDim.x * blockDim.x;
  for (i64 g0 = (i64)blockIdx.x * blockDim.x + (threadIdx.x & ~31); g0 < total; g0 += gsize) {
    const i64 g = g0 + lane;
    const bool valid = g < total;
    int b = -1, it = 0, cur = 0, fcur = 0;
    unsigned keep = 0u, relm = 0u;
    bool first_pend = false;
    if (valid) {
      b = nb == 1 ? 0 : flat_owner(base, nb, g);
      const i64 slot = g - base[b];
      const i64 o = (i64)b * pb.n_items;
      cur = pb.ctl[b * 8];
      fcur = pb.ctl[b * 8 + 1];
      const double tau = __dadd_rn(pb.tau[(i64)b * 4 + 2], pb.tau[(i64)b * 4 + 1]);
      it = ((fcur ? pb.farq1 : pb.farq0) + o)[slot];
      const unsigned m = pb.far_mask[o + it];
      const i64 v0 = (i64)b * pb.n + pb.item_first[it];
      unsigned mm = m;
      while (mm) {
        const int k = __ffs(mm) - 1;
        mm &= mm - 1u;
        if (pb.dist[(v0 + k) * pb.ds] < tau) relm |= 1u << k;
      }
      keep = m & ~relm;
      pb.far_mask[o + it] = keep;
      if (!keep) pb.infar[o + it] = 0u;
      if (relm) first_pend = atomicOr(&pb.pend_mask[o + it], relm) == 0u;
      if (slot == 0) pb.tau[(i64)b * 4] = tau;
    }
    const int b0 = __shfl_sync(FULL, b, 0);
    if (__all_sync(FULL, !valid || b == b0)) {  // one source: one append per list and warp
      const unsigned kb = __ballot_sync(FULL, valid && keep != 0u), pbm = __ballot_sync(FULL, first_pend);
      const int cur0 = __shfl_sync(FULL, cur, 0), fcur0 = __shfl_sync(FULL, fcur, 0);
      u64 base_k = 0, base_p = 0;
      if (lane == 0 && b0 >= 0) {
        if (kb) base_k = atomicAdd(&pb.counters[(i64)b0 * 8 + 4 + (fcur0 ^ 1)], (u64)__popc(kb));
        if (pbm) base_p = atomicAdd(&pb.counters[(i64)b0 * 8 + cur0], (u64)__popc(pbm));
      }
      base_k = __shfl_sync(FULL, base_k, 0);
      base_p = __shfl_sync(FULL, base_p, 0);
      const i64 o = (i64)max(b0, 0) * pb.n_items;
      if (valid && keep) ((fcur0 ? pb.farq0 : pb.farq1) + o)[base_k + __popc(kb & ((1u << lane) - 1u))] = it;
      if (first_pend) ((cur0 ? pb.nearq1 : pb.nearq0) + o)[base_p + __popc(pbm & ((1u << lane) - 1u))] = it;
    } else if (valid) {
      const i64 o = (i64)b * pb.n_items;
      if (keep) ((fcur ? pb.farq0 : pb.farq1) + o)[atomicAdd(&pb.counters[(i64)b * 8 + 4 + (fcur ^ 1)], 1ull)] = it;
      if (first_pend) ((cur ? pb.nearq1 : pb.nearq0) + o)[atomicAdd(&pb.counters[(i64)b * 8 + cur], 1ull)] = it;
    }
  }
  round_tail(pb, tail);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant: ONE cooperative launch runs up to `max_rounds` rounds; phases are separated by grid-wide
// barriers instead of kernel boundaries (a round costs two or three grid.sync() instead of ~5 launches).
template <bool WARP, int MODE>
__global__ void __launch_bounds__(PUSH_BLOCK) nearfar_persistent_kernel(PP pb, int max_rounds) {
  cg::grid_group grid = cg::this_grid();
  const bool first = blockIdx.x == 0 && threadIdx.x == 0;
  const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 gsize = (i64)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int nb = pb.nb;  // <= 32 sources in lock step
  unsigned curb = 0, fcurb = 0;  // bit b = ping-pong index of source b
  for (int b = 0; b < nb; ++b) {
    curb |= (unsigned)(pb.ctl[b * 8 + 0] & 1) << b;
    fcurb |= (unsigned)(pb.ctl[b * 8 + 1] & 1) << b;
  }
  int rounds = 0;
  unsigned pushes = 0, advances = 0;  // per-source activity of the LAST round (for the counters below)
  int done = 0;
  for (int r = 0; r < max_rounds; ++r) {
    // ---- phase A: decide each source's mode from its (stable) counters; prepare
    unsigned mode1 = 0, mode2 = 0;
    for (int b = 0; b < nb; ++b) {
      const PP p = pp_view(pb, b);
      const int cur = (curb >> b) & 1, fcur = (fcurb >> b) & 1;
      const i64 n_near = (i64)__ldcg(&p.counters[cur]);
      const i64 n_far = (i64)__ldcg(&p.counters[4 + fcur]);
      if (n_near > 0) {
        mode1 |= 1u << b;
        if (first) {
          p.counters[cur ^ 1] = 0;
          p.ctl[5] += 1;
        }
        const i32* near_cur = nq(p, cur);
        for (i64 slot = gtid; slot < n_near; slot += gsize)
          p.cur_mask[slot] = atomicExch(&p.pend_mask[__ldcg(&near_cur[slot])], 0u);
      } else if (n_far > 0) {
        mode2 |= 1u << b;
        if (first) {
          p.tau[2] = __longlong_as_double(-1LL);
          p.counters[4 + (fcur ^ 1)] = 0;
        }
      }
      if (first && (n_near > 0 || n_far > 0)) p.ctl[4] += 1;
    }
    if ((mode1 | mode2) == 0u) {
      done = 1;
      break;
    }
    ++rounds;
    grid.sync();
    // ---- phase B: push (mode 1) / smallest waiting value (mode 2)
    for (int b = 0; b < nb; ++b) {
      const PP p = pp_view(pb, b);
      const int cur = (curb >> b) & 1, fcur = (fcurb >> b) & 1;
      if ((mode1 >> b) & 1u) {
        if (WARP && nb > 1) {
          // every source gets its own team of warps (warp w serves source w % nb) so that the sources advance
          // concurrently instead of one after the other
          const i64 gw = gtid >> 5, nw = gsize >> 5;
          if ((int)(gw % nb) == b) push2d_warp_body<MODE>(p, nq(p, cur), cur, nq(p, cur ^ 1), fq(p, fcur), fcur, gw / nb,
                                                    (nw - b + nb - 1) / nb);
        } else {
          push2d_body<WARP, MODE>(p, nq(p, cur), cur, nq(p, cur ^ 1), fq(p, fcur), fcur);
        }
      } else if ((mode2 >> b) & 1u) {
        const i64 n_far = (i64)__ldcg(&p.counters[4 + fcur]);
        const i32* far_cur = fq(p, fcur);
        u64 best = ~0ull;
        for (i64 slot = gtid >> 5; slot < n_far; slot += gsize >> 5) {
          const int it = __ldcg(&far_cur[slot]);
          const unsigned m = __ldcg(&p.far_mask[it]);
          if ((m >> lane) & 1u) {
            const u64 bb = (u64)__double_as_longlong(__ldcg(&p.dist[(i64)(p.item_first[it] + lane) * p.ds]));
            best = bb < best ? bb : best;
          }
        }
        for (int o = 16; o; o >>= 1) {
          const u64 other = __shfl_xor_sync(FULL, best, o);
          best = other < best ? other : best;
        }
        if (lane == 0 && best != ~0ull) atomicMin((u64*)&p.tau[2], best);
      }
    }
    grid.sync();
    // ---- phase C: release below the new threshold (mode 2)
    if (mode2) {
      for (int b = 0; b < nb; ++b) {
        if (!((mode2 >> b) & 1u)) continue;
        const PP p = pp_view(pb, b);
        const int cur = (curb >> b) & 1, fcur = (fcurb >> b) & 1;
        const i64 n_far = (i64)__ldcg(&p.counters[4 + fcur]);
        const i32* far_cur = fq(p, fcur);
        const double tau = __dadd_rn(__ldcg(&p.tau[2]), __ldcg(&p.tau[1]));
        i32* far_next = fq(p, fcur ^ 1);
        i32* near_next = nq(p, cur);
        for (i64 slot = gtid >> 5; slot < n_far; slot += gsize >> 5) {
          const int it = __ldcg(&far_cur[slot]);
          const unsigned m = __ldcg(&p.far_mask[it]);
          const bool mine = (m >> lane) & 1u;
          const bool rel = mine && __ldcg(&p.dist[(i64)(p.item_first[it] + lane) * p.ds]) < tau;
          const unsigned relm = __ballot_sync(FULL, rel);
          if (lane == 0) {
            const unsigned keep = m & ~relm;
            p.far_mask[it] = keep;
            if (keep)
              far_next[atomicAdd(&p.counters[4 + (fcur ^ 1)], 1ull)] = it;
            else
              p.infar[it] = 0u;
            if (relm) {
              const unsigned old = atomicOr(&p.pend_mask[it], relm);
              if (old == 0u) near_next[atomicAdd(&p.counters[cur], 1ull)] = it;
            }
          }
        }
        if (first) p.tau[0] = tau;
      }
      grid.sync();
    }
    curb ^= mode1;
    fcurb ^= mode2;
    pushes = mode1;
    advances = mode2;
  }
  (void)pushes;
  (void)advances;
  (void)rounds;
  if (first) {
    for (int b = 0; b < nb; ++b) {
      pb.ctl[b * 8 + 0] = (curb >> b) & 1;
      pb.ctl[b * 8 + 1] = (fcurb >> b) & 1;
      pb.ctl[b * 8 + 3] = done;
    }
  }
}

// ---- packed mode epilogue: split the pairs into dist / prev; halo-coupled nodes take the predecessor of their
// twin (update_halo!: p[h2] = p[h1], bfm.jl:59), resolved in a few passes for twin chains.
__global__ void unpack_kernel(PP pb, i64 n, double* __restrict__ dist_out_base) {
  const PP p = pp_view(pb, blockIdx.y);
  const int source = p.source;
  double* dist_out = dist_out_base + (i64)blockIdx.y * n;
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = p.dist[2 * i];
  dist_out[i] = d;
  const u64 k = p.keys[2 * i + 1];
  if (i == source || k == KEY_NONE) return;  // keeps the halo-init / unset value like the reference
  if (k & KEY_HALO)
    p.prev[i] = -3 - (i32)(k & 0xffffffffull);
  else
    p.prev[i] = (i32)(k & 0xffffffffull);
}
__global__ void halo_prev_fix_kernel(PP pb, const i32* __restrict__ h2, i64 rows, int last) {
  const PP p = pp_view(pb, blockIdx.y);
  const int source = p.source;
  const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int b = h2[r];
  const i32 pv = p.prev[b];
  if (pv > -3) return;
  const int t = -3 - pv;
  const i32 pt = p.prev[t];
  if (t == source)
    p.prev[b] = source;
  else if (pt > -3 && pt >= 0)
    p.prev[b] = pt;
  else if (last)
    p.prev[b] = t;  // unresolved chain: the twin itself is a valid zero-weight predecessor
}

int ensure_push_workspace(rt_mesh* h, int nb, bool packed, bool need_bdist) {
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  if (!m.push_ready) {
    RT_TRY(m.node_item.alloc(m.n));
    node_item_kernel<<<grid_for(m.n_items * 32, 256), 256, 0, s>>>(m.item_first.p, m.n_items, m.node_item.p);
    RT_TRY(m.hn_index.alloc(m.n));
    RT_CUDA(cudaMemsetAsync(m.hn_index.p, 0xff, m.n * sizeof(i32), s));
    if (m.n_hn) hn_index_kernel<<<grid_for(m.n_hn, 256), 256, 0, s>>>(m.hn_node.p, m.n_hn, m.hn_index.p);
    for (int k = 0; k < 2; ++k) RT_TRY(m.unresolved[k].alloc(m.n));
    RT_TRY(m.pending_prev.alloc(m.n));
    RT_CUDA(cudaStreamSynchronize(s));
    m.push_ready = true;
  }
  if (m.push_nb < nb) {  // per-source state for a batch of nb sources solved in lock step
    const size_t B = (size_t)nb;
    RT_TRY(m.pend_mask.alloc(B * m.n_items));
    RT_TRY(m.far_mask.alloc(B * m.n_items));
    RT_TRY(m.infar_u.alloc(B * m.n_items));
    RT_TRY(m.cur_mask.alloc(B * m.n_items));
    for (int k = 0; k < 2; ++k) {
      RT_TRY(m.nearq[k].alloc(B * m.n_items));
      RT_TRY(m.farq[k].alloc(B * m.n_items));
    }
    RT_TRY(m.tau.alloc(B * 4));
    RT_TRY(m.ctl.alloc(B * 8));
    RT_TRY(m.bcounters.alloc(B * 8));
    RT_TRY(m.bsources.alloc(B));
    RT_TRY(m.bprev.alloc(B * m.n));
    m.bdist.release();
    m.dp.release();
    m.push_nb = nb;
  }
  if (!m.flat.p) RT_TRY(m.flat.alloc(2 * FLAT_STRIDE + 2));  // + the ticket word of round_tail
  {
    const size_t want = m.push_nb > 1 ? (size_t)std::min<i64>((i64)m.push_nb * m.n_items, (i64)1 << 27) : 1;
    if (m.flat_b.n < want) RT_TRY(m.flat_b.alloc(want));
  }
  if (packed && !m.dp.p) RT_TRY(m.dp.alloc(2 * (size_t)m.n * (size_t)m.push_nb));
  // plain travel-time tables: only the separate-pass mode keeps them here, or a caller that passes no dist buffer
  if ((!packed || need_bdist) && !m.bdist.p) RT_TRY(m.bdist.alloc((size_t)m.push_nb * m.n));
  return RT_OK;
}


const void* persistent_kernel_for(bool warp, int mode) {
  if (warp) {
    if (mode == MODE_DUAL) return (const void*)nearfar_persistent_kernel<true, MODE_DUAL>;
    if (mode == MODE_F32) return (const void*)nearfar_persistent_kernel<true, MODE_F32>;
    return (const void*)nearfar_persistent_kernel<true, MODE_F64>;
  }
  if (mode == MODE_DUAL) return (const void*)nearfar_persistent_kernel<false, MODE_DUAL>;
  if (mode == MODE_F32) return (const void*)nearfar_persistent_kernel<false, MODE_F32>;
  return (const void*)nearfar_persistent_kernel<false, MODE_F64>;
}
void launch_push_dc(bool warp, int mode, dim3 grid, cudaStream_t s, const PP& p, int tail) {
  if (warp) {
    if (mode == MODE_DUAL)
      push2d_dc_kernel<true, MODE_DUAL><<<grid, PUSH_BLOCK, 0, s>>>(p, tail);
    else if (mode == MODE_F32)
      push2d_dc_kernel<true, MODE_F32><<<grid, PUSH_BLOCK, 0, s>>>(p, tail);
    else
      push2d_dc_kernel<true, MODE_F64><<<grid, PUSH_BLOCK, 0, s>>>(p, tail);
  } else {
    if (mode == MODE_DUAL)
      push2d_dc_kernel<false, MODE_DUAL><<<grid, PUSH_BLOCK, 0, s>>>(p, tail);
    else if (mode == MODE_F32)
      push2d_dc_kernel<false, MODE_F32><<<grid, PUSH_BLOCK, 0, s>>>(p, tail);
    else
      push2d_dc_kernel<false, MODE_F64><<<grid, PUSH_BLOCK, 0, s>>>(p, tail);
  }
}

}  // namespace

int bfm2d_ensure_workspace(rt_mesh* h);

int bfm2d_solve_push_impl(rt_mesh* h, const double* U_dev, bool dual_arg, const i64* sources, i64 nsrc, double* dist_dev,
                          i32* prev_dev, rt_stats* stats);

int bfm2d_solve_push(rt_mesh* h, const double* U_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                     rt_stats* stats) {
  return bfm2d_solve_push_impl(h, U_dev, false, sources, nsrc, dist_dev, prev_dev, stats);
}
// U2_dev: [n x 2] column-major (dual_velocity)
int bfm2d_solve_push_dual(rt_mesh* h, const double* U2_dev, const i64* sources, i64 nsrc, double* dist_dev,
                          i32* prev_dev, rt_stats* stats) {
  return bfm2d_solve_push_impl(h, U2_dev, true, sources, nsrc, dist_dev, prev_dev, stats);
}

int bfm2d_solve_push_impl(rt_mesh* h, const double* U_dev, bool dual_arg, const i64* sources, i64 nsrc, double* dist_dev,
                          i32* prev_dev, rt_stats* stats) {
  Mesh2D& m = *h->m2;
  const bool dual = dual_arg, f32 = h->f32;
  const int mode = dual ? MODE_DUAL : (f32 ? MODE_F32 : MODE_F64);
  if (dual) RT_ARG(m.has_polar, "the dual-velocity relax needs gr.r (mesh adopted without theta / r)");
  if (mode != MODE_F64)
    RT_ARG(h->opts.packed_prev != 0 && h->opts.cta_units == 0 && h->opts.profile_timers == 0,
           "dual velocity / precision 32 in the near-far schedule need packed_prev=1, cta_units=0, profile_timers=0");
  if (f32) RT_TRY(mesh2d_prepare_f32(h));
  cudaStream_t s = h->stream;
  const i64 n = m.n;
  const bool timers = h->opts.profile_timers != 0;
  const bool packed = h->opts.packed_prev != 0;
  // batch width: small meshes leave the GPU idle per round, so many sources advance in lock step: every launch of a
  // round serves the concatenated near / far lists of all of them (flat index space, see round_begin_kernel)
  const int warp_units =
      (h->opts.warp_units == 1 || (h->opts.warp_units < 0 && m.graph_edges / std::max<i64>(n, 1) < 1500)) ? 1 : 0;
  int nb = 1;
  if (!timers && packed && nsrc > 1 && n <= 4000000) {
    const i64 per_src = n * 32 + m.n_items * 40;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const i64 budget = std::max<i64>((i64)6 << 30, (i64)(free_b / 3));
    // the long-column units keep one grid row per source (<= 32); the warp-per-item units are flattened
    const i64 cap = warp_units ? MAX_NB : 32;
    nb = (int)std::min<i64>(std::min<i64>(nsrc, h->opts.batch > 0 ? std::min<i64>(h->opts.batch, cap) : (warp_units ? 512 : 32)),
                            std::max<i64>(1, budget / per_src));
    nb = std::max(1, nb);
  }
  if (timers || !packed) RT_TRY(bfm2d_ensure_workspace(h));  // pinned counter mirror of the host-driven paths
  RT_TRY(ensure_push_workspace(h, nb, packed, dist_dev == nullptr));
  nb = std::min(nb, m.push_nb);
  PP p;
  std::memset((void*)&p, 0, sizeof(p));  // (the struct is also the cache key of the captured round graph: no stray padding)
  p.x = f32 ? m.xf.p : m.x.p;
  p.z = f32 ? m.zf.p : m.z.p;
  p.U = U_dev;
  p.U1 = U_dev;
  p.U2 = dual ? U_dev + m.n : U_dev;
  p.r = m.r.p;
  p.e2n_off = m.e2n_off.p;
  p.e2n_idx = m.e2n_idx.p;
  p.g_off = m.g_off.p;
  p.g_idx = m.g_idx.p;
  p.item_first = m.item_first.p;
  p.node_item = m.node_item.p;
  p.hn_node = m.hn_node.p;
  p.hn_off = m.hn_off.p;
  p.hn_part = m.hn_part.p;
  p.hn_index = m.hn_index.p;
  p.n_hn = (int)m.n_hn;
  p.source = -1;
  p.dist = packed ? m.dp.p : m.bdist.p;
  p.ds = packed ? 2 : 1;
  p.keys = packed ? (u64*)m.dp.p : nullptr;
  p.prev = m.bprev.p;
  p.pend_mask = m.pend_mask.p;
  p.far_mask = m.far_mask.p;
  p.infar = m.infar_u.p;
  p.cur_mask = m.cur_mask.p;
  p.counters = m.bcounters.p;
  p.tau = m.tau.p;
  p.nearq0 = m.nearq[0].p;
  p.nearq1 = m.nearq[1].p;
  p.farq0 = m.farq[0].p;
  p.farq1 = m.farq[1].p;
  p.ctl = m.ctl.p;
  p.nb = 1;
  p.warp_units = warp_units;
  p.flat = m.flat.p;
  p.ticket = reinterpret_cast<unsigned*>(m.flat.p + 2 * FLAT_STRIDE);
  p.flat_b = m.flat_b.p;
  p.flat_cap = (i64)m.flat_b.n;
  p.cta_units = h->opts.cta_units;
  p.group_screen = h->opts.group_screen;
  p.compact = h->opts.compact;
  p.n = n;
  p.n_items = m.n_items;
  p.sources = m.bsources.p;
  p.tgt_off = nullptr;
  p.tgt_idx = nullptr;
  if (warp_units && !m.tgt_tried) {  // de-duplicated target lists of the work items, once per mesh
    m.tgt_tried = true;
    DevBuf<i64> off;
    size_t sb = 0;
    if (off.alloc(m.n_items + 1) == RT_OK) {
      const unsigned gi = (unsigned)std::min<i64>((m.n_items + 3) / 4, 148 * 32);
      item_targets_kernel<0><<<gi, 128, 0, s>>>(p, off.p, nullptr);
      cudaMemsetAsync(off.p + m.n_items, 0, sizeof(i64), s);
      cub::DeviceScan::ExclusiveSum(nullptr, sb, off.p, off.p, m.n_items + 1, s);
      DevBuf<uint8_t> tmp;
      i64 total = -1;
      if (tmp.alloc(sb) == RT_OK) {
        cub::DeviceScan::ExclusiveSum(tmp.p, sb, off.p, off.p, m.n_items + 1, s);
        cudaMemcpyAsync(&total, off.p + m.n_items, sizeof(i64), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
      }
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      m.tgt_missing = true;  // refined below: are there items without a list (column longer than TGT_CAP)?
      if (total > 0 && (size_t)total * sizeof(i32) < free_b / 8 && m.tgt_idx.alloc((size_t)total) == RT_OK) {
        item_targets_kernel<1><<<gi, 128, 0, s>>>(p, off.p, m.tgt_idx.p);
        if (m.tgt_off.alloc(m.n_items + 1) == RT_OK) {
          cudaMemcpyAsync(m.tgt_off.p, off.p, (m.n_items + 1) * sizeof(i64), cudaMemcpyDeviceToDevice, s);
          std::vector<i64> hoff(m.n_items + 1);
          cudaMemcpyAsync(hoff.data(), off.p, (m.n_items + 1) * sizeof(i64), cudaMemcpyDeviceToHost, s);
          cudaStreamSynchronize(s);
          m.tgt_missing = false;
          for (i64 q = 0; q < m.n_items; ++q)
            if (hoff[q + 1] == hoff[q]) {
              m.tgt_missing = true;
              break;
            }
        } else {
          m.tgt_idx.release();
        }
      }
    }
    RT_CUDA(cudaGetLastError());
  }
  if (warp_units && m.tgt_off.p && m.tgt_idx.p && h->opts.target_lists) {
    p.tgt_off = m.tgt_off.p;
    p.tgt_idx = m.tgt_idx.p;
  }

  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  const i64 max_blocks = (i64)sm_count * 16;
  i64 coop_blocks = 0;
  {
    int coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
    const void* pk = persistent_kernel_for(p.warp_units != 0, mode);
    const cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk, PUSH_BLOCK, 0);
    if (coop && oe == cudaSuccess) coop_blocks = (i64)per_sm * sm_count;
  }
  cudaEvent_t ev0, ev1, evr0, evr1;
  RT_CUDA(cudaEventCreate(&ev0));
  RT_CUDA(cudaEventCreate(&ev1));
  RT_CUDA(cudaEventCreate(&evr0));
  RT_CUDA(cudaEventCreate(&evr1));
  rt_stats st = {};
  st.graph_edges = m.graph_edges;
  int rc = RT_OK;

  // bucket width: option "delta" [s], or delta_factor x (mean travel time across a cell)
  double delta = h->opts.delta;
  if (!(delta > 0.0)) {
    RT_CUDA(cudaMemsetAsync(m.tau.p + 3, 0, sizeof(double), s));
    wdiag_kernel<<<grid_for(m.nel, 256), 256, 0, s>>>(p, m.nel, m.tau.p + 3);
    double wsum = 0.0;
    RT_CUDA(cudaMemcpyAsync(&wsum, m.tau.p + 3, sizeof(double), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    const double wmean = wsum > 0.0 ? wsum / (double)m.nel : 1.0;
    delta = wmean * (h->opts.delta_factor > 0.0 ? h->opts.delta_factor : 1.0);
  }
  // persistent kernel wins while the frontier is small (few grid-wide barriers beat 5 launches per round); large
  // meshes are better served by hardware block scheduling of the batched launches
  const bool small_mesh = (i64)n * nb <= 1500000 || (nb > 1 && n <= 1500000);
  // (measured on the 591 k-node mesh of BASELINE configs[2]: a persistent round costs ~17 us + 10 us per source, a
  // launch-sequence round ~36 us + 1.5 us per source: the persistent kernel only wins for one to three sources)
  const bool use_persistent = !timers && coop_blocks > 0 && nb <= 32 &&
                              (h->opts.persistent == 1 || (h->opts.persistent < 0 && small_mesh && nb <= 3));
  std::vector<int> hsrc(nb);
  std::vector<int> hctl((size_t)nb * 8);
  std::vector<u64> hcnt((size_t)nb * 8);

  for (i64 s0 = 0; s0 < nsrc && rc == RT_OK; s0 += nb) {
    const int B = (int)std::min<i64>(nb, nsrc - s0);
    for (int b = 0; b < B; ++b) {
      const i64 src1 = sources[s0 + b];
      if (src1 < 1 || src1 > n) {
        rt_set_error("source %lld out of range 1..%lld", (long long)src1, (long long)n);
        rc = RT_ERR_ARG;
        break;
      }
      hsrc[b] = (int)(src1 - 1);
    }
    if (rc != RT_OK) break;
    p.nb = B;
    p.source = hsrc[0];
    cudaEventRecord(ev0, s);
    cudaMemcpyAsync(m.bsources.p, hsrc.data(), B * sizeof(int), cudaMemcpyHostToDevice, s);
    cudaMemsetAsync(m.pend_mask.p, 0, (size_t)B * m.n_items * sizeof(unsigned), s);
    cudaMemsetAsync(m.far_mask.p, 0, (size_t)B * m.n_items * sizeof(unsigned), s);
    cudaMemsetAsync(m.infar_u.p, 0, (size_t)B * m.n_items * sizeof(unsigned), s);
    push_init_kernel<<<dim3(grid_for(n, 256), B), 256, 0, s>>>(p, n, delta);
    st.total_launches += 1;
    i64 rounds = 0;
    if (use_persistent) {
      bool all_done = false;
      int max_rounds = 8192;
      while (!all_done) {
        void* args[] = {(void*)&p, (void*)&max_rounds};
        const void* kfn = persistent_kernel_for(p.warp_units != 0, mode);
        cudaError_t le = cudaLaunchCooperativeKernel(kfn, dim3((unsigned)coop_blocks), dim3(PUSH_BLOCK), args, 0, s);
        if (le != cudaSuccess) {
          rc = RT_ERR_CUDA;
          break;
        }
        st.total_launches += 1;
        cudaMemcpyAsync(hctl.data(), m.ctl.p, (size_t)B * 8 * sizeof(int), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) {
          rc = RT_ERR_CUDA;
          break;
        }
        all_done = true;
        for (int b = 0; b < B; ++b) all_done = all_done && hctl[b * 8 + 3];
      }
    } else if (!timers) {
      // device-controlled rounds, host sync every `check_every` rounds
      const int R = h->opts.check_every > 1 ? h->opts.check_every : 32;
      const unsigned gsmall = (unsigned)(sm_count * 2);
      const unsigned gfar = (unsigned)(sm_count * (B > 32 ? 8 : 2));
      // long-column units: one grid row per source; warp-per-item units: one flat grid over all sources
      const dim3 gpush = p.warp_units ? dim3((unsigned)max_blocks, 1) : dim3((unsigned)std::max<i64>(sm_count, max_blocks / B), B);
      bool all_done = false;
      i64 enq_rounds = 0;
      // the R-round launch sequence is the same every time (which phase runs is decided on the device): it is captured
      // once into a CUDA graph and replayed, which removes the per-launch host cost and most of the inter-kernel gaps
      i64 seq_launches = 0;
      const bool fused = h->opts.fuse_begin != 0 && B <= 32;
      cudaMemsetAsync(p.ticket, 0, sizeof(unsigned), s);
      auto enqueue_rounds = [&](cudaStream_t q) {
        int af = 1;  // nothing is pending before the first round; the last round of a sequence always runs the far kernels
        seq_launches = 0;
        for (int r = 0; r < R; ++r) {
          // fused: the round control of round r + 1 runs in the tail of the last kernel of round r (round_tail); only
          // the first round of the sequence launches it as a kernel of its own
          const bool far_now = r % FAR_EVERY == FAR_EVERY - 1 || r == R - 1;
          if (!fused || r == 0) {
            if (B <= 32)
              round_begin_kernel<32><<<1, 32, 0, q>>>(p, af);
            else
              round_begin_kernel<MAX_NB><<<1, MAX_NB, 0, q>>>(p, af);
            seq_launches += 1;
          }
          prep_dc_kernel<<<gsmall, 256, 0, q>>>(p);
          const int ptail = fused && !far_now ? 0 : -1;
          if (p.tgt_off) {  // (items without a list are walked inside the same kernel, out of line)
            if (mode == MODE_DUAL)
              push2d_tgt_dc_kernel<MODE_DUAL><<<gpush, PUSH_BLOCK, 0, q>>>(p, ptail);
            else if (mode == MODE_F32)
              push2d_tgt_dc_kernel<MODE_F32><<<gpush, PUSH_BLOCK, 0, q>>>(p, ptail);
            else
              push2d_tgt_dc_kernel<MODE_F64><<<gpush, PUSH_BLOCK, 0, q>>>(p, ptail);
          } else {
            launch_push_dc(p.warp_units != 0, mode, gpush, q, p, ptail);
          }
          seq_launches += 2;
          af = 0;
          if (far_now) {
            far_min_dc_kernel<<<gfar, 256, 0, q>>>(p);
            far_release_dc_kernel<<<gfar, 256, 0, q>>>(p, fused && r < R - 1 ? 1 : -1);
            seq_launches += 2;
            af = 1;
          }
        }
      };
      cudaGraphExec_t gexec = nullptr;
      if (h->opts.use_graph) {
        std::vector<char> key(sizeof(PP) + 4 * sizeof(int));
        std::memcpy(key.data(), &p, sizeof(PP));
        const int kv[4] = {B, mode, R + (fused ? 1 << 16 : 0), (int)gpush.x};
        std::memcpy(key.data() + sizeof(PP), kv, sizeof(kv));
        if (m.round_graph && key != m.round_graph_key) {
          cudaGraphExecDestroy((cudaGraphExec_t)m.round_graph);
          m.round_graph = nullptr;
        }
        if (!m.round_graph) {
          cudaGraph_t g = nullptr;
          if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            enqueue_rounds(s);
            if (cudaStreamEndCapture(s, &g) == cudaSuccess && g) {
              cudaGraphExec_t ge = nullptr;
              if (cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) {
                m.round_graph = (void*)ge;
                m.round_graph_key = key;
                m.round_graph_launches = seq_launches;
              }
              cudaGraphDestroy(g);
            }
          }
          cudaGetLastError();  // a failed capture falls back to plain launches
        }
        gexec = (cudaGraphExec_t)m.round_graph;
        if (gexec) seq_launches = m.round_graph_launches;
      }
      while (!all_done) {
        if (gexec) {
          if (cudaGraphLaunch(gexec, s) != cudaSuccess) {
            rc = RT_ERR_CUDA;
            break;
          }
        } else {
          enqueue_rounds(s);
        }
        st.total_launches += seq_launches;
        enq_rounds += R;
        if (enq_rounds > ((i64)1 << 22)) {  // a solve needs 1e3 - 1e5 rounds: never spin forever on a logic error
          rt_set_error("near-far schedule did not converge within %lld rounds", (long long)enq_rounds);
          rc = RT_ERR_CUDA;
          break;
        }
        cudaMemcpyAsync(hctl.data(), m.ctl.p, (size_t)B * 8 * sizeof(int), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) {
          rc = RT_ERR_CUDA;
          break;
        }
        all_done = true;
        for (int b = 0; b < B; ++b) all_done = all_done && hctl[b * 8 + 3];
      }
    } else {
      // host-driven rounds with CUDA-event timers around every push launch (single source)
      int cur = 0, fcur = 0;
      i64 n_near = 1, n_far = 0;
      u64* ch = m.counters_host;
      while (n_near > 0 || n_far > 0) {
        const int nxt = cur ^ 1;
        bool pushed = false;
        if (n_near > 0) {
          cudaMemsetAsync(p.counters + nxt, 0, sizeof(u64), s);
          prep_kernel<<<grid_for(n_near, 256), 256, 0, s>>>(p, m.nearq[cur].p, cur);
          cudaEventRecord(evr0, s);
          if (p.warp_units)
            push2d_kernel<true, MODE_F64><<<(unsigned)std::min<i64>((n_near + 3) / 4, max_blocks), PUSH_BLOCK, 0, s>>>(
                p, m.nearq[cur].p, cur, m.nearq[nxt].p, m.farq[fcur].p, fcur);
          else
            push2d_kernel<false, MODE_F64><<<(unsigned)std::min<i64>(p.cta_units ? n_near * PUSH_GY : std::max<i64>(n_near * (PUSH_GE / 4), std::min<i64>(n_near * (PUSH_GE / 4) * PUSH_SPLIT, PUSH_FILL / 4)), max_blocks), PUSH_BLOCK, 0, s>>>(
                p, m.nearq[cur].p, cur, m.nearq[nxt].p, m.farq[fcur].p, fcur);
          cudaEventRecord(evr1, s);
          pushed = true;
          st.total_launches += 2;
          st.relax_launches += 1;
          cur = nxt;
        } else {
          cudaMemsetAsync(m.tau.p + 2, 0xff, sizeof(double), s);
          cudaMemsetAsync(p.counters + 4 + (fcur ^ 1), 0, sizeof(u64), s);
          cudaMemsetAsync(p.counters + cur, 0, sizeof(u64), s);
          far_min_kernel<<<grid_for(n_far * 32, 256), 256, 0, s>>>(p, m.farq[fcur].p, fcur);
          far_release_kernel<<<grid_for(n_far * 32, 256), 256, 0, s>>>(p, m.farq[fcur].p, fcur, m.farq[fcur ^ 1].p,
                                                                       m.nearq[cur].p, cur);
          st.total_launches += 2;
          fcur ^= 1;
        }
        cudaMemcpyAsync(ch, p.counters, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) {
          rc = RT_ERR_CUDA;
          break;
        }
        if (pushed) {
          float ms = 0.f;
          cudaEventElapsedTime(&ms, evr0, evr1);
          st.relax_ms += ms;
        }
        n_near = (i64)ch[cur];
        n_far = (i64)ch[4 + fcur];
        ++rounds;
      }
    }
    if (rc != RT_OK) break;
    cudaMemcpyAsync(hcnt.data(), p.counters, (size_t)B * 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      rc = RT_ERR_CUDA;
      break;
    }
    if (!timers)
      for (int b = 0; b < B; ++b) {
        rounds = std::max<i64>(rounds, hctl[b * 8 + 4]);
        st.relax_launches += hctl[b * 8 + 5];  // push rounds that had work (total_launches counts what was enqueued)
      }
    st.sweeps += rounds;
    for (int b = 0; b < B; ++b) {
      st.relaxed_edges += (i64)hcnt[b * 8 + 2];
      st.vertex_updates += (i64)hcnt[b * 8 + 3];
      if (timers) {
        st.screened_edges += (i64)hcnt[b * 8 + 6];
        st.exact_edges += (i64)hcnt[b * 8 + 7];
      }
    }
    // ---- predecessors
    cudaEventRecord(evr0, s);
    double* dist_out = dist_dev ? dist_dev + s0 * n : m.bdist.p;
    i64 n_un = 0;
    int ucur = 0;
    if (packed) {
      // the keys already hold a consistent predecessor per node: split the pairs, resolve the halo couplings
      if (m.n_hinit)
        prev_halo_init_kernel<<<dim3(grid_for(m.n_hinit, 256), B), 256, 0, s>>>(p, m.hinit_node.p, m.hinit_val.p,
                                                                                m.n_hinit);
      unpack_kernel<<<dim3(grid_for(n, 256), B), 256, 0, s>>>(p, n, dist_out);
      st.total_launches += 2;
      if (m.halo_rows > 0)
        for (int pass = 0; pass < 4; ++pass) {
          halo_prev_fix_kernel<<<dim3(grid_for(m.halo_rows, 256), B), 256, 0, s>>>(p, m.halo_h2.p, m.halo_rows,
                                                                                   pass == 3);
          st.total_launches += 1;
        }
    } else {
      // separate tightness pass (single source): dist lives in bdist, prev in bprev
      u64* ch = m.counters_host;
      if (m.n_hinit)
        prev_halo_init_kernel<<<dim3(grid_for(m.n_hinit, 256), 1), 256, 0, s>>>(p, m.hinit_node.p, m.hinit_val.p,
                                                                                m.n_hinit);
      cudaMemsetAsync(p.counters + 6, 0, sizeof(u64), s);
      prev_tight_kernel<<<(unsigned)std::min<i64>((m.n_items + 7) / 8, (i64)sm_count * 8), 256, 0, s>>>(
          p, m.n_items, hsrc[0], m.unresolved[0].p);
      st.total_launches += 2;
      cudaMemcpyAsync(ch, p.counters, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        rc = RT_ERR_CUDA;
        break;
      }
      n_un = (i64)ch[6];
      for (int iter = 0; iter < 64 && n_un > 0; ++iter) {
        cudaMemsetAsync(p.counters + 7, 0, sizeof(u64), s);
        prev_resolve_kernel<<<grid_for(n_un, 128), 128, 0, s>>>(p, m.unresolved[ucur].p, n_un, hsrc[0],
                                                               m.unresolved[ucur ^ 1].p, m.pending_prev.p);
        prev_apply_kernel<<<grid_for(n_un, 256), 256, 0, s>>>(p, m.unresolved[ucur].p, n_un, m.pending_prev.p,
                                                              m.unresolved[ucur ^ 1].p);
        st.total_launches += 2;
        cudaMemcpyAsync(ch, p.counters, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
        if (cudaStreamSynchronize(s) != cudaSuccess) {
          rc = RT_ERR_CUDA;
          break;
        }
        const i64 left = (i64)ch[7];
        ucur ^= 1;
        if (left == n_un) break;  // no progress: the rest keeps its halo-init / unset predecessor
        n_un = left;
      }
      if (rc != RT_OK) break;
      if (n_un > 0) prev_giveup_kernel<<<grid_for(n_un, 256), 256, 0, s>>>(p, m.unresolved[ucur].p, n_un);
      if (dist_dev) cudaMemcpyAsync(dist_out, m.bdist.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s);
    }
    if (h->opts.canonical_prev) {
      // reference predecessors, exact ties included (canonical_prev.cu); the travel-time tables are final here
      for (int b = 0; b < B && rc == RT_OK; ++b) {
        i64 launches = 0;
        const double* dsrc = (packed || dist_dev) ? dist_out + (i64)b * n : m.bdist.p;
        rc = canonical_prev_2d(h, p.x, p.z, p.U1, dual ? p.U2 : nullptr, mode, dsrc, hsrc[b], m.bprev.p + (i64)b * n,
                               nullptr, &launches);
        st.total_launches += launches;
      }
      if (rc != RT_OK) break;
    }
    cudaEventRecord(evr1, s);
    cudaEventRecord(ev1, s);
    if (prev_dev)
      cudaMemcpyAsync(prev_dev + s0 * n, m.bprev.p, (size_t)B * n * sizeof(i32), cudaMemcpyDeviceToDevice, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      rc = RT_ERR_CUDA;
      break;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    st.kernel_ms += ms;
    cudaEventElapsedTime(&ms, evr0, evr1);
    st.prev_ms += ms;
  }
  cudaError_t e = cudaGetLastError();
  if (rc == RT_ERR_CUDA || e != cudaSuccess) {
    rt_set_error("CUDA failure in bfm2d_solve_push: %s", cudaGetErrorString(e));
    rc = RT_ERR_CUDA;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cudaEventDestroy(evr0);
  cudaEventDestroy(evr1);
  if (stats) *stats = st;
  return rc;
}
