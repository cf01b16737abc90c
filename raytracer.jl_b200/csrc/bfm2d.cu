// bfm2d.cu -- Bellman-Ford-Moore travel-time relaxation on the two-level annulus graph (sm_100a).
//
// Takes over bfm(G, halo, source, gr, U) src/SSSP/bfm.jl:1-52 and its helpers _relax! :161-210,
// update_halo! :54-62, init_halo_path! :64-70, init_Q! :74-80, update_Q! :82-98.
//
// Schedule 0 ("jacobi") keeps the reference's sweep structure (double-buffered dist0 -> dist, halo rule,
// frontier rebuild) so that dist AND prev equal the single-thread reference bit for bit, exact ties included
// (first minimiser in scan order wins, strict `>` against the incumbent).  What changes is the mapping:
//   * a warp owns a work item = up to 32 consecutive nodes that share one G column; the candidate scan
//     (column -> elements -> e2n lists) is read once per warp with coalesced id loads; lanes are laid out as
//     (target, part) so that small items still fill 32 lanes, and parts are merged with a lexicographic
//     (delta, scan position) shuffle reduction that reproduces the serial tie rule;
//   * the frontier follows update_Q! literally but at element granularity: an improved node touches every
//     element of its G column (push side), a node is active iff an element that contains it was touched (pull
//     side), an item is relaxed iff one of its nodes is active (a superset relaxes nothing new: bit-identical);
//   * edge weights 2*len/(U_i+U_j) are computed in registers with round-to-nearest intrinsics (no FMA).
// All arithmetic on the value path is IEEE fp64 with the reference's operation order.
#include "mesh2d.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RELAX_BLOCK = 256;

struct P2 {
  const double* __restrict__ x;
  const double* __restrict__ z;
  const double* __restrict__ U;
  // dual-velocity relax (_relax!(..., U::Matrix), bfm.jl:113-159): U1 = U[:,1] (below), U2 = U[:,2] (above), r = gr.r
  const double* __restrict__ U1;
  const double* __restrict__ U2;
  const double* __restrict__ r;
  const i32* __restrict__ e2n_off;
  const i32* __restrict__ e2n_idx;
  const i64* __restrict__ g_off;
  const i32* __restrict__ g_idx;
  const i32* __restrict__ n2e_off;
  const i32* __restrict__ n2e_idx;
  const i32* __restrict__ item_first;
  double* dist;
  double* dist0;
  i32* prev;
  uint8_t* dirty;
  u64* counters;
  const uint8_t* __restrict__ allowed;  // restricted continuation (rt_bfm_continue): only these nodes change; null = all
};
__device__ __forceinline__ bool node_allowed(const P2& p, int i) { return !p.allowed || p.allowed[i]; }

// dGi + 2.0 * sqrt(0 + dx*dx + dz*dz) / (Ui + Uj)   (bfm.jl:186, GridAnnulus.jl:808-815), no contraction
// exact_cand2<F32> (exact.h); F32 = the Float32 relax of src/SSSP/bfm_gpu.jl:487-526
template <bool F32>
__device__ __forceinline__ double cand_delta(double dj, double xi, double zi, double Ui, double xj, double zj,
                                             double Uj) {
  return exact_cand2<F32>(dj, xi, zi, Ui, xj, zj, Uj);
}

// One warp per active work item.  DUAL: the velocity pair of an edge is chosen by the radial order of its ends
// (head_idx = (r_i > r_j) + 1, tail_idx = (head_idx == 1) + 1, bfm.jl:137-138); muladd(2, len/(Ut+Uh), d) there
// equals d + (2 len)/(Ut+Uh) bit for bit because the factor 2 is exact with or without FMA.
template <bool DUAL, bool F32>
__global__ void __launch_bounds__(RELAX_BLOCK) relax2d_kernel(P2 p, const i32* __restrict__ active, int cur) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const i64 n_active = (i64)p.counters[cur];
  u64 evals = 0, updates = 0;
  for (i64 w = (i64)blockIdx.x * wpb + (threadIdx.x >> 5); w < n_active; w += (i64)gridDim.x * wpb) {
    const int item = active[w];
    const int v0 = p.item_first[item];
    const int t = p.item_first[item + 1] - v0;
    int lg = 0;
    while ((1 << lg) < t) ++lg;  // tl = 2^lg >= t, tl in {1..32}
    const int tl = 1 << lg;
    const int parts = 32 >> lg;
    const int ti = lane & (tl - 1);
    const int part = lane >> lg;
    const int i = v0 + min(ti, t - 1);
    const double xi = p.x[i], zi = p.z[i];
    const double Ui = DUAL ? 0.0 : p.U[i];
    const double Ui1 = DUAL ? p.U1[i] : 0.0, Ui2 = DUAL ? p.U2[i] : 0.0, ri = DUAL ? p.r[i] : 0.0;
    double best = p.dist0[i];
    int bpos = -1, bid = -1;  // -1 = incumbent dist0[i]; wins every tie (strict `di > delta`)
    int pos_base = 0;
    const i64 c0 = p.g_off[v0], c1 = p.g_off[v0 + 1];
    for (i64 c = c0; c < c1; ++c) {
      const int el = p.g_idx[c];
      const int s = p.e2n_off[el];
      const int m = p.e2n_off[el + 1] - s;
#pragma unroll 2
      for (int k = part; k < m; k += parts) {
        const int j = p.e2n_idx[s + k];
        const double dj = p.dist0[j];
        if (!(dj < best)) continue;  // dj + w >= dj >= best (also dj == Inf)
        const double xj = p.x[j], zj = p.z[j];
        double Ut = Ui, Uj;  // tail (this node) and head (candidate) velocities
        if (DUAL) {
          const bool down = ri > p.r[j];  // head_idx = 2 (above value of the candidate), tail_idx = 1
          Ut = down ? Ui1 : Ui2;
          Uj = down ? p.U2[j] : p.U1[j];
        } else {
          Uj = p.U[j];
        }
        {
          const double dx = __dsub_rn(xi, xj), dz = __dsub_rn(zi, zj);
          const double d2 = __fma_rn(dx, dx, dz * dz);
          if (screen_cannot_improve_t<F32>(best, dj, d2, __dadd_rn(Ut, Uj))) continue;
        }
        const double delta = cand_delta<F32>(dj, xi, zi, Ut, xj, zj, Uj);
        if (delta < best) {
          best = delta;
          bpos = pos_base + k;
          bid = j;
        }
      }
      pos_base += m;
    }
    // merge the parts: lexicographic (delta, scan position); incumbent position -1 is the smallest
    for (int off = 16; off >= tl; off >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, off);
      const int op = __shfl_xor_sync(FULL, bpos, off);
      const int oi = __shfl_xor_sync(FULL, bid, off);
      if (ob < best || (ob == best && op < bpos)) {
        best = ob;
        bpos = op;
        bid = oi;
      }
    }
    if (part == 0 && ti < t && node_allowed(p, i)) {
      p.dist[i] = best;
      if (bpos >= 0) p.prev[i] = bid;
    }
    if (lane == 0) {
      evals += (u64)t * (u64)pos_base;
      updates += (u64)t;
    }
  }
  if (lane == 0 && updates) {
    atomicAdd(&p.counters[2], evals);
    atomicAdd(&p.counters[3], updates);
  }
}

// push side of update_Q! (bfm.jl:88-95): an improved node touches every element of its G column
__device__ __forceinline__ void touch_column(const P2& p, int node, int lane0, int stride) {
  for (i64 c = p.g_off[node] + lane0; c < p.g_off[node + 1]; c += stride) p.dirty[p.g_idx[c]] = 1;
}

// update_halo! first half of the rows: (orig_k -> twin_k); every twin is written by exactly one row.
__global__ void halo_phase1_kernel(P2 p, const i32* __restrict__ h1, const i32* __restrict__ h2, i64 H) {
  i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= H) return;
  const int a = h1[k], b = h2[k];
  const double da = p.dist[a];
  if (da < p.dist0[a] && p.dist[b] > da && node_allowed(p, b)) {
    p.dist[b] = da;
    p.prev[b] = p.prev[a];
  }
}
// second half of the rows: (twin -> orig), grouped by orig, applied in ascending row order by one thread.
__global__ void halo_phase2_kernel(P2 p, const i32* __restrict__ orig, const i32* __restrict__ off,
                                   const i32* __restrict__ twin, i64 n_orig) {
  i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_orig) return;
  const int o = orig[g];
  double d_o = p.dist[o];
  int p_o = -2;
  for (int q = off[g]; q < off[g + 1]; ++q) {
    const int b = twin[q];
    const double db = p.dist[b];
    if (db < p.dist0[b] && d_o > db) {
      d_o = db;
      p_o = p.prev[b];
    }
  }
  if (p_o != -2 && node_allowed(p, o)) {
    p.dist[o] = d_o;
    p.prev[o] = p_o;
  }
}
// generic halo (arbitrary row structure): literal serial loop, one thread.
__global__ void halo_serial_kernel(P2 p, const i32* __restrict__ h1, const i32* __restrict__ h2, i64 rows) {
  if (blockIdx.x || threadIdx.x) return;
  for (i64 k = 0; k < rows; ++k) {
    const int a = h1[k], b = h2[k];
    if (p.dist[a] < p.dist0[a] && p.dist[b] > p.dist[a] && node_allowed(p, b)) {
      p.dist[b] = p.dist[a];
      p.prev[b] = p.prev[a];
    }
  }
}

// update_Q! + copyto!(dist0, dist) restricted to the vertices that can have changed: the active items ...
__global__ void commit_items_kernel(P2 p, const i32* __restrict__ active, int cur) {
  const i64 n_active = (i64)p.counters[cur];
  const i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_active) return;  // warp-uniform
  const int item = active[w];
  const int v0 = p.item_first[item];
  const int t = p.item_first[item + 1] - v0;
  bool improved = false;
  if (lane < t) {
    const int i = v0 + lane;
    const double d = p.dist[i];
    if (d < p.dist0[i]) {
      p.dist0[i] = d;
      improved = true;
    }
  }
  // all nodes of an item share one G column: touch it once, cooperatively
  if (__any_sync(FULL, improved)) touch_column(p, v0, lane, 32);
}
// ... and the targets of the halo rows.
__global__ void commit_halo_kernel(P2 p, const i32* __restrict__ h2, i64 rows) {
  i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= rows) return;
  const int b = h2[k];
  const double d = p.dist[b];
  if (d < p.dist0[b]) {
    p.dist0[b] = d;
    touch_column(p, b, 0, 1);
  }
}

// pull side of update_Q!: node j joins the frontier iff one of the elements that CONTAIN it was touched;
// an item is active iff any of its nodes is.  One warp per item, warp-aggregated append by lane 0.
__global__ void activate_kernel(P2 p, i64 n_items, i32* __restrict__ next_active, int nxt) {
  const i64 it = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (it >= n_items) return;  // warp-uniform
  const int v0 = p.item_first[it];
  const int t = p.item_first[it + 1] - v0;
  bool act = false;
  if (lane < t) {
    const int v = v0 + lane;
    if (node_allowed(p, v))
      for (int q = p.n2e_off[v]; q < p.n2e_off[v + 1]; ++q)
        if (p.dirty[p.n2e_idx[q]]) {
          act = true;
          break;
        }
  }
  if (__any_sync(FULL, act) && lane == 0) {
    const u64 slot = atomicAdd(&p.counters[nxt], 1ull);
    next_active[slot] = (i32)it;
  }
}

__global__ void init_state_kernel(double* __restrict__ dist, double* __restrict__ dist0, i32* __restrict__ prev,
                                  i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  dist[i] = INF;
  dist0[i] = INF;
  prev[i] = -1;
}
__global__ void init_source_kernel(P2 p, const i32* __restrict__ hnode, const i32* __restrict__ hval, i64 nh,
                                   int source) {
  i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nh) p.prev[hnode[k]] = hval[k];
  if (k == 0) {
    p.dist[source] = 0.0;
    p.dist0[source] = 0.0;
    touch_column(p, source, 0, 1);  // init_Q! (bfm.jl:74-80): the frontier starts as the source's star patch
  }
}

// rt_bfm_continue: a restart treats the seeds as nodes that have just improved.  (1) their travel times cross their halo
// rows (update_halo! with "is a seed" in place of "improved in this sweep": serial row order, allowed targets only, a node
// set this way counts as a seed for the rows after it) -- without this a phase that restarts on a discontinuity never
// enters the layer below, whose twins only receive values through the halo rule; (2) the frontier is the star patch of
// every seeded node (update_Q! for the seeds).
__global__ void seed_flag_kernel(const i32* __restrict__ seeds, i64 nseeds, uint8_t* __restrict__ seeded) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nseeds) seeded[seeds[k]] = 1;
}
__global__ void seed_halo_serial_kernel(P2 p, const i32* __restrict__ h1, const i32* __restrict__ h2, i64 rows,
                                        uint8_t* __restrict__ seeded) {
  if (blockIdx.x || threadIdx.x) return;
  for (i64 k = 0; k < rows; ++k) {
    const int a = h1[k], b = h2[k];
    if (!seeded[a] || !node_allowed(p, b)) continue;
    if (p.dist[b] > p.dist[a]) {
      p.dist[b] = p.dist[a];
      p.prev[b] = p.prev[a];
      seeded[b] = 1;
    }
  }
}
__global__ void seed_touch_kernel(P2 p, const uint8_t* __restrict__ seeded, i64 n) {
  const i64 k = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k < n && seeded[k]) touch_column(p, (int)k, threadIdx.x & 31, 32);
}
__global__ void prev_from_i64_kernel(const i64* __restrict__ src, i32* __restrict__ dst, i64 n) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (i32)(src[i] - 1);
}

int ensure_workspace(rt_mesh* h) {
  Mesh2D& m = *h->m2;
  if (m.ws_ready) return RT_OK;
  RT_TRY(m.dist.alloc(m.n));
  RT_TRY(m.dist0.alloc(m.n));
  RT_TRY(m.prev.alloc(m.n));
  RT_TRY(m.dirty.alloc(m.nel));
  RT_TRY(m.act[0].alloc(m.n_items));
  RT_TRY(m.act[1].alloc(m.n_items));
  RT_TRY(m.counters.alloc(8));
  RT_CUDA(cudaMallocHost((void**)&m.counters_host, 8 * sizeof(u64)));
  m.ws_ready = true;
  return RT_OK;
}

}  // namespace

int bfm2d_ensure_workspace(rt_mesh* h) { return ensure_workspace(h); }

// cont != null: one continuation solve from the state already loaded into the workspace (rt_bfm_continue)
struct Continuation {
  const uint8_t* allowed_dev;
  const i32* seeds_dev;
  i64 nseeds;
};
int bfm2d_solve_impl(rt_mesh* h, const double* U_dev, bool dual, const i64* sources, i64 nsrc, double* dist_dev,
                     i32* prev_dev, rt_stats* stats, const Continuation* cont = nullptr);

// precision = 32: Float32-rounded copies of the coordinates (Float32.(gr.x), bfm_gpu.jl:176-177), built once
int mesh2d_prepare_f32(rt_mesh* h) {
  Mesh2D& m = *h->m2;
  if (m.xf.n == (size_t)m.n) return RT_OK;
  RT_TRY(m.xf.alloc(m.n));
  RT_TRY(m.zf.alloc(m.n));
  RT_TRY(round_to_f32_device(m.x.p, m.xf.p, m.n, h->stream));
  RT_TRY(round_to_f32_device(m.z.p, m.zf.p, m.n, h->stream));
  return RT_OK;
}

int bfm2d_solve(rt_mesh* h, const double* U_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                rt_stats* stats) {
  if (h->opts.schedule == 1) return bfm2d_solve_push(h, U_dev, sources, nsrc, dist_dev, prev_dev, stats);
  return bfm2d_solve_impl(h, U_dev, false, sources, nsrc, dist_dev, prev_dev, stats);
}

// U2_dev: [n x 2] column-major (Julia Matrix): U[:,1] then U[:,2].
int bfm2d_solve_dual(rt_mesh* h, const double* U2_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                     rt_stats* stats) {
  RT_ARG(h->m2->has_polar, "the dual-velocity relax needs gr.r (mesh adopted without theta / r)");
  if (h->f32) {
    rt_set_error("precision = 32 is not available for the dual-velocity relax");
    return RT_ERR_UNSUPPORTED;
  }
  if (h->opts.schedule == 1) return bfm2d_solve_push_dual(h, U2_dev, sources, nsrc, dist_dev, prev_dev, stats);
  return bfm2d_solve_impl(h, U2_dev, true, sources, nsrc, dist_dev, prev_dev, stats);
}

int bfm2d_solve_impl(rt_mesh* h, const double* U_dev, bool dual, const i64* sources, i64 nsrc, double* dist_dev,
                     i32* prev_dev, rt_stats* stats, const Continuation* cont) {
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  RT_TRY(ensure_workspace(h));
  const i64 n = m.n;
  const bool f32 = h->f32;
  if (f32) RT_TRY(mesh2d_prepare_f32(h));
  P2 p;
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
  p.n2e_off = m.n2e_off.p;
  p.n2e_idx = m.n2e_idx.p;
  p.item_first = m.item_first.p;
  p.dist = m.dist.p;
  p.dist0 = m.dist0.p;
  p.prev = m.prev.p;
  p.dirty = m.dirty.p;
  p.counters = m.counters.p;
  p.allowed = cont ? cont->allowed_dev : nullptr;

  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  const int wpb = RELAX_BLOCK / 32;
  const i64 max_relax_blocks = (i64)sm_count * 8;  // 8 x 256 threads per SM, grid-stride beyond that

  cudaEvent_t ev0, ev1, evr0, evr1;
  RT_CUDA(cudaEventCreate(&ev0));
  RT_CUDA(cudaEventCreate(&ev1));
  RT_CUDA(cudaEventCreate(&evr0));
  RT_CUDA(cudaEventCreate(&evr1));
  rt_stats st = {};
  st.graph_edges = m.graph_edges;
  int rc = RT_OK;

  for (i64 si = 0; si < nsrc && rc == RT_OK; ++si) {
    const i64 src1 = cont ? 1 : sources[si];
    if (src1 < 1 || src1 > n) {
      rt_set_error("source %lld out of range 1..%lld", (long long)src1, (long long)n);
      rc = RT_ERR_ARG;
      break;
    }
    cudaEventRecord(ev0, s);
    cudaMemsetAsync(m.dirty.p, 0, m.nel, s);
    cudaMemsetAsync(m.counters.p, 0, 8 * sizeof(u64), s);
    if (cont) {  // dist / prev were loaded by the caller; seeds cross their halo rows; dist0 = copy; frontier
      DevBuf<uint8_t> seeded;
      if (seeded.alloc(n) != RT_OK) {
        rc = RT_ERR_CUDA;
        break;
      }
      cudaMemsetAsync(seeded.p, 0, n, s);
      if (cont->nseeds) seed_flag_kernel<<<grid_for(cont->nseeds, 256), 256, 0, s>>>(cont->seeds_dev, cont->nseeds, seeded.p);
      if (m.halo_rows > 0)
        seed_halo_serial_kernel<<<1, 32, 0, s>>>(p, m.halo_h1.p, m.halo_h2.p, m.halo_rows, seeded.p);
      cudaMemcpyAsync(p.dist0, p.dist, n * sizeof(double), cudaMemcpyDeviceToDevice, s);
      seed_touch_kernel<<<grid_for(n * 32, 256), 256, 0, s>>>(p, seeded.p, n);
      cudaStreamSynchronize(s);  // `seeded` goes out of scope
    } else {
      init_state_kernel<<<grid_for(n, 256), 256, 0, s>>>(p.dist, p.dist0, p.prev, n);
      init_source_kernel<<<grid_for(std::max<i64>(m.n_hinit, 1), 256), 256, 0, s>>>(p, m.hinit_node.p, m.hinit_val.p,
                                                                                   m.n_hinit, (int)(src1 - 1));
    }
    int cur = 0;
    activate_kernel<<<grid_for(m.n_items * 32, 256), 256, 0, s>>>(p, m.n_items, m.act[cur].p, cur);
    cudaMemsetAsync(m.dirty.p, 0, m.nel, s);
    st.total_launches += 3;
    cudaMemcpyAsync(m.counters_host, m.counters.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      rc = RT_ERR_CUDA;
      break;
    }
    i64 n_active = (i64)m.counters_host[cur];
    while (n_active > 0) {
      const int nxt = cur ^ 1;
      const i64 rb = std::min<i64>((n_active + wpb - 1) / wpb, max_relax_blocks);
      if (h->opts.profile_timers) cudaEventRecord(evr0, s);
      if (dual)
        relax2d_kernel<true, false><<<(unsigned)rb, RELAX_BLOCK, 0, s>>>(p, m.act[cur].p, cur);
      else if (f32)
        relax2d_kernel<false, true><<<(unsigned)rb, RELAX_BLOCK, 0, s>>>(p, m.act[cur].p, cur);
      else
        relax2d_kernel<false, false><<<(unsigned)rb, RELAX_BLOCK, 0, s>>>(p, m.act[cur].p, cur);
      if (h->opts.profile_timers) cudaEventRecord(evr1, s);
      if (m.halo_rows > 0) {
        if (m.halo_structured) {
          halo_phase1_kernel<<<grid_for(m.H, 256), 256, 0, s>>>(p, m.halo_h1.p, m.halo_h2.p, m.H);
          halo_phase2_kernel<<<grid_for(m.n_h2_orig, 256), 256, 0, s>>>(p, m.h2_orig.p, m.h2_off.p, m.h2_twin.p,
                                                                        m.n_h2_orig);
          st.total_launches += 2;
        } else {
          halo_serial_kernel<<<1, 32, 0, s>>>(p, m.halo_h1.p, m.halo_h2.p, m.halo_rows);
          st.total_launches += 1;
        }
      }
      commit_items_kernel<<<grid_for(n_active * 32, 256), 256, 0, s>>>(p, m.act[cur].p, cur);
      if (m.halo_rows > 0) {
        commit_halo_kernel<<<grid_for(m.halo_rows, 256), 256, 0, s>>>(p, m.halo_h2.p, m.halo_rows);
        st.total_launches += 1;
      }
      cudaMemsetAsync(m.counters.p + nxt, 0, sizeof(u64), s);
      activate_kernel<<<grid_for(m.n_items * 32, 256), 256, 0, s>>>(p, m.n_items, m.act[nxt].p, nxt);
      cudaMemsetAsync(m.dirty.p, 0, m.nel, s);
      cudaMemcpyAsync(m.counters_host, m.counters.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
      st.total_launches += 3;
      st.relax_launches += 1;
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        rc = RT_ERR_CUDA;
        break;
      }
      if (h->opts.profile_timers) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evr0, evr1);
        st.relax_ms += ms;
      }
      st.sweeps += 1;
      cur = nxt;
      n_active = (i64)m.counters_host[cur];
    }
    if (rc != RT_OK) break;
    st.relaxed_edges += (i64)m.counters_host[2];
    st.vertex_updates += (i64)m.counters_host[3];
    cudaEventRecord(ev1, s);
    if (dist_dev)
      cudaMemcpyAsync(dist_dev + si * n, m.dist.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (prev_dev) cudaMemcpyAsync(prev_dev + si * n, m.prev.p, n * sizeof(i32), cudaMemcpyDeviceToDevice, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      rc = RT_ERR_CUDA;
      break;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    st.kernel_ms += ms;
  }
  cudaError_t e = cudaGetLastError();
  if (rc == RT_ERR_CUDA || e != cudaSuccess) {
    rt_set_error("CUDA failure in bfm2d_solve: %s", cudaGetErrorString(e));
    rc = RT_ERR_CUDA;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cudaEventDestroy(evr0);
  cudaEventDestroy(evr1);
  if (stats) *stats = st;
  return rc;
}

// Restricted continuation -- the inner loop of bfm_multiphase (src/SSSP/bfm_multiphase.jl:118-150) on the graph of bfm:
// Jacobi sweeps from the caller's (dist, prev) state in which only nodes with allowed[i] != 0 are relaxed, activated or
// set by the halo rule; the frontier starts as the allowed nodes of the seeds' star patches.  Host arrays, 1-based ids.
int bfm2d_continue(rt_mesh* h, const double* U_dev, const uint8_t* allowed, const i64* seeds, i64 nseeds,
                   double* dist_io, i64* prev_io, rt_stats* stats) {
  Mesh2D& m = *h->m2;
  cudaStream_t s = h->stream;
  RT_TRY(ensure_workspace(h));
  const i64 n = m.n;
  if (h->f32) {
    rt_set_error("rt_bfm_continue is Float64 only");
    return RT_ERR_UNSUPPORTED;
  }
  std::vector<i32> hs(nseeds);
  for (i64 k = 0; k < nseeds; ++k) {
    RT_ARG(seeds[k] >= 1 && seeds[k] <= n, "seed out of range");
    hs[k] = (i32)(seeds[k] - 1);
  }
  DevBuf<uint8_t> dal;
  DevBuf<i32> dseeds;
  DevBuf<i64> dprev64;
  if (allowed) RT_TRY(dal.upload(allowed, n, s));
  RT_TRY(dseeds.upload(hs.data(), nseeds, s));
  RT_TRY(dprev64.upload(prev_io, n, s));
  RT_CUDA(cudaMemcpyAsync(m.dist.p, dist_io, n * sizeof(double), cudaMemcpyHostToDevice, s));
  prev_from_i64_kernel<<<grid_for(n, 256), 256, 0, s>>>(dprev64.p, m.prev.p, n);
  RT_CUDA(cudaGetLastError());
  Continuation c{allowed ? dal.p : nullptr, dseeds.p, nseeds};
  DevBuf<double> dout;
  DevBuf<i32> pout;
  RT_TRY(dout.alloc(n));
  RT_TRY(pout.alloc(n));
  RT_TRY(bfm2d_solve_impl(h, U_dev, false, nullptr, 1, dout.p, pout.p, stats, &c));
  RT_CUDA(cudaMemcpyAsync(dist_io, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  RT_TRY(prev_to_host_i64(pout.p, n, prev_io, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

// partition_grid(gr) src/topology/topology.jl:183-206, find_layer_number :137-147: id > 0 = "Layer_id", id < 0 =
// "Boundary_(-id)"; the radius is rounded to two digits first (Julia: round(r; digits = 2) = round(r * 100) / 100, ties to even)
namespace {
__constant__ double c_rl[7] = {6371.0 - 20.0, 6371.0 - 35.0, 6371.0 - 210.0, 6371.0 - 410.0,
                               6371.0 - 660.0, 6371.0 - 2740.0, 6371.0 - 2891.5};
__global__ void partition_kernel(const double* __restrict__ r, i64 n, int* __restrict__ id) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ri = __ddiv_rn(rint(__dmul_rn(r[i], 100.0)), 100.0);
  int out = 0;
  for (int k = 0; k < 7; ++k)
    if (ri == c_rl[k]) out = -(k + 1);
  if (out == 0) {
    if (ri > c_rl[0]) {
      out = 1;
    } else if (ri < c_rl[6]) {
      out = 8;
    } else {
      for (int k = 0; k < 6; ++k)
        if (c_rl[k] > ri && ri > c_rl[k + 1]) out = k + 2;
    }
  }
  id[i] = out;
}
}  // namespace

int mesh2d_partition(const rt_mesh* h, int32_t* id_out) {
  const Mesh2D& m = *h->m2;
  RT_ARG(m.has_polar && id_out, "partition_grid needs gr.r and an output array");
  DevBuf<int> id;
  RT_TRY(id.alloc(m.n));
  partition_kernel<<<grid_for(m.n, 256), 256, 0, h->stream>>>(m.r.p, m.n, id.p);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaMemcpyAsync(id_out, id.p, m.n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  RT_CUDA(cudaStreamSynchronize(h->stream));
  return RT_OK;
}
