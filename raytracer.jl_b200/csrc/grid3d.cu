// grid3d.cu -- 3-D structured grid: coordinates, implicit star-L adjacency and the BFM relaxation (sm_100a).
//
// Takes over grid(c0,c1,nnods) src/StructuredGrid.jl:35-45, CartesianIndex :90-96 (x fastest),
// spherical2cart :225-235, nodal_incidence(gr; neighbour_levels) :177-223, distance3D :239-243,
// edge_weight src/SSSP/weights.jl:20 and the control flow of BFM / foo! / goo! src/Dijsktra.jl:294-343,376-403.
//
// The reference materialises Dict{Int,Set{Int}} with ~125 entries per node; here the adjacency is implicit:
// star-L of a box grid is the clipped (2*2^L+1)^3 window (every expansion round of nodal_incidence unions the
// neighbours' CURRENT sets, so the radius doubles per level; self included once L >= 1), so no adjacency arrays
// exist at all.  A CTA owns one 8x4x4 tile of nodes, stages the tile plus its halo of (X,Y,Z,U,dist0) in
// shared memory once, and every thread scans its window in ascending linear id (the canonical scan order:
// the reference iterates a Julia Set whose order is not reproducible).  Frontier = active tile list.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int TX = 8, TY = 4, TZ = 4;
constexpr int TILE_THREADS = TX * TY * TZ;
constexpr unsigned FULL = 0xffffffffu;

}  // namespace

struct TpSlot;  // per-source state of a tile-pull solve running on its own stream (batches)

struct Grid3D {
  i64 nn[3] = {0, 0, 0};
  i64 n = 0;
  int star_levels = 1;
  int w = 2;  // window half width = 2^star_levels (StructuredGrid.jl:204-212)
  int self = 1;
  int coord_system = 0;
  double c0[3], c1[3];
  DevBuf<double> X, Y, Z;
  DevBuf<double> Xf, Yf, Zf;  // Float32-rounded coordinates (precision = 32), built on first use
  DevBuf<double> ax, ay, az;  // raw axis coordinates gr.x, gr.y, gr.z (getindex / closest_point), built on first use
  i64 tn[3] = {0, 0, 0};
  i64 n_tiles = 0;
  i64 graph_edges = 0;
  // workspace
  DevBuf<double> dist, dist0;
  DevBuf<i32> prev;
  DevBuf<unsigned> improved;  // per tile: 0 = nothing improved, else bit 31 | bounding box of the improved nodes
  DevBuf<i32> act[2];
  DevBuf<u64> counters;
  u64* counters_host = nullptr;
  bool ws_ready = false;
  // ---- near-far (push) schedule: items = runs of 32 consecutive x-nodes of one (y, z) grid line
  i64 nbx = 0, n_items = 0;
  DevBuf<unsigned> pend_mask, far_mask, infar, cur_mask;
  DevBuf<i32> nearq[2], farq[2];
  DevBuf<double> tau;
  DevBuf<int> ctl;
  DevBuf<i32> unresolved;
  bool push_ready = false;
  // ---- near-far, tile-pull rounds (tile = the 8x4x4 block of the Jacobi kernel; 4 mask words per tile, word = z-plane)
  DevBuf<unsigned> tpend, trel;    // [n_tiles * 4] improved since the last release / released in the current round
  DevBuf<unsigned> tmark;          // [n_tiles] round stamp of the last activation as a target tile
  DevBuf<unsigned> tmaxhi;         // [n_tiles] 1 + high word of the tile's largest travel time
  bool pull_ready = false;
  void* tp_graph = nullptr;        // cudaGraphExec_t of check_every rounds
  std::vector<TpSlot*> tp_slots;   // batches: several sources in flight, one slot + stream each
  std::vector<char> tp_graph_key;
  CanonWs* canon = nullptr;  // canonical-predecessor pass (option canonical_prev)
};

namespace {

struct P3 {
  const double* __restrict__ X;
  const double* __restrict__ Y;
  const double* __restrict__ Z;
  const double* __restrict__ U;
  double* dist;
  double* dist0;
  i32* prev;
  unsigned* improved;
  u64* counters;
  int nx, ny, nz;
  int tnx, tny, tnz;
  int w, self;
  int wmode;  // 0: weights.jl:20, 1: Dijsktra.jl:388 (exact.h)
};

// Julia Base lerpi (LinRange element): t = j/d; (1-t)*a + t*b
__device__ __forceinline__ double lerpi_dev(i64 j, i64 d, double a, double b) {
  if (d <= 0) return a;
  const double t = __ddiv_rn((double)j, (double)d);
  return __dadd_rn(__dmul_rn(__dsub_rn(1.0, t), a), __dmul_rn(t, b));
}

__global__ void coords3d_kernel(double c0x, double c0y, double c0z, double c1x, double c1y, double c1z, int nx,
                                int ny, int nz, int coord_system, double* __restrict__ X, double* __restrict__ Y,
                                double* __restrict__ Z) {
  const i64 I = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 n = (i64)nx * ny * nz;
  if (I >= n) return;
  const int i = (int)(I % nx), j = (int)((I / nx) % ny), k = (int)(I / ((i64)nx * ny));
  const double a = lerpi_dev(i, nx - 1, c0x, c1x);
  const double b = lerpi_dev(j, ny - 1, c0y, c1y);
  const double c = lerpi_dev(k, nz - 1, c0z, c1z);
  if (coord_system == 0) {
    X[I] = a;
    Y[I] = b;
    Z[I] = c;
  } else {
    // spherical2cart(theta=a, phi=b, r=c): x = r*cos(phi)*sin(theta), y = r*sin(phi)*sin(theta), z = r*cos(theta)
    const double st = sin(a), ct = cos(a), sp = sin(b), cp = cos(b);
    X[I] = __dmul_rn(__dmul_rn(c, cp), st);
    Y[I] = __dmul_rn(__dmul_rn(c, sp), st);
    Z[I] = __dmul_rn(c, ct);
  }
}

// dist0[J] + distance3D(pI,pJ) * (1/abs(UI+UJ)) * 2   (weights.jl:20, StructuredGrid.jl:239-241)
// exact_cand3<F32> (exact.h); F32: every operation rounded to Float32 (precision = 32)
template <bool F32 = false>
__device__ __forceinline__ double cand3(double dj, double xi, double yi, double zi, double ui, double xj,
                                        double yj, double zj, double uj, int wmode) {
  return exact_cand3<F32>(dj, xi, yi, zi, ui, xj, yj, zj, uj, wmode);
}

// One CTA per active tile; dynamic smem = 5 * SX*SY*SZ doubles.
template <bool F32>
__global__ void __launch_bounds__(TILE_THREADS) relax3d_kernel(P3 p, const i32* __restrict__ active, int cur) {
  extern __shared__ double sm[];
  const int w = p.w;
  const int SX = TX + 2 * w, SY = TY + 2 * w, SZ = TZ + 2 * w;
  const int SN = SX * SY * SZ;
  double* sX = sm;
  double* sY = sX + SN;
  double* sZ = sY + SN;
  double* sU = sZ + SN;
  double* sD = sU + SN;
  const i64 n_active = (i64)p.counters[cur];
  const int tid = threadIdx.x;
  const int lx = tid % TX, ly = (tid / TX) % TY, lz = tid / (TX * TY);
  u64 evals = 0, updates = 0;
  for (i64 a = blockIdx.x; a < n_active; a += gridDim.x) {
    const int tile = active[a];
    const int tx = tile % p.tnx, ty = (tile / p.tnx) % p.tny, tz = tile / (p.tnx * p.tny);
    const int ox = tx * TX - w, oy = ty * TY - w, oz = tz * TZ - w;  // global coords of smem cell (0,0,0)
    __syncthreads();  // previous iteration's readers are done
    for (int c = tid; c < SN; c += TILE_THREADS) {
      const int cx = c % SX, cy = (c / SX) % SY, cz = c / (SX * SY);
      const int gx = ox + cx, gy = oy + cy, gz = oz + cz;
      if (gx >= 0 && gx < p.nx && gy >= 0 && gy < p.ny && gz >= 0 && gz < p.nz) {
        const i64 J = (i64)gx + (i64)p.nx * ((i64)gy + (i64)p.ny * gz);
        sX[c] = p.X[J];
        sY[c] = p.Y[J];
        sZ[c] = p.Z[J];
        sU[c] = p.U[J];
        sD[c] = p.dist0[J];
      }
    }
    __syncthreads();
    const int gx = tx * TX + lx, gy = ty * TY + ly, gz = tz * TZ + lz;
    if (gx < p.nx && gy < p.ny && gz < p.nz) {
      const i64 I = (i64)gx + (i64)p.nx * ((i64)gy + (i64)p.ny * gz);
      const int ci = (lx + w) + SX * ((ly + w) + SY * (lz + w));
      const double xi = sX[ci], yi = sY[ci], zi = sZ[ci], ui = sU[ci];
      double best = sD[ci];
      i64 bid = -1;
      const int x0 = max(0, gx - w), x1 = min(p.nx - 1, gx + w);
      const int y0 = max(0, gy - w), y1 = min(p.ny - 1, gy + w);
      const int z0 = max(0, gz - w), z1 = min(p.nz - 1, gz + w);
      for (int zz = z0; zz <= z1; ++zz)
        for (int yy = y0; yy <= y1; ++yy) {
          const int rowc = SX * ((yy - oy) + SY * (zz - oz)) - ox;
          for (int xx = x0; xx <= x1; ++xx) {
            const int c = rowc + xx;
            if (!p.self && c == ci) continue;
            const double dj = sD[c];
            if (!(dj < best)) continue;  // dj + w >= dj >= best (also dj == Inf)
            const double xj = sX[c], yj = sY[c], zj = sZ[c], uj = sU[c];
            {
              const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj), dz = __dsub_rn(zi, zj);
              const double d2 = __fma_rn(dx, dx, __fma_rn(dy, dy, dz * dz));
              if (screen_cannot_improve_t<F32>(best, dj, d2, screen_ssum3(fabs(__dadd_rn(ui, uj)), p.wmode))) continue;
            }
            const double delta = cand3<F32>(dj, xi, yi, zi, ui, xj, yj, zj, uj, p.wmode);
            if (delta < best) {
              best = delta;
              bid = (i64)xx + (i64)p.nx * ((i64)yy + (i64)p.ny * zz);
            }
          }
        }
      p.dist[I] = best;
      if (bid >= 0) p.prev[I] = (i32)bid;
      evals += (u64)(x1 - x0 + 1) * (u64)(y1 - y0 + 1) * (u64)(z1 - z0 + 1) - (p.self ? 0u : 1u);
      updates += 1;
    }
  }
  // block-level sum of the counters
  for (int o = 16; o; o >>= 1) {
    evals += __shfl_xor_sync(FULL, evals, o);
    updates += __shfl_xor_sync(FULL, updates, o);
  }
  if ((tid & 31) == 0 && updates) {
    atomicAdd(&p.counters[2], evals);
    atomicAdd(&p.counters[3], updates);
  }
}

// goo! + copyto!(dist0, dist) over the active tiles; flags tiles that hold an improved node.
__global__ void __launch_bounds__(TILE_THREADS) commit3d_kernel(P3 p, const i32* __restrict__ active, int cur) {
  const i64 n_active = (i64)p.counters[cur];
  const int tid = threadIdx.x;
  const int lx = tid % TX, ly = (tid / TX) % TY, lz = tid / (TX * TY);
  for (i64 a = blockIdx.x; a < n_active; a += gridDim.x) {
    const int tile = active[a];
    const int tx = tile % p.tnx, ty = (tile / p.tnx) % p.tny, tz = tile / (p.tnx * p.tny);
    const int gx = tx * TX + lx, gy = ty * TY + ly, gz = tz * TZ + lz;
    bool imp = false;
    if (gx < p.nx && gy < p.ny && gz < p.nz) {
      const i64 I = (i64)gx + (i64)p.nx * ((i64)gy + (i64)p.ny * gz);
      const double d = p.dist[I];
      if (d < p.dist0[I]) {
        p.dist0[I] = d;
        imp = true;
      }
    }
    // bounding box (tile-local coordinates) of the improved nodes: lets activate3d skip neighbours that the
    // +-w window of those nodes cannot reach
    __shared__ int s_box[6];
    if (tid < 6) s_box[tid] = (tid & 1) ? -1 : 99;  // [minx, maxx, miny, maxy, minz, maxz]
    __syncthreads();
    if (imp) {
      atomicMin(&s_box[0], lx);
      atomicMax(&s_box[1], lx);
      atomicMin(&s_box[2], ly);
      atomicMax(&s_box[3], ly);
      atomicMin(&s_box[4], lz);
      atomicMax(&s_box[5], lz);
    }
    __syncthreads();
    if (tid == 0 && s_box[1] >= 0)
      p.improved[tile] = 0x80000000u | (unsigned)s_box[0] | ((unsigned)s_box[1] << 3) | ((unsigned)s_box[2] << 6) |
                         ((unsigned)s_box[3] << 8) | ((unsigned)s_box[4] << 10) | ((unsigned)s_box[5] << 12);
    __syncthreads();
  }
}

// a tile joins the frontier iff a tile within reach of the window (ceil(w/T) tiles per axis) improved
__global__ void activate3d_kernel(P3 p, i64 n_tiles, i32* __restrict__ next_active, int nxt) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  bool act = false;
  if (t < n_tiles) {
    const int tx = (int)(t % p.tnx), ty = (int)((t / p.tnx) % p.tny), tz = (int)(t / ((i64)p.tnx * p.tny));
    const int rx = (p.w + TX - 1) / TX, ry = (p.w + TY - 1) / TY, rz = (p.w + TZ - 1) / TZ;
    for (int zz = max(0, tz - rz); zz <= min(p.tnz - 1, tz + rz) && !act; ++zz)
      for (int yy = max(0, ty - ry); yy <= min(p.tny - 1, ty + ry) && !act; ++yy)
        for (int xx = max(0, tx - rx); xx <= min(p.tnx - 1, tx + rx); ++xx) {
          const unsigned b = p.improved[(i64)xx + (i64)p.tnx * ((i64)yy + (i64)p.tny * zz)];
          if (!b) continue;
          // improved box of that tile, grown by the window half width, against this tile's node range
          const int x0 = xx * TX + (int)(b & 7u) - p.w, x1 = xx * TX + (int)((b >> 3) & 7u) + p.w;
          const int y0 = yy * TY + (int)((b >> 6) & 3u) - p.w, y1 = yy * TY + (int)((b >> 8) & 3u) + p.w;
          const int z0 = zz * TZ + (int)((b >> 10) & 3u) - p.w, z1 = zz * TZ + (int)((b >> 12) & 3u) + p.w;
          if (x0 <= tx * TX + TX - 1 && x1 >= tx * TX && y0 <= ty * TY + TY - 1 && y1 >= ty * TY &&
              z0 <= tz * TZ + TZ - 1 && z1 >= tz * TZ) {
            act = true;
            break;
          }
        }
  }
  const unsigned ball = __ballot_sync(FULL, act);
  if (ball) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(ball) - 1;
    u64 base = 0;
    if (lane == leader) base = atomicAdd(&p.counters[nxt], (u64)__popc(ball));
    base = __shfl_sync(FULL, base, leader);
    if (act) next_active[base + __popc(ball & ((1u << lane) - 1u))] = (i32)t;
  }
}

__global__ void init3d_kernel(double* __restrict__ dist, double* __restrict__ dist0, i32* __restrict__ prev, i64 n,
                              i64 source) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = (i == source) ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
  dist[i] = v;
  dist0[i] = v;
  prev[i] = -1;
}

int ensure_ws3(rt_mesh* h) {
  Grid3D& g = *h->g3;
  if (g.ws_ready) return RT_OK;
  RT_TRY(g.dist.alloc(g.n));
  RT_TRY(g.dist0.alloc(g.n));
  RT_TRY(g.prev.alloc(g.n));
  RT_TRY(g.improved.alloc(g.n_tiles));
  RT_TRY(g.act[0].alloc(g.n_tiles));
  RT_TRY(g.act[1].alloc(g.n_tiles));
  RT_TRY(g.counters.alloc(16));
  RT_CUDA(cudaMallocHost((void**)&g.counters_host, 16 * sizeof(u64)));
  g.ws_ready = true;
  return RT_OK;
}

}  // namespace

int grid3d_build(rt_mesh* h, const double c0[3], const double c1[3], const i64 nn[3], int star_levels,
                 int coord_system) {
  RT_ARG(c0 && c1 && nn, "null argument");
  RT_ARG(nn[0] >= 1 && nn[1] >= 1 && nn[2] >= 1, "grid needs at least one node per axis");
  RT_ARG(star_levels >= 0 && star_levels <= 3, "star_levels must be in 0..3");
  if (star_levels == 3) {
    // radius 2^3 = 8: a 17^3 window (4913 candidates per node); the tile + halo of the relax kernel would need
    // 384 KB of shared memory.  Not built: the reference itself cannot hold that Dict beyond toy grids.
    rt_set_error("neighbour_levels = 3 (17^3 window) is not supported; use 0, 1 or 2");
    return RT_ERR_UNSUPPORTED;
  }
  RT_ARG(coord_system == 0 || coord_system == 1, "coord_system must be 0 or 1");
  const i64 n = nn[0] * nn[1] * nn[2];
  RT_ARG(n < (i64)2000000000 && nn[0] < 65536 * 16 && nn[1] < 65536 * 16 && nn[2] < 65536 * 16, "grid too large");
  Grid3D* gp = new Grid3D();
  h->g3 = gp;
  h->kind = 3;
  Grid3D& g = *gp;
  for (int d = 0; d < 3; ++d) {
    g.nn[d] = nn[d];
    g.c0[d] = c0[d];
    g.c1[d] = c1[d];
  }
  g.n = n;
  g.star_levels = star_levels;
  g.w = 1 << star_levels;
  g.self = star_levels >= 1 ? 1 : 0;
  g.coord_system = coord_system;
  g.tn[0] = (nn[0] + TX - 1) / TX;
  g.tn[1] = (nn[1] + TY - 1) / TY;
  g.tn[2] = (nn[2] + TZ - 1) / TZ;
  g.n_tiles = g.tn[0] * g.tn[1] * g.tn[2];
  RT_TRY(g.X.alloc(n));
  RT_TRY(g.Y.alloc(n));
  RT_TRY(g.Z.alloc(n));
  coords3d_kernel<<<grid_for(n, 256), 256, 0, h->stream>>>(c0[0], c0[1], c0[2], c1[0], c1[1], c1[2], (int)nn[0],
                                                          (int)nn[1], (int)nn[2], coord_system, g.X.p, g.Y.p, g.Z.p);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaStreamSynchronize(h->stream));
  // E_graph = prod_d sum_i |clipped window_d(i)|  (- n when self is excluded)
  i64 s[3];
  for (int d = 0; d < 3; ++d) {
    s[d] = 0;
    for (i64 i = 0; i < nn[d]; ++i)
      s[d] += std::min(nn[d] - 1, i + g.w) - std::max<i64>(0, i - g.w) + 1;
  }
  g.graph_edges = s[0] * s[1] * s[2] - (g.self ? 0 : n);
  return RT_OK;
}

int grid3d_export(const rt_mesh* h, double* X, double* Y, double* Z) {
  const Grid3D& g = *h->g3;
  if (X) RT_CUDA(cudaMemcpy(X, g.X.p, g.n * sizeof(double), cudaMemcpyDeviceToHost));
  if (Y) RT_CUDA(cudaMemcpy(Y, g.Y.p, g.n * sizeof(double), cudaMemcpyDeviceToHost));
  if (Z) RT_CUDA(cudaMemcpy(Z, g.Z.p, g.n * sizeof(double), cudaMemcpyDeviceToHost));
  return RT_OK;
}

void tp_slots_free(Grid3D& g);

void grid3d_free(rt_mesh* h) {
  if (h->g3) {
    tp_slots_free(*h->g3);
    if (h->g3->counters_host) cudaFreeHost(h->g3->counters_host);
    if (h->g3->canon) canon_ws_free(h->g3->canon);
    if (h->g3->tp_graph) cudaGraphExecDestroy((cudaGraphExec_t)h->g3->tp_graph);
    delete h->g3;
  }
  h->g3 = nullptr;
}

int grid3d_coords(const rt_mesh* h, const double** X, const double** Y, const double** Z, const double** none) {
  *X = h->g3->X.p;
  *Y = h->g3->Y.p;
  *Z = h->g3->Z.p;
  *none = nullptr;
  return RT_OK;
}

int grid3d_n(const rt_mesh* h, i64* n) {
  *n = h->g3->n;
  return RT_OK;
}


// =========================================================================================================
// Near-far push schedule on the implicit star-L graph (see bfm2d_push.cu for the scheme).  Work item = 32
// consecutive x-nodes of one (y, z) grid line; a released item pushes into the (32 + 2w) x (2w+1)^2 block around
// it: lanes walk the target x positions (coalesced loads of X, Y, Z, U, dist), each target is reached by the <= 2w+1
// released sources within +-w in x, broadcast from shared memory.
namespace {

struct Q3 {
  const double* __restrict__ X;
  const double* __restrict__ Y;
  const double* __restrict__ Z;
  const double* __restrict__ U;
  double* dist;
  i32* prev;
  unsigned* pend_mask;
  unsigned* far_mask;
  unsigned* infar;
  unsigned* cur_mask;
  u64* counters;  // [0],[1] near counts [2] evals [3] releases [4],[5] far counts [6] unresolved
  double* tau;    // [0] tau [1] delta [2] min far bits [3] scratch
  i32* nearq[2];
  i32* farq[2];
  int* ctl;       // [0] cur [1] fcur [2] mode [3] done [4] rounds [5] push rounds
  int nx, ny, nz, nbx;
  int w, self;
  int wmode;  // 0: weights.jl:20, 1: Dijsktra.jl:388 (exact.h)
  FastDiv fd_nx, fd_ny, fd_nbx, fd_W;  // node id -> (x, line), line -> (y, z), item -> (bx, line), unit -> (slot, dz)
};

// append to a list through its (hot) counter: one atomic per group of converged lanes instead of one per lane
__device__ __forceinline__ void list_append(i32* __restrict__ list, u64* counter, i32 v) {
  const unsigned act = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(act) - 1;
  u64 base = 0;
  if (lane == leader) base = atomicAdd(counter, (u64)__popc(act));
  base = __shfl_sync(act, base, leader);
  list[base + __popc(act & ((1u << lane) - 1u))] = v;
}

__device__ __forceinline__ void enqueue3(const Q3& p, i64 J, double d, double tau, i32* near_next, int nxt,
                                         i32* far_list, int fcur) {
  unsigned line, ix;
  p.fd_nx.divmod((unsigned)J, line, ix);
  const int it = (int)((ix >> 5) + (unsigned)p.nbx * line);
  const unsigned bit = 1u << (ix & 31);
  if (d < tau) {
    const unsigned old = atomicOr(&p.pend_mask[it], bit);
    if (old == 0u) list_append(near_next, &p.counters[nxt], it);
    if (__ldcg(&p.far_mask[it]) & bit) atomicAnd(&p.far_mask[it], ~bit);
  } else {
    atomicOr(&p.far_mask[it], bit);
    if (atomicExch(&p.infar[it], 1u) == 0u) list_append(far_list, &p.counters[4 + fcur], it);
  }
}

constexpr int P3_BLOCK = 128;

// Warp-level work unit = (near-list slot, dz): the warp keeps the 32 sources of its item in a private shared-memory
// slab (no block barrier), then walks the (32 + 2w) target columns of that z-plane; for each lane the loads of all
// 2w+1 target rows are issued back to back (memory-level parallelism) before any of them is evaluated.
template <bool F32>
__device__ __forceinline__ void push3d_body(const Q3& p, const i32* near_cur, int cur, i32* near_next,
                                            i32* far_list, int fcur) {
  __shared__ double sm_src[P3_BLOCK / 32][5][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double(*S)[32] = sm_src[warp];
  const int w = p.w, W = 2 * w + 1;
  const i64 n_near = (i64)__ldcg(&p.counters[cur]);
  const double tau = __ldcg(&p.tau[0]);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
  u64 evals = 0;
  for (i64 unit = gw; unit < n_near * W; unit += nw) {
    unsigned slot_u, dz_u, line_u, bx_u, sy_u, sz_u;
    p.fd_W.divmod((unsigned)unit, slot_u, dz_u);  // n_near * W < 2^32 (n_items * 7 at the largest accepted grid)
    const i64 slot = slot_u;
    const int dzi = (int)dz_u - w;
    const int it = __ldcg(&near_cur[slot]);
    p.fd_nbx.divmod((unsigned)it, line_u, bx_u);
    p.fd_ny.divmod(line_u, sz_u, sy_u);
    const int bx = (int)bx_u, sy = (int)sy_u, sz = (int)sz_u;
    const int tz = sz + dzi;
    if (tz < 0 || tz >= p.nz) continue;  // warp-uniform
    const unsigned mask = __ldcg(&p.cur_mask[slot]);
    if (mask == 0u) continue;
    __syncwarp();
    {
      const int gx = bx * 32 + lane;
      if (((mask >> lane) & 1u) && gx < p.nx) {
        const i64 I = (i64)gx + (i64)p.nx * ((i64)sy + (i64)p.ny * sz);
        S[0][lane] = p.X[I];
        S[1][lane] = p.Y[I];
        S[2][lane] = p.Z[I];
        S[3][lane] = p.U[I];
        S[4][lane] = __ldcg(&p.dist[I]);
      } else {
        S[4][lane] = INF;
      }
    }
    __syncwarp();
    // targets of this unit: rows y0..y1 times the columns within +-w of the released span (a thin wavefront
    // releases only one or two nodes of an x-line per round, so the block is usually ~5 x 5, not 36 x 5)
    const int first = __ffs(mask) - 1, last = 31 - __clz(mask);
    const int xa = max(0, bx * 32 + first - w), xb = min(p.nx - 1, bx * 32 + last + w);
    const int y0 = max(0, sy - w), y1 = min(p.ny - 1, sy + w);
    const int ncol = xb - xa + 1, nt = ncol * (y1 - y0 + 1);
    const float rcol = 1.0f / (float)ncol;
    for (int t = lane; t < nt; t += 32) {
      const int r = __float2int_rz(((float)t + 0.5f) * rcol);  // exact: t < 36 * 7, ncol <= 46
      const int tx = xa + (t - r * ncol), ty = y0 + r;
      const i64 J = (i64)tx + (i64)p.nx * ((i64)ty + (i64)p.ny * tz);
      const double dj = __ldcg(&p.dist[J]);
      const double xj = p.X[J], yj = p.Y[J], zj = p.Z[J], uj = p.U[J];
      double best = dj;
      const int q0 = max(tx - w, bx * 32) - bx * 32, q1 = min(tx + w, bx * 32 + 31) - bx * 32;
      // released sources inside the window of this target
      unsigned wm = (mask >> q0) & (0xffffffffu >> (31 - (q1 - q0)));
      while (wm) {
        const int q = q0 + __ffs(wm) - 1;
        wm &= wm - 1u;
        const double di = S[4][q];
        if (!(di < best)) continue;
        if (!p.self && ty == sy && dzi == 0 && bx * 32 + q == tx) continue;
        const double dx = __dsub_rn(S[0][q], xj), dy = __dsub_rn(S[1][q], yj), dz = __dsub_rn(S[2][q], zj);
        const double d2 = __fma_rn(dx, dx, __fma_rn(dy, dy, dz * dz));
        if (screen_cannot_improve_t<F32>(best, di, d2, screen_ssum3(fabs(__dadd_rn(S[3][q], uj)), p.wmode))) continue;
        const double delta = cand3<F32>(di, S[0][q], S[1][q], S[2][q], S[3][q], xj, yj, zj, uj, p.wmode);
        best = delta < best ? delta : best;
      }
      if (best < dj) {
        const u64 bits = (u64)__double_as_longlong(best);
        const u64 old = atomicMin((u64*)&p.dist[J], bits);
        if (bits < old) enqueue3(p, J, best, tau, near_next, cur ^ 1, far_list, fcur);
      }
    }
    if (lane == 0) {
      // evaluations of this z-plane: per released source, clipped x-extent times clipped y-extent
      u64 e = 0;
      const int ycnt = y1 - y0 + 1;
      if (bx * 32 + first - w >= 0 && bx * 32 + last + w <= p.nx - 1) {
        e = (u64)__popc(mask) * (u64)(2 * w + 1) * (u64)ycnt;  // interior item: full x-extent for every source
      } else {
        for (int q = first; q <= last; ++q)
          if ((mask >> q) & 1u) {
            const int gx = bx * 32 + q;
            e += (u64)(min(p.nx - 1, gx + w) - max(0, gx - w) + 1) * (u64)ycnt;
          }
      }
      evals += e;
      if (dzi == 0) atomicAdd(&p.counters[3], (u64)__popc(mask));
    }
  }
  if (lane == 0 && evals) atomicAdd(&p.counters[2], evals);
}

// after_far: the far kernels were enqueued since the previous round_begin (they only are every FAR3_EVERY-th
// round); a requested threshold advance (mode 2) waits for them, prep / push return at once meanwhile
constexpr int FAR3_EVERY = 3;
#ifndef TP_DELTA_FACTOR
#define TP_DELTA_FACTOR 8.0
#endif
#ifndef TP_EARLY
#define TP_EARLY 0.0  // x n^(2/3) releases per round (measured on 216^3 / 368^3: no gain at 0.25 .. 4)
#endif
__global__ void round_begin3_kernel(Q3 p, int after_far) {
  int* c = p.ctl;
  if (c[3]) return;
  if (c[2] == 2 && !after_far) return;
  if (c[2] == 1)
    c[0] ^= 1;
  else if (c[2] == 2)
    c[1] ^= 1;
  const int cur = c[0], fcur = c[1];
  const u64 n_near = p.counters[cur], n_far = p.counters[4 + fcur];
  if (n_near == 0 && n_far == 0) {
    c[2] = 0;
    c[3] = 1;
    return;
  }
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
__global__ void prep3_kernel(Q3 p) {
  if (p.ctl[2] != 1) return;
  const int cur = p.ctl[0];
  const i32* near_cur = p.nearq[cur];
  const i64 n = (i64)p.counters[cur];
  for (i64 slot = (i64)blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += (i64)gridDim.x * blockDim.x)
    p.cur_mask[slot] = atomicExch(&p.pend_mask[near_cur[slot]], 0u);
}
template <bool F32>
__global__ void __launch_bounds__(P3_BLOCK) push3d_kernel(Q3 p) {
  if (p.ctl[2] != 1) return;
  const int cur = p.ctl[0], fcur = p.ctl[1];
  push3d_body<F32>(p, p.nearq[cur], cur, p.nearq[cur ^ 1], p.farq[fcur], fcur);
}
__device__ __forceinline__ i64 item_node0(const Q3& p, int it) {
  const int bx = it % p.nbx;
  const i64 line = it / p.nbx;
  return (i64)bx * 32 + (i64)p.nx * line;
}
__global__ void far_min3_kernel(Q3 p) {
  if (p.ctl[2] != 2) return;
  const int fcur = p.ctl[1];
  const i32* far_cur = p.farq[fcur];
  const i64 nslots = (i64)p.counters[4 + fcur];
  const int lane = threadIdx.x & 31;
  u64 best = ~0ull;
  for (i64 slot = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < nslots;
       slot += ((i64)gridDim.x * blockDim.x) >> 5) {
    const int it = far_cur[slot];
    const unsigned m = p.far_mask[it];
    if ((m >> lane) & 1u) {
      const u64 b = (u64)__double_as_longlong(p.dist[item_node0(p, it) + lane]);
      best = b < best ? b : best;
    }
  }
  for (int o = 16; o; o >>= 1) {
    const u64 other = __shfl_xor_sync(FULL, best, o);
    best = other < best ? other : best;
  }
  if (lane == 0 && best != ~0ull) atomicMin((u64*)&p.tau[2], best);
}
__global__ void far_release3_kernel(Q3 p) {
  if (p.ctl[2] != 2) return;
  const int cur = p.ctl[0], fcur = p.ctl[1];
  const i32* far_cur = p.farq[fcur];
  i32* far_next = p.farq[fcur ^ 1];
  i32* near_next = p.nearq[cur];
  const i64 nslots = (i64)p.counters[4 + fcur];
  const int lane = threadIdx.x & 31;
  const double tau = __dadd_rn(p.tau[2], p.tau[1]);
  for (i64 slot = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < nslots;
       slot += ((i64)gridDim.x * blockDim.x) >> 5) {
    const int it = far_cur[slot];
    const unsigned m = p.far_mask[it];
    const bool mine = (m >> lane) & 1u;
    const bool rel = mine && p.dist[item_node0(p, it) + lane] < tau;
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
      if (slot == 0) p.tau[0] = tau;
    }
  }
}
__global__ void push3_init_kernel(Q3 p, i64 n, i64 source, double delta) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    p.dist[i] = (i == source) ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
    p.prev[i] = -1;
  }
  if (i == 0) {
    const int ix = (int)(source % p.nx);
    const i64 line = source / p.nx;
    const int it = (int)((ix >> 5) + (i64)p.nbx * line);
    p.pend_mask[it] = 1u << (ix & 31);
    p.nearq[0][0] = it;
    p.counters[0] = 1ull;
    p.tau[0] = delta;
    p.tau[1] = delta;
  }
}
// mean travel time along the cell diagonal (node -> node + (1,1,1)): sizes the bucket
__global__ void wdiag3_kernel(Q3 p, double* __restrict__ sum, u64* __restrict__ cnt) {
  const i64 I = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 7;  // sample every 7th node
  double wt = 0.0;
  const i64 n = (i64)p.nx * p.ny * p.nz;
  if (I < n) {
    const int i = (int)(I % p.nx), j = (int)((I / p.nx) % p.ny), k = (int)(I / ((i64)p.nx * p.ny));
    const int i2 = min(i + 1, p.nx - 1), j2 = min(j + 1, p.ny - 1), k2 = min(k + 1, p.nz - 1);
    const i64 J = (i64)i2 + (i64)p.nx * ((i64)j2 + (i64)p.ny * k2);
    if (J != I) wt = cand3(0.0, p.X[I], p.Y[I], p.Z[I], p.U[I], p.X[J], p.Y[J], p.Z[J], p.U[J], p.wmode);
    if (!(wt == wt) || wt > 1e300) wt = 0.0;
  }
  const unsigned has = __ballot_sync(FULL, wt > 0.0);
  for (int o = 16; o; o >>= 1) wt += __shfl_xor_sync(FULL, wt, o);
  if ((threadIdx.x & 31) == 0 && has) {
    atomicAdd(sum, wt);
    atomicAdd(cnt, (u64)__popc(has));
  }
}
// predecessors: first candidate in ascending linear id (the canonical scan order) with dist[J] < dist[I] that is
// bit-exactly tight.  Thread per node, window read straight from global memory (L2-resident neighbourhood).
template <bool F32>
__global__ void prev_tight3_kernel(Q3 p, i64 n, i64 source) {
  const i64 I = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= n) return;
  const double di = p.dist[I];
  if (!(di < __longlong_as_double(0x7ff0000000000000LL)) || I == source) return;
  const int i = (int)(I % p.nx), j = (int)((I / p.nx) % p.ny), k = (int)(I / ((i64)p.nx * p.ny));
  const int w = p.w;
  const double xi = p.X[I], yi = p.Y[I], zi = p.Z[I], ui = p.U[I];
  for (int zz = max(0, k - w); zz <= min(p.nz - 1, k + w); ++zz)
    for (int yy = max(0, j - w); yy <= min(p.ny - 1, j + w); ++yy)
      for (int xx = max(0, i - w); xx <= min(p.nx - 1, i + w); ++xx) {
        const i64 J = (i64)xx + (i64)p.nx * ((i64)yy + (i64)p.ny * zz);
        const double dj = p.dist[J];
        if (!(dj < di)) continue;
        const double xj = p.X[J], yj = p.Y[J], zj = p.Z[J], uj = p.U[J];
        const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj), dz = __dsub_rn(zi, zj);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (!screen_maybe_tight_t<F32>(di, dj, d2, screen_ssum3(fabs(__dadd_rn(ui, uj)), p.wmode))) continue;
        if (cand3<F32>(dj, xi, yi, zi, ui, xj, yj, zj, uj, p.wmode) == di) {
          p.prev[I] = (i32)J;
          return;
        }
      }
}

int grid3d_prepare_f32(rt_mesh* h) {
  Grid3D& g = *h->g3;
  if (g.Xf.n == (size_t)g.n) return RT_OK;
  RT_TRY(g.Xf.alloc(g.n));
  RT_TRY(g.Yf.alloc(g.n));
  RT_TRY(g.Zf.alloc(g.n));
  RT_TRY(round_to_f32_device(g.X.p, g.Xf.p, g.n, h->stream));
  RT_TRY(round_to_f32_device(g.Y.p, g.Yf.p, g.n, h->stream));
  RT_TRY(round_to_f32_device(g.Z.p, g.Zf.p, g.n, h->stream));
  return RT_OK;
}

int ensure_push3(rt_mesh* h) {
  Grid3D& g = *h->g3;
  if (g.push_ready) return RT_OK;
  g.nbx = (g.nn[0] + 31) / 32;
  g.n_items = g.nbx * g.nn[1] * g.nn[2];
  RT_TRY(g.pend_mask.alloc(g.n_items));
  RT_TRY(g.far_mask.alloc(g.n_items));
  RT_TRY(g.infar.alloc(g.n_items));
  RT_TRY(g.cur_mask.alloc(g.n_items));
  for (int k = 0; k < 2; ++k) {
    RT_TRY(g.nearq[k].alloc(g.n_items));
    RT_TRY(g.farq[k].alloc(g.n_items));
  }
  RT_TRY(g.tau.alloc(4));
  RT_TRY(g.ctl.alloc(8));
  g.push_ready = true;
  return RT_OK;
}

}  // namespace

int bfm3d_solve_push(rt_mesh* h, const double* U_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                     rt_stats* stats) {
  Grid3D& g = *h->g3;
  cudaStream_t s = h->stream;
  RT_TRY(ensure_ws3(h));
  RT_TRY(ensure_push3(h));
  const i64 n = g.n;
  const bool f32 = h->f32;
  if (f32) RT_TRY(grid3d_prepare_f32(h));
  Q3 p;
  p.X = f32 ? g.Xf.p : g.X.p;
  p.Y = f32 ? g.Yf.p : g.Y.p;
  p.Z = f32 ? g.Zf.p : g.Z.p;
  p.U = U_dev;
  p.dist = g.dist.p;
  p.prev = g.prev.p;
  p.pend_mask = g.pend_mask.p;
  p.far_mask = g.far_mask.p;
  p.infar = g.infar.p;
  p.cur_mask = g.cur_mask.p;
  p.counters = g.counters.p;
  p.tau = g.tau.p;
  p.nearq[0] = g.nearq[0].p;
  p.nearq[1] = g.nearq[1].p;
  p.farq[0] = g.farq[0].p;
  p.farq[1] = g.farq[1].p;
  p.ctl = g.ctl.p;
  p.nx = (int)g.nn[0];
  p.ny = (int)g.nn[1];
  p.nz = (int)g.nn[2];
  p.nbx = (int)g.nbx;
  p.w = g.w;
  p.self = g.self;
  p.wmode = h->opts.weight3d;
  p.fd_nx = FastDiv((unsigned)g.nn[0]);
  p.fd_ny = FastDiv((unsigned)g.nn[1]);
  p.fd_nbx = FastDiv((unsigned)g.nbx);
  p.fd_W = FastDiv((unsigned)(2 * g.w + 1));
  RT_ARG(g.n_items * (2 * g.w + 1) < ((i64)1 << 32), "grid too large for the near-far schedule");
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  const unsigned gbig = (unsigned)(sm_count * 16), gsmall = (unsigned)(sm_count * 2);
  u64* ch = g.counters_host;
  cudaEvent_t ev0, ev1, evr0, evr1;
  RT_CUDA(cudaEventCreate(&ev0));
  RT_CUDA(cudaEventCreate(&ev1));
  RT_CUDA(cudaEventCreate(&evr0));
  RT_CUDA(cudaEventCreate(&evr1));
  rt_stats st = {};
  st.graph_edges = g.graph_edges;
  int rc = RT_OK;
  const bool timers = h->opts.profile_timers != 0;
  double delta = h->opts.delta;
  if (!(delta > 0.0)) {
    RT_CUDA(cudaMemsetAsync(g.tau.p + 3, 0, sizeof(double), s));
    RT_CUDA(cudaMemsetAsync(g.counters.p + 7, 0, sizeof(u64), s));
    wdiag3_kernel<<<grid_for((n + 6) / 7, 256), 256, 0, s>>>(p, g.tau.p + 3, g.counters.p + 7);
    double wsum = 0.0;
    u64 wc = 0;
    RT_CUDA(cudaMemcpyAsync(&wsum, g.tau.p + 3, sizeof(double), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaMemcpyAsync(&wc, g.counters.p + 7, sizeof(u64), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
    const double wmean = wc ? wsum / (double)wc : 1.0;
    // measured on B200 (216^3): rounds are latency-bound, 8 cell diagonals per bucket minimise the solve time
    delta = wmean * (h->opts.delta_factor > 0.0 ? h->opts.delta_factor : 8.0);
  }
  for (i64 si = 0; si < nsrc && rc == RT_OK; ++si) {
    const i64 src1 = sources[si];
    if (src1 < 1 || src1 > n) {
      rt_set_error("source %lld out of range 1..%lld", (long long)src1, (long long)n);
      rc = RT_ERR_ARG;
      break;
    }
    const i64 src = src1 - 1;
    cudaEventRecord(ev0, s);
    cudaMemsetAsync(g.counters.p, 0, 8 * sizeof(u64), s);
    cudaMemsetAsync(g.pend_mask.p, 0, g.n_items * sizeof(unsigned), s);
    cudaMemsetAsync(g.far_mask.p, 0, g.n_items * sizeof(unsigned), s);
    cudaMemsetAsync(g.infar.p, 0, g.n_items * sizeof(unsigned), s);
    cudaMemsetAsync(g.ctl.p, 0, 8 * sizeof(int), s);
    push3_init_kernel<<<grid_for(n, 256), 256, 0, s>>>(p, n, src, delta);
    st.total_launches += 1;
    int hctl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int R = timers ? 1 : (h->opts.check_every > 1 ? h->opts.check_every : 32);
    int after_far = 1;
    i64 enq_rounds = 0;
    while (!hctl[3]) {
      for (int r = 0; r < R; ++r) {
        round_begin3_kernel<<<1, 1, 0, s>>>(p, after_far);
        prep3_kernel<<<gsmall, 256, 0, s>>>(p);
        if (timers) cudaEventRecord(evr0, s);
        if (f32)
          push3d_kernel<true><<<gbig, P3_BLOCK, 0, s>>>(p);
        else
          push3d_kernel<false><<<gbig, P3_BLOCK, 0, s>>>(p);
        if (timers) cudaEventRecord(evr1, s);
        st.total_launches += 3;
        after_far = 0;
        if (r % FAR3_EVERY == FAR3_EVERY - 1 || r == R - 1) {
          far_min3_kernel<<<gsmall, 256, 0, s>>>(p);
          far_release3_kernel<<<gsmall, 256, 0, s>>>(p);
          st.total_launches += 2;
          after_far = 1;
        }
      }
      enq_rounds += R;
      if (enq_rounds > ((i64)1 << 22)) {  // a solve needs 1e2 - 1e4 rounds: never spin forever on a logic error
        rc = RT_ERR_CUDA;
        break;
      }
      cudaMemcpyAsync(hctl, g.ctl.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        rc = RT_ERR_CUDA;
        break;
      }
      if (timers && hctl[2] == 1) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evr0, evr1);
        st.relax_ms += ms;
      }
    }
    if (rc != RT_OK) break;
    st.sweeps += hctl[4];
    st.relax_launches += hctl[5];

    cudaEventRecord(evr0, s);
    if (f32)
      prev_tight3_kernel<true><<<grid_for(n, 128), 128, 0, s>>>(p, n, src);
    else
      prev_tight3_kernel<false><<<grid_for(n, 128), 128, 0, s>>>(p, n, src);
    st.total_launches += 1;
    if (h->opts.canonical_prev) {  // the reference schedule's predecessors, ties included (canonical_prev.cu)
      Grid3Desc gd{p.X, p.Y, p.Z, p.nx, p.ny, p.nz, p.w, p.self, p.wmode};
      i64 launches = 0;
      rc = canonical_prev_3d(h, &g.canon, gd, U_dev, f32, g.dist.p, src, g.prev.p, &launches);
      st.total_launches += launches;
      if (rc != RT_OK) break;
    }
    cudaEventRecord(evr1, s);
    cudaMemcpyAsync(ch, g.counters.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
    cudaEventRecord(ev1, s);
    if (dist_dev) cudaMemcpyAsync(dist_dev + si * n, g.dist.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (prev_dev) cudaMemcpyAsync(prev_dev + si * n, g.prev.p, n * sizeof(i32), cudaMemcpyDeviceToDevice, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      rc = RT_ERR_CUDA;
      break;
    }
    st.relaxed_edges += (i64)ch[2];  // pushes only; the tightness pass is timed separately (prev_ms)
    st.vertex_updates += (i64)ch[3];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    st.kernel_ms += ms;
    cudaEventElapsedTime(&ms, evr0, evr1);
    st.prev_ms += ms;
  }
  cudaError_t e = cudaGetLastError();
  if (rc == RT_ERR_CUDA || e != cudaSuccess) {
    rt_set_error("CUDA failure in bfm3d_solve_push: %s", cudaGetErrorString(e));
    rc = RT_ERR_CUDA;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cudaEventDestroy(evr0);
  cudaEventDestroy(evr1);
  if (stats) *stats = st;
  return rc;
}

// =========================================================================================================
// Near-far schedule as tile-pull rounds (option tile_pull = 1, the default for schedule 1 on 3-D grids).
//
// Same label-correcting scheme as the push units above (threshold tau, bucket width delta, every node that improved
// and lies below tau is released in the next round), but the round is turned around: the TARGETS of a round pull from
// the released sources.  Pending / released sets are 128-bit masks per tile (tile = the 8x4x4 block of the Jacobi
// kernel; word = z-plane of the tile = warp of the CTA, bit = lane).  A tile within reach of a released node builds,
// from the released words of the <= 27 tiles around it, one bit mask per x-row of tile + halo ("is this neighbour a
// released source" becomes a shift and a mask), stages (travel time, X, Y, Z, U) of the released cells only, and every
// thread relaxes its own node from the released sources of its window.  A node is written by exactly one thread of
// one CTA: no atomics on the travel times, no per-node work lists; "improved" is one ballot per warp.  A source's
// travel time is read from dist[] itself: any value it ever held is a valid label, and a source that improves during
// the round is pending again.
//
// Pruning (all exact-safe: weights are >= 0): a tile whose largest travel time lies below the smallest released one
// of a source tile is not visited for it (tmaxhi: high words, rounded up); a thread whose own travel time lies below
// every released one of the staged block, or whose window holds no released source in a z-plane, skips the scan.
//
// Per round two launches: tp_release (scan the tile masks: nodes with pending improvement and dist < tau are released,
// the tiles within reach stamped and appended to the target list) and tp_pull (the relaxation; its last CTA does the
// round control: threshold advance to min(pending) + delta when nothing was released, termination).  The travel times
// are the least fixed point of the same monotone relaxation, hence bit-identical to the other schedules; predecessors
// come from the same tightness pass (prev_tight3_kernel / canonical_prev_3d).
namespace {

struct T3 {
  const double* __restrict__ X;
  const double* __restrict__ Y;
  const double* __restrict__ Z;
  const double* __restrict__ U;
  double* dist;
  unsigned* tpend;   // [n_tiles * 4] improved since the last release
  unsigned* trel;    // [n_tiles * 4] released in the current round
  unsigned* tmark;   // [n_tiles] round stamp of the last activation as a target tile
  unsigned* tmaxhi;  // [n_tiles] 1 + high word of the largest travel time of the tile (0xffffffff: never visited)
  i32* act;          // target tiles of the current round
  // [0] target tiles [1] released this round [2] nominal candidates [3] releases (total) [4] finished pull CTAs
  // [5] tile visits [7] min pending bits [8] screened [9] exact
  u64* counters;
  double* tau;  // [0] tau [1] delta
  int* ctl;     // [0] round stamp [3] done [4] rounds [5] rounds with releases
  int nx, ny, nz, tnx, tny, tnz;
  i64 n_tiles;
  int w, self, wmode, count;
  int ctl_tail;  // 1: the last CTA of tp_pull does the round control, 0: tp_ctl_kernel does
  u64 early;  // releases per round below which the threshold advances although the bucket is not empty
  FastDiv fd_tnx, fd_tny, fd_SY;
};

constexpr int TPR_BLOCK = 256;
static_assert(TX == 8 && TY == 4 && TZ == 4, "the tile masks assume 8 x 4 x 4 tiles (word = z-plane, byte = y-row)");

__global__ void __launch_bounds__(TPR_BLOCK) tp_release_kernel(T3 p) {
  if (p.ctl[3]) return;
  __shared__ int s_list[TPR_BLOCK];
  __shared__ int s_cnt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  const i64 t = (i64)blockIdx.x * TPR_BLOCK + tid;
  bool work = false;
  if (t < p.n_tiles) {
    const uint4 pw = __ldcg(reinterpret_cast<const uint4*>(p.tpend) + t);
    const uint4 rw = __ldcg(reinterpret_cast<const uint4*>(p.trel) + t);
    work = (pw.x | pw.y | pw.z | pw.w | rw.x | rw.y | rw.z | rw.w) != 0u;
  }
  const unsigned ball = __ballot_sync(FULL, work);
  if (ball) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&s_cnt, __popc(ball));
    base = __shfl_sync(FULL, base, 0);
    if (work) s_list[base + __popc(ball & ((1u << lane) - 1u))] = (int)t;
  }
  __syncthreads();
  const int cnt = s_cnt;
  if (cnt == 0) return;
  const unsigned stamp = (unsigned)__ldcg(&p.ctl[0]);
  const double tau = __ldcg(&p.tau[0]);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const int w = p.w;
  u64 fmin = ~0ull, evals = 0;
  unsigned nrel = 0;
  for (int it = warp; it < cnt; it += TPR_BLOCK / 32) {
    const int tile = s_list[it];
    unsigned line, txu, tzu, tyu;
    p.fd_tnx.divmod((unsigned)tile, line, txu);
    p.fd_tny.divmod(line, tzu, tyu);
    const int tx = (int)txu, ty = (int)tyu, tz = (int)tzu;
    const uint4 pw = __ldcg(reinterpret_cast<const uint4*>(p.tpend) + tile);
    const uint4 rw = __ldcg(reinterpret_cast<const uint4*>(p.trel) + tile);
    const unsigned pk[4] = {pw.x, pw.y, pw.z, pw.w};
    const int gx = tx * TX + (lane & 7), gy = ty * TY + (lane >> 3);
    double d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // pending bits only exist for nodes of the grid
      const i64 J = (i64)gx + (i64)p.nx * ((i64)gy + (i64)p.ny * (tz * TZ + k));
      d[k] = ((pk[k] >> lane) & 1u) ? __ldcg(&p.dist[J]) : INF;
    }
    unsigned nk[4];
    unsigned hmin = 0xffffffffu;
    const int exy = (min(p.nx - 1, gx + w) - max(0, gx - w) + 1) * (min(p.ny - 1, gy + w) - max(0, gy - w) + 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool pend = (pk[k] >> lane) & 1u;
      const bool near = pend && d[k] < tau;
      nk[k] = __ballot_sync(FULL, near);
      if (near) {
        hmin = min(hmin, (unsigned)__double2hiint(d[k]));
        const int gz = tz * TZ + k;
        evals += (u64)(exy * (min(p.nz - 1, gz + w) - max(0, gz - w) + 1) - (p.self ? 0 : 1));
      }
      if (pend && !near) {
        const u64 b = (u64)__double_as_longlong(d[k]);
        fmin = b < fmin ? b : fmin;
      }
    }
    const unsigned any = nk[0] | nk[1] | nk[2] | nk[3];
    if (lane == 0) {
      if (any | rw.x | rw.y | rw.z | rw.w) reinterpret_cast<uint4*>(p.trel)[tile] = make_uint4(nk[0], nk[1], nk[2], nk[3]);
      if (any)
        reinterpret_cast<uint4*>(p.tpend)[tile] = make_uint4(pk[0] & ~nk[0], pk[1] & ~nk[1], pk[2] & ~nk[2], pk[3] & ~nk[3]);
      nrel += (unsigned)(__popc(nk[0]) + __popc(nk[1]) + __popc(nk[2]) + __popc(nk[3]));
    }
    if (any) {
      const unsigned tminhi = __reduce_min_sync(FULL, hmin);  // high word of the smallest released travel time (>= 0)
      if (lane < 27) {
        // bounding box of the released nodes (tile-local), grown by the window half width, against the 27 tiles around
        const unsigned fold = (any | (any >> 8) | (any >> 16) | (any >> 24)) & 0xffu;
        const int bx0 = __ffs(fold) - 1, bx1 = 31 - __clz(fold);
        const unsigned rows = ((any & 0xffu) ? 1u : 0u) | ((any & 0xff00u) ? 2u : 0u) | ((any & 0xff0000u) ? 4u : 0u) |
                              ((any & 0xff000000u) ? 8u : 0u);
        const int by0 = __ffs(rows) - 1, by1 = 31 - __clz(rows);
        const int bz0 = nk[0] ? 0 : nk[1] ? 1 : nk[2] ? 2 : 3, bz1 = nk[3] ? 3 : nk[2] ? 2 : nk[1] ? 1 : 0;
        const int ddx = lane % 3 - 1, ddy = (lane / 3) % 3 - 1, ddz = lane / 9 - 1;
        const int ux = tx + ddx, uy = ty + ddy, uz = tz + ddz;
        if (ux >= 0 && ux < p.tnx && uy >= 0 && uy < p.tny && uz >= 0 && uz < p.tnz) {
          const int x0 = tx * TX + bx0 - w, x1 = tx * TX + bx1 + w;
          const int y0 = ty * TY + by0 - w, y1 = ty * TY + by1 + w;
          const int z0 = tz * TZ + bz0 - w, z1 = tz * TZ + bz1 + w;
          if (x0 <= ux * TX + TX - 1 && x1 >= ux * TX && y0 <= uy * TY + TY - 1 && y1 >= uy * TY &&
              z0 <= uz * TZ + TZ - 1 && z1 >= uz * TZ) {
            const int nb = ux + p.tnx * (uy + p.tny * uz);
            // every travel time of that tile already lies below every released one of this tile: nothing to improve
            if (__ldcg(&p.tmaxhi[nb]) > tminhi)
              if (atomicExch(&p.tmark[nb], stamp) != stamp) p.act[atomicAdd(&p.counters[0], 1ull)] = nb;
          }
        }
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    const u64 other = __shfl_xor_sync(FULL, fmin, o);
    fmin = other < fmin ? other : fmin;
    evals += __shfl_xor_sync(FULL, evals, o);
  }
  if (lane == 0) {
    if (fmin != ~0ull) atomicMin(&p.counters[7], fmin);
    if (nrel) atomicAdd(&p.counters[1], (u64)nrel);
    if (evals) atomicAdd(&p.counters[2], evals);
  }
}

// end of a round: threshold advance to min(pending) + delta when nothing was released, termination, counters reset
__device__ __forceinline__ void tp_round_control(const T3& p) {
  const u64 n_rel = __ldcg(&p.counters[1]);
  const u64 mp = __ldcg(&p.counters[7]);
  p.ctl[4] += 1;
  if (n_rel > 0) {
    p.ctl[5] += 1;
    p.counters[3] += n_rel;
    p.counters[5] += __ldcg(&p.counters[0]);
    // the tail of a bucket (a few stragglers per round) may share its rounds with the head of the next bucket: any
    // release order reaches the same fixed point (option early_advance)
    if (n_rel < p.early && mp != ~0ull) {
      const double t2 = __dadd_rn(__longlong_as_double((long long)mp), p.tau[1]);
      if (t2 > p.tau[0]) p.tau[0] = t2;
    }
  } else if (mp != ~0ull) {
    p.tau[0] = __dadd_rn(__longlong_as_double((long long)mp), p.tau[1]);
  } else {
    p.ctl[3] = 1;
  }
  p.ctl[0] += 1;
  p.counters[0] = 0ull;
  p.counters[1] = 0ull;
  p.counters[7] = ~0ull;
}

// dynamic smem: 5 * SN doubles (travel time, X, Y, Z, U of the released cells of tile + halo) + SY * SZ row masks +
// SZ plane masks
template <bool F32>
__global__ void __launch_bounds__(TILE_THREADS) tp_pull_kernel(T3 p) {
  extern __shared__ double sm[];
  __shared__ unsigned s_bminhi;
  if (p.ctl[3]) return;
  const int w = p.w, W = 2 * w + 1;
  const int SX = TX + 2 * w, SY = TY + 2 * w, SZ = TZ + 2 * w;
  const int SN = SX * SY * SZ, NROW = SY * SZ;
  double* sR = sm;
  double* sX = sR + SN;
  double* sY = sX + SN;
  double* sZ = sY + SN;
  double* sU = sZ + SN;
  unsigned* rowm = reinterpret_cast<unsigned*>(sU + SN);
  unsigned* planem = rowm + NROW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lx = tid % TX, ly = (tid / TX) % TY, lz = tid / (TX * TY);
  const i64 n_act = (i64)__ldcg(&p.counters[0]);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const unsigned wmask = (1u << W) - 1u, xmask = (1u << SX) - 1u;
  u64 screened = 0, exact = 0;
  for (i64 a = blockIdx.x; a < n_act; a += gridDim.x) {
    const int tile = __ldcg(&p.act[a]);
    unsigned line, txu, tzu, tyu;
    p.fd_tnx.divmod((unsigned)tile, line, txu);
    p.fd_tny.divmod(line, tzu, tyu);
    const int tx = (int)txu, ty = (int)tyu, tz = (int)tzu;
    const int ox = tx * TX - w, oy = ty * TY - w, oz = tz * TZ - w;  // grid coordinates of staged cell (0, 0, 0)
    __syncthreads();  // the previous tile's readers are done
    if (tid < SZ) planem[tid] = 0u;
    if (tid == 0) {
      s_bminhi = 0xffffffffu;
      p.tmaxhi[tile] = 0u;
    }
    // own node
    const int gx = tx * TX + lx, gy = ty * TY + ly, gz = tz * TZ + lz;
    const bool inside = gx < p.nx && gy < p.ny && gz < p.nz;
    const i64 I = (i64)gx + (i64)p.nx * ((i64)gy + (i64)p.ny * gz);
    double di = INF, xi = 0.0, yi = 0.0, zi = 0.0, ui = 0.0;
    if (inside) {
      di = __ldcg(&p.dist[I]);
      xi = p.X[I];
      yi = p.Y[I];
      zi = p.Z[I];
      ui = p.U[I];
    }
    __syncthreads();
    // phase 1: one mask per x-row of tile + halo from the released words of the (up to three) tiles the row crosses
    for (int row = tid; row < NROW; row += TILE_THREADS) {
      unsigned czu, cyu;
      p.fd_SY.divmod((unsigned)row, czu, cyu);
      const int gyy = oy + (int)cyu, gzz = oz + (int)czu;
      unsigned m = 0u;
      if (gyy >= 0 && gyy < p.ny && gzz >= 0 && gzz < p.nz) {
        const i64 base = ((i64)(gzz / TZ) * p.tny + (gyy / TY)) * p.tnx;
        const int word = gzz % TZ, sh8 = 8 * (gyy % TY);
#pragma unroll
        for (int dxx = -1; dxx <= 1; ++dxx) {
          const int txx = tx + dxx;
          if (txx < 0 || txx >= p.tnx) continue;
          const unsigned byte = (__ldcg(&p.trel[(base + txx) * 4 + word]) >> sh8) & 0xffu;
          const int sh = dxx * TX + w;  // staged x of that tile's first node
          m |= sh >= 0 ? byte << sh : byte >> (-sh);
        }
        m &= xmask;
      }
      rowm[row] = m;
      if (m) atomicOr(&planem[czu], m);
    }
    __syncthreads();
    // phase 2: travel time and coordinates of the released cells (16 lanes per x-row, SX <= 16)
    {
      const int cx = tid & 15;
      if (cx < SX) {
        for (int row = tid >> 4; row < NROW; row += TILE_THREADS / 16) {
          if (!((rowm[row] >> cx) & 1u)) continue;
          unsigned czu, cyu;
          p.fd_SY.divmod((unsigned)row, czu, cyu);
          const i64 J = (i64)(ox + cx) + (i64)p.nx * ((i64)(oy + (int)cyu) + (i64)p.ny * (oz + (int)czu));
          const int c = cx + SX * row;
          const double r = __ldcg(&p.dist[J]);
          sR[c] = r;
          sX[c] = p.X[J];
          sY[c] = p.Y[J];
          sZ[c] = p.Z[J];
          sU[c] = p.U[J];
          atomicMin(&s_bminhi, (unsigned)__double2hiint(r));
        }
      }
    }
    __syncthreads();
    bool improved = false;
    double best = di;
    // a travel time below every released one of the block cannot improve (high words: conservative)
    if (inside && (unsigned)__double2hiint(di) >= s_bminhi) {
      for (int zz = 0; zz < W; ++zz) {
        if (!((planem[lz + zz] >> lx) & wmask)) continue;
        const int row0 = ly + SY * (lz + zz);
        for (int yy = 0; yy < W; ++yy) {
          const int row = row0 + yy;
          unsigned m = (rowm[row] >> lx) & wmask;
          const int c0 = lx + SX * row;
          while (m) {
            const int c = c0 + __ffs(m) - 1;
            m &= m - 1u;
            const double dj = sR[c];
            if (!(dj < best)) continue;  // also the node itself
            const double xj = sX[c], yj = sY[c], zj = sZ[c], uj = sU[c];
            const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj), dz = __dsub_rn(zi, zj);
            const double d2 = __fma_rn(dx, dx, __fma_rn(dy, dy, dz * dz));
            if (p.count) screened += 1;
            if (screen_cannot_improve_t<F32>(best, dj, d2, screen_ssum3(fabs(__dadd_rn(ui, uj)), p.wmode))) continue;
            if (p.count) exact += 1;
            const double delta = cand3<F32>(dj, xi, yi, zi, ui, xj, yj, zj, uj, p.wmode);
            best = delta < best ? delta : best;
          }
        }
      }
      if (best < di) {
        p.dist[I] = best;
        improved = true;
      }
    }
    const unsigned imp = __ballot_sync(FULL, improved);
    const unsigned wmax = __reduce_max_sync(FULL, inside ? (unsigned)__double2hiint(best) : 0u);
    if (lane == 0) {
      if (imp) p.tpend[(i64)tile * 4 + warp] |= imp;  // this CTA is the only writer of the tile's words
      atomicMax(&p.tmaxhi[tile], wmax + 1u);          // reset by thread 0 at the start of the visit (two barriers ago)
    }
  }
  if (p.count) {
    for (int o = 16; o; o >>= 1) {
      screened += __shfl_xor_sync(FULL, screened, o);
      exact += __shfl_xor_sync(FULL, exact, o);
    }
    if (lane == 0 && screened) {
      atomicAdd(&p.counters[8], screened);
      atomicAdd(&p.counters[9], exact);
    }
  }
  if (!p.ctl_tail) return;  // the round control runs as a kernel of its own (tp_ctl_kernel)
  // round control by the last CTA to finish
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const u64 ticket = atomicAdd(&p.counters[4], 1ull);
    if (ticket == (u64)gridDim.x - 1ull) {
      __threadfence();
      tp_round_control(p);
      p.counters[4] = 0ull;
      __threadfence();
    }
  }
}

// the round control as a one-thread kernel (option ctl_tail = 0): no ticket atomics in the pull kernel
__global__ void tp_ctl_kernel(T3 p) {
  if (p.ctl[3]) return;
  tp_round_control(p);
}

__global__ void tp_init_kernel(T3 p, i32* __restrict__ prev, i64 n, i64 source, double delta) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    p.dist[i] = (i == source) ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
    prev[i] = -1;
  }
  if (i == 0) {
    const int sx = (int)(source % p.nx), sy = (int)((source / p.nx) % p.ny), sz = (int)(source / ((i64)p.nx * p.ny));
    const i64 tile = (sx / TX) + (i64)p.tnx * ((sy / TY) + (i64)p.tny * (sz / TZ));
    p.tpend[tile * 4 + (sz % TZ)] = 1u << ((sx % TX) + TX * (sy % TY));
    p.ctl[0] = 1;
    p.tau[0] = delta;
    p.tau[1] = delta;
    p.counters[7] = ~0ull;
  }
}

int ensure_pull3(rt_mesh* h) {
  Grid3D& g = *h->g3;
  if (g.pull_ready) return RT_OK;
  RT_TRY(g.tpend.alloc(g.n_tiles * 4));
  RT_TRY(g.trel.alloc(g.n_tiles * 4));
  RT_TRY(g.tmark.alloc(g.n_tiles));
  RT_TRY(g.tmaxhi.alloc(g.n_tiles));
  RT_TRY(g.tau.alloc(4));
  RT_TRY(g.ctl.alloc(8));
  g.pull_ready = true;
  return RT_OK;
}

}  // namespace

int bfm3d_solve_pull_batch(rt_mesh* h, const double* U_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                           rt_stats* stats);

int bfm3d_solve_pull(rt_mesh* h, const double* U_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                     rt_stats* stats) {
  // a batch keeps several sources in flight (option batch: 0 = automatic, 1 = one after the other, k = k slots)
  if (nsrc >= 2 && !h->opts.profile_timers && h->opts.batch != 1)
    return bfm3d_solve_pull_batch(h, U_dev, sources, nsrc, dist_dev, prev_dev, stats);
  Grid3D& g = *h->g3;
  cudaStream_t s = h->stream;
  RT_TRY(ensure_ws3(h));
  RT_TRY(ensure_pull3(h));
  const i64 n = g.n;
  const bool f32 = h->f32;
  if (f32) RT_TRY(grid3d_prepare_f32(h));
  T3 p;
  std::memset((void*)&p, 0, sizeof(T3));  // its bytes are the key of the cached graph
  p.X = f32 ? g.Xf.p : g.X.p;
  p.Y = f32 ? g.Yf.p : g.Y.p;
  p.Z = f32 ? g.Zf.p : g.Z.p;
  p.U = U_dev;
  p.dist = g.dist.p;
  p.tpend = g.tpend.p;
  p.trel = g.trel.p;
  p.tmark = g.tmark.p;
  p.tmaxhi = g.tmaxhi.p;
  p.act = g.act[0].p;
  p.counters = g.counters.p;
  p.tau = g.tau.p;
  p.ctl = g.ctl.p;
  p.nx = (int)g.nn[0];
  p.ny = (int)g.nn[1];
  p.nz = (int)g.nn[2];
  p.tnx = (int)g.tn[0];
  p.tny = (int)g.tn[1];
  p.tnz = (int)g.tn[2];
  p.n_tiles = g.n_tiles;
  p.w = g.w;
  p.self = g.self;
  p.wmode = h->opts.weight3d;
  p.count = h->opts.profile_timers != 0;
  p.ctl_tail = h->opts.fuse_begin != 0;
  p.early = (u64)((h->opts.early_advance >= 0.0 ? h->opts.early_advance : TP_EARLY) * std::pow((double)g.n, 2.0 / 3.0));
  p.fd_tnx = FastDiv((unsigned)g.tn[0]);
  p.fd_tny = FastDiv((unsigned)g.tn[1]);
  p.fd_SY = FastDiv((unsigned)(TY + 2 * g.w));
  RT_ARG(g.n_tiles < ((i64)1 << 31), "grid too large for the near-far schedule");
  // the tightness pass and the bucket-width probe take the push schedule's parameter block
  Q3 q = {};
  q.X = p.X;
  q.Y = p.Y;
  q.Z = p.Z;
  q.U = U_dev;
  q.dist = g.dist.p;
  q.prev = g.prev.p;
  q.nx = p.nx;
  q.ny = p.ny;
  q.nz = p.nz;
  q.w = g.w;
  q.self = g.self;
  q.wmode = p.wmode;
  const int w = g.w;
  const size_t SN = (size_t)(TX + 2 * w) * (TY + 2 * w) * (TZ + 2 * w);
  const size_t smem = 5 * SN * sizeof(double) + (size_t)((TY + 2 * w) * (TZ + 2 * w) + (TZ + 2 * w)) * sizeof(unsigned);
  RT_CUDA(cudaFuncSetAttribute(tp_pull_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RT_CUDA(cudaFuncSetAttribute(tp_pull_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  int per_sm = 1;
  if (f32)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tp_pull_kernel<true>, TILE_THREADS, smem);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tp_pull_kernel<false>, TILE_THREADS, smem);
  const unsigned gpull = (unsigned)std::min<i64>(g.n_tiles, (i64)sm_count * std::max(per_sm, 1));
  const unsigned grel = grid_for(g.n_tiles, TPR_BLOCK);
  cudaEvent_t ev0, ev1, evr0, evr1;
  RT_CUDA(cudaEventCreate(&ev0));
  RT_CUDA(cudaEventCreate(&ev1));
  RT_CUDA(cudaEventCreate(&evr0));
  RT_CUDA(cudaEventCreate(&evr1));
  rt_stats st = {};
  st.graph_edges = g.graph_edges;
  int rc = RT_OK;
  const bool timers = h->opts.profile_timers != 0;
  double delta = h->opts.delta;
  if (!(delta > 0.0)) {
    cudaMemsetAsync(g.tau.p + 3, 0, sizeof(double), s);
    cudaMemsetAsync(g.counters.p + 7, 0, sizeof(u64), s);
    wdiag3_kernel<<<grid_for((n + 6) / 7, 256), 256, 0, s>>>(q, g.tau.p + 3, g.counters.p + 7);
    double wsum = 0.0;
    u64 wc = 0;
    cudaMemcpyAsync(&wsum, g.tau.p + 3, sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&wc, g.counters.p + 7, sizeof(u64), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = RT_ERR_CUDA;
    const double wmean = wc ? wsum / (double)wc : 1.0;
    delta = wmean * (h->opts.delta_factor > 0.0 ? h->opts.delta_factor : TP_DELTA_FACTOR);
  }
  const int R = timers ? 1 : (h->opts.check_every > 1 ? h->opts.check_every : 32);
  auto enqueue_rounds = [&](cudaStream_t qs) {
    for (int r = 0; r < R; ++r) {
      tp_release_kernel<<<grel, TPR_BLOCK, 0, qs>>>(p);
      if (timers) cudaEventRecord(evr0, qs);
      if (f32)
        tp_pull_kernel<true><<<gpull, TILE_THREADS, smem, qs>>>(p);
      else
        tp_pull_kernel<false><<<gpull, TILE_THREADS, smem, qs>>>(p);
      if (timers) cudaEventRecord(evr1, qs);
      if (!p.ctl_tail) tp_ctl_kernel<<<1, 1, 0, qs>>>(p);
    }
  };
  cudaGraphExec_t gexec = nullptr;
  if (h->opts.use_graph && !timers && rc == RT_OK) {
    std::vector<char> key(sizeof(T3) + 4 * sizeof(int));
    std::memcpy(key.data(), &p, sizeof(T3));
    const int kv[4] = {R, (int)gpull, (int)f32, (int)smem};
    std::memcpy(key.data() + sizeof(T3), kv, sizeof(kv));
    if (g.tp_graph && key != g.tp_graph_key) {
      cudaGraphExecDestroy((cudaGraphExec_t)g.tp_graph);
      g.tp_graph = nullptr;
    }
    if (!g.tp_graph) {
      cudaGraph_t cg = nullptr;
      if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        enqueue_rounds(s);
        if (cudaStreamEndCapture(s, &cg) == cudaSuccess && cg) {
          cudaGraphExec_t ge = nullptr;
          if (cudaGraphInstantiate(&ge, cg, 0) == cudaSuccess) {
            g.tp_graph = (void*)ge;
            g.tp_graph_key = key;
          }
          cudaGraphDestroy(cg);
        }
      }
      cudaGetLastError();  // a failed capture falls back to plain launches
    }
    gexec = (cudaGraphExec_t)g.tp_graph;
  }
  u64* ch = g.counters_host;
  for (i64 si = 0; si < nsrc && rc == RT_OK; ++si) {
    const i64 src1 = sources[si];
    if (src1 < 1 || src1 > n) {
      rt_set_error("source %lld out of range 1..%lld", (long long)src1, (long long)n);
      rc = RT_ERR_ARG;
      break;
    }
    const i64 src = src1 - 1;
    cudaEventRecord(ev0, s);
    cudaMemsetAsync(g.counters.p, 0, 16 * sizeof(u64), s);
    cudaMemsetAsync(g.tmaxhi.p, 0xff, g.n_tiles * sizeof(unsigned), s);
    cudaMemsetAsync(g.tpend.p, 0, g.n_tiles * 4 * sizeof(unsigned), s);
    cudaMemsetAsync(g.trel.p, 0, g.n_tiles * 4 * sizeof(unsigned), s);
    cudaMemsetAsync(g.tmark.p, 0, g.n_tiles * sizeof(unsigned), s);
    cudaMemsetAsync(g.ctl.p, 0, 8 * sizeof(int), s);
    tp_init_kernel<<<grid_for(n, 256), 256, 0, s>>>(p, g.prev.p, n, src, delta);
    st.total_launches += 1;
    int hctl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    i64 enq_rounds = 0;
    while (!hctl[3]) {
      if (gexec) {
        if (cudaGraphLaunch(gexec, s) != cudaSuccess) {
          rc = RT_ERR_CUDA;
          break;
        }
      } else {
        enqueue_rounds(s);
      }
      st.total_launches += (p.ctl_tail ? 2 : 3) * R;
      enq_rounds += R;
      if (enq_rounds > ((i64)1 << 22)) {  // a solve needs 1e2 - 1e4 rounds: never spin forever on a logic error
        rt_set_error("near-far schedule did not converge within %lld rounds", (long long)enq_rounds);
        rc = RT_ERR_CUDA;
        break;
      }
      cudaMemcpyAsync(hctl, g.ctl.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        rc = RT_ERR_CUDA;
        break;
      }
      if (timers) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evr0, evr1);
        st.relax_ms += ms;
      }
    }
    if (rc != RT_OK) break;
    st.sweeps += hctl[4];
    st.relax_launches += hctl[5];

    cudaEventRecord(evr0, s);
    if (f32)
      prev_tight3_kernel<true><<<grid_for(n, 128), 128, 0, s>>>(q, n, src);
    else
      prev_tight3_kernel<false><<<grid_for(n, 128), 128, 0, s>>>(q, n, src);
    st.total_launches += 1;
    if (h->opts.canonical_prev) {  // the reference schedule's predecessors, ties included (canonical_prev.cu)
      Grid3Desc gd{p.X, p.Y, p.Z, p.nx, p.ny, p.nz, p.w, p.self, p.wmode};
      i64 launches = 0;
      rc = canonical_prev_3d(h, &g.canon, gd, U_dev, f32, g.dist.p, src, g.prev.p, &launches);
      st.total_launches += launches;
      if (rc != RT_OK) break;
    }
    cudaEventRecord(evr1, s);
    cudaMemcpyAsync(ch, g.counters.p, 16 * sizeof(u64), cudaMemcpyDeviceToHost, s);
    cudaEventRecord(ev1, s);
    if (dist_dev) cudaMemcpyAsync(dist_dev + si * n, g.dist.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (prev_dev) cudaMemcpyAsync(prev_dev + si * n, g.prev.p, n * sizeof(i32), cudaMemcpyDeviceToDevice, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      rc = RT_ERR_CUDA;
      break;
    }
    st.relaxed_edges += (i64)ch[2];  // nominal: every (released source, target) pair; the tightness pass is in prev_ms
    st.vertex_updates += (i64)ch[3];
    st.screened_edges += (i64)ch[8];
    st.exact_edges += (i64)ch[9];
    if (std::getenv("RT_TP_DEBUG"))
      std::fprintf(stderr, "tile-pull: %llu tile visits (%lld tiles), %llu releases, %d rounds\n", ch[5], (long long)g.n_tiles,
                   ch[3], hctl[4]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    st.kernel_ms += ms;
    cudaEventElapsedTime(&ms, evr0, evr1);
    st.prev_ms += ms;
  }
  cudaError_t e = cudaGetLastError();
  if (rc == RT_ERR_CUDA || e != cudaSuccess) {
    rt_set_error("CUDA failure in bfm3d_solve_pull: %s", cudaGetErrorString(e));
    rc = RT_ERR_CUDA;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cudaEventDestroy(evr0);
  cudaEventDestroy(evr1);
  if (stats) *stats = st;
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Batches on the 3-D grid: up to TP_SLOTS sources in flight, each with its own solver state and stream.  A single solve
// leaves the machine idle in its sparse rounds (latency-bound) and at ~0.6 of the issue rate in its dense ones; the
// rounds of another source fill those gaps (measured with separate handles, tools/probe_concurrent3d.py: 22.7 -> 18.3 ms
// per source at 216^3).  The kernels are the single-source ones; only the host loop differs: every running slot gets
// one CUDA-graph replay (check_every rounds) per turn, a slot that has converged runs its tightness pass, copies its
// tables out and takes the next source of the list.
struct TpSlot {
  DevBuf<double> dist, tau;
  DevBuf<i32> prev, act;
  DevBuf<unsigned> tpend, trel, tmark, tmaxhi;
  DevBuf<u64> counters;
  DevBuf<int> ctl;
  int* hctl = nullptr;   // pinned
  u64* hcnt = nullptr;   // pinned
  cudaStream_t stream = nullptr;
  cudaEvent_t evp0 = nullptr, evp1 = nullptr;
  void* graph = nullptr;
  std::vector<char> key;
  i64 si = -1;           // index of the source this slot is solving, -1 = idle
  i64 rounds_enq = 0;
};

void tp_slots_free(Grid3D& g) {
  for (TpSlot* S : g.tp_slots) {
    if (!S) continue;
    if (S->graph) cudaGraphExecDestroy((cudaGraphExec_t)S->graph);
    if (S->hctl) cudaFreeHost(S->hctl);
    if (S->hcnt) cudaFreeHost(S->hcnt);
    if (S->evp0) cudaEventDestroy(S->evp0);
    if (S->evp1) cudaEventDestroy(S->evp1);
    if (S->stream) cudaStreamDestroy(S->stream);
    delete S;
  }
  g.tp_slots.clear();
}

#ifndef TP_SLOTS
#define TP_SLOTS 4
#endif

namespace {
int tp_slot_ensure(Grid3D& g, size_t k) {
  while (g.tp_slots.size() <= k) g.tp_slots.push_back(nullptr);
  if (g.tp_slots[k]) return RT_OK;
  TpSlot* S = new TpSlot();
  g.tp_slots[k] = S;
  RT_TRY(S->dist.alloc(g.n));
  RT_TRY(S->prev.alloc(g.n));
  RT_TRY(S->tpend.alloc(g.n_tiles * 4));
  RT_TRY(S->trel.alloc(g.n_tiles * 4));
  RT_TRY(S->tmark.alloc(g.n_tiles));
  RT_TRY(S->tmaxhi.alloc(g.n_tiles));
  RT_TRY(S->act.alloc(g.n_tiles));
  RT_TRY(S->counters.alloc(16));
  RT_TRY(S->tau.alloc(4));
  RT_TRY(S->ctl.alloc(8));
  RT_CUDA(cudaMallocHost((void**)&S->hctl, 8 * sizeof(int)));
  RT_CUDA(cudaMallocHost((void**)&S->hcnt, 16 * sizeof(u64)));
  RT_CUDA(cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking));
  RT_CUDA(cudaEventCreate(&S->evp0));
  RT_CUDA(cudaEventCreate(&S->evp1));
  return RT_OK;
}
}  // namespace

int bfm3d_solve_pull_batch(rt_mesh* h, const double* U_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                           rt_stats* stats) {
  Grid3D& g = *h->g3;
  RT_TRY(ensure_ws3(h));
  RT_TRY(ensure_pull3(h));
  const i64 n = g.n;
  const bool f32 = h->f32;
  if (f32) RT_TRY(grid3d_prepare_f32(h));
  for (i64 si = 0; si < nsrc; ++si)
    if (sources[si] < 1 || sources[si] > n) {
      rt_set_error("source %lld out of range 1..%lld", (long long)sources[si], (long long)n);
      return RT_ERR_ARG;
    }
  RT_ARG(g.n_tiles < ((i64)1 << 31), "grid too large for the near-far schedule");
  const int K = (int)std::min<i64>(nsrc, h->opts.batch >= 2 ? std::min(h->opts.batch, 8) : TP_SLOTS);
  for (int k = 0; k < K; ++k) RT_TRY(tp_slot_ensure(g, (size_t)k));
  // parameter block shared by the slots (zeroed: its bytes are the key of the cached graphs)
  T3 p0;
  std::memset((void*)&p0, 0, sizeof(T3));
  p0.X = f32 ? g.Xf.p : g.X.p;
  p0.Y = f32 ? g.Yf.p : g.Y.p;
  p0.Z = f32 ? g.Zf.p : g.Z.p;
  p0.U = U_dev;
  p0.nx = (int)g.nn[0];
  p0.ny = (int)g.nn[1];
  p0.nz = (int)g.nn[2];
  p0.tnx = (int)g.tn[0];
  p0.tny = (int)g.tn[1];
  p0.tnz = (int)g.tn[2];
  p0.n_tiles = g.n_tiles;
  p0.w = g.w;
  p0.self = g.self;
  p0.wmode = h->opts.weight3d;
  p0.count = 0;
  p0.ctl_tail = h->opts.fuse_begin != 0;
  p0.early = (u64)((h->opts.early_advance >= 0.0 ? h->opts.early_advance : TP_EARLY) * std::pow((double)g.n, 2.0 / 3.0));
  p0.fd_tnx = FastDiv((unsigned)g.tn[0]);
  p0.fd_tny = FastDiv((unsigned)g.tn[1]);
  p0.fd_SY = FastDiv((unsigned)(TY + 2 * g.w));
  Q3 q0 = {};
  q0.X = p0.X;
  q0.Y = p0.Y;
  q0.Z = p0.Z;
  q0.U = U_dev;
  q0.nx = p0.nx;
  q0.ny = p0.ny;
  q0.nz = p0.nz;
  q0.w = g.w;
  q0.self = g.self;
  q0.wmode = p0.wmode;
  const int w = g.w;
  const size_t SN = (size_t)(TX + 2 * w) * (TY + 2 * w) * (TZ + 2 * w);
  const size_t smem = 5 * SN * sizeof(double) + (size_t)((TY + 2 * w) * (TZ + 2 * w) + (TZ + 2 * w)) * sizeof(unsigned);
  RT_CUDA(cudaFuncSetAttribute(tp_pull_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RT_CUDA(cudaFuncSetAttribute(tp_pull_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  int per_sm = 1;
  if (f32)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tp_pull_kernel<true>, TILE_THREADS, smem);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tp_pull_kernel<false>, TILE_THREADS, smem);
  const unsigned gpull = (unsigned)std::min<i64>(g.n_tiles, (i64)sm_count * std::max(per_sm, 1));
  const unsigned grel = grid_for(g.n_tiles, TPR_BLOCK);
  // bucket width (same rule as the single-source path)
  double delta = h->opts.delta;
  if (!(delta > 0.0)) {
    cudaStream_t s = h->stream;
    cudaMemsetAsync(g.tau.p + 3, 0, sizeof(double), s);
    cudaMemsetAsync(g.counters.p + 7, 0, sizeof(u64), s);
    wdiag3_kernel<<<grid_for((n + 6) / 7, 256), 256, 0, s>>>(q0, g.tau.p + 3, g.counters.p + 7);
    double wsum = 0.0;
    u64 wc = 0;
    cudaMemcpyAsync(&wsum, g.tau.p + 3, sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&wc, g.counters.p + 7, sizeof(u64), cudaMemcpyDeviceToHost, s);
    RT_CUDA(cudaStreamSynchronize(s));
    const double wmean = wc ? wsum / (double)wc : 1.0;
    delta = wmean * (h->opts.delta_factor > 0.0 ? h->opts.delta_factor : TP_DELTA_FACTOR);
  }
  const int R = h->opts.check_every > 1 ? h->opts.check_every : 32;
  const int launches_per_round = p0.ctl_tail ? 2 : 3;
  std::vector<T3> P((size_t)K, p0);
  auto enqueue_rounds = [&](const T3& p, cudaStream_t qs) {
    for (int r = 0; r < R; ++r) {
      tp_release_kernel<<<grel, TPR_BLOCK, 0, qs>>>(p);
      if (f32)
        tp_pull_kernel<true><<<gpull, TILE_THREADS, smem, qs>>>(p);
      else
        tp_pull_kernel<false><<<gpull, TILE_THREADS, smem, qs>>>(p);
      if (!p.ctl_tail) tp_ctl_kernel<<<1, 1, 0, qs>>>(p);
    }
  };
  for (int k = 0; k < K; ++k) {
    TpSlot& S = *g.tp_slots[(size_t)k];
    T3& p = P[(size_t)k];
    p.dist = S.dist.p;
    p.tpend = S.tpend.p;
    p.trel = S.trel.p;
    p.tmark = S.tmark.p;
    p.tmaxhi = S.tmaxhi.p;
    p.act = S.act.p;
    p.counters = S.counters.p;
    p.tau = S.tau.p;
    p.ctl = S.ctl.p;
    S.si = -1;
    if (h->opts.use_graph) {
      std::vector<char> key(sizeof(T3) + 4 * sizeof(int));
      std::memcpy(key.data(), &p, sizeof(T3));
      const int kv[4] = {R, (int)gpull, (int)f32, (int)smem};
      std::memcpy(key.data() + sizeof(T3), kv, sizeof(kv));
      if (S.graph && key != S.key) {
        cudaGraphExecDestroy((cudaGraphExec_t)S.graph);
        S.graph = nullptr;
      }
      if (!S.graph) {
        cudaGraph_t cg = nullptr;
        if (cudaStreamBeginCapture(S.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          enqueue_rounds(p, S.stream);
          if (cudaStreamEndCapture(S.stream, &cg) == cudaSuccess && cg) {
            cudaGraphExec_t ge = nullptr;
            if (cudaGraphInstantiate(&ge, cg, 0) == cudaSuccess) {
              S.graph = (void*)ge;
              S.key = key;
            }
            cudaGraphDestroy(cg);
          }
        }
        cudaGetLastError();  // a failed capture falls back to plain launches
      }
    } else if (S.graph) {
      cudaGraphExecDestroy((cudaGraphExec_t)S.graph);
      S.graph = nullptr;
    }
  }
  rt_stats st = {};
  st.graph_edges = g.graph_edges;
  int rc = RT_OK;
  cudaEvent_t ev0, ev1;
  RT_CUDA(cudaEventCreate(&ev0));
  RT_CUDA(cudaEventCreate(&ev1));
  cudaEventRecord(ev0, h->stream);
  i64 next = 0, finished = 0;
  auto start = [&](int k) {  // slot k takes source `next`
    TpSlot& S = *g.tp_slots[(size_t)k];
    const T3& p = P[(size_t)k];
    cudaStream_t s = S.stream;
    S.si = next++;
    S.rounds_enq = 0;
    cudaMemsetAsync(S.counters.p, 0, 16 * sizeof(u64), s);
    cudaMemsetAsync(S.tmaxhi.p, 0xff, g.n_tiles * sizeof(unsigned), s);
    cudaMemsetAsync(S.tpend.p, 0, g.n_tiles * 4 * sizeof(unsigned), s);
    cudaMemsetAsync(S.trel.p, 0, g.n_tiles * 4 * sizeof(unsigned), s);
    cudaMemsetAsync(S.tmark.p, 0, g.n_tiles * sizeof(unsigned), s);
    cudaMemsetAsync(S.ctl.p, 0, 8 * sizeof(int), s);
    tp_init_kernel<<<grid_for(n, 256), 256, 0, s>>>(p, S.prev.p, n, sources[S.si] - 1, delta);
    st.total_launches += 1;
  };
  for (int k = 0; k < K && next < nsrc; ++k) start(k);
  while (finished < nsrc && rc == RT_OK) {
    for (int k = 0; k < K; ++k) {  // one replay of R rounds for every running slot
      TpSlot& S = *g.tp_slots[(size_t)k];
      if (S.si < 0) continue;
      if (S.graph) {
        if (cudaGraphLaunch((cudaGraphExec_t)S.graph, S.stream) != cudaSuccess) rc = RT_ERR_CUDA;
      } else {
        enqueue_rounds(P[(size_t)k], S.stream);
      }
      st.total_launches += (i64)launches_per_round * R;
      S.rounds_enq += R;
      cudaMemcpyAsync(S.hctl, S.ctl.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, S.stream);
    }
    for (int k = 0; k < K && rc == RT_OK; ++k) {
      TpSlot& S = *g.tp_slots[(size_t)k];
      if (S.si < 0) continue;
      if (cudaStreamSynchronize(S.stream) != cudaSuccess) {
        rc = RT_ERR_CUDA;
        break;
      }
      if (!S.hctl[3]) {
        if (S.rounds_enq > ((i64)1 << 22)) {
          rt_set_error("near-far schedule did not converge within %lld rounds", (long long)S.rounds_enq);
          rc = RT_ERR_CUDA;
        }
        continue;
      }
      // converged: predecessors, tables out, counters; then the slot takes the next source
      const i64 si = S.si;
      const i64 src = sources[si] - 1;
      Q3 q = q0;
      q.dist = S.dist.p;
      q.prev = S.prev.p;
      cudaEventRecord(S.evp0, S.stream);
      if (f32)
        prev_tight3_kernel<true><<<grid_for(n, 128), 128, 0, S.stream>>>(q, n, src);
      else
        prev_tight3_kernel<false><<<grid_for(n, 128), 128, 0, S.stream>>>(q, n, src);
      st.total_launches += 1;
      if (h->opts.canonical_prev) {  // one workspace, the handle's stream: serialised
        if (cudaStreamSynchronize(S.stream) != cudaSuccess) {
          rc = RT_ERR_CUDA;
          break;
        }
        Grid3Desc gd{q0.X, q0.Y, q0.Z, q0.nx, q0.ny, q0.nz, g.w, g.self, q0.wmode};
        i64 launches = 0;
        rc = canonical_prev_3d(h, &g.canon, gd, U_dev, f32, S.dist.p, src, S.prev.p, &launches);
        st.total_launches += launches;
        if (rc != RT_OK) break;
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
          rc = RT_ERR_CUDA;
          break;
        }
      }
      cudaEventRecord(S.evp1, S.stream);
      cudaMemcpyAsync(S.hcnt, S.counters.p, 16 * sizeof(u64), cudaMemcpyDeviceToHost, S.stream);
      if (dist_dev) cudaMemcpyAsync(dist_dev + si * n, S.dist.p, n * sizeof(double), cudaMemcpyDeviceToDevice, S.stream);
      if (prev_dev) cudaMemcpyAsync(prev_dev + si * n, S.prev.p, n * sizeof(i32), cudaMemcpyDeviceToDevice, S.stream);
      if (cudaStreamSynchronize(S.stream) != cudaSuccess) {
        rc = RT_ERR_CUDA;
        break;
      }
      st.sweeps += S.hctl[4];
      st.relax_launches += S.hctl[5];
      st.relaxed_edges += (i64)S.hcnt[2];
      st.vertex_updates += (i64)S.hcnt[3];
      float ms = 0.f;
      cudaEventElapsedTime(&ms, S.evp0, S.evp1);
      st.prev_ms += ms;
      finished += 1;
      S.si = -1;
      if (next < nsrc) start(k);
    }
  }
  for (int k = 0; k < K; ++k) cudaStreamSynchronize(g.tp_slots[(size_t)k]->stream);
  cudaEventRecord(ev1, h->stream);  // the handle's stream was idle throughout: ev0 .. ev1 brackets the batch in host order
  cudaStreamSynchronize(h->stream);
  {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    st.kernel_ms = ms;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cudaError_t e = cudaGetLastError();
  if (rc == RT_ERR_CUDA || e != cudaSuccess) {
    if (e != cudaSuccess) rt_set_error("CUDA failure in bfm3d_solve_pull_batch: %s", cudaGetErrorString(e));
    rc = RT_ERR_CUDA;
  }
  if (stats) *stats = st;
  return rc;
}

int bfm3d_solve(rt_mesh* h, const double* U_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                rt_stats* stats) {
  if (h->opts.schedule == 1 && h->opts.tile_pull)
    return bfm3d_solve_pull(h, U_dev, sources, nsrc, dist_dev, prev_dev, stats);
  if (h->opts.schedule == 1) return bfm3d_solve_push(h, U_dev, sources, nsrc, dist_dev, prev_dev, stats);
  Grid3D& g = *h->g3;
  cudaStream_t s = h->stream;
  RT_TRY(ensure_ws3(h));
  const i64 n = g.n;
  const bool f32 = h->f32;
  if (f32) RT_TRY(grid3d_prepare_f32(h));
  P3 p;
  p.X = f32 ? g.Xf.p : g.X.p;
  p.Y = f32 ? g.Yf.p : g.Y.p;
  p.Z = f32 ? g.Zf.p : g.Z.p;
  p.U = U_dev;
  p.dist = g.dist.p;
  p.dist0 = g.dist0.p;
  p.prev = g.prev.p;
  p.improved = g.improved.p;
  p.counters = g.counters.p;
  p.nx = (int)g.nn[0];
  p.ny = (int)g.nn[1];
  p.nz = (int)g.nn[2];
  p.tnx = (int)g.tn[0];
  p.tny = (int)g.tn[1];
  p.tnz = (int)g.tn[2];
  p.w = g.w;
  p.self = g.self;
  p.wmode = h->opts.weight3d;
  const int w = g.w;
  const size_t smem = (size_t)5 * (TX + 2 * w) * (TY + 2 * w) * (TZ + 2 * w) * sizeof(double);
  RT_CUDA(cudaFuncSetAttribute(relax3d_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RT_CUDA(cudaFuncSetAttribute(relax3d_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
  const i64 max_blocks = (i64)sm_count * 16;

  cudaEvent_t ev0, ev1, evr0, evr1;
  RT_CUDA(cudaEventCreate(&ev0));
  RT_CUDA(cudaEventCreate(&ev1));
  RT_CUDA(cudaEventCreate(&evr0));
  RT_CUDA(cudaEventCreate(&evr1));
  rt_stats st = {};
  st.graph_edges = g.graph_edges;
  int rc = RT_OK;
  for (i64 si = 0; si < nsrc && rc == RT_OK; ++si) {
    const i64 src1 = sources[si];
    if (src1 < 1 || src1 > n) {
      rt_set_error("source %lld out of range 1..%lld", (long long)src1, (long long)n);
      rc = RT_ERR_ARG;
      break;
    }
    const i64 src = src1 - 1;
    cudaEventRecord(ev0, s);
    init3d_kernel<<<grid_for(n, 256), 256, 0, s>>>(p.dist, p.dist0, p.prev, n, src);
    cudaMemsetAsync(g.improved.p, 0, g.n_tiles * sizeof(unsigned), s);
    cudaMemsetAsync(g.counters.p, 0, 8 * sizeof(u64), s);
    // initial active set = window of the source (BFM: union!(active, G[source])) -> tiles around its tile
    {
      const i64 sx = src % g.nn[0], sy = (src / g.nn[0]) % g.nn[1], sz = src / (g.nn[0] * g.nn[1]);
      const i64 st_tile = (sx / TX) + g.tn[0] * ((sy / TY) + g.tn[1] * (sz / TZ));
      const unsigned lx = (unsigned)(sx % TX), ly = (unsigned)(sy % TY), lz = (unsigned)(sz % TZ);
      const unsigned one = 0x80000000u | lx | (lx << 3) | (ly << 6) | (ly << 8) | (lz << 10) | (lz << 12);
      cudaMemcpyAsync(g.improved.p + st_tile, &one, sizeof(unsigned), cudaMemcpyHostToDevice, s);
    }
    int cur = 0;
    activate3d_kernel<<<grid_for(g.n_tiles, 256), 256, 0, s>>>(p, g.n_tiles, g.act[cur].p, cur);
    cudaMemsetAsync(g.improved.p, 0, g.n_tiles * sizeof(unsigned), s);
    cudaMemcpyAsync(g.counters_host, g.counters.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
    st.total_launches += 2;
    if (cudaStreamSynchronize(s) != cudaSuccess) {
      rc = RT_ERR_CUDA;
      break;
    }
    i64 n_active = (i64)g.counters_host[cur];
    while (n_active > 0) {
      const int nxt = cur ^ 1;
      const unsigned nb = (unsigned)std::min<i64>(n_active, max_blocks);
      if (h->opts.profile_timers) cudaEventRecord(evr0, s);
      if (f32)
        relax3d_kernel<true><<<nb, TILE_THREADS, smem, s>>>(p, g.act[cur].p, cur);
      else
        relax3d_kernel<false><<<nb, TILE_THREADS, smem, s>>>(p, g.act[cur].p, cur);
      if (h->opts.profile_timers) cudaEventRecord(evr1, s);
      commit3d_kernel<<<nb, TILE_THREADS, 0, s>>>(p, g.act[cur].p, cur);
      cudaMemsetAsync(g.counters.p + nxt, 0, sizeof(u64), s);
      activate3d_kernel<<<grid_for(g.n_tiles, 256), 256, 0, s>>>(p, g.n_tiles, g.act[nxt].p, nxt);
      cudaMemsetAsync(g.improved.p, 0, g.n_tiles * sizeof(unsigned), s);
      cudaMemcpyAsync(g.counters_host, g.counters.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, s);
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
      n_active = (i64)g.counters_host[cur];
    }
    if (rc != RT_OK) break;
    st.relaxed_edges += (i64)g.counters_host[2];
    st.vertex_updates += (i64)g.counters_host[3];
    cudaEventRecord(ev1, s);
    if (dist_dev) cudaMemcpyAsync(dist_dev + si * n, g.dist.p, n * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (prev_dev) cudaMemcpyAsync(prev_dev + si * n, g.prev.p, n * sizeof(i32), cudaMemcpyDeviceToDevice, s);
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
    rt_set_error("CUDA failure in bfm3d_solve: %s", cudaGetErrorString(e));
    rc = RT_ERR_CUDA;
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cudaEventDestroy(evr0);
  cudaEventDestroy(evr1);
  if (stats) *stats = st;
  return rc;
}

// =========================================================================================================
// The rest of the 3-D grid surface of src/StructuredGrid.jl: axes, getindex (linear and Cartesian), connectivity,
// closest_point.  All of it works on the RAW axis coordinates (gr.x, gr.y, gr.z) exactly like the reference:
// gr[I] = Point(gr.x[i], gr.y[j], gr.z[k]) (:77-81) is NOT mapped through spherical2cart.
namespace {

__global__ void axes3d_kernel(double c0x, double c0y, double c0z, double c1x, double c1y, double c1z, int nx, int ny,
                              int nz, double* __restrict__ ax, double* __restrict__ ay, double* __restrict__ az) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nx) ax[t] = lerpi_dev(t, nx - 1, c0x, c1x);
  if (t < ny) ay[t] = lerpi_dev(t, ny - 1, c0y, c1y);
  if (t < nz) az[t] = lerpi_dev(t, nz - 1, c0z, c1z);
}

// gr[I] (:77-81) with CartesianIndex(gr, I) (:90-96): i = mod(I-1, nx)+1; k = cld(I, nx*ny); j = cld(I - nx*ny*(k-1), nx)
__global__ void points3d_kernel(const double* __restrict__ ax, const double* __restrict__ ay,
                                const double* __restrict__ az, i64 nx, i64 ny, const i64* __restrict__ ids, i64 count,
                                double* __restrict__ xyz, i64* __restrict__ ijk) {
  const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  const i64 I = ids[q];
  const i64 i = (I - 1) % nx + 1;
  const i64 k = (I + nx * ny - 1) / (nx * ny);
  const i64 r = I - nx * ny * (k - 1);
  const i64 j = (r + nx - 1) / nx;
  if (xyz) {
    xyz[3 * q + 0] = ax[i - 1];
    xyz[3 * q + 1] = ay[j - 1];
    xyz[3 * q + 2] = az[k - 1];
  }
  if (ijk) {
    ijk[3 * q + 0] = i;
    ijk[3 * q + 1] = j;
    ijk[3 * q + 2] = k;
  }
}

// connectivity(gr, iel) (:146-168) with cornerindex_ijk (:106-112): 8 corner ids of hex `iel` (1-based), ordered
// (idx, idx+1, idx+1+nx, idx+nx, idx+nxny, idx+nxny+1, idx+nxny+1+nx, idx+nxny+nx)
__global__ void connectivity3d_kernel(i64 nx, i64 ny, i64 ex, i64 ey, i64 first, i64 count, i64* __restrict__ out) {
  const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  const i64 iel = first + q;
  const i64 i = (iel - 1) % ex + 1;
  const i64 j = ((iel + ex - 1) / ex - 1) % ey + 1;
  const i64 k = (iel + ex * ey - 1) / (ex * ey);
  const i64 idx = i + (j - 1) * nx + (k - 1) * nx * ny;
  const i64 nxny = nx * ny;
  i64* o = out + 8 * q;
  o[0] = idx;
  o[1] = idx + 1;
  o[2] = idx + 1 + nx;
  o[3] = idx + nx;
  o[4] = idx + nxny;
  o[5] = idx + nxny + 1;
  o[6] = idx + nxny + 1 + nx;
  o[7] = idx + nxny + nx;
}

// closest_point(gr, x, y, z) (:257-270): argmin over the linear index of distance3D(gr[i], p) with strict `<`
// (first index wins).  The rounded distance fl(sqrt(fl(fl(a_i + b_j) + c_k))) with a_i = fl(fl(x_i - px)^2) (b_j, c_k
// alike) is monotone in each of a_i, b_j, c_k, so its minimum D* is attained at the per-axis minima and the FIRST
// linear index that attains it (z slowest, x fastest) is found axis by axis: the smallest k with
// val(a_min, b_min, c_k) == D*, then the smallest j with val(a_min, b_j, c_k0) == D*, then the smallest i with
// val(a_i, b_j0, c_k0) == D*.  O(nx + ny + nz) per query instead of a sweep over all nodes; one thread per query.
__device__ __forceinline__ double sq_diff(double a, double pa) {
  const double d = __dsub_rn(a, pa);
  return __dmul_rn(d, d);
}
__device__ __forceinline__ double dist_abc(double a, double b, double c) {
  return __dsqrt_rn(__dadd_rn(__dadd_rn(a, b), c));
}
__global__ void closest3d_kernel(const double* __restrict__ ax, const double* __restrict__ ay,
                                 const double* __restrict__ az, int nx, int ny, int nz, const double* __restrict__ pq,
                                 i64 npts, i64* __restrict__ index) {
  const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npts) return;
  const double qa = pq[3 * q], qb = pq[3 * q + 1], qc = pq[3 * q + 2];
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  double amin = INF, bmin = INF, cmin = INF;
  for (int i = 0; i < nx; ++i) amin = fmin(amin, sq_diff(ax[i], qa));  // fmin drops NaN: handled by the D test below
  for (int j = 0; j < ny; ++j) bmin = fmin(bmin, sq_diff(ay[j], qb));
  for (int k = 0; k < nz; ++k) cmin = fmin(cmin, sq_diff(az[k], qc));
  const double D = dist_abc(amin, bmin, cmin);
  i64 out = -1;  // no distance compares below Inf (NaN or infinite query): the reference returns its initial index
  if (qa == qa && qb == qb && qc == qc && D < INF) {
    int k0 = 0, j0 = 0, i0 = 0;
    while (k0 < nz - 1 && dist_abc(amin, bmin, sq_diff(az[k0], qc)) != D) ++k0;
    const double ck = sq_diff(az[k0], qc);
    while (j0 < ny - 1 && dist_abc(amin, sq_diff(ay[j0], qb), ck) != D) ++j0;
    const double bj = sq_diff(ay[j0], qb);
    while (i0 < nx - 1 && dist_abc(sq_diff(ax[i0], qa), bj, ck) != D) ++i0;
    out = (i64)i0 + (i64)nx * ((i64)j0 + (i64)ny * k0);
  }
  index[q] = out;
}

int ensure_axes3(const rt_mesh* h) {
  Grid3D& g = *h->g3;
  if (g.ax.n == (size_t)g.nn[0] && g.ay.n == (size_t)g.nn[1] && g.az.n == (size_t)g.nn[2]) return RT_OK;
  RT_TRY(g.ax.alloc(g.nn[0]));
  RT_TRY(g.ay.alloc(g.nn[1]));
  RT_TRY(g.az.alloc(g.nn[2]));
  const i64 m = std::max(g.nn[0], std::max(g.nn[1], g.nn[2]));
  axes3d_kernel<<<grid_for(m, 256), 256, 0, h->stream>>>(g.c0[0], g.c0[1], g.c0[2], g.c1[0], g.c1[1], g.c1[2],
                                                        (int)g.nn[0], (int)g.nn[1], (int)g.nn[2], g.ax.p, g.ay.p, g.az.p);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaStreamSynchronize(h->stream));
  return RT_OK;
}

}  // namespace

int grid3d_axes(const rt_mesh* h, double* x, double* y, double* z) {
  RT_TRY(ensure_axes3(h));
  const Grid3D& g = *h->g3;
  if (x) RT_CUDA(cudaMemcpy(x, g.ax.p, g.nn[0] * sizeof(double), cudaMemcpyDeviceToHost));
  if (y) RT_CUDA(cudaMemcpy(y, g.ay.p, g.nn[1] * sizeof(double), cudaMemcpyDeviceToHost));
  if (z) RT_CUDA(cudaMemcpy(z, g.az.p, g.nn[2] * sizeof(double), cudaMemcpyDeviceToHost));
  return RT_OK;
}

int grid3d_points(const rt_mesh* h, const i64* ids, i64 count, double* xyz, i64* ijk) {
  RT_ARG(ids && count >= 0, "bad getindex arguments");
  const Grid3D& g = *h->g3;
  for (i64 q = 0; q < count; ++q) RT_ARG(ids[q] >= 1 && ids[q] <= g.n, "linear index out of range (BoundsError)");
  if (count == 0) return RT_OK;
  RT_TRY(ensure_axes3(h));
  DevBuf<i64> dids, dijk;
  DevBuf<double> dxyz;
  RT_TRY(dids.upload(ids, count));
  if (xyz) RT_TRY(dxyz.alloc(3 * count));
  if (ijk) RT_TRY(dijk.alloc(3 * count));
  points3d_kernel<<<grid_for(count, 256), 256>>>(g.ax.p, g.ay.p, g.az.p, g.nn[0], g.nn[1], dids.p, count,
                                                 xyz ? dxyz.p : nullptr, ijk ? dijk.p : nullptr);
  RT_CUDA(cudaGetLastError());
  if (xyz) RT_CUDA(cudaMemcpy(xyz, dxyz.p, 3 * count * sizeof(double), cudaMemcpyDeviceToHost));
  if (ijk) RT_CUDA(cudaMemcpy(ijk, dijk.p, 3 * count * sizeof(i64), cudaMemcpyDeviceToHost));
  return RT_OK;
}

int grid3d_connectivity(const rt_mesh* h, i64 first_el, i64 count, i64* e2n) {
  const Grid3D& g = *h->g3;
  const i64 ex = g.nn[0] - 1, ey = g.nn[1] - 1, ez = g.nn[2] - 1;
  const i64 nel = (ex > 0 && ey > 0 && ez > 0) ? ex * ey * ez : 0;
  RT_ARG(e2n && count >= 0 && first_el >= 1 && first_el + count - 1 <= nel, "element range outside 1..prod(nels)");
  const i64 slab = (i64)1 << 24;
  DevBuf<i64> out;
  RT_TRY(out.alloc(8 * std::min(count, slab)));
  for (i64 o = 0; o < count; o += slab) {
    const i64 c = std::min(slab, count - o);
    connectivity3d_kernel<<<grid_for(c, 256), 256>>>(g.nn[0], g.nn[1], ex, ey, first_el + o, c, out.p);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaMemcpy(e2n + 8 * o, out.p, 8 * c * sizeof(i64), cudaMemcpyDeviceToHost));
  }
  return RT_OK;
}

int grid3d_closest(const rt_mesh* h, const double* px, const double* py, const double* pz, i64 npts, i64* out) {
  RT_ARG(px && py && pz && out && npts >= 0, "bad closest_point arguments");
  if (npts == 0) return RT_OK;
  RT_TRY(ensure_axes3(h));
  const Grid3D& g = *h->g3;
  cudaStream_t s = h->stream;
  std::vector<double> pq(3 * npts);
  for (i64 q = 0; q < npts; ++q) {
    pq[3 * q] = px[q];
    pq[3 * q + 1] = py[q];
    pq[3 * q + 2] = pz[q];
  }
  DevBuf<double> dpq;
  DevBuf<i64> index;
  RT_TRY(dpq.upload(pq.data(), 3 * npts, s));
  RT_TRY(index.alloc(npts));
  closest3d_kernel<<<grid_for(npts, 64), 64, 0, s>>>(g.ax.p, g.ay.p, g.az.p, (int)g.nn[0], (int)g.nn[1], (int)g.nn[2],
                                                     dpq.p, npts, index.p);
  RT_CUDA(cudaGetLastError());
  std::vector<i64> hi(npts);
  RT_CUDA(cudaMemcpyAsync(hi.data(), index.p, npts * sizeof(i64), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  for (i64 q = 0; q < npts; ++q) out[q] = hi[q] < 0 ? -1 : hi[q] + 1;  // -1 as in the reference (NaN query)
  return RT_OK;
}
