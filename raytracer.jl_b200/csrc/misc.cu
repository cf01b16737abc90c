// misc.cu -- the small kernels either side of the solver: velocity interpolation, closest_point, id
// conversion and recontruct_path.
#include <algorithm>

#include "common.cuh"

namespace {

// src/GridAnnulus.jl:73 (also :262,:297,:912; src/ShortestPath.jl:77): the seven discontinuity radii
__constant__ double RL_DEV[7] = {6371.0 - 20.0,  6371.0 - 35.0,   6371.0 - 210.0, 6371.0 - 410.0,
                                 6371.0 - 660.0, 6371.0 - 2740.0, 6371.0 - 2891.5};

// Float32.(v) of src/SSSP/bfm_gpu.jl:173-177, kept in fp64 storage
__global__ void round_f32_kernel(const double* __restrict__ in, double* __restrict__ out, i64 n) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    out[i] = (double)__double2float_rn(in[i]);
}

// interpolate_velocity: src/utils.jl:38-44 (buffer < 0) / src/ShortestPath.jl:74-90 (buffer >= 0).
// Interpolations.jl gridded linear: i = clamp(searchsortedlast(knots, x), 1, nk-1);
// f = (x - k[i]) / (k[i+1] - k[i]);  v = (1 - f) * y[i] + f * y[i+1]   (no FMA)
__global__ void interp_kernel(const double* __restrict__ kr, const double* __restrict__ kv, i64 nk,
                              const double* __restrict__ r, i64 n, double buffer, double* __restrict__ out,
                              int* __restrict__ bad) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xq = r[i];
  if (buffer >= 0.0) {
#pragma unroll
    for (int k = 0; k < 7; ++k)
      if (xq == RL_DEV[k]) {
        xq = __dadd_rn(xq, buffer);
        break;
      }
  }
  if (!(xq >= kr[0] && xq <= kr[nk - 1])) {
    *bad = 1;
    out[i] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  // searchsortedlast: number of knots <= xq
  i64 lo = 0, hi = nk;
  while (lo < hi) {
    const i64 mid = (lo + hi) >> 1;
    if (kr[mid] <= xq)
      lo = mid + 1;
    else
      hi = mid;
  }
  i64 idx = lo;  // 1-based index of the last knot <= xq
  if (idx < 1) idx = 1;
  if (idx > nk - 1) idx = nk - 1;
  const double k0 = kr[idx - 1], k1 = kr[idx];
  const double f = __ddiv_rn(__dsub_rn(xq, k0), __dsub_rn(k1, k0));
  out[i] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, f), kv[idx - 1]), __dmul_rn(f, kv[idx]));
}

// dual_velocity(r, interpolant; buffer) src/utils.jl:51-66: on a discontinuity radius V[i,1] = itp(r - buffer),
// V[i,2] = itp(r + buffer), elsewhere both = itp(r).  out: [n x 2] column-major.
__device__ __forceinline__ double lerp_knots(const double* __restrict__ kr, const double* __restrict__ kv, i64 nk,
                                             double xq, int* bad) {
  if (!(xq >= kr[0] && xq <= kr[nk - 1])) {
    *bad = 1;
    return __longlong_as_double(0x7ff8000000000000LL);
  }
  i64 lo = 0, hi = nk;
  while (lo < hi) {
    const i64 mid = (lo + hi) >> 1;
    if (kr[mid] <= xq)
      lo = mid + 1;
    else
      hi = mid;
  }
  i64 idx = lo;
  if (idx < 1) idx = 1;
  if (idx > nk - 1) idx = nk - 1;
  const double k0 = kr[idx - 1], k1 = kr[idx];
  const double f = __ddiv_rn(__dsub_rn(xq, k0), __dsub_rn(k1, k0));
  return __dadd_rn(__dmul_rn(__dsub_rn(1.0, f), kv[idx - 1]), __dmul_rn(f, kv[idx]));
}
__global__ void dual_velocity_kernel(const double* __restrict__ kr, const double* __restrict__ kv, i64 nk,
                                     const double* __restrict__ r, i64 n, double buffer, double* __restrict__ out,
                                     int* __restrict__ bad) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ri = r[i];
  bool on_layer = false;
#pragma unroll
  for (int k = 0; k < 7; ++k) on_layer = on_layer || (ri == RL_DEV[k]);
  if (on_layer) {
    out[i] = lerp_knots(kr, kv, nk, __dsub_rn(ri, buffer), bad);
    out[n + i] = lerp_knots(kr, kv, nk, __dadd_rn(ri, buffer), bad);
  } else {
    const double v = lerp_knots(kr, kv, nk, ri, bad);
    out[i] = v;
    out[n + i] = v;
  }
}

// closest_point (src/GridAnnulus.jl:823-840): argmin_i sqrt((a_i-pa)^2 + (b_i-pb)^2), FIRST index on ties.
// Pass 1: 64-bit atomicMin on the bit pattern of the (non-negative) distance; pass 2: atomicMin on the index
// among the nodes that attain it.  One grid row (blockIdx.y) per query point.
__device__ __forceinline__ double cp_dist(double a, double b, double pa, double pb) {
  const double da = __dsub_rn(a, pa), db = __dsub_rn(b, pb);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db)));
}
__global__ void closest_pass1_kernel(const double* __restrict__ a, const double* __restrict__ b, i64 n,
                                     const double* __restrict__ pa, const double* __restrict__ pb,
                                     u64* __restrict__ best) {
  const int q = blockIdx.y;
  const double qa = pa[q], qb = pb[q];
  u64 m = ~0ull;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const double d = cp_dist(a[i], b[i], qa, qb);
    if (d == d) {  // NaN never wins (`di < dist` is false for NaN)
      const u64 bits = (u64)__double_as_longlong(d);
      m = bits < m ? bits : m;
    }
  }
  for (int o = 16; o; o >>= 1) {
    const u64 other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other < m ? other : m;
  }
  if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(&best[q], m);
}
__global__ void closest_pass2_kernel(const double* __restrict__ a, const double* __restrict__ b, i64 n,
                                     const double* __restrict__ pa, const double* __restrict__ pb,
                                     const u64* __restrict__ best, u64* __restrict__ index) {
  const int q = blockIdx.y;
  const double qa = pa[q], qb = pb[q];
  const u64 target = best[q];
  u64 m = ~0ull;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const double d = cp_dist(a[i], b[i], qa, qb);
    if (d == d && (u64)__double_as_longlong(d) == target) m = (u64)i < m ? (u64)i : m;
  }
  for (int o = 16; o; o >>= 1) {
    const u64 other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other < m ? other : m;
  }
  if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(&index[q], m);
}

__global__ void prev_i32_to_i64_kernel(const i32* __restrict__ src, i64* __restrict__ dst, i64 count) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = (i64)src[i] + 1;  // -1 (never set) -> 0
}
__global__ void prev_i64_to_i32_kernel(const i64* __restrict__ src, i32* __restrict__ dst, i64 count) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = (i32)(src[i] - 1);
}

// recontruct_path(prev, source, receiver) src/SSSP/ssspm.jl:30-40: one thread per receiver, two passes
// (count, then fill).  prev is 0-based int32, -1 = never set.  len = -1 if the chase fails.
__global__ void path_len_kernel(const i32* __restrict__ prev, i64 n, i32 source, const i32* __restrict__ recv,
                                i64 nrec, i64* __restrict__ len) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrec) return;
  const i32 rcv = recv[k];
  if (rcv < 0 || rcv >= n) {
    len[k] = -1;
    return;
  }
  i64 l = 1;
  i32 ip = (rcv == source) ? source : prev[rcv];
  i64 steps = 0;
  while (ip != source) {
    if (ip < 0 || ip >= n || ++steps > n) {
      len[k] = -1;
      return;
    }
    ++l;
    ip = prev[ip];
  }
  len[k] = l + 1;
}
__global__ void path_fill_kernel(const i32* __restrict__ prev, i32 source, const i32* __restrict__ recv, i64 nrec,
                                 const i64* __restrict__ off, i64* __restrict__ out) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrec) return;
  i64 o = off[k];
  const i64 end = off[k + 1];
  const i32 rcv = recv[k];
  out[o++] = (i64)rcv + 1;
  i32 ip = (rcv == source) ? source : prev[rcv];
  while (ip != source && o < end - 1) {
    out[o++] = (i64)ip + 1;
    ip = prev[ip];
  }
  out[o] = (i64)source + 1;
}


// recontruct_path(D, source, receiver) src/SSSP/ssspm.jl:14-28 (the struct method): the chase runs `while ipath ∉ path`
// -- until a node repeats -- and `source` is appended afterwards, whether or not the chase passed through it.  The
// number of distinct nodes of the sequence x0 = receiver, x(k+1) = prev[x(k)] is mu + lambda (tail + cycle), found
// with Brent's cycle detection; an unset entry (the reference: BoundsError / UndefRefError) gives len = -1.
__device__ __forceinline__ i32 chase(const i32* __restrict__ prev, i64 n, i32 v) {
  return (v < 0 || v >= n) ? -1 : prev[v];
}
__global__ void path_len_guarded_kernel(const i32* __restrict__ prev, i64 n, const i32* __restrict__ recv, i64 nrec,
                                        i64* __restrict__ len) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrec) return;
  const i32 x0 = recv[k];
  len[k] = -1;
  if (x0 < 0 || x0 >= n) return;
  i64 power = 1, lam = 1;
  i32 tortoise = x0, hare = chase(prev, n, x0);
  while (tortoise != hare) {
    if (hare < 0) return;
    if (power == lam) {
      tortoise = hare;
      power *= 2;
      lam = 0;
    }
    hare = chase(prev, n, hare);
    ++lam;
  }
  tortoise = hare = x0;
  for (i64 q = 0; q < lam; ++q) hare = chase(prev, n, hare);
  i64 mu = 0;
  while (tortoise != hare) {
    tortoise = chase(prev, n, tortoise);
    hare = chase(prev, n, hare);
    ++mu;
  }
  len[k] = mu + lam + 1;
}
__global__ void path_fill_guarded_kernel(const i32* __restrict__ prev, i32 source, const i32* __restrict__ recv,
                                         i64 nrec, const i64* __restrict__ off, i64* __restrict__ out) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrec) return;
  i64 o = off[k];
  const i64 end = off[k + 1];
  i32 ip = recv[k];
  while (o < end - 1) {
    out[o++] = (i64)ip + 1;
    ip = prev[ip];
  }
  out[o] = (i64)source + 1;
}

// travel_times(D, gr, receivers) src/utils.jl:4-8 for a batch: out[s, k] = dist[s, receivers[k]]
__global__ void travel_times_kernel(const double* __restrict__ dist, i64 n, i64 nsrc, const i64* __restrict__ recv,
                                    i64 nrec, double* __restrict__ out) {
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsrc * nrec) return;
  const i64 s = t / nrec, k = t - s * nrec;
  out[t] = dist[s * n + (recv[k] - 1)];
}

// polardistance3D(a, b) src/StructuredGrid.jl:245-255: spherical2cart (:225-230) of both (theta, phi, r) triples,
// then distance3D (:239-241)
__global__ void polardistance3d_kernel(const double* __restrict__ a, const double* __restrict__ b, i64 count,
                                       double* __restrict__ out) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  double c[2][3];
  const double* pts[2] = {a + 3 * k, b + 3 * k};
  for (int q = 0; q < 2; ++q) {
    const double th = pts[q][0], ph = pts[q][1], r = pts[q][2];
    c[q][0] = __dmul_rn(__dmul_rn(r, cos(ph)), sin(th));
    c[q][1] = __dmul_rn(__dmul_rn(r, sin(ph)), sin(th));
    c[q][2] = __dmul_rn(r, cos(th));
  }
  const double dx = __dsub_rn(c[0][0], c[1][0]), dy = __dsub_rn(c[0][1], c[1][1]), dz = __dsub_rn(c[0][2], c[1][2]);
  out[k] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

}  // namespace

int interp_velocity_device(const double* kr_h, const double* kv_h, i64 nk, const double* r_dev, i64 n,
                           double buffer, double* out_dev) {
  RT_ARG(kr_h && kv_h && nk >= 2 && r_dev && out_dev && n >= 0, "bad interpolation arguments");
  for (i64 k = 0; k + 1 < nk; ++k) RT_ARG(kr_h[k] < kr_h[k + 1], "knots must be strictly ascending");
  DevBuf<double> kr, kv;
  DevBuf<int> bad;
  RT_TRY(kr.upload(kr_h, nk));
  RT_TRY(kv.upload(kv_h, nk));
  RT_TRY(bad.alloc(1));
  RT_TRY(bad.zero());
  if (n) interp_kernel<<<grid_for(n, 256), 256>>>(kr.p, kv.p, nk, r_dev, n, buffer, out_dev, bad.p);
  RT_CUDA(cudaGetLastError());
  int hb = 0;
  RT_CUDA(cudaMemcpy(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (hb) {
    rt_set_error("interpolation point outside the knots (BoundsError in the reference)");
    return RT_ERR_RANGE;
  }
  return RT_OK;
}

int dual_velocity_device(const double* kr_h, const double* kv_h, i64 nk, const double* r_dev, i64 n, double buffer,
                         double* out_dev) {
  for (i64 k = 0; k + 1 < nk; ++k) RT_ARG(kr_h[k] < kr_h[k + 1], "knots must be strictly ascending");
  DevBuf<double> kr, kv;
  DevBuf<int> bad;
  RT_TRY(kr.upload(kr_h, nk));
  RT_TRY(kv.upload(kv_h, nk));
  RT_TRY(bad.alloc(1));
  RT_TRY(bad.zero());
  if (n) dual_velocity_kernel<<<grid_for(n, 256), 256>>>(kr.p, kv.p, nk, r_dev, n, buffer, out_dev, bad.p);
  RT_CUDA(cudaGetLastError());
  int hb = 0;
  RT_CUDA(cudaMemcpy(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (hb) {
    rt_set_error("interpolation point outside the knots (BoundsError in the reference)");
    return RT_ERR_RANGE;
  }
  return RT_OK;
}

int closest_point_device(const double* a_dev, const double* b_dev, i64 n, const double* pa, const double* pb,
                         i64 npts, i64* out, cudaStream_t s) {
  RT_ARG(pa && pb && out && npts >= 0, "bad closest_point arguments");
  if (npts == 0) return RT_OK;
  DevBuf<double> dpa, dpb;
  DevBuf<u64> best, index;
  RT_TRY(dpa.upload(pa, npts, s));
  RT_TRY(dpb.upload(pb, npts, s));
  RT_TRY(best.alloc(npts));
  RT_TRY(index.alloc(npts));
  RT_CUDA(cudaMemsetAsync(best.p, 0xff, npts * sizeof(u64), s));
  RT_CUDA(cudaMemsetAsync(index.p, 0xff, npts * sizeof(u64), s));
  const unsigned bx = (unsigned)std::min<i64>(grid_for(n, 256), 1184);
  for (i64 q0 = 0; q0 < npts; q0 += 32768) {  // one grid row per query: gridDim.y <= 65535
    const i64 nq = std::min<i64>(32768, npts - q0);
    dim3 grid(bx, (unsigned)nq);
    closest_pass1_kernel<<<grid, 256, 0, s>>>(a_dev, b_dev, n, dpa.p + q0, dpb.p + q0, best.p + q0);
    closest_pass2_kernel<<<grid, 256, 0, s>>>(a_dev, b_dev, n, dpa.p + q0, dpb.p + q0, best.p + q0, index.p + q0);
  }
  RT_CUDA(cudaGetLastError());
  std::vector<u64> hi(npts);
  RT_CUDA(cudaMemcpyAsync(hi.data(), index.p, npts * sizeof(u64), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  for (i64 q = 0; q < npts; ++q) out[q] = hi[q] == ~0ull ? -1 : (i64)hi[q] + 1;  // -1 as in the reference
  return RT_OK;
}

int prev_to_host_i64(const i32* prev_dev, i64 count, i64* out, cudaStream_t s) {
  // convert in slabs so that the staging buffer stays small
  const i64 slab = (i64)1 << 26;
  DevBuf<i64> tmp;
  RT_TRY(tmp.alloc(std::min(count, slab)));
  for (i64 o = 0; o < count; o += slab) {
    const i64 c = std::min(slab, count - o);
    prev_i32_to_i64_kernel<<<grid_for(c, 256), 256, 0, s>>>(prev_dev + o, tmp.p, c);
    RT_CUDA(cudaMemcpyAsync(out + o, tmp.p, c * sizeof(i64), cudaMemcpyDeviceToHost, s));
    RT_CUDA(cudaStreamSynchronize(s));
  }
  return RT_OK;
}

int prev_to_host_i64_staged(const i32* prev_dev, i64 count, i64* out, i64* stage64, cudaStream_t s) {
  prev_i32_to_i64_kernel<<<grid_for(count, 256), 256, 0, s>>>(prev_dev, stage64, count);
  RT_CUDA(cudaMemcpyAsync(out, stage64, count * sizeof(i64), cudaMemcpyDeviceToHost, s));
  return RT_OK;
}

// prev_dev: 0-based int32 table on the device.  receivers: host, 1-based.  guarded = 1: the struct method
// recontruct_path(D, source, receiver) (ssspm.jl:14-28) instead of the vector method (:30-40).
// Scratch of the path sweeps, kept per host thread and device: a receiver sweep makes two calls per source, and four
// cudaMalloc / cudaFree pairs per call cost more than the pointer chase itself (measured: up to 46 ms per call once other
// libraries have grown the process's address space, against 0.3 ms of kernels).  The work stays on the legacy default
// stream: a caller that filled prev_dev on its own (blocking) stream keeps the implicit ordering it had.
namespace {
struct PathScratch {
  int device = -1;
  DevBuf<i32> recv;
  DevBuf<i64> len, off, out;
};
thread_local PathScratch g_paths;
template <typename T>
int grow(DevBuf<T>& b, size_t count) {
  if (b.n >= count && b.p) return RT_OK;
  size_t cap = 1024;
  while (cap < count) cap *= 2;
  return b.alloc(cap);
}
}  // namespace

int reconstruct_paths_device(const i32* prev_dev, i64 n, i64 source, const i64* receivers, i64 nrec,
                             i64* path_off, i64* path_idx, i64 cap, int guarded) {
  RT_ARG(prev_dev && receivers && path_off && nrec >= 0, "bad path arguments");
  RT_ARG(source >= 1 && source <= n, "source out of range");
  std::vector<i32> r32(nrec);
  for (i64 k = 0; k < nrec; ++k) {
    RT_ARG(receivers[k] >= 1 && receivers[k] <= n, "receiver out of range");
    r32[k] = (i32)(receivers[k] - 1);
  }
  PathScratch& ps = g_paths;
  int dev = 0;
  RT_CUDA(cudaGetDevice(&dev));
  if (ps.device != dev) {  // another device is current on this thread: the scratch lives where the table lives
    ps.recv.release();
    ps.len.release();
    ps.off.release();
    ps.out.release();
    ps.device = dev;
  }
  cudaStream_t s = cudaStreamLegacy;
  RT_TRY(grow(ps.recv, (size_t)nrec));
  RT_TRY(grow(ps.len, (size_t)nrec));
  if (nrec) {
    RT_CUDA(cudaMemcpyAsync(ps.recv.p, r32.data(), nrec * sizeof(i32), cudaMemcpyHostToDevice, s));
    if (guarded)
      path_len_guarded_kernel<<<grid_for(nrec, 128), 128, 0, s>>>(prev_dev, n, ps.recv.p, nrec, ps.len.p);
    else
      path_len_kernel<<<grid_for(nrec, 128), 128, 0, s>>>(prev_dev, n, (i32)(source - 1), ps.recv.p, nrec, ps.len.p);
  }
  RT_CUDA(cudaGetLastError());
  std::vector<i64> hl(nrec);
  if (nrec) RT_CUDA(cudaMemcpyAsync(hl.data(), ps.len.p, nrec * sizeof(i64), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  path_off[0] = 0;
  for (i64 k = 0; k < nrec; ++k) {
    if (hl[k] < 0) {
      rt_set_error(guarded ? "recontruct_path: the chase from receiver %lld (source %lld) reads an unset predecessor"
                           : "recontruct_path: receiver %lld does not reach source %lld",
                   (long long)receivers[k], (long long)source);
      return RT_ERR_NOPATH;
    }
    path_off[k + 1] = path_off[k] + hl[k];
  }
  if (!path_idx) return RT_OK;
  RT_ARG(cap >= path_off[nrec], "path_idx capacity too small");
  if (nrec == 0) return RT_OK;
  RT_TRY(grow(ps.off, (size_t)nrec + 1));
  RT_TRY(grow(ps.out, (size_t)std::max<i64>(path_off[nrec], 1)));
  RT_CUDA(cudaMemcpyAsync(ps.off.p, path_off, (nrec + 1) * sizeof(i64), cudaMemcpyHostToDevice, s));
  if (guarded)
    path_fill_guarded_kernel<<<grid_for(nrec, 128), 128, 0, s>>>(prev_dev, (i32)(source - 1), ps.recv.p, nrec, ps.off.p, ps.out.p);
  else
    path_fill_kernel<<<grid_for(nrec, 128), 128, 0, s>>>(prev_dev, (i32)(source - 1), ps.recv.p, nrec, ps.off.p, ps.out.p);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaMemcpyAsync(path_idx, ps.out.p, path_off[nrec] * sizeof(i64), cudaMemcpyDeviceToHost, s));
  RT_CUDA(cudaStreamSynchronize(s));
  return RT_OK;
}

int travel_times_device(const double* dist_dev, i64 n, i64 nsrc, const i64* receivers, i64 nrec, double* out) {
  RT_ARG(dist_dev && receivers && out && n > 0 && nsrc >= 0 && nrec >= 0, "bad travel_times arguments");
  for (i64 k = 0; k < nrec; ++k) RT_ARG(receivers[k] >= 1 && receivers[k] <= n, "receiver out of range");
  if (nsrc * nrec == 0) return RT_OK;
  DevBuf<i64> recv;
  DevBuf<double> dout;
  RT_TRY(recv.upload(receivers, nrec));
  RT_TRY(dout.alloc(nsrc * nrec));
  travel_times_kernel<<<grid_for(nsrc * nrec, 256), 256>>>(dist_dev, n, nsrc, recv.p, nrec, dout.p);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaMemcpy(out, dout.p, nsrc * nrec * sizeof(double), cudaMemcpyDeviceToHost));
  return RT_OK;
}

int polardistance3d_device(const double* a, const double* b, i64 count, double* out) {
  RT_ARG(a && b && out && count >= 0, "bad polardistance3D arguments");
  if (count == 0) return RT_OK;
  DevBuf<double> da, db, dout;
  RT_TRY(da.upload(a, 3 * count));
  RT_TRY(db.upload(b, 3 * count));
  RT_TRY(dout.alloc(count));
  polardistance3d_kernel<<<grid_for(count, 256), 256>>>(da.p, db.p, count, dout.p);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaMemcpy(out, dout.p, count * sizeof(double), cudaMemcpyDeviceToHost));
  return RT_OK;
}

int prev_host_to_device_i32(const i64* prev, i64 n, DevBuf<i32>& out) {
  DevBuf<i64> tmp;
  RT_TRY(tmp.upload(prev, n));
  RT_TRY(out.alloc(n));
  prev_i64_to_i32_kernel<<<grid_for(n, 256), 256>>>(tmp.p, out.p, n);
  RT_CUDA(cudaGetLastError());
  RT_CUDA(cudaDeviceSynchronize());
  return RT_OK;
}

int round_to_f32_device(const double* in, double* out, i64 n, cudaStream_t s) {
  if (n <= 0) return RT_OK;
  round_f32_kernel<<<(unsigned)std::min<i64>((n + 255) / 256, 148 * 16), 256, 0, s>>>(in, out, n);
  RT_CUDA(cudaGetLastError());
  return RT_OK;
}
