// api.cu -- the extern "C" surface declared in include/rt_sssp.h: handle management, argument checks and
// host <-> device staging around the kernels.  No compute happens on the host.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

int dual_velocity_device(const double* kr, const double* kv, i64 nk, const double* r_dev, i64 n, double buffer,
                         double* out_dev);
int prev_to_host_i64_staged(const i32* prev_dev, i64 count, i64* out, i64* stage64, cudaStream_t s);
int reconstruct_paths_device(const i32* prev_dev, i64 n, i64 source, const i64* receivers, i64 nrec, i64* path_off,
                             i64* path_idx, i64 cap, int guarded);
int prev_host_to_device_i32(const i64* prev, i64 n, DevBuf<i32>& out);

static thread_local char g_err[512] = "";

void rt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int make_handle(rt_mesh** out) {
  RT_ARG(out, "null output handle");
  *out = nullptr;
  int dev = 0;
  RT_CUDA(cudaGetDevice(&dev));
  rt_mesh* h = new rt_mesh();
  h->device = dev;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    rt_set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
    return RT_ERR_CUDA;
  }
  *out = h;
  return RT_OK;
}

static i64 mesh_n(const rt_mesh* m) {
  i64 n = 0;
  if (m->kind == 3) grid3d_n(m, &n);
  if (m->kind == 2) {
    i64 s[8];
    mesh2d_sizes(m, s);
    n = s[0];
  }
  return n;
}

extern "C" {

const char* rt_last_error(void) { return g_err; }
const char* rt_version(void) { return "rt_sssp 0.1 (sm_100a)"; }

int rt_device_count(int* count) {
  RT_ARG(count, "null count");
  *count = 0;
  RT_CUDA(cudaGetDeviceCount(count));
  return RT_OK;
}

int rt_set_device(int device) {
  RT_CUDA(cudaSetDevice(device));
  return RT_OK;
}

int rt_annulus_build(int64_t ntheta, int64_t nr, double spacing, rt_mesh** out) {
  RT_ARG(ntheta >= 3 && nr >= 2 && spacing > 0.0, "init_annulus needs ntheta >= 3, nr >= 2, spacing > 0");
  RT_TRY(make_handle(out));
  int rc = annulus_build_device(*out, ntheta, nr, spacing);
  if (rc != RT_OK) {
    rt_mesh_free(*out);
    *out = nullptr;
  }
  return rc;
}

int rt_mesh_sizes(const rt_mesh* m, int64_t sizes[8]) {
  RT_ARG(m && sizes, "null argument");
  if (m->kind == 2) return mesh2d_sizes(m, sizes);
  std::memset(sizes, 0, 8 * sizeof(int64_t));
  sizes[0] = mesh_n(m);
  return RT_OK;
}

int rt_mesh_export(const rt_mesh* m, double* x, double* z, double* theta, double* r, int64_t* e2n_off,
                   int64_t* e2n_idx, int64_t* G_colptr, int64_t* G_rowval, int64_t* halo, int64_t* nbr_off,
                   int64_t* nbr_idx, int8_t* el_type) {
  RT_ARG(m && m->kind == 2, "rt_mesh_export needs a 2-D mesh");
  RT_CUDA(cudaSetDevice(m->device));
  return mesh2d_export(m, x, z, theta, r, e2n_off, e2n_idx, G_colptr, G_rowval, halo, nbr_off, nbr_idx, el_type);
}

int rt_mesh_from_arrays(int64_t n, int64_t nel, const int64_t* e2n_off, const int64_t* e2n_idx,
                        const int64_t* G_colptr, const int64_t* G_rowval, const int64_t* halo, int64_t halo_rows,
                        const double* x, const double* z, const double* theta, const double* r, rt_mesh** out) {
  RT_TRY(make_handle(out));
  int rc = mesh2d_from_host(*out, n, nel, e2n_off, e2n_idx, G_colptr, G_rowval, halo, halo_rows, x, z, theta, r);
  if (rc != RT_OK) {
    rt_mesh_free(*out);
    *out = nullptr;
  }
  return rc;
}

int rt_grid3d_build(const double c0[3], const double c1[3], const int64_t nn[3], int star_levels,
                    int coord_system, rt_mesh** out) {
  RT_TRY(make_handle(out));
  int rc = grid3d_build(*out, c0, c1, nn, star_levels, coord_system);
  if (rc != RT_OK) {
    rt_mesh_free(*out);
    *out = nullptr;
  }
  return rc;
}

int rt_grid3d_export(const rt_mesh* m, double* X, double* Y, double* Z) {
  RT_ARG(m && m->kind == 3, "rt_grid3d_export needs a 3-D grid");
  RT_CUDA(cudaSetDevice(m->device));
  return grid3d_export(m, X, Y, Z);
}

int rt_mesh_free(rt_mesh* m) {
  if (!m) return RT_OK;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  mesh2d_free(m);
  grid3d_free(m);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
  return RT_OK;
}

int rt_interp_velocity(const double* knots_r, const double* knots_v, int64_t nk, const double* r, int64_t n,
                       double buffer, double* out) {
  RT_ARG(knots_r && knots_v && r && out && n >= 0 && nk >= 2, "bad interpolation arguments");
  DevBuf<double> dr, dout;
  RT_TRY(dr.upload(r, n));
  RT_TRY(dout.alloc(n));
  RT_TRY(interp_velocity_device(knots_r, knots_v, nk, dr.p, n, buffer, dout.p));
  if (n) RT_CUDA(cudaMemcpy(out, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  return RT_OK;
}

int rt_interp_velocity_dev(const double* knots_r, const double* knots_v, int64_t nk, const double* r_dev,
                           int64_t n, double buffer, double* out_dev) {
  return interp_velocity_device(knots_r, knots_v, nk, r_dev, n, buffer, out_dev);
}

int rt_mesh_coords_dev(const rt_mesh* m, const double** a, const double** b, const double** c, const double** d) {
  RT_ARG(m && a && b && c && d, "null argument");
  if (m->kind == 2) return mesh2d_coords(m, a, b, c, d);
  if (m->kind == 3) return grid3d_coords(m, a, b, c, d);
  rt_set_error("mesh handle is empty");
  return RT_ERR_ARG;
}

int rt_interpolate_cells(rt_mesh* m, const int8_t* el_type, double* V) {
  RT_ARG(m && m->kind == 2 && V, "rt_interpolate_cells needs a 2-D mesh and a velocity array");
  RT_CUDA(cudaSetDevice(m->device));
  i64 sz[8];
  mesh2d_sizes(m, sz);
  DevBuf<double> dV;
  RT_TRY(dV.upload(V, sz[0]));
  RT_CUDA(cudaDeviceSynchronize());
  RT_TRY(mesh2d_interpolate_cells(m, el_type, dV.p));
  RT_CUDA(cudaMemcpy(V, dV.p, sz[0] * sizeof(double), cudaMemcpyDeviceToHost));
  return RT_OK;
}

int rt_nodal_adjacency(rt_mesh* m, int64_t* deg, int64_t* list_off, int64_t* list_idx, int64_t cap) {
  RT_ARG(m && m->kind == 2, "rt_nodal_adjacency needs a 2-D mesh");
  RT_CUDA(cudaSetDevice(m->device));
  return mesh2d_nodal_adjacency(m, deg, list_off, list_idx, cap);
}

int rt_rcm(rt_mesh* m, int64_t* perm_out) {
  RT_ARG(m && m->kind == 2 && perm_out, "rt_rcm needs a 2-D mesh and an output array");
  RT_CUDA(cudaSetDevice(m->device));
  return mesh2d_rcm(m, perm_out);
}

int rt_partition_grid(const rt_mesh* m, int32_t* id_out) {
  RT_ARG(m && m->kind == 2, "rt_partition_grid needs a 2-D mesh");
  RT_CUDA(cudaSetDevice(m->device));
  return mesh2d_partition(m, id_out);
}

int rt_bfm_continue(rt_mesh* m, const double* U, const uint8_t* allowed, const int64_t* seeds, int64_t nseeds,
                    double* dist_inout, int64_t* prev_inout, rt_stats* stats) {
  RT_ARG(m && m->kind == 2 && U && dist_inout && prev_inout && nseeds >= 0 && (nseeds == 0 || seeds),
         "rt_bfm_continue needs a 2-D mesh, U, the (dist, prev) state and the seed nodes");
  RT_CUDA(cudaSetDevice(m->device));
  m->f32 = false;
  i64 sz[8];
  mesh2d_sizes(m, sz);
  DevBuf<double> dU;
  RT_TRY(dU.upload(U, sz[0], m->stream));
  return bfm2d_continue(m, dU.p, allowed, seeds, nseeds, dist_inout, prev_inout, stats);
}

int rt_sssp_nodal(rt_mesh* m, const double* U, int64_t source, int algorithm, double* dist_out, int64_t* prev_out,
                  rt_stats* stats) {
  RT_ARG(m && m->kind == 2 && U, "rt_sssp_nodal needs a 2-D mesh and a velocity array");
  RT_ARG(algorithm == 0 || algorithm == 1, "algorithm must be 0 (dijkstra) or 1 (radius_stepping)");
  RT_CUDA(cudaSetDevice(m->device));
  i64 sz[8];
  mesh2d_sizes(m, sz);
  DevBuf<double> dU;
  RT_TRY(dU.upload(U, sz[0], m->stream));
  return mesh2d_sssp_nodal(m, dU.p, source, algorithm, dist_out, prev_out, stats);
}

int rt_closest_point3d(const rt_mesh* m, const double* px, const double* py, const double* pz, int64_t npts,
                       int64_t* index_out) {
  RT_ARG(m && m->kind == 3, "rt_closest_point3d needs a 3-D grid");
  RT_CUDA(cudaSetDevice(m->device));
  return grid3d_closest(m, px, py, pz, npts, index_out);
}

int rt_grid3d_axes(const rt_mesh* m, double* x, double* y, double* z) {
  RT_ARG(m && m->kind == 3, "rt_grid3d_axes needs a 3-D grid");
  RT_CUDA(cudaSetDevice(m->device));
  return grid3d_axes(m, x, y, z);
}

int rt_grid3d_points(const rt_mesh* m, const int64_t* I, int64_t count, double* xyz, int64_t* ijk) {
  RT_ARG(m && m->kind == 3, "rt_grid3d_points needs a 3-D grid");
  RT_CUDA(cudaSetDevice(m->device));
  return grid3d_points(m, I, count, xyz, ijk);
}

int rt_grid3d_connectivity(const rt_mesh* m, int64_t first_el, int64_t count, int64_t* e2n) {
  RT_ARG(m && m->kind == 3, "rt_grid3d_connectivity needs a 3-D grid");
  RT_CUDA(cudaSetDevice(m->device));
  return grid3d_connectivity(m, first_el, count, e2n);
}

int rt_polardistance3d(const double* a, const double* b, int64_t count, double* out) {
  return polardistance3d_device(a, b, count, out);
}

int rt_travel_times_dev(const double* dist_dev, int64_t n, int64_t nsrc, const int64_t* receivers, int64_t nrec,
                        double* out) {
  return travel_times_device(dist_dev, n, nsrc, receivers, nrec, out);
}

int rt_travel_times(const double* dist, int64_t n, int64_t nsrc, const int64_t* receivers, int64_t nrec, double* out) {
  RT_ARG(dist && n > 0 && nsrc >= 0, "bad travel_times arguments");
  DevBuf<double> d;
  RT_TRY(d.upload(dist, (size_t)n * (size_t)nsrc));
  return travel_times_device(d.p, n, nsrc, receivers, nrec, out);
}

int rt_closest_point(const rt_mesh* m, const double* pa, const double* pb, int64_t npts, int system,
                     int64_t* index_out) {
  RT_ARG(m && m->kind == 2, "rt_closest_point needs a 2-D mesh (3-D grids: rt_closest_point3d)");
  RT_ARG(system == 0 || system == 1, "system must be 0 (:cartesian) or 1 (:polar)");
  RT_CUDA(cudaSetDevice(m->device));
  return mesh2d_closest(m, pa, pb, npts, system, index_out);
}

int rt_bfm_solve_dev(rt_mesh* m, const double* U_dev, const int64_t* sources, int64_t nsrc, int precision,
                     double* dist_dev, int32_t* prev_dev, rt_stats* stats) {
  RT_ARG(m && U_dev && sources && nsrc >= 0, "null argument");
  RT_ARG(precision == 64 || precision == 32, "precision must be 64 (bfm) or 32 (the Float32 path of bfm_gpu)");
  RT_ARG(m->kind == 2 || m->kind == 3, "mesh handle is empty");
  RT_CUDA(cudaSetDevice(m->device));
  m->f32 = precision == 32;
  if (m->f32) {
    // Float32.(U) (bfm_gpu.jl:173); the coordinates are rounded once per mesh inside the solvers
    const i64 n = mesh_n(m);
    if (m->Uf.n != (size_t)n) RT_TRY(m->Uf.alloc(n));
    RT_TRY(round_to_f32_device(U_dev, m->Uf.p, n, m->stream));
    U_dev = m->Uf.p;
  }
  const int rc = m->kind == 2 ? bfm2d_solve(m, U_dev, sources, nsrc, dist_dev, prev_dev, stats)
                              : bfm3d_solve(m, U_dev, sources, nsrc, dist_dev, prev_dev, stats);
  m->f32 = false;
  return rc;
}

int rt_bfm_solve(rt_mesh* m, const double* U, const int64_t* sources, int64_t nsrc, int precision,
                 double* dist_out, int64_t* prev_out, rt_stats* stats) {
  RT_ARG(m && U && sources && nsrc >= 0, "null argument");
  RT_CUDA(cudaSetDevice(m->device));
  const i64 n = mesh_n(m);
  cudaStream_t cs = m->stream;
  // sources go through in chunks so that small meshes can advance several sources in lock step while the
  // staging buffers stay bounded (<= 32 tables, <= ~2 GB)
  // (2-D meshes of <= 4 M nodes advance up to 512 sources in lock step, see bfm2d_push.cu)
  const bool wide = m->kind == 2 && n <= 4000000;
  i64 chunk = std::max<i64>(1, std::min<i64>(std::min<i64>(nsrc, wide ? 512 : 32),
                                             ((i64)1 << (wide ? 29 : 28)) / std::max<i64>(n, 1)));
  if (m->stage_U.n != (size_t)n) RT_TRY(m->stage_U.alloc(n));
  if (dist_out && m->stage_dist.n < (size_t)(chunk * n)) RT_TRY(m->stage_dist.alloc(chunk * n));
  if (prev_out && m->stage_prev.n < (size_t)(chunk * n)) {
    RT_TRY(m->stage_prev.alloc(chunk * n));
    RT_TRY(m->stage_prev64.alloc(chunk * n));
  }
  RT_CUDA(cudaMemcpyAsync(m->stage_U.p, U, n * sizeof(double), cudaMemcpyHostToDevice, cs));
  rt_stats total = {};
  for (i64 s = 0; s < nsrc; s += chunk) {
    const i64 c = std::min(chunk, nsrc - s);
    rt_stats st = {};
    RT_TRY(rt_bfm_solve_dev(m, m->stage_U.p, sources + s, c, precision, dist_out ? m->stage_dist.p : nullptr,
                            prev_out ? m->stage_prev.p : nullptr, &st));
    if (dist_out)
      RT_CUDA(cudaMemcpyAsync(dist_out + s * n, m->stage_dist.p, c * n * sizeof(double), cudaMemcpyDeviceToHost, cs));
    if (prev_out) RT_TRY(prev_to_host_i64_staged(m->stage_prev.p, c * n, prev_out + s * n, m->stage_prev64.p, cs));
    RT_CUDA(cudaStreamSynchronize(cs));
    total.sweeps += st.sweeps;
    total.relaxed_edges += st.relaxed_edges;
    total.vertex_updates += st.vertex_updates;
    total.graph_edges = st.graph_edges;
    total.kernel_ms += st.kernel_ms;
    total.relax_ms += st.relax_ms;
    total.relax_launches += st.relax_launches;
    total.total_launches += st.total_launches;
    total.prev_ms += st.prev_ms;
    total.screened_edges += st.screened_edges;
    total.exact_edges += st.exact_edges;
  }
  if (stats) *stats = total;
  return RT_OK;
}

int rt_bfm_solve_dual(rt_mesh* m, const double* U2, const int64_t* sources, int64_t nsrc, double* dist_out,
                      int64_t* prev_out, rt_stats* stats) {
  RT_ARG(m && m->kind == 2 && U2 && sources && nsrc >= 0, "rt_bfm_solve_dual needs a 2-D mesh, U[n x 2] and sources");
  RT_CUDA(cudaSetDevice(m->device));
  m->f32 = false;  // an earlier precision = 32 call that failed half way must not leak into this Float64 solve
  const i64 n = mesh_n(m);
  cudaStream_t cs = m->stream;
  DevBuf<double> dU;
  RT_TRY(dU.upload(U2, 2 * n, cs));
  if (dist_out && m->stage_dist.n < (size_t)n) RT_TRY(m->stage_dist.alloc(n));
  if (prev_out && m->stage_prev.n < (size_t)n) {
    RT_TRY(m->stage_prev.alloc(n));
    RT_TRY(m->stage_prev64.alloc(n));
  }
  rt_stats total = {};
  for (i64 s = 0; s < nsrc; ++s) {
    rt_stats st = {};
    RT_TRY(bfm2d_solve_dual(m, dU.p, sources + s, 1, dist_out ? m->stage_dist.p : nullptr,
                            prev_out ? m->stage_prev.p : nullptr, &st));
    if (dist_out)
      RT_CUDA(cudaMemcpyAsync(dist_out + s * n, m->stage_dist.p, n * sizeof(double), cudaMemcpyDeviceToHost, cs));
    if (prev_out) RT_TRY(prev_to_host_i64_staged(m->stage_prev.p, n, prev_out + s * n, m->stage_prev64.p, cs));
    RT_CUDA(cudaStreamSynchronize(cs));
    total.sweeps += st.sweeps;
    total.relaxed_edges += st.relaxed_edges;
    total.vertex_updates += st.vertex_updates;
    total.graph_edges = st.graph_edges;
    total.kernel_ms += st.kernel_ms;
    total.relax_ms += st.relax_ms;
    total.relax_launches += st.relax_launches;
    total.total_launches += st.total_launches;
    total.prev_ms += st.prev_ms;
  }
  if (stats) *stats = total;
  return RT_OK;
}

int rt_bfm_solve_multi(rt_mesh* const* meshes, int ndev, const double* U, const int64_t* sources, int64_t nsrc,
                       int precision, double* dist_out, int64_t* prev_out, rt_stats* stats) {
  RT_ARG(meshes && ndev >= 1 && U && sources && nsrc >= 0, "null argument");
  for (int d = 0; d < ndev; ++d) RT_ARG(meshes[d] && meshes[d]->kind == meshes[0]->kind, "replicas must be meshes of one kind");
  const i64 n = mesh_n(meshes[0]);
  for (int d = 1; d < ndev; ++d) {
    RT_ARG(mesh_n(meshes[d]) == n, "replicas must have the same number of nodes");
    for (int e = 0; e < d; ++e) RT_ARG(meshes[e] != meshes[d], "every replica needs its own handle");
  }
  const i64 k = (nsrc + ndev - 1) / ndev;
  std::vector<int> rc(ndev, RT_OK);
  std::vector<std::string> err(ndev);
  std::vector<rt_stats> st(ndev);
  std::vector<std::thread> workers;
  for (int d = 0; d < ndev; ++d) {
    workers.emplace_back([&, d]() {
      const i64 lo = std::min<i64>(nsrc, d * k), hi = std::min<i64>(nsrc, (d + 1) * k);
      st[d] = rt_stats{};
      if (hi <= lo) return;
      rc[d] = rt_bfm_solve(meshes[d], U, sources + lo, hi - lo, precision, dist_out ? dist_out + lo * n : nullptr,
                           prev_out ? prev_out + lo * n : nullptr, &st[d]);
      if (rc[d] != RT_OK) err[d] = rt_last_error();  // the message lives in the worker's thread-local buffer
    });
  }
  for (auto& w : workers) w.join();
  rt_stats total = {};
  for (int d = 0; d < ndev; ++d) {
    if (rc[d] != RT_OK) {
      rt_set_error("replica %d (device %d): %s", d, meshes[d]->device, err[d].c_str());
      return rc[d];
    }
    total.sweeps += st[d].sweeps;
    total.relaxed_edges += st[d].relaxed_edges;
    total.vertex_updates += st[d].vertex_updates;
    total.graph_edges = std::max(total.graph_edges, st[d].graph_edges);
    total.kernel_ms = std::max(total.kernel_ms, st[d].kernel_ms);
    total.relax_ms = std::max(total.relax_ms, st[d].relax_ms);
    total.relax_launches += st[d].relax_launches;
    total.total_launches += st[d].total_launches;
    total.prev_ms = std::max(total.prev_ms, st[d].prev_ms);
  }
  if (stats) *stats = total;
  return RT_OK;
}

int rt_dual_velocity(const double* knots_r, const double* knots_v, int64_t nk, const double* r, int64_t n,
                     double buffer, double* out) {
  RT_ARG(knots_r && knots_v && r && out && n >= 0 && nk >= 2 && buffer >= 0.0, "bad dual_velocity arguments");
  DevBuf<double> dr, dout;
  RT_TRY(dr.upload(r, n));
  RT_TRY(dout.alloc(2 * n));
  RT_TRY(dual_velocity_device(knots_r, knots_v, nk, dr.p, n, buffer, dout.p));
  if (n) RT_CUDA(cudaMemcpy(out, dout.p, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
  return RT_OK;
}

int rt_set_option(rt_mesh* m, const char* key, double value) {
  RT_ARG(m && key, "null argument");
  if (!std::strcmp(key, "schedule")) {
    RT_ARG(value == 0.0 || value == 1.0, "schedule must be 0 (jacobi) or 1 (near-far)");
    m->opts.schedule = (int)value;
  } else if (!std::strcmp(key, "profile_timers")) {
    m->opts.profile_timers = value != 0.0;
  } else if (!std::strcmp(key, "check_every")) {
    RT_ARG(value >= 0.0 && value <= 1024.0, "check_every must be in 0..1024");
    m->opts.check_every = (int)value;
  } else if (!std::strcmp(key, "delta")) {
    RT_ARG(value >= 0.0, "delta must be >= 0");
    m->opts.delta = value;
  } else if (!std::strcmp(key, "cta_units")) {
    m->opts.cta_units = value != 0.0;
  } else if (!std::strcmp(key, "warp_units")) {
    m->opts.warp_units = value < 0.0 ? -1 : (value != 0.0);
  } else if (!std::strcmp(key, "batch")) {
    RT_ARG(value >= 0.0 && value <= 1024.0, "batch must be in 0..1024");
    m->opts.batch = (int)value;
  } else if (!std::strcmp(key, "packed_prev")) {
    m->opts.packed_prev = value != 0.0;
  } else if (!std::strcmp(key, "persistent")) {
    m->opts.persistent = value < 0.0 ? -1 : (value != 0.0);
  } else if (!std::strcmp(key, "weight3d")) {
    RT_ARG(value == 0.0 || value == 1.0, "weight3d must be 0 (weights.jl:20) or 1 (Dijsktra.jl:388)");
    m->opts.weight3d = (int)value;
  } else if (!std::strcmp(key, "use_graph")) {
    m->opts.use_graph = value != 0.0;
  } else if (!std::strcmp(key, "compact")) {
    m->opts.compact = value != 0.0;
  } else if (!std::strcmp(key, "group_screen")) {
    m->opts.group_screen = value != 0.0;
  } else if (!std::strcmp(key, "target_lists")) {
    m->opts.target_lists = value != 0.0;
  } else if (!std::strcmp(key, "canonical_prev")) {
    m->opts.canonical_prev = value != 0.0;
  } else if (!std::strcmp(key, "early_advance")) {
    m->opts.early_advance = value;
  } else if (!std::strcmp(key, "fuse_begin")) {
    m->opts.fuse_begin = value != 0.0;
  } else if (!std::strcmp(key, "tile_pull")) {
    m->opts.tile_pull = value != 0.0;
  } else if (!std::strcmp(key, "delta_factor")) {
    RT_ARG(value >= 0.0, "delta_factor must be >= 0");
    m->opts.delta_factor = value;
  } else {
    rt_set_error("unknown option '%s'", key);
    return RT_ERR_ARG;
  }
  return RT_OK;
}

int rt_reconstruct_paths(const int64_t* prev, int64_t n, int64_t source, const int64_t* receivers, int64_t nrec,
                         int64_t* path_off, int64_t* path_idx, int64_t cap) {
  RT_ARG(prev && n > 0, "null prev table");
  DevBuf<i32> dprev;
  RT_TRY(prev_host_to_device_i32(prev, n, dprev));
  return reconstruct_paths_device(dprev.p, n, source, receivers, nrec, path_off, path_idx, cap, 0);
}

int rt_reconstruct_paths_guarded(const int64_t* prev, int64_t n, int64_t source, const int64_t* receivers,
                                 int64_t nrec, int64_t* path_off, int64_t* path_idx, int64_t cap) {
  RT_ARG(prev && n > 0, "null prev table");
  DevBuf<i32> dprev;
  RT_TRY(prev_host_to_device_i32(prev, n, dprev));
  return reconstruct_paths_device(dprev.p, n, source, receivers, nrec, path_off, path_idx, cap, 1);
}

int rt_reconstruct_paths_dev(const int32_t* prev_dev, int64_t n, int64_t source, const int64_t* receivers,
                             int64_t nrec, int64_t* path_off, int64_t* path_idx, int64_t cap) {
  return reconstruct_paths_device(prev_dev, n, source, receivers, nrec, path_off, path_idx, cap, 0);
}

}  // extern "C"
