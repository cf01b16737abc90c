/* rt_sssp.h -- C ABI of librt_sssp.so: the B200-native (sm_100a) drop-in for the shortest-path-method hot
 * path of albert-de-montserrat/RayTracer.jl.
 *
 * The reference has no FFI boundary of its own (it is pure Julia); the boundary replaced here is the set of
 * exported Julia functions of src/RayTracer.jl:24-34.  Each entry point below names the reference function
 * whose work it takes over.  julia/RayTracerB200.jl binds these with `ccall`; the Python package
 * raytracer.jl_b200 binds them with ctypes.
 *
 * Conventions
 *  - every node / element id crossing the ABI is a 1-based int64 exactly as in Julia; 0 in a `prev` table means
 *    "never set" (the reference leaves those entries undefined, src/SSSP/bfm.jl:12);
 *  - all pointers are HOST pointers unless the function name ends in `_dev`; outputs go into caller-allocated
 *    buffers whose capacity is passed explicitly; the library never frees caller memory;
 *  - every function returns an int status (RT_OK == 0); rt_last_error() returns a thread-local message;
 *    nothing throws across the boundary;
 *  - there is NO CPU fallback: without a CUDA device every compute entry point returns RT_ERR_CUDA.
 */
#ifndef RT_SSSP_H
#define RT_SSSP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_OK 0
#define RT_ERR_ARG 1      /* invalid argument (bad id, null pointer, capacity too small) */
#define RT_ERR_CUDA 2     /* CUDA runtime / no device */
#define RT_ERR_RANGE 3    /* interpolation point outside the knots (BoundsError in the reference) */
#define RT_ERR_NOPATH 4   /* recontruct_path would not terminate (receiver unreachable / cycle) */
#define RT_ERR_UNSUPPORTED 5

typedef struct rt_mesh rt_mesh; /* opaque: a graph resident in HBM (2-D annulus mesh or 3-D structured grid) */

/* Counters of one rt_bfm_solve call (summed over its sources).  Replaces the reference's only log line,
 * println("Converged in $it iterations") src/SSSP/bfm.jl:49 (it == sweeps + 1 there). */
typedef struct rt_stats {
  int64_t sweeps;          /* relaxation sweeps executed                                                    */
  int64_t relaxed_edges;   /* E_relaxed: candidate (edge) evaluations of the relax / push kernels           */
  int64_t vertex_updates;  /* active-vertex updates (vertices relaxed, summed over sweeps)                   */
  int64_t graph_edges;     /* E_graph: sum over vertices of |scan list| (reference multiplicities, per source)*/
  double kernel_ms;        /* device time of the solve loop(s), CUDA events on the solver stream            */
  double relax_ms;         /* device time of the relax kernel alone (0 unless profiling timers are enabled)  */
  int64_t relax_launches;  /* launches of the relax kernel                                                  */
  int64_t total_launches;  /* all kernel launches issued by the solve(s)                                    */
  double prev_ms;          /* device time of the predecessor (tightness) pass of the near-far schedule      */
  int64_t screened_edges;  /* E_evaluated: candidates that reached the algebraic screen (not discarded by the   */
                           /* d_from >= bound comparison); filled with profile_timers = 1 (2-D near-far, 3-D tile-pull), else 0 */
  int64_t exact_edges;     /* candidates that passed the screen and paid the exact sqrt/div evaluation (same)  */
} rt_stats;

/* ---- library ------------------------------------------------------------------------------------------ */
const char* rt_last_error(void);
const char* rt_version(void);
int rt_device_count(int* count);
int rt_set_device(int device); /* call before building meshes on this host thread */

/* ---- 2-D annulus: init_annulus(ntheta, nr; spacing) src/GridAnnulus.jl:57-70 ---------------------------- */
/* Runs primary_grid :72-142, secondary_nodes :607-698, constrain2layers! :296-321,
 * discontinuous_boundaries :910-968 and element_incidence :420-452; the mesh stays resident on the device. */
int rt_annulus_build(int64_t ntheta, int64_t nr, double spacing, rt_mesh** out);

/* sizes[8] = { n, nel, sum|e2n|, nnz(G), halo_rows (2H), sum|neighbours|, ntheta, nr(+7) }.
 * For a 3-D grid: { n, 0, 0, 0, 0, 0, 0, 0 }. */
int rt_mesh_sizes(const rt_mesh* m, int64_t sizes[8]);

/* Materialise Grid2D / SparseMatrixCSC{Bool,Int64} / halo::Matrix{Int64} for Julia (any pointer may be NULL).
 * e2n_off has nel+1 0-based offsets into e2n_idx; G_colptr is the 1-based Julia colptr (n+1); halo is the
 * (2H x 2) matrix in column-major order. */
int rt_mesh_export(const rt_mesh* m, double* x, double* z, double* theta, double* r, int64_t* e2n_off,
                   int64_t* e2n_idx, int64_t* G_colptr, int64_t* G_rowval, int64_t* halo, int64_t* nbr_off,
                   int64_t* nbr_idx, int8_t* el_type);

/* Adopt a graph that was built elsewhere (by the reference in Julia, or by a test):  the arrays that
 * bfm(G, halo, source, gr, U) reads -- G.colptr/G.rowval, gr.e2n (flattened), gr.x, gr.z, halo.
 * theta / r may be NULL (only needed by rt_mesh_export and rt_closest_point).
 * Precondition: the node neighbourhood N(i) = union of e2n[e] over the column e of G[:, i] is SYMMETRIC (j in N(i) <=>
 * i in N(j)), as every graph of init_annulus is.  The solver's frontier is element-granular (a superset of the
 * reference's per-node queue) and the near-far schedule pushes i -> j along the same lists that the reference pulls
 * j <- i; on an asymmetric graph sweeps, predecessors and even travel times could differ from the reference's queue
 * semantics.  Ids must be in range (checked on the 64-bit values before narrowing); the symmetry is not checked. */
int rt_mesh_from_arrays(int64_t n, int64_t nel, const int64_t* e2n_off, const int64_t* e2n_idx,
                        const int64_t* G_colptr, const int64_t* G_rowval, const int64_t* halo, int64_t halo_rows,
                        const double* x, const double* z, const double* theta, const double* r, rt_mesh** out);

/* ---- 3-D structured grid: grid(c0,c1,nnods) src/StructuredGrid.jl:35-45 + nodal_incidence :177-223 ------- */
/* coord_system 0: Cartesian axes; 1: axes are (theta, phi, r) mapped through spherical2cart :225-235.
 * star_levels = neighbour_levels: 0 = 26-neighbourhood; L >= 1 = the clipped (2*2^L + 1)^3 window incl. self -- every
 * expansion round of nodal_incidence (:204-212) unions the neighbours' current sets, so the radius doubles per level
 * (5^3 at L = 1, 9^3 at L = 2).  L = 3 (17^3) returns RT_ERR_UNSUPPORTED.
 * Edge weight: rt_set_option("weight3d", 0) (default) = distance3D(p1,p2) * (1/abs(U1+U2)) * 2 (src/SSSP/weights.jl:20);
 * ("weight3d", 1) = fw(a,b) / abs(U[a]+U[b]) * 0.5, the expression inside BFM/foo! (src/Dijsktra.jl:388) and dijsktra
 * (:44).  The solver's control flow is BFM/foo!/goo! (src/Dijsktra.jl:294-343, 376-403) in both modes. */
int rt_grid3d_build(const double c0[3], const double c1[3], const int64_t nn[3], int star_levels,
                    int coord_system, rt_mesh** out);
int rt_grid3d_export(const rt_mesh* m, double* X, double* Y, double* Z); /* Cartesian node coordinates */

/* gr.x, gr.y, gr.z: the nodal ranges collect(LinRange(c0[d], c1[d], nnods[d])) (:38-40); any pointer may be NULL. */
int rt_grid3d_axes(const rt_mesh* m, double* x, double* y, double* z);
/* getindex(gr, I) (:77-81) for `count` linear indices (1-based): xyz[count x 3] = Point(gr.x[i], gr.y[j], gr.z[k]),
 * ijk[count x 3] = CartesianIndex(gr, I) (:90-96).  Either output may be NULL.  (gr[i, j, k] is axes[i], [j], [k].) */
int rt_grid3d_points(const rt_mesh* m, const int64_t* I, int64_t count, double* xyz, int64_t* ijk);
/* connectivity(gr) / connectivity(gr, iel) (:121-168): the 8 corner ids of the hexes first_el .. first_el+count-1
 * (1-based), e2n[count x 8] row-major in the reference's corner order. */
int rt_grid3d_connectivity(const rt_mesh* m, int64_t first_el, int64_t count, int64_t* e2n);
/* closest_point(gr, x, y, z) (:257-270) for a batch of query points: first index of the minimum of
 * distance3D(gr[i], p) on the RAW axis coordinates (the reference does not apply spherical2cart here). */
int rt_closest_point3d(const rt_mesh* m, const double* px, const double* py, const double* pz, int64_t npts,
                       int64_t* index_out);
/* polardistance3D(a, b) (:245-255) for `count` pairs of (theta, phi, r) triples: a, b [count x 3] row-major. */
int rt_polardistance3d(const double* a, const double* b, int64_t count, double* out);

int rt_mesh_free(rt_mesh* m);

/* ---- velocity: interpolate_velocity(r, interpolant) src/utils.jl:38-44 ----------------------------------- */
/* Gridded linear interpolation on (knots_r, knots_v), knots ascending.  buffer < 0: utils.jl semantics;
 * buffer >= 0: src/ShortestPath.jl:74-90 (points exactly on a discontinuity radius are evaluated at r+buffer). */
int rt_interp_velocity(const double* knots_r, const double* knots_v, int64_t nk, const double* r, int64_t n,
                       double buffer, double* out);

/* interpolate!(V, gr) src/Interpolations/interpolation.jl:5-18: re-interpolate the velocity of every secondary
 * node from the corner values of its cell (bilinear.jl:1-17 for quads in (theta, r) space, barycentric.jl:1-15 for
 * the centre triangles); a node shared by two cells keeps the value of the cell with the larger id, as the
 * reference's sequential loop does.  V is a host array [n], updated in place.  el_type[nel] (0 = :Quad,
 * 1 = :Tri) may be NULL for a mesh built by rt_annulus_build. */
int rt_interpolate_cells(rt_mesh* m, const int8_t* el_type, double* V);

/* Same, r_dev / out_dev are DEVICE pointers (knots stay on the host: they are 6372 entries). */
int rt_interp_velocity_dev(const double* knots_r, const double* knots_v, int64_t nk, const double* r_dev,
                           int64_t n, double buffer, double* out_dev);

/* Device pointers to the node coordinates of a resident mesh: 2-D -> (x, z, theta, r) i.e. gr.x, gr.z, gr.theta,
 * gr.r (theta / r NULL if the mesh was adopted without them); 3-D -> (X, Y, Z, NULL). */
int rt_mesh_coords_dev(const rt_mesh* m, const double** a, const double** b, const double** c, const double** d);

/* ---- node-to-node topology (src/topology/topology.jl, src/SSSP/rcm.jl) ------------------------------------- */
/* nodal_incidence(gr::Grid2D) src/GridAnnulus.jl:763-804 as the CSR container SparseAdjencyList{list,deg,idx} of
 * sparse_adjacency_list (topology.jl:94-111): star-0 neighbours (nodes sharing a cell, self excluded, each once).
 * deg[n] = nodal_degree (topology.jl:70-77); list_off[n+1] 0-based offsets (the reference's idx = list_off + 1);
 * list_idx 1-based ids in (ascending cell, list order) -- the reference iterates a Julia Set, whose order is not
 * reproducible.  Two-call pattern: list_idx == NULL writes only deg / list_off.  Any pointer may be NULL. */
int rt_nodal_adjacency(rt_mesh* m, int64_t* deg, int64_t* list_off, int64_t* list_idx, int64_t cap);

/* symrcm(adjgr, degrees) src/SSSP/rcm.jl:2-46 on that adjacency: BFS from the minimum-degree node (next component:
 * next minimum-degree unplaced node), children ordered by (position of their earliest parent, node id) where the
 * reference uses Set iteration order, result reversed.  perm_out[n] 1-based: position k of the reordered mesh holds
 * old node perm_out[k], i.e. `gr.x .= gr.x[prm]` of reorder! (rcm.jl:62-85). */
int rt_rcm(rt_mesh* m, int64_t* perm_out);

/* ---- closest_point(gr, px, pz; system) src/GridAnnulus.jl:823-840 ---------------------------------------- */
/* system 0 = :cartesian (x,z), 1 = :polar ((theta, r) treated as Cartesian).  First index of the minimum. */
int rt_closest_point(const rt_mesh* m, const double* pa, const double* pb, int64_t npts, int system,
                     int64_t* index_out);

/* ---- solver: bfm(G, halo, source, gr, U) src/SSSP/bfm.jl:1-52 -------------------------------------------- */
/* Solves nsrc independent single-source problems on the same mesh and velocity.  dist_out / prev_out are
 * [nsrc x n] row-major (source-major) host buffers == BellmanFordMoore(prev, dist) per source
 * (src/SSSP/ssspm.jl:3-10).  Either output may be NULL.  stats may be NULL.
 * precision: 64 = the Float64 arithmetic of bfm (src/SSSP/bfm.jl); 32 = the Float32 arithmetic of bfm_gpu
 * (src/SSSP/bfm_gpu.jl:170-205, 487-526): x, z (X, Y, Z) and U are rounded to Float32 and every operation of the
 * relax is rounded to Float32; dist_out then holds Float32 values widened to double (narrowing them is exact).
 * Both precisions run in both schedules and give results bit-identical to the respective arithmetic. */
int rt_bfm_solve(rt_mesh* m, const double* U, const int64_t* sources, int64_t nsrc, int precision,
                 double* dist_out, int64_t* prev_out, rt_stats* stats);

/* Device-resident variant: U_dev [n] doubles and dist_dev [nsrc x n] doubles / prev_dev [nsrc x n] int32
 * (0-based, -1 = never set) are DEVICE pointers on the mesh's device.  Used by benchmarks and by sharded
 * multi-GPU drivers that gather tables with NCCL. */
int rt_bfm_solve_dev(rt_mesh* m, const double* U_dev, const int64_t* sources, int64_t nsrc, int precision,
                     double* dist_dev, int32_t* prev_dev, rt_stats* stats);

/* Single-process multi-GPU batch (many earthquakes shard by source; no exchange inside the relaxation).
 * meshes[0..ndev) are replicas of the SAME mesh, one per device (built by the caller after rt_set_device(d) with
 * rt_annulus_build / rt_mesh_from_arrays / rt_grid3d_build); replica d solves the contiguous block
 * sources[d*k .. (d+1)*k), k = ceil(nsrc / ndev), on its own host thread and stream and writes its rows of the
 * [nsrc x n] host tables directly -- the "gather" is the host buffer the caller (Julia) owns.  stats: sums over the
 * replicas, kernel_ms = max over replicas.  Options (schedule ...) are taken from each replica's own handle. */
int rt_bfm_solve_multi(rt_mesh* const* meshes, int ndev, const double* U, const int64_t* sources, int64_t nsrc,
                       int precision, double* dist_out, int64_t* prev_out, rt_stats* stats);

/* One process per GPU (torchrun, mpirun, Distributed.jl): the source batch is sharded over the ranks, every rank solves
 * its contiguous block sources[first .. first+count) (rt_comm_shard) on its own replica of the mesh, and the
 * travel-time / predecessor tables are gathered on EVERY rank with one ncclAllGather each over NVLink / NVSwitch
 * (no collective inside the relaxation).  NCCL is bound at run time (dlopen libnccl.so.2); without it these entry points
 * return RT_ERR_UNSUPPORTED.  rt_comm_unique_id: rank 0 creates the 128-byte ncclUniqueId, the caller broadcasts it with
 * whatever it already has (MPI, torch.distributed, a file), then every rank calls rt_comm_init after rt_set_device.
 * U_dev [n] doubles, dist_dev [nsrc x n] doubles, prev_dev [nsrc x n] int32 (0-based, -1 = never set; may be NULL) are
 * DEVICE pointers; the tables come back in the order of `sources` on every rank.  stats: the local solve's counters,
 * prev_ms = device time of the gather. */
typedef struct rt_comm rt_comm;
int rt_comm_unique_id(unsigned char id[128]);
int rt_comm_init(const unsigned char id[128], int rank, int world, rt_comm** out);
int rt_comm_shard(int64_t nsrc, int rank, int world, int64_t* first, int64_t* count);
int rt_bfm_solve_sharded(rt_comm* c, rt_mesh* m, const double* U_dev, const int64_t* sources, int64_t nsrc,
                         int precision, double* dist_dev, int32_t* prev_dev, rt_stats* stats);
/* The same with HOST buffers (a caller without device memory of its own): U [n] in, dist_out [nsrc x n] doubles and
 * prev_out [nsrc x n] 1-based int64 (0 = never set; may be NULL) out, as rt_bfm_solve returns them, on every rank. */
int rt_bfm_solve_sharded_host(rt_comm* c, rt_mesh* m, const double* U, const int64_t* sources, int64_t nsrc,
                              int precision, double* dist_out, int64_t* prev_out, rt_stats* stats);
int rt_comm_destroy(rt_comm* c);

/* Dual-velocity variant: bfm with U::Matrix -> _relax!(..., U::Matrix) src/SSSP/bfm.jl:113-159.  U2 is the
 * [n x 2] matrix of dual_velocity (column-major: U[:,1] "below" values, then U[:,2] "above" values); for an edge
 * between node i and candidate j the pair is U[i, tail] + U[j, head] with head = (r_i > r_j) + 1, tail = 3 - head.
 * Runs in the schedule selected by rt_set_option("schedule") like rt_bfm_solve (travel times bit-identical in
 * both; near-far needs packed_prev = 1, the default); outputs as rt_bfm_solve.  Needs a mesh that carries gr.r. */
int rt_bfm_solve_dual(rt_mesh* m, const double* U2, const int64_t* sources, int64_t nsrc, double* dist_out,
                      int64_t* prev_out, rt_stats* stats);

/* dual_velocity(r, interpolant; buffer) src/utils.jl:51-66 -> out[n x 2] column-major. */
int rt_dual_velocity(const double* knots_r, const double* knots_v, int64_t nk, const double* r, int64_t n,
                     double buffer, double* out);

/* ---- multiphase / layer-restricted propagation (src/topology/topology.jl:150-206, src/SSSP/bfm_multiphase.jl) ------ */
/* partition_grid(gr) topology.jl:183-206: id_out[n], k > 0 = "Layer_k" (find_layer_number :137-147, k = 1..8), -k =
 * "Boundary_k" (k = 1..7: the discontinuity radii R - {20, 35, 210, 410, 660, 2740, 2891.5}); radii are rounded to two
 * digits first, as the reference does. */
int rt_partition_grid(const rt_mesh* m, int32_t* id_out);
/* The inner loop of bfm_multiphase (bfm_multiphase.jl:118-150) on the graph of bfm: reference sweeps (Jacobi schedule)
 * that CONTINUE from the caller's state dist_inout / prev_inout [n] and only relax, activate or halo-update nodes with
 * allowed[i] != 0 (`ID[Gi] in current_level`; allowed == NULL: every node); the frontier starts as the allowed nodes of
 * the star patches of seeds[0 .. nseeds).  Everything else of a phase loop (which layers, which boundary velocity,
 * where to restart) stays with the caller: see bfm_multiphase in julia/RayTracerB200.jl and raytracer.jl_b200/api.py.
 * The reference routine is unfinished (it calls the undefined _relax_bfm! / fillfalse!); the relax step used here is
 * _relax!(U::Vector) of bfm.jl:161-210. */
int rt_bfm_continue(rt_mesh* m, const double* U, const uint8_t* allowed, const int64_t* seeds, int64_t nseeds,
                    double* dist_inout, int64_t* prev_inout, rt_stats* stats);

/* ---- alternative solvers behind the Dijkstra / RadiusStepping result structs (src/SSSP/ssspm.jl:3-10) ------------- */
/* algorithm 0: dijkstra(G::Dict, source, gr, U) src/SSSP/dijkstra.jl:68-136; algorithm 1: radius_stepping(Gsp, source,
 * gr, U) src/SSSP/radius_stepping.jl:7-46.  Both work on the star-0 node graph nodal_incidence(gr) (src/GridAnnulus.jl:
 * 763-804; rt_nodal_adjacency exports it) WITHOUT halo coupling, with the weight 2*distance/abs(U[i]+U[j]).  The travel
 * times are the least fixed point of that graph (bit-identical to either reference loop); predecessors follow the settle
 * order: the tight neighbour with the smallest (travel time, id), coincident duplicates at equal time resolved in the
 * order they were fixed.  dist_out / prev_out: [n] host arrays (1-based ids, 0 = never set, Inf = unreachable). */
int rt_sssp_nodal(rt_mesh* m, const double* U, int64_t source, int algorithm, double* dist_out, int64_t* prev_out,
                  rt_stats* stats);

/* Solver options of one mesh handle, key / value.  Results-relevant:
 *   "schedule"        0 = Jacobi sweeps exactly as the reference (default), 1 = work-efficient near-far ordering (travel
 *                     times bit-identical; predecessors tight, identical to the reference except on exact ties)
 *   "canonical_prev"  1 = the near-far schedule returns the reference's predecessors, exact ties included (about one
 *                     extra sweep over the graph)
 *   "weight3d"        3-D edge weight expression (see rt_grid3d_build)
 * Execution strategy only (every combination gives the same bits; defaults are the measured best):
 *   "delta" [s] / "delta_factor"  bucket width of the near-far schedule (0 = automatic)
 *   "batch"           2-D: sources advanced in the same launches (0 = automatic, <= 1024); 3-D tile-pull rounds: sources
 *                     of a call in flight at once, each in its own slot and stream (0 = automatic = 4, 1 = one after the
 *                     other, k <= 8)
 *   "persistent", "warp_units", "cta_units", "packed_prev", "compact", "group_screen", "target_lists", "use_graph",
 *   "fuse_begin"      2-D near-far kernel variants / launch sequence
 *   "tile_pull"       3-D near-far: 1 = tile-pull rounds (default), 0 = push units;  "early_advance" (tile-pull rounds)
 *   "check_every"     rounds between host convergence checks;  "profile_timers" 1 = fill rt_stats.relax_ms,
 *                     screened_edges, exact_edges (adds a host synchronisation per round).
 * An unknown key returns RT_ERR_ARG. */
int rt_set_option(rt_mesh* m, const char* key, double value);

/* ---- paths: recontruct_path(prev, source, receiver) src/SSSP/ssspm.jl:30-40 ------------------------------- */
/* Two-call pattern: with path_idx == NULL only path_off[nrec+1] (0-based offsets) is written.  Each path is
 * [receiver, ..., source].  Returns RT_ERR_NOPATH if a chase does not reach `source` within n steps. */
int rt_reconstruct_paths(const int64_t* prev, int64_t n, int64_t source, const int64_t* receivers, int64_t nrec,
                         int64_t* path_off, int64_t* path_idx, int64_t cap);

/* recontruct_path(D, source, receiver) src/SSSP/ssspm.jl:14-28, the method on the result structs: the chase runs until
 * a node repeats (`while ipath ∉ path`), then `source` is appended; reading an unset predecessor (the reference:
 * BoundsError) returns RT_ERR_NOPATH.  Same two-call pattern. */
int rt_reconstruct_paths_guarded(const int64_t* prev, int64_t n, int64_t source, const int64_t* receivers,
                                 int64_t nrec, int64_t* path_off, int64_t* path_idx, int64_t cap);

/* travel_times(D, gr, receivers) src/utils.jl:4-8 for a batch of sources: out[nsrc x nrec] = dist[s, receivers[k]]
 * gathered on the device from [nsrc x n] tables (host table / table resident on the device). */
int rt_travel_times(const double* dist, int64_t n, int64_t nsrc, const int64_t* receivers, int64_t nrec, double* out);
int rt_travel_times_dev(const double* dist_dev, int64_t n, int64_t nsrc, const int64_t* receivers, int64_t nrec,
                        double* out);

/* Same with the predecessor table resident on the device (int32, 0-based, -1 = never set) as written by
 * rt_bfm_solve_dev; receivers / outputs are host arrays with 1-based ids. */
int rt_reconstruct_paths_dev(const int32_t* prev_dev, int64_t n, int64_t source, const int64_t* receivers,
                             int64_t nrec, int64_t* path_off, int64_t* path_idx, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* RT_SSSP_H */
