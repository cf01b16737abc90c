// mesh2d.cuh -- device-resident two-level annulus graph (node -> elements -> nodes) and solver workspace.
//
// Data layout in HBM (all ids 0-based int32 internally; 1-based int64 only at the C ABI):
//   x, z, theta, r        double[n]        node coordinates (gr.x, gr.z, gr.theta, gr.r)
//   e2n_off / e2n_idx     int32[nel+1] / int32[sum|e2n|]   element -> nodes in the reference's list order
//   g_off / g_idx         int64[n+1]  / int32[nnz(G)]      node -> elements, ascending (SparseMatrixCSC column)
//   n2e_off / n2e_idx     int32[n+1]  / int32[sum|e2n|]    node -> elements that CONTAIN it (transpose of e2n)
//   item_first            int32[n_items+1]  work items: runs of <= 32 consecutive node ids (32-aligned) whose
//                         G columns are identical -> one warp relaxes one item and shares the candidate scan
//   halo                  orig[H], twin[H] (+ CSR of twin rows grouped by orig for the second half of the rows)
#pragma once
#include "common.cuh"

struct Mesh2D {
  i64 n = 0, nel = 0, sum_e2n = 0, nnzG = 0, halo_rows = 0, sum_nbr = 0, ntheta = 0, nr = 0;
  DevBuf<double> x, z, theta, r;
  DevBuf<double> xf, zf;  // Float32-rounded coordinates (precision = 32), built on first use
  bool has_polar = false;
  DevBuf<i32> e2n_off, e2n_idx;
  DevBuf<i64> g_off;
  DevBuf<i32> g_idx;
  DevBuf<i32> n2e_off, n2e_idx;
  i64 n_items = 0;
  DevBuf<i32> item_first;
  // halo
  i64 H = 0;
  bool halo_structured = true;
  DevBuf<i32> halo_h1, halo_h2;  // all 2H rows (0-based), used by the generic serial path and by commit
  DevBuf<i32> h2_orig, h2_off, h2_twin;  // second-half rows grouped by their target (orig) in row order
  i64 n_h2_orig = 0;
  DevBuf<i32> hinit_node, hinit_val;  // init_halo_path! result (last writer wins, serial order)
  i64 n_hinit = 0;
  i64 graph_edges = 0;
  // extras kept on the host for rt_mesh_export of built meshes
  std::vector<i64> nbr_off_h, nbr_idx_h;
  std::vector<int8_t> el_type_h;
  // solver workspace (allocated on first solve)
  DevBuf<double> dist, dist0;
  DevBuf<i32> prev;
  DevBuf<uint8_t> dirty;
  DevBuf<i32> act[2];
  DevBuf<u64> counters;  // [0],[1]: active counts (ping-pong); [2]: evals; [3]: vertex updates; [4]: improved flag
  u64* counters_host = nullptr;  // pinned
  bool ws_ready = false;
  // ---- near-far (push) schedule
  DevBuf<i32> node_item;               // node -> work item
  DevBuf<unsigned> pend_mask, far_mask;  // per item: nodes with an unpropagated improvement (near / far)
  DevBuf<unsigned> infar_u;            // per item: already in the far list
  DevBuf<unsigned> cur_mask;           // per near-list slot: the nodes released this round
  DevBuf<i32> nearq[2], farq[2];
  DevBuf<i32> hn_index;                // node -> row in hn_off, -1 if the node has no halo partner
  DevBuf<i32> hn_node, hn_off, hn_part;  // halo partners: sorted unique nodes -> CSR of partner nodes
  i64 n_hn = 0;
  DevBuf<double> tau;                  // [0] current threshold, [1] delta, [2] min far dist (as u64 bits)
  DevBuf<i32> unresolved[2];
  DevBuf<int> pending_prev;
  DevBuf<int> ctl;
  DevBuf<double> dp;                   // packed (travel time bits, predecessor key) pairs, 16 B per node
  bool push_ready = false;
  int push_nb = 0;                     // batch width the per-source state is allocated for
  DevBuf<u64> bcounters;               // [nb x 8]
  DevBuf<int> bsources;
  DevBuf<i32> bprev;                   // [nb x n]
  DevBuf<double> bdist;                // [nb x n]
  DevBuf<i64> flat;                    // flattened near / far slot prefixes of a batch round
  DevBuf<i32> flat_b;                  // owner source per global near-list slot of a batch round
  CanonWs* canon = nullptr;
  void* round_graph = nullptr;         // cudaGraphExec_t of the near-far round sequence (bfm2d_push.cu)
  std::vector<char> round_graph_key;
  i64 round_graph_launches = 0;
  DevBuf<i64> tgt_off;                 // de-duplicated target lists of the work items (short-column meshes)
  DevBuf<i32> tgt_idx;
  bool tgt_tried = false;
  bool tgt_missing = true;             // some items have no list (column longer than the builder's cap)
};

// Finishes a Mesh2D whose primary arrays (x,z,e2n_*,g_*) are already on the device: builds n2e, work items,
// halo tables, E_graph.  halo_host: (2H x 2) column-major 1-based (may be null if halo_rows == 0).
int mesh2d_finalize(rt_mesh* h, const i64* halo_host);
int mesh2d_prepare_f32(rt_mesh* h);
int bfm2d_solve_push_dual(rt_mesh* h, const double* U2_dev, const i64* sources, i64 nsrc, double* dist_dev,
                          i32* prev_dev, rt_stats* stats);
// reference predecessors (ties included) from converged travel times; see canonical_prev.cu
int canonical_prev_2d(rt_mesh* h, const double* x, const double* z, const double* U1, const double* U2, int mode,
                      const double* dist, int source, i32* prev, i64* levels_out, i64* launches_out);
int bfm2d_solve_push(rt_mesh* h, const double* U_dev, const i64* sources, i64 nsrc, double* dist_dev, i32* prev_dev,
                     rt_stats* stats);
