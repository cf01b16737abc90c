// comm.cu -- multi-GPU batches inside the library: one process per GPU, sources sharded over the
// ranks, the travel-time / predecessor tables gathered with ONE ncclAllGather each over NVLink / NVSwitch.  There is no
// collective inside the relaxation: a single-source solve does not shard.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a Julia process picks up the system library, a torchrun process
// the copy PyTorch already loaded; the shared object has no link-time NCCL dependency, so single-GPU users need none.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

struct NcclId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclId, int);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*GetErrorStringFn)(int);

struct NcclApi {
  void* lib = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllGatherFn all_gather = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn get_error_string = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return RT_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    rt_set_error("NCCL is not available (dlopen libnccl.so.2: %s)", dlerror());
    return RT_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.lib = lib;
  a.get_unique_id = (GetUniqueIdFn)dlsym(lib, "ncclGetUniqueId");
  a.comm_init_rank = (CommInitRankFn)dlsym(lib, "ncclCommInitRank");
  a.all_gather = (AllGatherFn)dlsym(lib, "ncclAllGather");
  a.comm_destroy = (CommDestroyFn)dlsym(lib, "ncclCommDestroy");
  a.get_error_string = (GetErrorStringFn)dlsym(lib, "ncclGetErrorString");
  if (!a.get_unique_id || !a.comm_init_rank || !a.all_gather || !a.comm_destroy || !a.get_error_string) {
    rt_set_error("libnccl lacks an expected symbol");
    return RT_ERR_UNSUPPORTED;
  }
  g_nccl = a;
  return RT_OK;
}

#define RT_NCCL(call)                                                                                  \
  do {                                                                                                 \
    const int _r = (call);                                                                             \
    if (_r != 0) {                                                                                     \
      rt_set_error("NCCL error %s at %s:%d", g_nccl.get_error_string(_r), __FILE__, __LINE__);         \
      return RT_ERR_CUDA;                                                                              \
    }                                                                                                  \
  } while (0)

}  // namespace

struct rt_comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1, device = 0;
  DevBuf<double> gd;  // padded gather buffers (only when nsrc is not a multiple of the world size)
  DevBuf<i32> gp;
  DevBuf<double> hU, hd;  // staging of the host-buffer front (rt_bfm_solve_sharded_host)
  DevBuf<i32> hp;
};

extern "C" {

int rt_comm_unique_id(unsigned char id[128]) {
  RT_ARG(id, "null id");
  RT_TRY(load_nccl());
  NcclId nid;
  RT_NCCL(g_nccl.get_unique_id(&nid));
  std::memcpy(id, nid.internal, 128);
  return RT_OK;
}

int rt_comm_init(const unsigned char id[128], int rank, int world, rt_comm** out) {
  RT_ARG(id && out && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
  *out = nullptr;
  RT_TRY(load_nccl());
  NcclId nid;
  std::memcpy(nid.internal, id, 128);
  rt_comm* c = new rt_comm();
  c->rank = rank;
  c->world = world;
  if (cudaGetDevice(&c->device) != cudaSuccess) {
    delete c;
    rt_set_error("no CUDA device");
    return RT_ERR_CUDA;
  }
  const int r = g_nccl.comm_init_rank(&c->comm, world, nid, rank);
  if (r != 0) {
    rt_set_error("ncclCommInitRank failed: %s", g_nccl.get_error_string(r));
    delete c;
    return RT_ERR_CUDA;
  }
  *out = c;
  return RT_OK;
}

int rt_comm_destroy(rt_comm* c) {
  if (!c) return RT_OK;
  if (c->comm && g_nccl.comm_destroy) g_nccl.comm_destroy(c->comm);
  delete c;
  return RT_OK;
}

int rt_comm_shard(int64_t nsrc, int rank, int world, int64_t* first, int64_t* count) {
  RT_ARG(nsrc >= 0 && world >= 1 && rank >= 0 && rank < world && first && count, "bad shard arguments");
  const i64 k = (nsrc + world - 1) / world;
  *first = std::min<i64>(nsrc, (i64)rank * k);
  *count = std::min<i64>(nsrc, (i64)(rank + 1) * k) - *first;
  return RT_OK;
}

int rt_bfm_solve_sharded(rt_comm* c, rt_mesh* m, const double* U_dev, const int64_t* sources, int64_t nsrc,
                         int precision, double* dist_dev, int32_t* prev_dev, rt_stats* stats) {
  RT_ARG(c && m && U_dev && sources && nsrc >= 0 && dist_dev, "null argument");
  RT_ARG(m->device == c->device, "the mesh must live on the communicator's device");
  RT_CUDA(cudaSetDevice(m->device));
  i64 sizes[8];
  RT_TRY(rt_mesh_sizes(m, sizes));
  const i64 n = sizes[0];
  const i64 k = (nsrc + c->world - 1) / c->world;  // rows per rank (the last ranks may own fewer)
  i64 first = 0, count = 0;
  RT_TRY(rt_comm_shard(nsrc, c->rank, c->world, &first, &count));
  const bool even = k * c->world == nsrc;
  double* gd = dist_dev;
  i32* gp = prev_dev;
  if (!even) {  // all ranks must contribute k rows: gather into padded buffers, copy the first nsrc rows out
    if (c->gd.n < (size_t)(k * c->world * n)) RT_TRY(c->gd.alloc((size_t)(k * c->world * n)));
    gd = c->gd.p;
    if (prev_dev) {
      if (c->gp.n < (size_t)(k * c->world * n)) RT_TRY(c->gp.alloc((size_t)(k * c->world * n)));
      gp = c->gp.p;
    }
  }
  // every rank solves its block straight into its slot of the gather buffer (in-place all-gather)
  double* my_d = gd + (i64)c->rank * k * n;
  i32* my_p = gp ? gp + (i64)c->rank * k * n : nullptr;
  rt_stats st = {};
  if (count > 0) RT_TRY(rt_bfm_solve_dev(m, U_dev, sources + first, count, precision, my_d, my_p, &st));
  cudaStream_t s = m->stream;  // the solver stream has drained (rt_bfm_solve_dev synchronises): gather on it
  cudaEvent_t e0, e1;
  RT_CUDA(cudaEventCreate(&e0));
  RT_CUDA(cudaEventCreate(&e1));
  cudaEventRecord(e0, s);
  int r = g_nccl.all_gather(my_d, gd, (size_t)(k * n) * sizeof(double), /*ncclChar*/ 0, c->comm, s);
  if (r == 0 && gp) r = g_nccl.all_gather(my_p, gp, (size_t)(k * n) * sizeof(i32), 0, c->comm, s);
  cudaEventRecord(e1, s);
  if (r != 0) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    rt_set_error("ncclAllGather failed: %s", g_nccl.get_error_string(r));
    return RT_ERR_CUDA;
  }
  if (!even) {
    cudaMemcpyAsync(dist_dev, gd, (size_t)nsrc * n * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (prev_dev) cudaMemcpyAsync(prev_dev, gp, (size_t)nsrc * n * sizeof(i32), cudaMemcpyDeviceToDevice, s);
  }
  const cudaError_t ce = cudaStreamSynchronize(s);
  float gather_ms = 0.f;
  cudaEventElapsedTime(&gather_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ce != cudaSuccess) {
    rt_set_error("CUDA failure in rt_bfm_solve_sharded: %s", cudaGetErrorString(ce));
    return RT_ERR_CUDA;
  }
  if (stats) {
    *stats = st;
    stats->prev_ms = gather_ms;  // this entry point reports the gather time here (the local solve's prev pass is in kernel_ms)
  }
  return RT_OK;
}

// Host-buffer front of rt_bfm_solve_sharded for callers that hold no device memory (a plain Julia process): U in,
// gathered tables out ([nsrc x n], 1-based int64 predecessors, 0 = never set, as rt_bfm_solve returns them).
int rt_bfm_solve_sharded_host(rt_comm* c, rt_mesh* m, const double* U, const int64_t* sources, int64_t nsrc,
                              int precision, double* dist_out, int64_t* prev_out, rt_stats* stats) {
  RT_ARG(c && m && U && sources && nsrc >= 0 && dist_out, "null argument");
  RT_CUDA(cudaSetDevice(m->device));
  i64 sizes[8];
  RT_TRY(rt_mesh_sizes(m, sizes));
  const i64 n = sizes[0];
  if (c->hU.n != (size_t)n) RT_TRY(c->hU.alloc(n));
  if (c->hd.n < (size_t)(nsrc * n)) RT_TRY(c->hd.alloc(nsrc * n));
  if (prev_out && c->hp.n < (size_t)(nsrc * n)) RT_TRY(c->hp.alloc(nsrc * n));
  RT_CUDA(cudaMemcpy(c->hU.p, U, n * sizeof(double), cudaMemcpyHostToDevice));
  RT_TRY(rt_bfm_solve_sharded(c, m, c->hU.p, sources, nsrc, precision, c->hd.p, prev_out ? c->hp.p : nullptr, stats));
  RT_CUDA(cudaMemcpy(dist_out, c->hd.p, (size_t)nsrc * n * sizeof(double), cudaMemcpyDeviceToHost));
  if (prev_out) RT_TRY(prev_to_host_i64(c->hp.p, nsrc * n, prev_out, m->stream));
  return RT_OK;
}

}  // extern "C"
