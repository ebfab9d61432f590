/* dawn_b200_rt.h — thin C-ABI runtime under the generated B200 stencils (libdawn_b200rt.so).
 *
 * Replaces, for the generated-code path, what the reference's generated CUDA code gets from
 * GridTools device storages and its driver-includes:
 *   storage allocation / host<->device movement   dawn/src/driver-includes/storage_runtime.hpp:69-97
 *                                                 dawn/src/driver-includes/cuda_utils.hpp:127-210
 *   event timers                                  dawn/src/driver-includes/timer_cuda.hpp:39-84
 *   field verification on the GPU                 dawn/src/driver-includes/cuda_verify.hpp:84-196
 *   layout transposition element-major<->level-major  cuda_utils.hpp:81-125 (reshape/reshape_back)
 *   neighbour-table construction                  cuda_utils.hpp:212-263 (generateNbhTable)
 * and adds what the reference lacks entirely: j-slab halo exchange between the GPUs of a box.
 *
 * Conventions: every function returns 0 on success or a B200_ERR_* code (the reference aborts the
 * process in gpuErrchk, cuda_utils.hpp:31-39); b200rt_last_error() gives the message of the last
 * failure on the calling thread.  Streams and events are opaque `void*` (cudaStream_t/cudaEvent_t).
 * Plain pointers and sizes only.
 */
#ifndef DAWN_B200_RT_H
#define DAWN_B200_RT_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  B200_OK = 0,
  B200_ERR_CUDA = 1,    /* a CUDA runtime call failed */
  B200_ERR_ARG = 2,     /* bad argument */
  B200_ERR_HALO = 3,    /* domain halo smaller than the stencil's access extent */
  B200_ERR_LAYOUT = 4,  /* field strides incompatible */
  B200_ERR_RUNTIME = 5, /* other run-time failure */
  B200_ERR_NCCL = 6
};

/* Device-resident field handed to a generated stencil: `data` addresses element (0,0,0) of the
 * storage (halo included); i-stride is 1 for fields with an i dimension.  For masked dimensions the
 * stride is ignored.  Unstructured dense fields: element-stride 1, `stride_k` = level stride;
 * sparse fields: [k][nbh][element], `stride_j` = neighbour stride (= level-local element count). */
typedef struct b200_field_t {
  void* data;
  int stride_j;
  int stride_k;
} b200_field_t;

const char* b200rt_last_error(void);
int b200rt_set_error(int code, const char* msg); /* returns code */

int b200rt_device_count(int* n);
int b200rt_set_device(int device);
int b200rt_device_synchronize(void);

int b200rt_malloc(void** dptr, size_t bytes);
int b200rt_free(void* dptr);
int b200rt_malloc_host(void** hptr, size_t bytes); /* pinned */
int b200rt_free_host(void* hptr);
int b200rt_memset(void* dptr, int value, size_t bytes, void* stream);
int b200rt_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int b200rt_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int b200rt_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
/* strided 2-D copies (rows of `width_bytes`, `height` rows), either direction by pointer kind */
int b200rt_memcpy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes,
                    size_t height, void* stream);

int b200rt_stream_create(void** stream);
int b200rt_stream_destroy(void* stream);
int b200rt_stream_synchronize(void* stream);

int b200rt_event_create(void** event);
int b200rt_event_destroy(void* event);
int b200rt_event_record(void* event, void* stream);
int b200rt_event_synchronize(void* event);
int b200rt_stream_wait_event(void* stream, void* event); /* later work of `stream` waits for `event` */
int b200rt_event_elapsed_ms(void* start, void* stop, float* ms);
int b200rt_check_last_launch(void); /* cudaGetLastError -> B200_ERR_CUDA */

/* Encodes a CUtensorMap (128 bytes, 64-byte aligned, written to `map`) for TMA loads of
 * (box_i x box_j x box_k) boxes of doubles out of a 3-D field: `base` must be 16-byte aligned and
 * both strides (in elements) even.  Out-of-range box elements are zero-filled and never read. */
int b200rt_tensor_map_3d(void* map, const void* base, int size_i, int size_j, int size_k,
                         int stride_j, int stride_k, int box_i, int box_j, int box_k);
/* Same for elements of `elem_bytes` (8: double, 4: float); strides are in elements and must be multiples
 * of 16 bytes, like box_i. */
int b200rt_tensor_map_3d_es(void* map, const void* base, int size_i, int size_j, int size_k,
                            int stride_j, int stride_k, int box_i, int box_j, int box_k,
                            int elem_bytes);

/* VerificationMetrics twin (driver-includes/verification_metrics.hpp:3-10), computed in ONE pass
 * over the two device arrays (the reference runs 2 kernels + 8 thrust reductions + 2 cudaMallocs per
 * field, cuda_verify.hpp:84-196). */
typedef struct b200_verification_metrics {
  int iteration;
  int is_valid;
  double max_rel_err, min_rel_err, max_abs_err, min_abs_err;
  long long n_failed;
} b200_verification_metrics;

/* isclose semantics of cuda_verify.hpp:58-80: |a-b| <= abs_tol + rel_tol*|b|, false whenever a NaN is
 * involved; errors as cuda_verify.hpp:25-45 (0 for equal values, else |b-a| and |(b-a)/b|: inf for a zero
 * reference value; initial values of the reductions as thrust's: max 0, min inf).
 * `dsl` is the tested field, `ref` the reference one; both device pointers of n elements. */
int b200rt_verify_field(const double* dsl, const double* ref, size_t n, double rel_tol,
                        double abs_tol, int iteration, b200_verification_metrics* out, void* stream);
int b200rt_verify_field_f32(const float* dsl, const float* ref, size_t n, double rel_tol,
                            double abs_tol, int iteration, b200_verification_metrics* out, void* stream);

/* Verification record: twin of the reference's MetricsSerialiser (driver-includes/to_json.{hpp,cpp}),
 * which merges per-field metrics into a nlohmann::json object passed to setup_<N>.  Here the record is
 * an opaque handle; its JSON text has the same shape:
 *   { "<stencil>": [ { "<field>": { "field_is_valid": b, "max_absolute_error": x, "max_relative_error": x,
 *                                   "min_absolute_error": x, "min_relative_error": x }, ... }, ... ] }
 * one array entry per iteration (metrics->iteration); a field already present in an iteration entry is
 * kept (nlohmann's object insert does not overwrite), exactly like to_json.cpp:9-33. */
typedef struct b200_metrics_record b200_metrics_record;
int b200rt_metrics_record_create(b200_metrics_record** rec);
int b200rt_metrics_record_add(b200_metrics_record* rec, const char* stencil, const char* field,
                              const b200_verification_metrics* metrics);
/* JSON text (keys sorted, like nlohmann's default object type); owned by the record, valid until the
 * next call on it. */
const char* b200rt_metrics_record_json(b200_metrics_record* rec);
int b200rt_metrics_record_write(b200_metrics_record* rec, const char* path);
int b200rt_metrics_record_destroy(b200_metrics_record* rec);

/* Layout transposition element-major [elem][k] <-> level-major [k][elem] on the device
 * (cuda_utils.hpp:81-125 does this on one host thread).  sparse: [elem][sparse][k] <-> [k][sparse][elem]. */
int b200rt_reshape(const double* in, double* out, int k_size, int num_elements, int sparse_size,
                   void* stream);
int b200rt_reshape_back(const double* in, double* out, int k_size, int num_elements, int sparse_size,
                        void* stream);
int b200rt_reshape_f32(const float* in, float* out, int k_size, int num_elements, int sparse_size,
                       void* stream);
int b200rt_reshape_back_f32(const float* in, float* out, int k_size, int num_elements, int sparse_size,
                            void* stream);

/* Neighbour table: row-major host table [elem][nbh] -> device SoA table [nbh][elem] (int32, -1
 * padded), bit-exact with generateNbhTable (cuda_utils.hpp:212-263). */
int b200rt_upload_nbh_table(const int* row_major, int num_elements, int nbh_size, int* d_table_soa,
                            void* stream);

/* ---- unstructured meshes: what the reference passes as dawn::GlobalGpuTriMesh
 * (driver-includes/cuda_utils.hpp:43-57: element counts + a map (chain, include_center) -> device
 * neighbour table) as a plain struct.  Location codes: 0 cells, 1 edges, 2 vertices.  Tables are
 * SoA [nbh][element] int32 on the device, -1 = missing neighbour (b200rt_upload_nbh_table). ---- */
typedef struct b200_nbh_table_t {
  int chain[8];
  int chain_len;
  int include_center;
  int nbh_size;
  const int* table;
} b200_nbh_table_t;

/* One entry of the reference's dawn::unstructured_domain map (driver-includes/
 * unstructured_domain.hpp:23-36): first element index of a subdomain.  location: 0 cells, 1 edges,
 * 2 vertices; subdomain: 0 LateralBoundary, 1000 Nudging, 2000 Interior, 3000 Halo, 4000 End. */
typedef struct b200_splitter_t {
  int location, subdomain, offset, index;
} b200_splitter_t;

typedef struct b200_mesh_t {
  int num_cells, num_edges, num_vertices;
  int num_tables;
  const b200_nbh_table_t* tables;
  int num_splitters;                 /* may be 0: stencils without iteration spaces need none */
  const b200_splitter_t* splitters;  /* copied by setup_<N>; set_splitter_index_<N> edits the copy */
} b200_mesh_t;

/* ---- a time step as one CUDA graph (the reference synchronises the device around
 * every generated run(), CudaCodeGen.cpp:435-450).  Everything the C ABI enqueues on `stream` between
 * begin and end - generated kernels, halo pack/unpack kernels, the ncclSend/ncclRecv group - becomes one
 * graph; launch replays it.  Run the sequence once eagerly first: the first call of a generated module
 * sets kernel attributes and allocates scratch temporaries, which must not happen while capturing.
 * A graph holding NCCL nodes has to be destroyed before its communicator. ---- */
typedef struct b200_graph b200_graph;
int b200rt_graph_begin(void* stream);
int b200rt_graph_end(void* stream, b200_graph** graph);
int b200rt_graph_launch(b200_graph* graph, void* stream);
int b200rt_graph_destroy(b200_graph* graph);

/* ---- j-slab halo exchange across the GPUs of one box (new: the reference has no communication
 * layer at all).  One process per GPU.  `comm` is an ncclComm_t created by the host
 * (e.g. from torch.distributed's unique id) or by b200rt_comm_init_rank. ---- */
int b200rt_nccl_unique_id(void* id128 /* 128 bytes out */);
int b200rt_comm_init_rank(void** comm, int nranks, const void* id128, int rank);
int b200rt_comm_destroy(void* comm);

/* Exchanges `width` j-rows of every k level of an ijk field with the slab neighbours rank-1 / rank+1
 * (non-periodic).  The field is the local slab INCLUDING halos: rows [0,halo) and [nj-halo,nj) are
 * halo rows.  Interior rows adjacent to each boundary are packed by one kernel, sent with
 * ncclSend/ncclRecv inside one group, and unpacked into the halo rows.  Scratch is owned by the plan. */
typedef struct b200_halo_plan b200_halo_plan;
int b200rt_halo_plan_create(b200_halo_plan** plan, void* comm, int rank, int nranks, int ni, int nj,
                            int nk, int stride_j, int stride_k, int halo_j, int width_minus,
                            int width_plus);
/* Same with different halo depths at the two ends: rows [0,halo_lo) and [nj-halo_hi,nj) are halo rows
 * (a slab at the edge of the global domain has no halo on that side; row slabs of an unstructured
 * triangle mesh exchange whole rows of edge/cell/vertex slots this way). */
int b200rt_halo_plan_create2(b200_halo_plan** plan, void* comm, int rank, int nranks, int ni, int nj,
                             int nk, int stride_j, int stride_k, int halo_lo, int halo_hi,
                             int width_minus, int width_plus);
int b200rt_halo_exchange(b200_halo_plan* plan, double* field, void* stream);
int b200rt_halo_plan_destroy(b200_halo_plan* plan);

/* Peer-memory variant of the same exchange (NVLink / NVSwitch, no NCCL on the data path): the landing
 * buffers of a plan are mapped by its slab neighbours through CUDA IPC.  One kernel packs my boundary
 * rows and stores them straight into the neighbours' landing buffers, then raises an `arrived` flag
 * there; a second kernel waits for the neighbours' flag, moves their rows into my halo rows and
 * returns a `credit` so that the next push may overwrite the buffer.  Flags are epoch counters kept
 * in device memory, so the two launches can be replayed from a CUDA graph.  Set-up, once:
 *   every rank:  b200rt_halo_plan_ipc_handle(plan, h)        -> 64 opaque bytes
 *   host:        hand every rank the handles of rank-1 and rank+1 (e.g. an all-gather)
 *   every rank:  b200rt_halo_plan_connect(plan, h_lo, h_hi)   (NULL at the ends of the slab chain)
 * All ranks must call b200rt_halo_exchange_p2p the same number of times, and must synchronise among
 * themselves (a host barrier) before any of them destroys its plan.  A wait that sees no neighbour
 * for ~10 s gives up instead of hanging the device; b200rt_halo_plan_p2p_status reports that. */
int b200rt_halo_plan_ipc_handle(b200_halo_plan* plan, void* handle64 /* 64 bytes out */);
int b200rt_halo_plan_connect(b200_halo_plan* plan, const void* handle_lo, const void* handle_hi);
int b200rt_halo_exchange_p2p(b200_halo_plan* plan, double* field, void* stream);
int b200rt_halo_plan_p2p_status(b200_halo_plan* plan);

/* ---- chained unstructured launches (codegen option ico_chain) -------------------------------------------
 * One launch holds up to two producer roles and a consumer role that gathers what they wrote; blocks run
 * in a host-built order (blockIdx.x = level block * sched_len + position) in which every consumer block
 * follows the producer blocks its elements reach through `tables` (device SoA tables
 * [nbh][n_consumer], consumer element -> producer element of role table_role[t]) by `lag_blocks` more
 * producer blocks.  Producers publish flags[level block][role offset + block] = epoch, consumers wait for
 * the flag range need[4 c .. 4 c + 3] = {lo0, hi0, lo1, hi1} of their block c.  A wait that gives up
 * (bounded spin) raises *error; b200rt_chain_plan_status synchronises the stream and reports it. */
typedef struct b200_chain_plan b200_chain_plan;
typedef struct b200_chain_view {
  const unsigned* sched; /* device */
  const int* need;       /* device */
  unsigned* flags;       /* device, level_blocks x (nb0 + nb1) */
  unsigned* error;       /* device */
  unsigned sched_len, nb0, nb1, nc, level_blocks;
} b200_chain_view;
int b200rt_chain_plan_create(b200_chain_plan** plan, int n_consumer, int block, int n_roles,
                             const int* n_producer, const int* const* tables, const int* table_role,
                             const int* table_nbh, int n_tables, int level_blocks, int lag_blocks,
                             void* stream);
int b200rt_chain_plan_view(const b200_chain_plan* plan, b200_chain_view* view);
unsigned b200rt_chain_plan_next_epoch(b200_chain_plan* plan); /* 1, 2, 3, ... one per launch */
int b200rt_chain_plan_status(b200_chain_plan* plan, void* stream);
int b200rt_chain_plan_destroy(b200_chain_plan* plan);

#ifdef __cplusplus
}
#endif
#endif
