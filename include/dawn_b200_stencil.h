/* dawn_b200_stencil.h — C ABI that EVERY translation unit emitted by the b200 backend exports, with
 * <N> = the stencil instantiation's name (IIR `metadata.stencilName`).
 *
 * It is the Cartesian counterpart of the C interface the reference's cuda-ico backend generates
 * (`setup_<N>` / `run_<N>` / `free_<N>`, dawn/src/dawn/CodeGen/Cuda-ico/CudaIcoCodeGen.cpp:807-947,
 * 1186-1228), made handle-based (the reference keeps the mesh, stream and temporaries in static class
 * members, :1230-1264, i.e. one instance per process).  The same translation unit also contains the
 * C++ class `dawn_generated::b200::<N>` with the reference's generated-class API (constructor from
 * `gridtools::dawn::domain`, `run(storages...)`, `get_/set_<global>`, `get_name`, `reset_meters`,
 * `get_total_time`; CudaCodeGen.cpp:144-174,318-368,412-452) for C++ drivers.
 *
 * This header documents the symbols with N spelled literally as `NAME`; bind them by name
 * (dlsym / ctypes / ISO_C_BINDING).  Fields are DEVICE resident (b200_field_t, dawn_b200_rt.h) and
 * passed in the stencil's API order (`run()` argument order = IIR `APIFieldIDs`).
 */
#ifndef DAWN_B200_STENCIL_H
#define DAWN_B200_STENCIL_H
#include "dawn_b200_rt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Creates an instance for a domain of isize x jsize x ksize points INCLUDING halos;
 * halos = {iminus, iplus, jminus, jplus, kminus, kplus} (NULL: 3,3,3,3,0,0 as domain.hpp:42-58).
 * `stream` is the cudaStream_t the timers record on (kernels run on the stream given to run). */
int setup_NAME(void** handle, int isize, int jsize, int ksize, const int* halos, void* stream);

/* Launches every kernel of the stencil on `stream`; asynchronous.  Returns B200_ERR_HALO if the
 * domain's halos are smaller than the stencil's access extents, B200_ERR_LAYOUT if fields of equal
 * dimensionality disagree on strides, B200_ERR_ARG on a wrong field count. */
int run_NAME(void* handle, const b200_field_t* fields, int nfields, void* stream);

/* Same, restricted to the interior rows [j_begin, j_end) (0-based, halos excluded).  Lets a driver
 * run the rows that do not depend on halo rows while a halo exchange is in flight and the boundary
 * strips afterwards (j-slab sharding across GPUs). */
int run_rows_NAME(void* handle, const b200_field_t* fields, int nfields, void* stream, int j_begin,
                  int j_end);

void free_NAME(void* handle);

/* Number of kernel launches issued by this instance so far. */
long launches_NAME(void* handle);

/* Cartesian modules: shape of the most recent kernel launch - out5 = {grid.x, grid.y, grid.z, levels per
 * block (k chunk), kernel kind: 0 plain, 1 TMA-staged tile kernel, 2 TMA column ring, 3 column cache}.
 * Diagnostic (the reference has no twin); tests use it to assert which kernel ran and that the
 * shared-memory ring over k reached its steady state. */
void last_launch_NAME(void* handle, int* out5);

/* One pair per global variable G of the stencil (reference: get_G()/set_G(), CodeGen.cpp:110-136). */
void set_NAME_G(void* handle, double value);
double get_NAME_G(void* handle);

/* JSON description: {"name", "grid", "float_type": "double"|"float", "fields":[{"name","dims",
 * "extent":[i-,i+,j-,j+,k-,k+]}], "written":[...], "k_parallel": bool, "tile": [points per block in i, j],
 * "globals":[...]} — field order is the run() order; `extent` equals the reference's `static constexpr
 * cartesian_extent <field>_extent` (CodeGen.cpp:509-522); `tile` lets a driver that splits a run by rows
 * (run_rows_NAME) cut along whole tiles. */
const char* b200_module_info_NAME(void);

#ifdef __cplusplus
}
#endif
#endif
