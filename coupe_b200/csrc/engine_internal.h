// engine_internal.h — what the host path (ffi.cu) uses of the engine (engine.cu) beyond the exported
// C ABI: the engine's own device columns, so that points narrowed on the host are copied straight
// into them, and a run that takes those columns as its prologue and leaves compact ids behind.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include <vector>

struct coupe_b200_ctx;
// One context per GPU of ONE process (coupe_b200_group_create): ranks of an in-process communicator.
struct coupe_b200_group {
  std::vector<coupe_b200_ctx *> ctx;
};

namespace cb_engine {

struct HostColumns {
  float *x[3];        // SoA f32 coordinate columns of the engine (npad floats each)
  size_t npad;
  void *w;            // device copy of the caller's weights (n * wbytes), null when wbytes == 0
  void *ids_compact;  // the run's compact part ids: u16 when iter_count <= 16, else u32
  double *pts_raw;    // RIB: device copy of the caller's AoS f64 points (n * dim), else null
};

// Bounding-box keys of columns filled from the host: order-preserving u32 keys of the f32 minima
// ([0..D)) and inverted keys of the maxima ([4..4+D)), KEY_EMPTY elsewhere (rcb_kernels.cuh: f2key).
struct Prefilled {
  uint32_t bbox_keys[8];
};

// Locks / unlocks the context for a whole host call (buffers and streams are per context).
void lock(coupe_b200_ctx *c);
void unlock(coupe_b200_ctx *c);
// Frees the device buffers of the host path (host_columns); context locked.
void release_host_buffers(coupe_b200_ctx *c);
// Called by coupe_b200_ctx_destroy before the context goes away (ffi.cu drops its per-context lanes).
extern void (*on_destroy)(coupe_b200_ctx *c);
int device_of(const coupe_b200_ctx *c);
int rank_of(const coupe_b200_ctx *c);
int world_of(const coupe_b200_ctx *c);

// Ensures the device buffers of a host call of n points; returns a coupe_err.  Context locked.
int host_columns(coupe_b200_ctx *c, size_t n, size_t dim, size_t wbytes, bool raw_points, size_t iter_count,
                 HostColumns *out);

// The run itself, context locked by the caller.  `pre` non-null: the columns hold the narrowed points
// (no AoS f64 copy exists on the device: points_dev is ignored); compact ids go to HostColumns::
// ids_compact and *id_bytes is 2 or 4.  Returns a coupe_err.
int run_locked(coupe_b200_ctx *c, bool rib, cudaStream_t st, const Prefilled *pre, int *id_bytes, uintptr_t dim,
               uintptr_t n, const double *points_dev, int wtype, const void *weights_dev, const void *wconst_host,
               uintptr_t iter_count, double tolerance);

}  // namespace cb_engine
