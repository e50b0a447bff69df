/*
 * coupe_b200.h — device-level C ABI of the B200 RCB/RIB engine.
 *
 * coupe.h (the reference-compatible boundary) is implemented on top of these
 * entry points; they are exported so that a host that already keeps its mesh
 * on the GPU (or a benchmark that wants the H2D/D2H copies outside the timed
 * region) can call the same CUDA path with DEVICE pointers.  Plain pointers
 * and sizes only; streams are passed as `void *` (a cudaStream_t).
 *
 * What each call replaces in the reference (coupe/src/algorithms/
 * recursive_bisection.rs): coupe_b200_rcb_device = rcb() :644-705 with
 * rcb_recurse :575-642 and par_rcb_split :456-573; coupe_b200_rib_device =
 * rib() :829-851 with OrientedBoundingBox::from_points (geometry.rs:210-228).
 */
#ifndef COUPE_B200_H
#define COUPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct coupe_b200_ctx coupe_b200_ctx;

/* Weight descriptions accepted by the device entry points. */
enum coupe_b200_wtype {
	COUPE_B200_W_I32 = 0, /* same values as enum coupe_type */
	COUPE_B200_W_I64 = 1,
	COUPE_B200_W_F64 = 2,
};

/* Counters of the last call on a context (all ranks hold the same values
 * except the *_local ones). */
typedef struct coupe_b200_stats {
	uint64_t n_local;        /* points on this rank */
	uint64_t n_global;       /* points over all ranks */
	uint32_t levels;         /* iter_count */
	uint32_t dense_sweeps;   /* first passes (one per level) */
	uint32_t refine_sweeps;  /* sparse refinement passes */
	uint32_t kernel_launches;/* CUDA kernels launched by the call */
	uint32_t collectives;    /* NCCL calls issued by the call */
	int32_t  weight_shift;   /* f64 weights: one accumulator unit is 2^-weight_shift (wide form: at the root node) */
	uint32_t host_syncs;     /* stream synchronisations inside the call */
	uint32_t flag_waits;     /* passes whose result the host polled for (mapped host memory) */
	uint32_t peer_exchange;  /* 1: histograms went through the peer-memory exchange, 0: NCCL / single GPU */
	uint32_t carry_free;     /* 1: integer weights small enough for 32-bit block-private sums (no carry chain in the sweeps) */
	uint32_t weight_wide;    /* f64 weights: 0 narrow form (i32 multiples of one unit), 1 wide form (i64, one unit per tree node) */
	uint32_t weight_rescales; /* f64 weights: root passes redone because the sampled weight statistics chose another form or scale
	                             than all the weights do, or because the wide form wanted a finer root unit (0..2) */
	double   matrix[9];      /* RIB: the obb_to_aabb matrix applied (row-major DxD) */
	double   dense_sweep_ms; /* option "time_sweeps": summed device time of the dense sweeps */
	double   refine_sweep_ms;/* option "time_sweeps": summed device time of the refinement sweeps */
	uint64_t refine_points;  /* points re-binned by the refinement sweeps (this rank) */
	double   exchange_wait_ms; /* multi-GPU, peer-memory exchange: time the walks of this rank waited for the other
	                              ranks' histograms (first block of every pass, SM cycles at the nominal clock) */
	uint32_t deferred_levels;  /* levels left undecided by their dense pass whose refinement read the list of deferred
	                              points written by the next level's dense sweep (no rescan of the idx words) */
	uint32_t list_refine_sweeps; /* ... refinement passes over such a list (counted in refine_sweeps too) */
} coupe_b200_stats;

/* One context per process and GPU.  `device` is a CUDA ordinal.  Returns a
 * coupe_err value. */
int coupe_b200_ctx_create(coupe_b200_ctx **out, int device);
void coupe_b200_ctx_destroy(coupe_b200_ctx *ctx);
/* CUDA ordinal the context was created on (-1 for NULL). */
int coupe_b200_ctx_device(const coupe_b200_ctx *ctx);

/* Multi-GPU: rank 0 calls coupe_b200_nccl_unique_id (128 bytes out), the host
 * framework broadcasts the bytes (torch.distributed), every rank then calls
 * coupe_b200_ctx_init_comm.  Points are sharded by the caller; each rank
 * passes its own shard to the calls below and receives ids for its shard. */
int coupe_b200_nccl_unique_id(void *out128);
int coupe_b200_ctx_init_comm(coupe_b200_ctx *ctx, const void *unique_id128, int rank, int world);

/*
 * RCB on device-resident data.
 *   part_dev     n uint64 part ids (device), written
 *   points_dev   n*dim doubles, AoS (device)
 *   weights_dev  n weights of `wtype` (device), or NULL for a constant weight
 *   wconst_host  when weights_dev is NULL: host pointer to ONE value of `wtype`
 * Returns a coupe_err value.  The call is synchronous with respect to the
 * host only at its end (the stream is synchronised before returning).
 */
int coupe_b200_rcb_device(coupe_b200_ctx *ctx, void *stream, uint64_t *part_dev, uintptr_t dim,
		uintptr_t n, const double *points_dev, int wtype, const void *weights_dev,
		const void *wconst_host, uintptr_t iter_count, double tolerance);

int coupe_b200_rib_device(coupe_b200_ctx *ctx, void *stream, uint64_t *part_dev, uintptr_t dim,
		uintptr_t n, const double *points_dev, int wtype, const void *weights_dev,
		const void *wconst_host, uintptr_t iter_count, double tolerance);

/*
 * The same two algorithms on HOST arrays with an explicit context (`partition`
 * holds n uintptr_t, as in coupe_rcb).  Host threads narrow the points to the f32
 * columns the kernels read while they copy them up (RIB: the f64 points go up, the
 * rotation precedes the narrowing) and widen the compact ids that come back.  This is what a language
 * binding that already holds plain slices calls (rust/coupe-gpu: `impl
 * Partition for GpuRcb`, replacing the body of coupe::Rcb::partition,
 * recursive_bisection.rs:805-811); coupe_rcb / coupe_rib of coupe.h are these
 * calls on the process-wide default context after unwrapping the coupe_data
 * sets.  coupe_b200_host_release frees the device staging buffers a context
 * accumulated through these calls (call it before coupe_b200_ctx_destroy).
 */
int coupe_b200_rcb_host(coupe_b200_ctx *ctx, uintptr_t *partition, uintptr_t dim, uintptr_t n,
		const double *points, int wtype, const void *weights, const void *wconst,
		uintptr_t iter_count, double tolerance);
int coupe_b200_rib_host(coupe_b200_ctx *ctx, uintptr_t *partition, uintptr_t dim, uintptr_t n,
		const double *points, int wtype, const void *weights, const void *wconst,
		uintptr_t iter_count, double tolerance);
void coupe_b200_host_release(coupe_b200_ctx *ctx);

/*
 * One process, several GPUs of one box.  A group holds one context per device (`devices` lists
 * `ndev` CUDA ordinals; ndev <= 0: every device of the box), ranks of an in-process NCCL
 * communicator with the peer-memory exchange mapped by plain peer access.  The *_host_group calls
 * shard the caller's HOST arrays by contiguous index ranges over the devices (one driving thread
 * per GPU, the copy threads split between them) and write all n part ids: the same results as one
 * GPU.  coupe_rcb / coupe_rib of coupe.h take this path when the environment variable
 * COUPE_B200_DEVICES is set ("all", or a comma-separated list of ordinals).
 */
typedef struct coupe_b200_group coupe_b200_group;
int coupe_b200_group_create(coupe_b200_group **out, const int *devices, int ndev);
void coupe_b200_group_destroy(coupe_b200_group *group);
int coupe_b200_group_size(const coupe_b200_group *group);
/* Context of device i of the group (statistics, options); owned by the group. */
coupe_b200_ctx *coupe_b200_group_ctx(coupe_b200_group *group, int i);
int coupe_b200_rcb_host_group(coupe_b200_group *group, uintptr_t *partition, uintptr_t dim, uintptr_t n,
		const double *points, int wtype, const void *weights, const void *wconst,
		uintptr_t iter_count, double tolerance);
int coupe_b200_rib_host_group(coupe_b200_group *group, uintptr_t *partition, uintptr_t dim, uintptr_t n,
		const double *points, int wtype, const void *weights, const void *wconst,
		uintptr_t iter_count, double tolerance);

/* Counters of the last call. */
int coupe_b200_last_stats(const coupe_b200_ctx *ctx, coupe_b200_stats *out);

/* Option "time_sweeps": device time (CUDA events on the call's stream) of every timed sweep of the
 * last call, in launch order: milliseconds, tree level, kind (0 dense sweep, 1 refinement sweep, 2 an
 * optimistic dense sweep that found the previous level undecided and returned at once).
 * Writes at most `cap` entries (arrays may be NULL) and returns how many sweeps were timed. */
uint32_t coupe_b200_last_sweep_times(const coupe_b200_ctx *ctx, double *ms, int32_t *level, int32_t *kind,
		uint32_t cap);

/*
 * Split tree of the last call, heap order (root 0, children 2i+1 / 2i+2),
 * 2^iter_count - 1 entries each, copied to HOST arrays (any may be NULL):
 * visited flag, f32 split position, left weight and node weight converted to
 * f64 the way the bisection compared them, bisection iterations.
 */
int coupe_b200_last_trace(coupe_b200_ctx *ctx, uint8_t *visited, float *split_pos,
		double *weight_left, double *sum, uint32_t *iters);

/* Pre-size the context's scratch buffers for shards of up to n points so that
 * later calls do not allocate. */
int coupe_b200_reserve(coupe_b200_ctx *ctx, uintptr_t n, uintptr_t dim, uintptr_t iter_count);

/* Tuning knobs for experiments (bench/profiles); safe defaults otherwise.  Unknown names return
 * COUPE_ERR_NOT_FOUND.
 *   kmax_a (1..10, 8)        bisection steps a dense sweep resolves (2^k bins per node)
 *   kmax_refine (1..10, 10)  ... a refinement sweep
 *   nb_smem_log2 (6..14, 14) histogram slots a block keeps in shared memory
 *   force_global (0)         accumulate with L2 atomics instead of shared-memory histograms
 *   trace (1)                keep the split tree for coupe_b200_last_trace
 *   time_sweeps (0)          1: CUDA events around the dense sweeps, 2: refinement sweeps too, 3: and print them
 *   peer_exchange (1)        multi-GPU: histograms over peer memory (0: NCCL all-reduces)
 *   sample_weights (1)       f64 weights: fixed-point form and scale from a sample, verified by the root sweep
 *   defer (0..2, 1)          levels left undecided by their dense pass: 0 refine by rescanning the idx words; 1 the next
 *                            level's dense sweep lists the points of the undecided bins (no rescan) where the level is
 *                            predicted to stay undecided — fewer bits per pass than the first levels, or undecided in the
 *                            context's previous call; 2 at every level.  The results do not depend on it.
 *   carve_fit (1), smem_pad (0), table_rep_max (3)   experiments on the shared-memory layout of the dense sweeps */
int coupe_b200_set_option(coupe_b200_ctx *ctx, const char *name, int64_t value);

/* Library / build identification string. */
const char *coupe_b200_version(void);

#ifdef __cplusplus
}
#endif

#endif
