/*
 * coupe_b200_mj.h — SURVEY.md §8f row N4: coupe's Multi-Jagged partitioner and axis_sort as device
 * code behind a plain C ABI.
 *
 *   MultiJagged { part_count, max_iter } + impl Partition   coupe/src/algorithms/multi_jagged.rs:354-366
 *   partition_scheme / compute_modifiers                      multi_jagged.rs:70-98, :136-148
 *   multi_jagged_recurse                                      multi_jagged.rs:181-220
 *   compute_split_positions                                   multi_jagged.rs:222-288
 *   axis_sort                                                 recursive_bisection.rs:815-827
 *   Grid::rcb (cartesian RCB, weighted median on prefix sums) coupe/src/cartesian/mod.rs:119-181, rcb.rs:52-266
 *
 * The reference exposes MultiJagged through the Rust `Partition` trait only (coupe-ffi has no entry
 * for it), so there is no reference C prototype to match: the entry points below are what a
 * `coupe-gpu` backend crate binds.  Device pointers are marked _dev.  Every function
 * returns a coupe_err value (include/coupe.h).
 *
 * What the reference leaves to the rayon schedule is pinned:
 * equal coordinates keep their previous order (the sort is stable), parts are numbered depth first
 * and left to right, and the weights of a node are summed in chunks of COUPE_B200_MJ_CHUNK consecutive
 * elements of its sorted slice.  Where the reference panics (a part left empty that still has to be
 * split, an all-zero total weight, part_count == 0) the call returns COUPE_ERR_CRASH.
 */
#ifndef COUPE_B200_MJ_H
#define COUPE_B200_MJ_H

#include <stdint.h>

#include "coupe_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* elements per fold chunk of compute_split_positions (multi_jagged.rs:241-248) */
#define COUPE_B200_MJ_CHUNK 1024

/*
 * Multi-Jagged on device-resident data.
 *   part_dev     n uint64 part ids in [0, part_count), written
 *   points_dev   n*dim doubles, AoS (dim 2 or 3)
 *   weights_dev  n doubles
 * n must be below 2^32.
 */
int coupe_b200_multi_jagged_device(coupe_b200_ctx *ctx, void *stream, uint64_t *part_dev, uintptr_t dim,
		uintptr_t n, const double *points_dev, const double *weights_dev, uintptr_t part_count,
		uintptr_t max_iter);

/* The same on host arrays (plain copies up and down). */
int coupe_b200_multi_jagged_host(coupe_b200_ctx *ctx, uint64_t *part, uintptr_t dim, uintptr_t n,
		const double *points, const double *weights, uintptr_t part_count, uintptr_t max_iter);

/*
 * axis_sort: sorts `permutation_dev` (n indices into the points, uint64 like the reference's usize)
 * by the `coord`-th coordinate of the points they name; equal coordinates keep their order.
 */
int coupe_b200_axis_sort_device(coupe_b200_ctx *ctx, void *stream, uintptr_t dim, uintptr_t n_points,
		const double *points_dev, uint64_t *permutation_dev, uintptr_t len, uintptr_t coord);

/* Leaves and levels of the partition scheme (multi_jagged.rs:70-98); COUPE_ERR_CRASH where the reference panics. */
int coupe_b200_mj_scheme(uintptr_t part_count, uintptr_t max_iter, uint64_t *leaves_out, uint64_t *levels_out);

/*
 * Cartesian RCB: coupe::Grid::new_2d / new_3d (sizes).rcb(partition, weights, iter_count)
 * (coupe/src/cartesian/mod.rs:119-181, rcb.rs:52-266).
 *   part_dev     one uint64 per cell, row major (x fastest), written: the id IterationResult::part_of gives
 *   sizes        `dim` grid sides {width, height[, depth]} (host array), all non-zero
 *   wtype        COUPE_INT64 or COUPE_DOUBLE; weights_dev: one weight per cell, row major
 *   threads      the size of the rayon pool the reference would run under: weighted_median (rcb.rs:64-68)
 *                cuts its search range into chunks of (max - min) / current_num_threads() elements, so the
 *                partition depends on it; must be >= 2 (with one thread the reference does not return)
 * The axis sums keep the reference's sequential order; the total weight (a parallel sum there) is the row
 * sums added in memory order.
 */
int coupe_b200_grid_rcb_device(coupe_b200_ctx *ctx, void *stream, uint64_t *part_dev, uintptr_t dim,
		const uint64_t *sizes, int wtype, const void *weights_dev, uintptr_t iter_count, uintptr_t threads);

int coupe_b200_grid_rcb_host(coupe_b200_ctx *ctx, uint64_t *part, uintptr_t dim, const uint64_t *sizes,
		int wtype, const void *weights, uintptr_t iter_count, uintptr_t threads);

/* Device time (ms, CUDA events) of the last multi_jagged call on this context's device: {total, sort passes, the rest}. */
int coupe_b200_mj_last_times(const coupe_b200_ctx *ctx, double *ms3);

#ifdef __cplusplus
}
#endif
#endif
