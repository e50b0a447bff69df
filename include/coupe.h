/*
 * coupe.h — C ABI of the RCB / RIB hot path, B200 implementation.
 *
 * Drop-in for the subset of coupe-ffi's header that covers this path: the
 * enums, the data-set constructors and coupe_rcb / coupe_rib have the same
 * names, values, argument meaning and error behaviour as the reference
 *   coupe-ffi/include/coupe.h:16-57    enum coupe_err
 *   coupe-ffi/include/coupe.h:64       coupe_strerror
 *   coupe-ffi/include/coupe.h:86-93    enum coupe_type
 *   coupe-ffi/include/coupe.h:105-188  coupe_data_{free,array,constant,fn}
 *   coupe-ffi/include/coupe.h:278-280  coupe_rcb
 *   coupe-ffi/include/coupe.h:294-296  coupe_rib
 * (Rust side: coupe-ffi/src/lib.rs:35-157, :255-364; coupe-ffi/src/data.rs).
 * The other algorithms of the reference header (hilbert, greedy, kk, ckk, fm,
 * adjacency structures) are out of scope and not exported.
 *
 * All pointers are HOST pointers; the library copies to the GPU, runs the
 * CUDA path and copies the part ids back.  There is no CPU fallback: without a
 * usable CUDA device every algorithm call returns COUPE_ERR_CRASH.
 */
#ifndef COUPE_H
#define COUPE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Return type of the algorithms (values 0..8, as in the reference). */
enum coupe_err {
	COUPE_ERR_OK,            /* no error */
	COUPE_ERR_ALLOC,         /* host or device allocation failed */
	COUPE_ERR_CRASH,         /* internal failure (CUDA / NCCL error, no device) */
	COUPE_ERR_BAD_DIMENSION, /* dimension is not 2 or 3 */
	COUPE_ERR_BAD_TYPE,      /* unsupported weight type tag */
	COUPE_ERR_BIPART_ONLY,   /* unused by rcb/rib, kept for value parity */
	COUPE_ERR_LEN_MISMATCH,  /* points and weights differ in length */
	COUPE_ERR_NOT_FOUND,     /* unused by rcb/rib, kept for value parity */
	COUPE_ERR_NEG_VALUES,    /* unused by rcb/rib, kept for value parity */
};

/* Static, nul-terminated message for an error code. */
const char *coupe_strerror(enum coupe_err err);

/* Opaque data set: an array, a repeated constant, or a callback. */
typedef struct coupe_data coupe_data;

enum coupe_type {
	COUPE_INT,    /* `int` (32 bit) */
	COUPE_INT64,  /* `int64_t` */
	COUPE_DOUBLE, /* `double` */
};

/* May be called with NULL; never frees the user memory the set points to. */
void coupe_data_free(coupe_data *data);

/* `len` elements read from `data`.  For point sets one element is
 * `dimension` consecutive doubles and `len` counts points. */
coupe_data *coupe_data_array(uintptr_t len, enum coupe_type type, const void *data);

/* `len` copies of `*value`. */
coupe_data *coupe_data_constant(uintptr_t len, enum coupe_type type, const void *value);

/* `len` elements, element i is `*i_th(context, i)`.  The callback may be
 * called from several threads, in any order, several times per index. */
coupe_data *coupe_data_fn(const void *context, uintptr_t len, enum coupe_type type,
		const void *(*i_th)(const void *, uintptr_t));

/*
 * Recursive coordinate bisection: at most 2^iter_count parts, ids written to
 * `partition` (N uintptr_t, fully overwritten when N > 0, ids start at 0).
 * Errors: dimension not in {2,3} -> BAD_DIMENSION; length mismatch ->
 * LEN_MISMATCH; unknown weight tag -> BAD_TYPE; device trouble -> ALLOC/CRASH.
 * Limit of this implementation (the reference recurses to any depth): the node
 * tables hold 2^iter_count entries on every GPU, so iter_count > 24 (more
 * than 2^24 parts) returns COUPE_ERR_ALLOC without touching `partition`.
 */
enum coupe_err coupe_rcb(uintptr_t *partition, uintptr_t dimension,
		const coupe_data *points, const coupe_data *weights,
		uintptr_t iter_count, double tolerance);

/* Recursive inertial bisection: rotate onto the principal inertia axis, then RCB. */
enum coupe_err coupe_rib(uintptr_t *partition, uintptr_t dimension,
		const coupe_data *points, const coupe_data *weights,
		uintptr_t iter_count, double tolerance);

#ifdef __cplusplus
}
#endif

#endif
