/*
 * coupe_b200_tools.h — the steps either side of the RCB/RIB hot path in the
 * reference's tool chain (SURVEY.md §8f rows N1–N3), as device kernels behind a
 * plain C ABI, plus the two binary file formats that make runs interchangeable
 * with the reference tools.
 *
 *   N1  cell barycentres from mesh topology      tools/lib/lib.rs:511-539
 *   N2  weight-gen distributions                  tools/bins/weight-gen.rs:116-153
 *       MeWe weight file                          mesh-io/src/weight.rs:11-111,147-205
 *   N3  MePe partition file                       mesh-io/src/partition.rs:45-86
 *       algorithm spec "rcb,ITER[,TOL]"           tools/lib/lib.rs:418-421
 *       part loads / imbalance                    coupe/src/imbalance.rs:14-78
 *
 * Device pointers are marked _dev; everything else is host memory.  Every
 * function returns a coupe_err value (include/coupe.h).  Streams are `void *`
 * (cudaStream_t); calls synchronise the stream before returning when they
 * hand a result back to the host.
 */
#ifndef COUPE_B200_TOOLS_H
#define COUPE_B200_TOOLS_H

#include <stdint.h>

#include "coupe_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * N1 — barycentres of `n_elems` elements with `nodes_per_elem` nodes each
 * (3 triangle, 4 quadrangle/tetrahedron, 8 hexahedron; mesh-io/src/lib.rs:46-55).
 *   elem_nodes_dev  n_elems*nodes_per_elem node indices (usize = uint64), element-major
 *   coords_dev      n_nodes*dim doubles, AoS (Mesh::coordinates)
 *   out_dev         n_elems*dim doubles, AoS: the PointND<D> array RCB takes
 * Arithmetic follows tools/lib/lib.rs:526-535 to the bit: per coordinate the
 * node values are added in node order starting from 0.0, then divided by the
 * node count as f64.  An out-of-range node index makes the call fail with
 * COUPE_ERR_CRASH (the reference panics on the slice index).
 */
int coupe_b200_barycentres_device(coupe_b200_ctx *ctx, void *stream, uintptr_t dim, uintptr_t n_elems,
		uintptr_t nodes_per_elem, const uint64_t *elem_nodes_dev, const double *coords_dev,
		uintptr_t n_nodes, double *out_dev);

/*
 * N2 — weight-gen "linear,AXIS,FROM,TO" (weight-gen.rs:122-137): min and max of
 * the axis coordinate over all points, alpha = (to-from)/(max-min) stepped
 * down with nextafter while to-from < alpha*(max-min), weight =
 * fma(x - min, alpha, from).  `out_dev` receives n doubles (one criterion: every
 * "-d" option of weight-gen is one call; RCB only reads criterion 0).
 * min/max/alpha are also returned to the host (any may be NULL).  n == 0 is
 * an error in the reference (`min_by(..).unwrap()` panics): COUPE_ERR_CRASH.
 */
int coupe_b200_weight_linear_device(coupe_b200_ctx *ctx, void *stream, uintptr_t dim, uintptr_t n,
		const double *points_dev, int axis, double from, double to, double *out_dev,
		double *min_out, double *max_out, double *alpha_out);

/* The alpha of the linear distribution alone (host arithmetic, weight-gen.rs:128-135). */
double coupe_b200_linear_alpha(double from, double to, double min, double max);

/*
 * N2 — "spike,HEIGHT,POS...": weight = sum over spikes of exp(ln(height) -
 * |pos - point|) (weight-gen.rs:138-151).  heights[n_spikes], positions
 * [n_spikes*dim] are HOST arrays.  exp/ln/sqrt are CUDA's f64 library
 * functions: results agree with the reference's libm to a few ulp, not to
 * the bit (tests use 1e-14 relative).
 */
int coupe_b200_weight_spike_device(coupe_b200_ctx *ctx, void *stream, uintptr_t dim, uintptr_t n,
		const double *points_dev, uintptr_t n_spikes, const double *heights, const double *positions,
		double *out_dev);

/* N2 — "constant,VALUE" and the "-i" conversion `criterion as i64` (saturating,
 * NaN -> 0; weight-gen.rs:179-181). */
int coupe_b200_weight_constant_device(coupe_b200_ctx *ctx, void *stream, uintptr_t n, double value,
		double *out_dev);
int coupe_b200_weight_to_i64_device(coupe_b200_ctx *ctx, void *stream, uintptr_t n,
		const double *in_dev, int64_t *out_dev);

/*
 * N3 — per-part loads and the reported imbalance (coupe/src/imbalance.rs:14-78):
 * loads[p] = sum of the weights of the points with part id p, imbalance =
 * max_p (loads[p] - ideal) / ideal with ideal = total / num_parts, 0.0 when
 * num_parts == 0 or ideal == 0.  Integer weights (COUPE_B200_W_I32 / _I64)
 * are summed exactly in i64 (`loads_out` holds num_parts int64); f64 weights
 * are summed in exact 64-bit fixed point (scale chosen from max|w| and n, see
 * DESIGN.md) and returned as doubles — the reference's own f64 sum depends on
 * rayon's reduction order, parity is 1e-12 relative.  A part id >=
 * num_parts fails with COUPE_ERR_CRASH (debug_assert / index panic in the
 * reference).  `loads_out` may be NULL.
 */
int coupe_b200_imbalance_device(coupe_b200_ctx *ctx, void *stream, uintptr_t n, const uint64_t *part_dev,
		uintptr_t num_parts, int wtype, const void *weights_dev, void *loads_out, double *imbalance_out);

/*
 * File formats (host).  MeWe (weight-gen(1) "WEIGHT FILE"): "MeWe", version 1,
 * flags (bit 0 = integers), U16 criterion count, U64 weight count, then
 * count*criteria little-endian I64 / F64.  MePe (mesh-part(1) "PARTITION
 * FILE"): "MePe", U64 count, count little-endian U64 ids.
 * Readers malloc the array (free with coupe_b200_free); error mapping:
 * bad magic or unsupported version -> COUPE_ERR_BAD_TYPE, I/O -> COUPE_ERR_CRASH.
 */
int coupe_b200_mewe_write(const char *path, int is_integer, uint16_t criterion_count, uint64_t count,
		const void *values);
int coupe_b200_mewe_read(const char *path, int *is_integer, uint16_t *criterion_count, uint64_t *count,
		void **values);
int coupe_b200_mepe_write(const char *path, uint64_t count, const uint64_t *ids);
int coupe_b200_mepe_read(const char *path, uint64_t *count, uint64_t **ids);
void coupe_b200_free(void *p);

/* "rcb,ITER[,TOL]" (tools/lib/lib.rs:418-421; TOL defaults to 0.05).  Anything
 * else, a missing ITER or trailing garbage -> COUPE_ERR_NOT_FOUND. */
int coupe_b200_parse_rcb_spec(const char *spec, uintptr_t *iter_count, double *tolerance);

#ifdef __cplusplus
}
#endif

#endif
