/* A C host of the reference-compatible ABI (include/coupe.h), written the way
 * a coupe-ffi user writes one (cf. coupe-ffi/examples/rcb.c): unit square,
 * constant COUPE_INT weight, one and two bisection levels; then a callback
 * data set.  Prints the part ids; exit status 0 iff every call returned OK. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "coupe.h"

static const double square[8] = {0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0};

static const void *ith_point(const void *ctx, uintptr_t i) {
	return (const double *)ctx + 2 * i;
}

static int run(uintptr_t iters, const coupe_data *points, const coupe_data *weights) {
	uintptr_t part[4] = {99, 99, 99, 99};
	enum coupe_err err = coupe_rcb(part, 2, points, weights, iters, 0.05);
	if (err != COUPE_ERR_OK) {
		fprintf(stderr, "coupe_rcb: %s\n", coupe_strerror(err));
		return 1;
	}
	printf("%lu %lu %lu %lu\n", (unsigned long)part[0], (unsigned long)part[1],
	       (unsigned long)part[2], (unsigned long)part[3]);
	return 0;
}

int main(void) {
	int one = 1, bad = 0;
	coupe_data *points = coupe_data_array(4, COUPE_DOUBLE, square);
	coupe_data *weights = coupe_data_constant(4, COUPE_INT, &one);
	coupe_data *fn_points = coupe_data_fn(square, 4, COUPE_DOUBLE, ith_point);
	if (!points || !weights || !fn_points) return 2;
	bad |= run(1, points, weights);
	bad |= run(2, points, weights);
	bad |= run(2, fn_points, weights);
	coupe_data_free(points);
	coupe_data_free(weights);
	coupe_data_free(fn_points);
	coupe_data_free(NULL);
	return bad;
}
