/* A C host of include/coupe_b200_tools.h: the parts that need no GPU (the
 * "rcb,ITER[,TOL]" spec, the MePe / MeWe codecs, the linear-weight alpha) used
 * the way tools/bins/mesh-part.rs and weight-gen.rs use their Rust
 * counterparts.  Exit status 0 iff everything round-trips. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "coupe.h"
#include "coupe_b200_tools.h"

int main(int argc, char **argv) {
	uintptr_t iters = 0;
	double tol = 0.0;
	uint64_t ids[5] = {3, 1, 4, 1, 5}, *back = NULL, count = 0;
	double w[3] = {0.5, 1.5, 2.5};
	void *wback = NULL;
	int is_int = -1;
	uint16_t crit = 0;
	char path[4096];
	if (argc < 2) return 2;
	if (coupe_b200_parse_rcb_spec("rcb,10,0.001", &iters, &tol) != COUPE_ERR_OK || iters != 10 || tol != 0.001) return 3;
	if (coupe_b200_parse_rcb_spec("hilbert,4", &iters, &tol) != COUPE_ERR_NOT_FOUND) return 4;
	snprintf(path, sizeof path, "%s/p.mepe", argv[1]);
	if (coupe_b200_mepe_write(path, 5, ids) != COUPE_ERR_OK) return 5;
	if (coupe_b200_mepe_read(path, &count, &back) != COUPE_ERR_OK || count != 5 || memcmp(ids, back, sizeof ids)) return 6;
	coupe_b200_free(back);
	snprintf(path, sizeof path, "%s/w.mewe", argv[1]);
	if (coupe_b200_mewe_write(path, 0, 1, 3, w) != COUPE_ERR_OK) return 7;
	if (coupe_b200_mewe_read(path, &is_int, &crit, &count, &wback) != COUPE_ERR_OK || is_int != 0 || crit != 1 ||
	    count != 3 || memcmp(w, wback, sizeof w))
		return 8;
	coupe_b200_free(wback);
	if (coupe_b200_linear_alpha(0.0, 100.0, 0.0, 50.0) != 2.0) return 9;
	printf("ok %lu %g\n", (unsigned long)iters, tol);
	return 0;
}
