/* Stub libgsl.so.0 for the reference's prebuilt count/solve binaries (test infrastructure).
 * The binaries import these four symbols (nm -D --undefined-only) from the read simulators
 * (common/read.h:289,497,534-562), which count/solve never call.  Reaching one is a bug. */
#include <stdlib.h>
void *gsl_ran_discrete_preproc(size_t k, const double *p) { (void)k; (void)p; abort(); }
size_t gsl_ran_discrete(void *r, void *g) { (void)r; (void)g; abort(); }
void gsl_ran_discrete_free(void *g) { (void)g; abort(); }
double gsl_ran_flat(void *r, double a, double b) { (void)r; (void)a; (void)b; abort(); }
