/* oracle/ctdr_oracle.c — TEST INFRASTRUCTURE, NOT THE PRODUCT.
 *
 * CPU restatement of the reference mesh forward projector of jakeoung/ShapeFromProjections:
 *   ctdr/cuda/diff_fp_for.cu:31-224   (forward, pass 1 and pass 2)
 *   ctdr/cuda/diff_fp_back.cu:47-295  (backward)
 *   ctdr/function/rasterizer.py:23-32 (bounding boxes)
 * in float and double (the reference dispatches both: diff_fp_for.cu:251, diff_fp_back.cu:326).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker or the timed CPU baseline.  The product
 * (shapefromprojections_b200/) never imports it and has no CPU fallback.
 *
 * Parity pin: the reference ships no golden vectors for this path, so this
 * restatement is pinned against the reference's own kernels compiled from
 * /root/reference/ctdr/cuda into oracle/_ref (oracle/build_ref.py) and executed on a B200:
 * tests/golden/ref_raster_*.npz hold those outputs (generator: oracle/gen_ref_golden.py) and
 * tests/test_oracle_golden.py requires bit-equality of im / cnt / fragment lists.
 *
 * Build: gcc -O2 -mfma -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_EPS 1e-15   /* diff_fp_for.cu:27, diff_fp_back.cu:28 */

#define REAL float
#define NAME(x) x##_f32
#define FMA(a, b, c) fmaf((a), (b), (c))
#include "ctdr_oracle_impl.inc"
#undef REAL
#undef NAME
#undef FMA

#define REAL double
#define NAME(x) x##_f64
#define FMA(a, b, c) fma((a), (b), (c))
#include "ctdr_oracle_impl.inc"
#undef REAL
#undef NAME
#undef FMA

int oracle_abi_version(void) { return 2; }

/* Thread control for the timed CPU baseline: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
 * silently time the "all host cores" arm on one core; bench.py sets the count explicitly and reports what
 * the OpenMP runtime actually uses. */
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_max_threads(void) { return omp_get_max_threads(); }
