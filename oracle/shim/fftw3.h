/* oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Declaration shim for the slice of the FFTW3 C API that the reference's hot
 * path uses, so that the UNMODIFIED reference sources under /root/reference/src
 * compile in a container that has no libfftw3:
 *
 *   fftw_plan_dft_1d / fftw_execute / fftw_destroy_plan
 *       reference call sites: src/process/ambiguity/Ambiguity.cpp:73-80,120-121,129,160
 *                             src/process/clutter/WienerHopf.cpp:31-44,72-73,80,104,145-146,153
 *   fftw_init_threads / fftw_plan_with_nthreads
 *       reference call site:  src/blah2.cpp:115-120 (not on the oracle's path; stubs)
 *
 * The implementation behind it (fftw_shim.cpp) is our own double-precision
 * any-length CPU FFT (mixed radix 2/3/4/5/small primes + Bluestein).  A DFT is
 * exact mathematics: any correct FP64 FFT reproduces FFTW's output to ~1e-15
 * relative, which is far inside the 1e-5 parity budget.
 */
#ifndef B200DD_ORACLE_FFTW3_SHIM_H
#define B200DD_ORACLE_FFTW3_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct b200dd_shim_plan_s *fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
void fftw_cleanup_threads(void);

#ifdef __cplusplus
}
#endif

#endif
