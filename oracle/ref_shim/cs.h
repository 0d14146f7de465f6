/* cs.h -- TEST INFRASTRUCTURE ONLY.  Declarations of the CSparse types / functions that the reference's alternative
 * (compiled, never executed: `USE_CSPARSE = false`, ISAM/isamlib/Cholesky.cpp:40) code path names, so that the
 * unmodified Cholesky.cpp compiles for oracle/_ref.  cholmod_shim.cpp defines them as aborting stubs. */
#ifndef POPUP_ORACLE_CS_SHIM_H
#define POPUP_ORACLE_CS_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct cs_sparse { int nzmax, m, n; int* p; int* i; double* x; int nz; } cs;
typedef struct cs_symbolic { int* pinv; int* q; int* parent; int* cp; int* leftmost; int m2; double lnz, unz; } css;
typedef struct cs_numeric { cs* L; cs* U; int* pinv; double* B; } csn;
cs* cs_spalloc(int m, int n, int nzmax, int values, int triplet);
cs* cs_spfree(cs* A);
cs* cs_transpose(const cs* A, int values);
cs* cs_multiply(const cs* A, const cs* B);
css* cs_sqr(int order, const cs* A, int qr);
csn* cs_qr(const cs* A, const css* S);
css* cs_schol(int order, const cs* A);
csn* cs_chol(const cs* A, const css* S);
css* cs_sfree(css* S);
csn* cs_nfree(csn* N);
void* cs_free(void* p);
int* cs_pinv(const int* p, int n);
int cs_gaxpy(const cs* A, const double* x, double* y);
int cs_lsolve(const cs* L, double* x);
int cs_ltsolve(const cs* L, double* x);
#ifdef __cplusplus
}
#endif
#endif
