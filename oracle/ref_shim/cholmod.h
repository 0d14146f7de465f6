/* cholmod.h -- TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
 *
 * SuiteSparse/CHOLMOD is a system library the reference links against (ISAM/isamlib/Cholesky.cpp:36-37, un-vendored,
 * version unpinned) and is not installed in the build container.  This header declares exactly the slice of the
 * CHOLMOD C API that Cholesky.cpp calls, with the documented semantics (column-compressed sparse matrices, P A P' = L L',
 * cholmod_solve systems CHOLMOD_P / CHOLMOD_L / CHOLMOD_Lt); cholmod_shim.cpp implements it with a weighted
 * minimum-degree ordering on compressed (indistinguishable) variables and an up-looking simplicial Cholesky.  It lets
 * oracle/_ref run the UNMODIFIED reference optimiser (Slam.cpp, Optimizer.cpp, Cholesky.cpp ...) end to end. */
#ifndef POPUP_ORACLE_CHOLMOD_SHIM_H
#define POPUP_ORACLE_CHOLMOD_SHIM_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHOLMOD_PATTERN 0
#define CHOLMOD_REAL 1
#define CHOLMOD_A 0
#define CHOLMOD_LDLt 1
#define CHOLMOD_LD 2
#define CHOLMOD_DLt 3
#define CHOLMOD_L 4
#define CHOLMOD_Lt 5
#define CHOLMOD_D 6
#define CHOLMOD_P 7
#define CHOLMOD_Pt 8
#define CHOLMOD_NATURAL 0
#define CHOLMOD_GIVEN 1
#define CHOLMOD_AMD 2
#define CHOLMOD_METIS 3
#define CHOLMOD_NESDIS 4
#define CHOLMOD_COLAMD 5

typedef struct cholmod_method_struct { int ordering; } cholmod_method;
typedef struct cholmod_common_struct {
  int nmethods;
  cholmod_method method[10];
  int postorder;
  int status;
  int supernodal;
} cholmod_common;

typedef struct cholmod_sparse_struct {
  size_t nrow, ncol, nzmax;
  void *p, *i, *nz, *x, *z;
  int stype, itype, xtype, dtype, sorted, packed;
} cholmod_sparse;

typedef struct cholmod_dense_struct {
  size_t nrow, ncol, nzmax, d;
  void *x, *z;
  int xtype, dtype;
} cholmod_dense;

typedef struct cholmod_factor_struct {
  size_t n, minor;
  void *Perm, *ColCount;
  size_t nzmax;
  void *p, *i, *x, *z, *nz;
  int ordering, is_ll, is_super, is_monotonic, itype, xtype, dtype;
  void* impl;   /* shim-private symbolic data */
} cholmod_factor;

int cholmod_start(cholmod_common* c);
int cholmod_finish(cholmod_common* c);
cholmod_sparse* cholmod_allocate_sparse(size_t nrow, size_t ncol, size_t nzmax, int sorted, int packed, int stype, int xtype, cholmod_common* c);
int cholmod_free_sparse(cholmod_sparse** A, cholmod_common* c);
cholmod_sparse* cholmod_transpose(cholmod_sparse* A, int values, cholmod_common* c);
cholmod_sparse* cholmod_ssmult(cholmod_sparse* A, cholmod_sparse* B, int stype, int values, int sorted, cholmod_common* c);
cholmod_factor* cholmod_analyze(cholmod_sparse* A, cholmod_common* c);
int cholmod_factorize(cholmod_sparse* A, cholmod_factor* L, cholmod_common* c);
int cholmod_change_factor(int to_xtype, int to_ll, int to_super, int to_packed, int to_monotonic, cholmod_factor* L, cholmod_common* c);
int cholmod_free_factor(cholmod_factor** L, cholmod_common* c);
cholmod_sparse* cholmod_factor_to_sparse(cholmod_factor* L, cholmod_common* c);
cholmod_dense* cholmod_zeros(size_t nrow, size_t ncol, int xtype, cholmod_common* c);
int cholmod_free_dense(cholmod_dense** X, cholmod_common* c);
int cholmod_sdmult(cholmod_sparse* A, int transpose, double alpha[2], double beta[2], cholmod_dense* X, cholmod_dense* Y, cholmod_common* c);
cholmod_dense* cholmod_solve(int sys, cholmod_factor* L, cholmod_dense* B, cholmod_common* c);

#ifdef __cplusplus
}
#endif
#endif
