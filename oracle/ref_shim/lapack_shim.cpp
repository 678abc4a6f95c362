// Test infrastructure only.  The reference's libtoast/src/toast_math_linearalgebra.cpp calls
// Fortran BLAS / LAPACK by name (dgemm_, dsyev_, ... with -DLAPACK_NAMES_UBACK).  This image has no
// system LAPACK, but scipy ships OpenBLAS with every symbol prefixed `scipy_`; these one-line
// forwarders let the reference's OWN cov_eigendecompose_diag (toast_map_cov.cpp:246-396) run
// here, so that the oracle's numpy restatement of it can be pinned.  Signatures are those the
// reference declares (toast_math_linearalgebra.cpp:25-27,157-158,260-262,363-365,440-443,
// 523-525); hidden Fortran string lengths are passed explicitly.
#include <cstddef>

extern "C" {
void scipy_dgemm_(char *, char *, int *, int *, int *, double *, double *, int *, double *,
                  int *, double *, double *, int *, size_t, size_t);
void scipy_dsyev_(char *, char *, int *, double *, int *, double *, double *, int *, int *,
                  size_t, size_t);
void scipy_dsymm_(char *, char *, int *, int *, double *, double *, int *, double *, int *,
                  double *, double *, int *, size_t, size_t);
void scipy_dsyrk_(char *, char *, int *, int *, double *, double *, int *, double *, double *,
                  int *, size_t, size_t);
void scipy_dgels_(char *, int *, int *, int *, double *, int *, double *, int *, double *,
                  int *, int *, size_t);
void scipy_dgelss_(int *, int *, int *, double *, int *, double *, int *, double *, double *,
                   int *, double *, int *, int *);

void dgemm_(char *ta, char *tb, int *m, int *n, int *k, double *al, double *a, int *lda,
            double *b, int *ldb, double *be, double *c, int *ldc) {
    scipy_dgemm_(ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc, 1, 1);
}
void dsyev_(char *jobz, char *uplo, int *n, double *a, int *lda, double *w, double *work,
            int *lwork, int *info) {
    scipy_dsyev_(jobz, uplo, n, a, lda, w, work, lwork, info, 1, 1);
}
void dsymm_(char *side, char *uplo, int *m, int *n, double *al, double *a, int *lda, double *b,
            int *ldb, double *be, double *c, int *ldc) {
    scipy_dsymm_(side, uplo, m, n, al, a, lda, b, ldb, be, c, ldc, 1, 1);
}
void dsyrk_(char *uplo, char *trans, int *n, int *k, double *al, double *a, int *lda,
            double *be, double *c, int *ldc) {
    scipy_dsyrk_(uplo, trans, n, k, al, a, lda, be, c, ldc, 1, 1);
}
// (the reference declares TRANS by value for this one)
void dgels_(char trans, int *m, int *n, int *nrhs, double *a, int *lda, double *b, int *ldb,
            double *work, int *lwork, int *info) {
    scipy_dgels_(&trans, m, n, nrhs, a, lda, b, ldb, work, lwork, info, 1);
}
void dgelss_(int *m, int *n, int *nrhs, double *a, int *lda, double *b, int *ldb, double *s,
             double *rcond, int *rank, double *work, int *lwork, int *info) {
    scipy_dgelss_(m, n, nrhs, a, lda, b, ldb, s, rcond, rank, work, lwork, info);
}
}
