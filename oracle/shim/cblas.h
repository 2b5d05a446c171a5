// Prototypes for the CBLAS entry points the reference's math wrappers call
// (no system cblas.h in this image; the symbols come from the OpenBLAS that
// ships inside the venv).  Test infrastructure only.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_sgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, float,
                 const float*, int, const float*, int, float, float*, int);
void cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, double,
                 const double*, int, const double*, int, double, double*, int);
void cblas_sgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, float, const float*, int,
                 const float*, int, float, float*, int);
void cblas_dgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int, int, double, const double*, int,
                 const double*, int, double, double*, int);
void cblas_sger(CBLAS_ORDER, int, int, float, const float*, int, const float*, int, float*, int);
void cblas_dger(CBLAS_ORDER, int, int, double, const double*, int, const double*, int, double*, int);
void cblas_saxpy(int, float, const float*, int, float*, int);
void cblas_daxpy(int, double, const double*, int, double*, int);
void cblas_sscal(int, float, float*, int);
void cblas_dscal(int, double, double*, int);
void cblas_saxpby(int, float, const float*, int, float, float*, int);
void cblas_daxpby(int, double, const double*, int, double, double*, int);
float cblas_sdot(int, const float*, int, const float*, int);
double cblas_ddot(int, const double*, int, const double*, int);
float cblas_sasum(int, const float*, int);
double cblas_dasum(int, const double*, int);
void cblas_scopy(int, const float*, int, float*, int);
void cblas_dcopy(int, const double*, int, double*, int);
void openblas_set_num_threads(int);
int openblas_get_num_threads(void);
#ifdef __cplusplus
}
#endif
