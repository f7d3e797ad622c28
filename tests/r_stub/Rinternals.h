#ifndef R_STUB_INTERNALS_H
#define R_STUB_INTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
enum { INTSXP = 13, REALSXP = 14, VECSXP = 19, RAWSXP = 24 };
extern SEXP R_NilValue;
extern double R_NaReal;
#define NA_REAL R_NaReal
#define ISNAN(x) ((x) != (x))
typedef void (*R_CFinalizer_t)(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...) __attribute__((noreturn));
SEXP Rf_install(const char*);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
int Rf_isNull(SEXP);
int Rf_length(SEXP);
R_xlen_t XLENGTH(SEXP);
int TYPEOF(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
int* INTEGER(SEXP);
double* REAL(SEXP);
unsigned char* RAW(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_mkNamed(unsigned int, const char**);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarLogical(int);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
#endif
