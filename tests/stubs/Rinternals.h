/* Stub of <Rinternals.h>: the part of the R C API that r/src/ruvnb_shim.c uses, with R's own prototypes (R 4.x), so that the
 * shim can be compile-checked where R is not installed.  Test infrastructure only. */
#ifndef STUB_RINTERNALS_H
#define STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue, R_NamesSymbol;
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...) __attribute__((noreturn));
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_mkNamed(SEXPTYPE, const char**);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_install(const char*);
SEXP Rf_ScalarReal(double);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_length(SEXP);
R_xlen_t Rf_xlength(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isMatrix(SEXP);
Rboolean Rf_isReal(SEXP);
SEXP Rf_lang2(SEXP, SEXP);
SEXP R_tryEval(SEXP, SEXP, int*);
extern SEXP R_GlobalEnv;
extern double R_NaReal;
#define NA_REAL R_NaReal
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
double* REAL(SEXP);
const char* CHAR(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
