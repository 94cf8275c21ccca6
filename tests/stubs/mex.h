/* Declaration-only stand-in for MATLAB's mex.h / matrix.h — NOT MathWorks' header and not usable to build a MEX file.
 * It declares just the part of the documented C Matrix / MEX API (R2018a interleaved-complex API names) that the
 * gateways under mex/ call, with the documented signatures, so that tests/test_mex_sources.py can type-check those
 * sources against include/vbmc_b200.h in an image that has no MATLAB.  Nothing here has a definition. */
#ifndef VBMC_B200_STUB_MEX_H
#define VBMC_B200_STUB_MEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef bool mxLogical;
typedef enum { mxUNKNOWN_CLASS = 0, mxLOGICAL_CLASS = 3, mxDOUBLE_CLASS = 6 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
void mexWarnMsgIdAndTxt(const char* id, const char* fmt, ...);
int mexAtExit(void (*fn)(void));
void mexLock(void);

bool mxIsEmpty(const mxArray* a);
bool mxIsStruct(const mxArray* a);
bool mxIsFinite(double v);
size_t mxGetM(const mxArray* a);
size_t mxGetN(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
double mxGetScalar(const mxArray* a);
double* mxGetDoubles(const mxArray* a);
void* mxGetData(const mxArray* a);
mxArray* mxGetField(const mxArray* s, mwIndex i, const char* name);
mxArray* mxGetFieldByNumber(const mxArray* s, mwIndex i, int field);
void mxSetFieldByNumber(mxArray* s, mwIndex i, int field, mxArray* value);
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateLogicalScalar(bool v);
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity flag);
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names);
#ifdef __cplusplus
}
#endif
#endif
