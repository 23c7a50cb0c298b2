/* Minimal declarations of the MATLAB / Octave MEX C API for a container that has neither MATLAB nor Octave (compile with
 * -DSNMFNAT_MEX_SHIM).  mex_host.cpp IMPLEMENTS them (a small working host: `make host`), so the gateways are linked and
 * executed by tests/test_mex_host.py.  With a real toolchain (`mex` or `mkoctfile --mex`) the real <mex.h> is used instead. */
#ifndef SNMFNAT_MEX_SHIM_H_
#define SNMFNAT_MEX_SHIM_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef bool mxLogical;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxUNKNOWN_CLASS, mxDOUBLE_CLASS, mxUINT64_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxSTRUCT_CLASS,
               mxCELL_CLASS, mxINT16_CLASS } mxClassID;
double* mxGetPr(const mxArray*);
void* mxGetData(const mxArray*);
double mxGetScalar(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
size_t mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
bool mxIsDouble(const mxArray*);
bool mxIsLogical(const mxArray*);
bool mxIsChar(const mxArray*);
bool mxIsStruct(const mxArray*);
bool mxIsCell(const mxArray*);
bool mxIsEmpty(const mxArray*);
mxLogical* mxGetLogicals(const mxArray*);
mxArray* mxGetField(const mxArray*, mwIndex, const char*);
mxArray* mxGetCell(const mxArray*, mwIndex);
int mxGetFieldNumber(const mxArray*, const char*);
int mxAddField(mxArray*, const char*);
void mxSetField(mxArray*, mwIndex, const char*, mxArray*);
mxArray* mxDuplicateArray(const mxArray*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
mxArray* mxCreateString(const char*);
mxArray* mxCreateLogicalMatrix(mwSize, mwSize);
mxArray* mxCreateCellMatrix(mwSize, mwSize);
void mxSetCell(mxArray*, mwIndex, mxArray*);
char* mxArrayToString(const mxArray*);
void mxFree(void*);
void mxDestroyArray(mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexCallMATLAB(int, mxArray**, int, mxArray**, const char*);
void mexLock(void);
int mexAtExit(void (*)(void));
int mexPrintf(const char*, ...);
/* the gateway entry point has C linkage, as in the real <mex.h> */
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
