// A minimal WORKING host for the MEX gateways: implements the part of the MATLAB / Octave MEX C API that mex_shim.h
// declares (heap mxArray with numeric / logical / char / struct / cell classes, mexCallMATLAB("rand", ...),
// mexErrMsgIdAndTxt as a C++ exception, mexAtExit) so that the gateways can be LINKED and EXECUTED without an
// interpreter.  tests/mexhost.py drives it through ctypes:  build mxArrays from NumPy, call a gateway's mexFunction
// through mexhost_call, read the outputs back.  It is test infrastructure (like oracle/): nothing in the product links it.
//
// rand:  rand('seed', s) switches to the Park-Miller "minimal standard" generator with state s (the stand-in of
//        tests/golden/ref_shadow/rand.m and oracle.snmf_oracle.park_miller); before any seed, rand(m, n) replays the stream
//        set with mexhost_set_rand_stream (init_buff.m:37-38 draws).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "mex_shim.h"

struct mxArray_tag {
  mxClassID cls = mxDOUBLE_CLASS;
  std::vector<mwSize> dims{0, 0};
  std::vector<unsigned char> data;        // numeric / logical / char (1 byte per char) payload, column-major
  std::vector<std::string> fields;        // struct: field names
  std::vector<mxArray*> kids;             // struct: fields x elements (field-major per element); cell: elements
};

namespace {
struct MexError : std::runtime_error {
  std::string id;
  MexError(const std::string& i, const std::string& m) : std::runtime_error(m), id(i) {}
};
size_t elsize(mxClassID c) {
  switch (c) {
    case mxDOUBLE_CLASS: case mxUINT64_CLASS: return 8;
    case mxINT16_CLASS: return 2;
    case mxLOGICAL_CLASS: case mxCHAR_CLASS: return 1;
    default: return 0;
  }
}
size_t numel(const mxArray* a) {
  size_t n = 1;
  for (mwSize d : a->dims) n *= d;
  return n;
}
long long g_live = 0;
mxArray* make(mxClassID c, const std::vector<mwSize>& dims) {
  mxArray* a = new mxArray_tag();
  a->cls = c;
  a->dims = dims;
  a->data.assign(numel(a) * elsize(c), 0);
  ++g_live;
  return a;
}
// rand state
bool g_seeded = false;
double g_state = 1.0;
std::vector<double> g_stream;
size_t g_pos = 0;
std::vector<void (*)(void)> g_atexit;
}  // namespace

extern "C" {

double* mxGetPr(const mxArray* a) { return a && a->cls == mxDOUBLE_CLASS ? (double*)a->data.data() : nullptr; }
void* mxGetData(const mxArray* a) { return a ? (void*)a->data.data() : nullptr; }
double mxGetScalar(const mxArray* a) {
  if (!a || numel(a) == 0) throw MexError("host:scalar", "mxGetScalar of an empty array");
  switch (a->cls) {
    case mxDOUBLE_CLASS: return *(const double*)a->data.data();
    case mxUINT64_CLASS: return (double)*(const uint64_t*)a->data.data();
    case mxINT16_CLASS: return (double)*(const int16_t*)a->data.data();
    case mxLOGICAL_CLASS: case mxCHAR_CLASS: return (double)a->data[0];
    default: throw MexError("host:scalar", "mxGetScalar of a non-numeric array");
  }
}
size_t mxGetM(const mxArray* a) { return a->dims.empty() ? 0 : a->dims[0]; }
size_t mxGetN(const mxArray* a) {
  size_t n = 1;
  for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i];
  return a->dims.size() < 2 ? 1 : n;
}
size_t mxGetNumberOfElements(const mxArray* a) { return numel(a); }
size_t mxGetNumberOfDimensions(const mxArray* a) { return a->dims.size(); }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
bool mxIsDouble(const mxArray* a) { return a && a->cls == mxDOUBLE_CLASS; }
bool mxIsLogical(const mxArray* a) { return a && a->cls == mxLOGICAL_CLASS; }
bool mxIsChar(const mxArray* a) { return a && a->cls == mxCHAR_CLASS; }
bool mxIsStruct(const mxArray* a) { return a && a->cls == mxSTRUCT_CLASS; }
bool mxIsCell(const mxArray* a) { return a && a->cls == mxCELL_CLASS; }
bool mxIsEmpty(const mxArray* a) { return !a || numel(a) == 0; }
mxLogical* mxGetLogicals(const mxArray* a) { return a && a->cls == mxLOGICAL_CLASS ? (mxLogical*)a->data.data() : nullptr; }

int mxGetFieldNumber(const mxArray* a, const char* name) {
  if (!mxIsStruct(a)) return -1;
  for (size_t i = 0; i < a->fields.size(); ++i)
    if (a->fields[i] == name) return (int)i;
  return -1;
}
mxArray* mxGetField(const mxArray* a, mwIndex idx, const char* name) {
  const int f = mxGetFieldNumber(a, name);
  if (f < 0 || idx >= numel(a)) return nullptr;
  return a->kids[idx * a->fields.size() + (size_t)f];
}
mxArray* mxGetCell(const mxArray* a, mwIndex idx) { return mxIsCell(a) && idx < a->kids.size() ? a->kids[idx] : nullptr; }
int mxAddField(mxArray* a, const char* name) {
  if (!mxIsStruct(a)) return -1;
  const size_t nf = a->fields.size(), ne = numel(a);
  std::vector<mxArray*> k((nf + 1) * ne, nullptr);
  for (size_t e = 0; e < ne; ++e)
    for (size_t f = 0; f < nf; ++f) k[e * (nf + 1) + f] = a->kids[e * nf + f];
  a->kids.swap(k);
  a->fields.push_back(name);
  return (int)nf;
}
void mxDestroyArray(mxArray* a) {
  if (!a) return;
  for (mxArray* k : a->kids) mxDestroyArray(k);
  delete a;
  --g_live;
}
void mxSetField(mxArray* a, mwIndex idx, const char* name, mxArray* v) {
  const int f = mxGetFieldNumber(a, name);
  if (f < 0 || idx >= numel(a)) throw MexError("host:field", std::string("mxSetField: no field ") + name);
  mxArray*& slot = a->kids[idx * a->fields.size() + (size_t)f];
  if (slot && slot != v) mxDestroyArray(slot);   // MATLAB leaks the old value here; a test host may as well free it
  slot = v;
}
void mxSetCell(mxArray* a, mwIndex idx, mxArray* v) {
  if (!mxIsCell(a) || idx >= a->kids.size()) throw MexError("host:cell", "mxSetCell: bad index");
  if (a->kids[idx] && a->kids[idx] != v) mxDestroyArray(a->kids[idx]);
  a->kids[idx] = v;
}
mxArray* mxDuplicateArray(const mxArray* a) {
  if (!a) return nullptr;
  mxArray* c = new mxArray_tag(*a);
  ++g_live;
  for (mxArray*& k : c->kids) k = mxDuplicateArray(k);
  return c;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity) { return make(mxDOUBLE_CLASS, {m, n}); }
mxArray* mxCreateDoubleScalar(double v) {
  mxArray* a = make(mxDOUBLE_CLASS, {1, 1});
  *(double*)a->data.data() = v;
  return a;
}
mxArray* mxCreateNumericArray(mwSize nd, const mwSize* dims, mxClassID c, mxComplexity) {
  return make(c, std::vector<mwSize>(dims, dims + nd));
}
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID c, mxComplexity) { return make(c, {m, n}); }
mxArray* mxCreateLogicalMatrix(mwSize m, mwSize n) { return make(mxLOGICAL_CLASS, {m, n}); }
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nf, const char** names) {
  mxArray* a = make(mxSTRUCT_CLASS, {m, n});
  for (int i = 0; i < nf; ++i) a->fields.push_back(names[i]);
  a->kids.assign((size_t)nf * m * n, nullptr);
  return a;
}
mxArray* mxCreateCellMatrix(mwSize m, mwSize n) {
  mxArray* a = make(mxCELL_CLASS, {m, n});
  a->kids.assign(m * n, nullptr);
  return a;
}
mxArray* mxCreateString(const char* s) {
  const size_t n = std::strlen(s);
  mxArray* a = make(mxCHAR_CLASS, {1, n});
  std::memcpy(a->data.data(), s, n);
  return a;
}
char* mxArrayToString(const mxArray* a) {
  if (!mxIsChar(a)) return nullptr;
  const size_t n = numel(a);
  char* s = (char*)std::malloc(n + 1);
  std::memcpy(s, a->data.data(), n);
  s[n] = 0;
  return s;
}
void mxFree(void* p) { std::free(p); }

void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw MexError(id ? id : "", buf);
}
int mexPrintf(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  const int n = std::vfprintf(stdout, fmt, ap);
  va_end(ap);
  return n;
}
void mexLock(void) {}
int mexAtExit(void (*fn)(void)) {
  g_atexit.push_back(fn);
  return 0;
}

int mexCallMATLAB(int nlhs, mxArray** plhs, int nrhs, mxArray** prhs, const char* name) {
  if (std::strcmp(name, "rand") != 0) throw MexError("host:call", std::string("mexCallMATLAB: no host function ") + name);
  if (nrhs >= 1 && mxIsChar(prhs[0])) {
    char* s = mxArrayToString(prhs[0]);
    const bool seed = std::strcmp(s, "seed") == 0 && nrhs == 2;
    mxFree(s);
    if (!seed) throw MexError("host:rand", "rand: only rand('seed', s) is supported");
    g_seeded = true;
    g_state = mxGetScalar(prhs[1]);
    return 0;
  }
  mwSize m = 1, n = 1;
  if (nrhs == 1) m = n = (mwSize)mxGetScalar(prhs[0]);
  if (nrhs >= 2) {
    m = (mwSize)mxGetScalar(prhs[0]);
    n = (mwSize)mxGetScalar(prhs[1]);
  }
  mxArray* out = mxCreateDoubleMatrix(m, n, mxREAL);
  double* o = mxGetPr(out);
  for (size_t i = 0; i < (size_t)m * n; ++i) {
    if (g_seeded) {
      g_state = std::fmod(16807.0 * g_state, 2147483647.0);
      o[i] = g_state / 2147483647.0;
    } else {
      if (g_pos >= g_stream.size()) {
        mxDestroyArray(out);
        throw MexError("host:rand", "rand: the stream set with mexhost_set_rand_stream is exhausted");
      }
      o[i] = g_stream[g_pos++];
    }
  }
  if (nlhs >= 1 && plhs) plhs[0] = out; else mxDestroyArray(out);
  return 0;
}

// ---------------------------------------------------------------- host control (used by tests/mexhost.py)
typedef void (*mexhost_fn)(int, mxArray**, int, const mxArray**);
// 0 = ok; 1 = the gateway raised mexErrMsgIdAndTxt (message in err); 2 = another C++ exception
int mexhost_call(mexhost_fn fn, int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs, char* err, int errlen) {
  try {
    fn(nlhs, plhs, nrhs, prhs);
    return 0;
  } catch (const MexError& e) {
    if (err && errlen > 0) std::snprintf(err, (size_t)errlen, "%s: %s", e.id.c_str(), e.what());
    return 1;
  } catch (const std::exception& e) {
    if (err && errlen > 0) std::snprintf(err, (size_t)errlen, "%s", e.what());
    return 2;
  }
}
void mexhost_set_rand_stream(const double* v, size_t n) {
  g_stream.assign(v, v + n);
  g_pos = 0;
  g_seeded = false;
}
void mexhost_shutdown(void) {
  for (auto it = g_atexit.rbegin(); it != g_atexit.rend(); ++it) (*it)();
  g_atexit.clear();
}
long long mexhost_live_arrays(void) { return g_live; }
int mexhost_class(const mxArray* a) { return (int)a->cls; }
int mexhost_num_fields(const mxArray* a) { return (int)a->fields.size(); }
const char* mexhost_field_name(const mxArray* a, int i) { return a->fields[(size_t)i].c_str(); }

}  // extern "C"
