/* ysm_pyfast.c -- native binding of the single-query call, Wrapper.match_scan(query, base_scans, penalty, do_fine)
 * (reference yag_slam/scan_matching.py:40-42; the reference's own binding is a pybind11 module).
 *
 * It is the host glue of karto_compat.Wrapper.match_scan written against the CPython C API: it reads the scans'
 * cached point readings and content tags, keeps them in regions of a persistent staging pool (least recently
 * used region replaced), fills the ysm_batch descriptor of one match and calls ysm_match_batch through the C
 * ABI (include/ysm.h) directly. No compute happens here; without libysm_b200.so there is nothing to call.
 * On a 30 us call the interpreted glue cost 6 us; this path costs about 2.
 *
 * The Python implementation (karto_compat.Wrapper.match_scan) stays as the specification of this logic -- the
 * CPU tests drive both against a stub library -- and as the path for calls this one declines (returns None):
 * more scans than regions, a scan longer than a region, scan objects without the expected attributes. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ysm.h"

#define PF_REGIONS 48         /* = Wrapper.POOL_REGIONS */
#define PF_REGION_POINTS 4096 /* = Wrapper.REGION_POINTS */

typedef int (*match_fn)(ysm_handle*, const ysm_batch*, ysm_result*, void*);

typedef struct {
  ysm_batch b;
  int32_t starts[PF_REGIONS], counts[PF_REGIONS], raw[PF_REGIONS], bidx[PF_REGIONS];
  uint64_t tags[PF_REGIONS];
  double pose[3];
  int32_t qidx[1], bptr[2];
  int ready;
} pf_desc;

typedef struct {
  match_fn fn;
  ysm_handle* handle;
  double* pool;
  uint64_t region_tag[PF_REGIONS], used[PF_REGIONS], clock;
  pf_desc desc[PF_REGIONS]; /* by number of base scans */
  ysm_result res;
  double* resf;       /* the caller's 16-double record (a numpy array it keeps alive) */
  PyObject* covv;     /* its 3 x 3 view of the covariance block: covv.copy() is the result's matrix */
  PyObject *pose_type, *result_type;
} pf_state;

static PyObject *s_points, *s_tag, *s_point_readings, *s_corrected_pose, *s_x, *s_y, *s_yaw, *s_ranges, *s_copy,
    *s_response, *s_covariance, *s_best_pose, *s_empty;

static void pf_free(PyObject* cap) {
  pf_state* S = (pf_state*)PyCapsule_GetPointer(cap, "ysm_pyfast");
  if (!S) return;
  Py_XDECREF(S->covv);
  Py_XDECREF(S->pose_type);
  Py_XDECREF(S->result_type);
  free(S->pool);
  free(S);
}

/* create(fn_address, handle_address, record_address, covv, Pose2, MatchResult) -> state */
static PyObject* pf_create(PyObject* self, PyObject* args) {
  unsigned long long fn, handle, resf;
  PyObject *covv, *pose_type, *result_type;
  if (!PyArg_ParseTuple(args, "KKKOOO", &fn, &handle, &resf, &covv, &pose_type, &result_type)) return NULL;
  if (!fn || !resf) {
    PyErr_SetString(PyExc_ValueError, "null function or record address");
    return NULL;
  }
  pf_state* S = (pf_state*)calloc(1, sizeof(pf_state));
  if (!S) return PyErr_NoMemory();
  if (posix_memalign((void**)&S->pool, 64, (size_t)PF_REGIONS * PF_REGION_POINTS * 16)) {
    free(S);
    return PyErr_NoMemory();
  }
  memset(S->pool, 0, (size_t)PF_REGIONS * PF_REGION_POINTS * 16);
  S->fn = (match_fn)(uintptr_t)fn;
  S->handle = (ysm_handle*)(uintptr_t)handle;
  S->resf = (double*)(uintptr_t)resf;
  Py_INCREF(covv); S->covv = covv;
  Py_INCREF(pose_type); S->pose_type = pose_type;
  Py_INCREF(result_type); S->result_type = result_type;
  return PyCapsule_New(S, "ysm_pyfast", pf_free);
}

static pf_desc* pf_descriptor(pf_state* S, int nb) {
  pf_desc* d = &S->desc[nb];
  if (!d->ready) {
    memset(d, 0, sizeof(*d));
    d->bptr[0] = 0; d->bptr[1] = nb;
    for (int i = 0; i < nb; i++) d->bidx[i] = i + 1;
    d->b.n_matches = 1; d->b.n_scans = nb + 1;
    d->b.n_points = (int64_t)PF_REGIONS * PF_REGION_POINTS;
    d->b.pool_xy = S->pool;
    d->b.scan_start = d->starts; d->b.scan_count = d->counts;
    d->b.query_scan = d->qidx; d->b.query_pose = d->pose;
    d->b.base_ptr = d->bptr; d->b.base_idx = nb ? d->bidx : NULL;
    d->b.pool_on_device = 0;
    d->b.scan_tag = d->tags;
    d->b.scan_raw_count = d->raw;
    d->ready = 1;
  }
  return d;
}

static int pf_double_attr(PyObject* o, PyObject* name, double* out) {
  PyObject* v = PyObject_GetAttr(o, name);
  if (!v) return -1;
  *out = PyFloat_AsDouble(v);
  Py_DECREF(v);
  return (*out == -1.0 && PyErr_Occurred()) ? -1 : 0;
}

/* the caller falls back to the interpreted path when this returns None */
#define PF_DECLINE() do { PyErr_Clear(); Py_XDECREF(seq); Py_RETURN_NONE; } while (0)

/* match(state, query, base_scans, penalty, do_fine) -> MatchResult | None (declined) | int (ysm error code) */
static PyObject* pf_match(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  if (nargs != 5) {
    PyErr_SetString(PyExc_TypeError, "match(state, query, base_scans, penalty, do_fine)");
    return NULL;
  }
  pf_state* S = (pf_state*)PyCapsule_GetPointer(args[0], "ysm_pyfast");
  if (!S) return NULL;
  PyObject* query = args[1];
  PyObject* seq = PySequence_Fast(args[2], "base_scans must be a sequence");
  if (!seq) return NULL;
  const Py_ssize_t nb = PySequence_Fast_GET_SIZE(seq);
  if (nb + 1 > PF_REGIONS) PF_DECLINE();
  const int penalty = PyObject_IsTrue(args[3]), fine = PyObject_IsTrue(args[4]);
  if (penalty < 0 || fine < 0) { Py_DECREF(seq); return NULL; }
  pf_desc* d = pf_descriptor(S, (int)nb);
  const uint64_t clock = ++S->clock;
  for (Py_ssize_t i = 0; i <= nb; i++) {
    PyObject* sc = i == 0 ? query : PySequence_Fast_GET_ITEM(seq, i - 1);
    PyObject* pts = PyObject_GetAttr(sc, s_points);
    if (!pts) PF_DECLINE();
    if (pts == Py_None) {  /* LocalizedRangeScan::Update: the scan refreshes its readings (and its content tag) */
      Py_DECREF(pts);
      pts = PyObject_CallMethodNoArgs(sc, s_point_readings);
      if (!pts) { Py_DECREF(seq); return NULL; }
    }
    PyObject* tago = PyObject_GetAttr(sc, s_tag);
    if (!tago) { Py_DECREF(pts); PF_DECLINE(); }
    const uint64_t tag = PyLong_AsUnsignedLongLong(tago);
    Py_DECREF(tago);
    if (PyErr_Occurred() || tag == 0) { Py_DECREF(pts); PF_DECLINE(); }
    const Py_ssize_t n = PyObject_Length(pts);
    if (n < 0 || n > PF_REGION_POINTS) { Py_DECREF(pts); PF_DECLINE(); }
    int r = -1;
    for (int k = 0; k < PF_REGIONS; k++)
      if (S->region_tag[k] == tag) { r = k; break; }
    if (r < 0) {
      /* a region no scan of this call sits in, least recently used first */
      for (int k = 0; k < PF_REGIONS; k++)
        if (S->used[k] != clock && (r < 0 || S->used[k] < S->used[r])) r = k;
      if (r < 0) { Py_DECREF(pts); PF_DECLINE(); }
      if (n > 0) {
        Py_buffer view;
        if (PyObject_GetBuffer(pts, &view, PyBUF_C_CONTIGUOUS) != 0) { Py_DECREF(pts); PF_DECLINE(); }
        if (view.len != n * 16) { PyBuffer_Release(&view); Py_DECREF(pts); PF_DECLINE(); }  /* float64 [n][2] */
        memcpy(S->pool + (size_t)r * PF_REGION_POINTS * 2, view.buf, (size_t)view.len);
        PyBuffer_Release(&view);
      }
      S->region_tag[r] = tag;
    }
    Py_DECREF(pts);
    S->used[r] = clock;
    d->starts[i] = r * PF_REGION_POINTS;
    d->counts[i] = (int32_t)n;
    d->tags[i] = tag;
  }
  {
    PyObject* p = PyObject_GetAttr(query, s_corrected_pose);
    if (!p) PF_DECLINE();
    const int bad = pf_double_attr(p, s_x, &d->pose[0]) || pf_double_attr(p, s_y, &d->pose[1]) ||
                    pf_double_attr(p, s_yaw, &d->pose[2]);
    Py_DECREF(p);
    if (bad) PF_DECLINE();
    PyObject* ranges = PyObject_GetAttr(query, s_ranges);
    if (!ranges) PF_DECLINE();
    const Py_ssize_t nraw = PyObject_Length(ranges);  /* (Karto tests the RAW reading count for its early return) */
    Py_DECREF(ranges);
    if (nraw < 0) PF_DECLINE();
    d->raw[0] = (int32_t)nraw;
  }
  Py_DECREF(seq);
  seq = NULL;
  d->b.do_penalize = penalty;
  d->b.do_refine = fine;
  int rc;
  Py_BEGIN_ALLOW_THREADS
  rc = S->fn(S->handle, &d->b, &S->res, NULL);
  Py_END_ALLOW_THREADS
  if (rc != 0) return PyLong_FromLong(rc);
  memcpy(S->resf, &S->res, 16 * sizeof(double));
  /* MatchResult(response, covariance 3 x 3, Pose2(x, y, yaw)), built attribute by attribute */
  PyObject *out = NULL, *pose = NULL, *cov = NULL, *v = NULL;
  pose = ((PyTypeObject*)S->pose_type)->tp_new((PyTypeObject*)S->pose_type, s_empty, NULL);
  if (!pose) goto fail;
  if (!(v = PyFloat_FromDouble(S->resf[1])) || PyObject_SetAttr(pose, s_x, v) < 0) goto fail;
  Py_CLEAR(v);
  if (!(v = PyFloat_FromDouble(S->resf[2])) || PyObject_SetAttr(pose, s_y, v) < 0) goto fail;
  Py_CLEAR(v);
  if (!(v = PyFloat_FromDouble(S->resf[3])) || PyObject_SetAttr(pose, s_yaw, v) < 0) goto fail;
  Py_CLEAR(v);
  cov = PyObject_CallMethodNoArgs(S->covv, s_copy);
  if (!cov) goto fail;
  out = ((PyTypeObject*)S->result_type)->tp_new((PyTypeObject*)S->result_type, s_empty, NULL);
  if (!out) goto fail;
  if (!(v = PyFloat_FromDouble(S->resf[0])) || PyObject_SetAttr(out, s_response, v) < 0) goto fail;
  Py_CLEAR(v);
  if (PyObject_SetAttr(out, s_covariance, cov) < 0 || PyObject_SetAttr(out, s_best_pose, pose) < 0) goto fail;
  Py_DECREF(cov);
  Py_DECREF(pose);
  return out;
fail:
  Py_XDECREF(v);
  Py_XDECREF(cov);
  Py_XDECREF(pose);
  Py_XDECREF(out);
  return NULL;
}

static PyMethodDef pf_methods[] = {
    {"create", pf_create, METH_VARARGS, "create(fn, handle, record, covv, Pose2, MatchResult) -> state"},
    {"match", (PyCFunction)(void (*)(void))pf_match, METH_FASTCALL,
     "match(state, query, base_scans, penalty, do_fine) -> MatchResult | None (declined) | int (error code)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef pf_module = {PyModuleDef_HEAD_INIT, "_ysm_pyfast",
                                       "native binding of Wrapper.match_scan over the ysm C ABI", -1, pf_methods};

PyMODINIT_FUNC PyInit__ysm_pyfast(void) {
#define PF_INTERN(var, text) if (!(var = PyUnicode_InternFromString(text))) return NULL
  PF_INTERN(s_points, "_points");
  PF_INTERN(s_tag, "_tag");
  PF_INTERN(s_point_readings, "point_readings");
  PF_INTERN(s_corrected_pose, "_corrected_pose");
  PF_INTERN(s_x, "x");
  PF_INTERN(s_y, "y");
  PF_INTERN(s_yaw, "yaw");
  PF_INTERN(s_ranges, "ranges");
  PF_INTERN(s_copy, "copy");
  PF_INTERN(s_response, "response");
  PF_INTERN(s_covariance, "covariance");
  PF_INTERN(s_best_pose, "best_pose");
  if (!(s_empty = PyTuple_New(0))) return NULL;
  return PyModule_Create(&pf_module);
}
