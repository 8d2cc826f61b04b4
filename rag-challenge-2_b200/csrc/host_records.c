/* _b2r_records — CPython helper that turns the rows a batched search returns into the reference's record
 * dictionaries (src/retrieval.py:1142-1157 hit construction, :1552-1599 aggregation, :1639-1716 fusion, for the
 * case every aggregation key holds ONE hit: one query text, basic search, distinct ids).  Host-side only: no CUDA,
 * no arithmetic of the search itself.  retrieval.py keeps the same loops in Python (`_RawHits.records`) and
 * tests/test_host_records.py holds the two equal, record for record and exception for exception.
 *
 * Why it exists: BASELINE config 2 returns 127 x 6 x 30 = 22,860 dictionaries per question batch; built from
 * Python bytecode they cost ~10x the device time of the searches.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

static PyObject *k_vector_similarity, *k_max_original_similarity, *k_page, *k_text, *k_hit_count,
    *k_retrieval_sources, *k_source_sha1, *k_query_sources, *k_basic, *g_round, *g_four, *g_one;

/* round(x, 4) exactly as float.__round__ computes it: the decimal k * 10^-4 nearest to the EXACT value of x
 * (ties to even), converted back to the nearest double.  k = nearbyint(x * 1e4) equals that k unless the exact
 * product lies within the rounding error of the multiplication (< 1.2e-7 for |x * 1e4| < 1e9) of a half-integer;
 * k / 1e4 is a correctly rounded division of two exactly representable numbers, i.e. the nearest double of the
 * decimal, which is what strtod returns.  Anything near a tie, huge or non-finite goes to the builtin. */
static PyObject *round4(double x) {
    if (isfinite(x)) {
        double t = x * 1e4;
        if (fabs(t) < 1e9) {
            double k = nearbyint(t);
            if (fabs(fabs(t - k) - 0.5) > 1e-5) return PyFloat_FromDouble(k / 1e4);
        }
    }
    PyObject *f = PyFloat_FromDouble(x);
    if (!f) return NULL;
    PyObject *r = PyObject_CallFunctionObjArgs(g_round, f, g_four, NULL);
    Py_DECREF(f);
    return r;
}

static PyObject *py_round4(PyObject *self, PyObject *arg) {
    double x = PyFloat_AsDouble(arg);
    if (x == -1.0 && PyErr_Occurred()) return NULL;
    return round4(x);
}

/* obj[key] with the exceptions of the Python expression (KeyError(key) for a dict that lacks it) */
static PyObject *getitem(PyObject *obj, PyObject *key) {
    if (PyDict_CheckExact(obj)) {
        PyObject *v = PyDict_GetItemWithError(obj, key);
        if (v) {
            Py_INCREF(v);
            return v;
        }
        if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, key);
        return NULL;
    }
    return PyObject_GetItem(obj, key);
}

static PyObject *seq_item(PyObject *seq, int64_t index) {
    if (PyList_CheckExact(seq)) {
        Py_ssize_t n = PyList_GET_SIZE(seq), i = (Py_ssize_t)index;
        if (i < 0) i += n;
        if (i < 0 || i >= n) {
            PyErr_SetString(PyExc_IndexError, "list index out of range");
            return NULL;
        }
        PyObject *v = PyList_GET_ITEM(seq, i);
        Py_INCREF(v);
        return v;
    }
    PyObject *idx = PyLong_FromLongLong(index);
    if (!idx) return NULL;
    PyObject *v = PyObject_GetItem(seq, idx);
    Py_DECREF(idx);
    return v;
}

static PyObject *one_list(PyObject *item) {
    PyObject *l = PyList_New(1);
    if (!l) return NULL;
    Py_INCREF(item);
    PyList_SET_ITEM(l, 0, item);
    return l;
}

/* Template of the records of one (report, question) row: every key in the reference's order, the two per-row
 * constants filled in.  A record is a copy of it (one table clone instead of eight insertions) with the six
 * per-hit values replaced. */
static PyObject *new_template(PyObject *sha1) {
    PyObject *t = PyDict_New();
    if (!t) return NULL;
    if (PyDict_SetItem(t, k_vector_similarity, Py_None) < 0 || PyDict_SetItem(t, k_max_original_similarity, Py_None) < 0 ||
        PyDict_SetItem(t, k_page, Py_None) < 0 || PyDict_SetItem(t, k_text, Py_None) < 0 ||
        PyDict_SetItem(t, k_hit_count, g_one) < 0 || PyDict_SetItem(t, k_retrieval_sources, Py_None) < 0 ||
        PyDict_SetItem(t, k_source_sha1, sha1) < 0 || PyDict_SetItem(t, k_query_sources, Py_None) < 0) {
        Py_DECREF(t);
        return NULL;
    }
    return t;
}

/* the record of a key that one basic hit of one query reached: both bonuses are exactly 1.0 (reference :1639-1661) */
static PyObject *new_record(PyObject *tmpl, PyObject *r, PyObject *page, PyObject *text, PyObject *query_text) {
    PyObject *rec = PyDict_Copy(tmpl);
    PyObject *sources = one_list(k_basic), *queries = one_list(query_text);
    if (!rec || !sources || !queries) goto fail;
    if (PyDict_SetItem(rec, k_vector_similarity, r) < 0 || PyDict_SetItem(rec, k_max_original_similarity, r) < 0 ||
        PyDict_SetItem(rec, k_page, page) < 0 || PyDict_SetItem(rec, k_text, text) < 0 ||
        PyDict_SetItem(rec, k_retrieval_sources, sources) < 0 || PyDict_SetItem(rec, k_query_sources, queries) < 0)
        goto fail;
    Py_DECREF(sources);
    Py_DECREF(queries);
    return rec;
fail:
    Py_XDECREF(rec);
    Py_XDECREF(sources);
    Py_XDECREF(queries);
    return NULL;
}

typedef struct {
    Py_buffer d, i;
    const float *scores;
    const int64_t *ids;
    Py_ssize_t k;
} Rows;

static int rows_open(Rows *rw, PyObject *D, PyObject *I, Py_ssize_t row) {
    memset(rw, 0, sizeof(*rw));
    if (PyObject_GetBuffer(D, &rw->d, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) return -1;
    if (PyObject_GetBuffer(I, &rw->i, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) {
        PyBuffer_Release(&rw->d);
        return -1;
    }
    int ok = rw->d.ndim == 2 && rw->i.ndim == 2 && rw->d.itemsize == 4 && rw->i.itemsize == 8 && rw->d.format &&
             rw->i.format && strcmp(rw->d.format, "f") == 0 &&
             (strcmp(rw->i.format, "l") == 0 || strcmp(rw->i.format, "q") == 0) &&
             rw->d.shape[0] == rw->i.shape[0] && rw->d.shape[1] == rw->i.shape[1] && row >= 0 && row < rw->d.shape[0];
    if (!ok) {
        PyBuffer_Release(&rw->d);
        PyBuffer_Release(&rw->i);
        PyErr_SetString(PyExc_TypeError, "expected D float32 [nq, k], I int64 [nq, k] (C-contiguous) and a valid row");
        return -1;
    }
    rw->k = rw->d.shape[1];
    rw->scores = (const float *)rw->d.buf + row * rw->k;
    rw->ids = (const int64_t *)rw->i.buf + row * rw->k;
    return 0;
}

static void rows_close(Rows *rw) {
    PyBuffer_Release(&rw->d);
    PyBuffer_Release(&rw->i);
}

/* A report's chunk list holds millions of small dictionaries (6.6M for BASELINE config 2): every hit is a chain of
 * dependent cache misses — list slot, dictionary object, its key table, the page and text objects — and built one
 * hit after the other the misses serialise (~2 us per record measured, twice the cost of the allocations).  The
 * hits of a row are independent, so walk the chain level by level for the whole row first, prefetching: the
 * misses of one level overlap.  Plain lists of plain dicts only (what json.load produces); prefetches never fault
 * and change nothing, every real access below still goes through the object protocol. */
#define PREFETCH_MAX 256
static void prefetch_hits(PyObject *chunks, const int64_t *ids, Py_ssize_t k, PyObject *pages, signed char *has) {
    for (Py_ssize_t j = 0; j < k && j < PREFETCH_MAX; ++j) has[j] = -1;   /* -1: not looked up yet */
    if (!PyList_CheckExact(chunks)) return;
    if (k > PREFETCH_MAX) k = PREFETCH_MAX;
    const Py_ssize_t n = PyList_GET_SIZE(chunks);
    PyObject **items = ((PyListObject *)chunks)->ob_item;
    PyObject *c[PREFETCH_MAX];
    for (Py_ssize_t j = 0; j < k; ++j)
        if (ids[j] >= 0 && ids[j] < n) __builtin_prefetch(items + ids[j]);
    for (Py_ssize_t j = 0; j < k; ++j) {
        c[j] = (ids[j] >= 0 && ids[j] < n) ? items[ids[j]] : NULL;
        if (c[j]) __builtin_prefetch(c[j]);
    }
    for (Py_ssize_t j = 0; j < k; ++j) {
        if (c[j] && PyDict_CheckExact(c[j])) {
            const char *keys = (const char *)((PyDictObject *)c[j])->ma_keys;
            __builtin_prefetch(keys);
            __builtin_prefetch(keys + 64);
            __builtin_prefetch(keys + 128);
        } else {
            c[j] = NULL;
        }
    }
    for (Py_ssize_t j = 0; j < k; ++j) {
        if (!c[j]) continue;
        PyObject *v = PyDict_GetItemWithError(c[j], k_page); /* borrowed; str keys: cannot fail */
        if (v) __builtin_prefetch(v);
        v = PyDict_GetItemWithError(c[j], k_text);
        if (v) __builtin_prefetch(v);
    }
    /* `page in pages` for the row, back to back: three more dependent misses per hit in a table of ~1M pages,
     * independent between hits, so a short loop lets the core overlap them.  Exact int keys in an exact dict
     * cannot raise; anything else is looked up where the Python loop does it. */
    if (pages && PyDict_CheckExact(pages))
        for (Py_ssize_t j = 0; j < k; ++j) {
            if (!c[j]) continue;
            PyObject *v = PyDict_GetItemWithError(c[j], k_page);
            if (v && PyLong_CheckExact(v)) has[j] = (signed char)(PyDict_Contains(pages, v) == 1);
        }
    if (PyErr_Occurred()) PyErr_Clear();
}

/* append_chunk_records(out, D, I, row, chunks, pages, sha1, query_text) -> None
 * One record per hit of row `row`, in rank order, appended to `out`.  An exception (missing chunk, missing parent
 * page) propagates with the records built before it already in `out` — the caller logs it and keeps them, as the
 * reference's per-task try / except does (:1441-1442). */
static PyObject *append_chunk_records(PyObject *self, PyObject *args) {
    PyObject *out, *D, *I, *chunks, *pages, *sha1, *query_text;
    Py_ssize_t row;
    if (!PyArg_ParseTuple(args, "O!OOnOOOO", &PyList_Type, &out, &D, &I, &row, &chunks, &pages, &sha1, &query_text))
        return NULL;
    Rows rw;
    if (rows_open(&rw, D, I, row) < 0) return NULL;
    PyObject *tmpl = new_template(sha1);
    int failed = tmpl == NULL;
    signed char known[PREFETCH_MAX];
    prefetch_hits(chunks, rw.ids, rw.k, pages, known);
    for (Py_ssize_t j = 0; j < rw.k && !failed; ++j) {
        PyObject *chunk = NULL, *page = NULL, *text = NULL, *r = NULL, *rec = NULL;
        failed = 1;
        chunk = seq_item(chunks, rw.ids[j]);
        if (!chunk) goto next;
        page = getitem(chunk, k_page);
        if (!page) goto next;
        int has = (j < PREFETCH_MAX && known[j] >= 0) ? known[j] : PySequence_Contains(pages, page);
        if (has < 0) goto next;
        if (!has) { /* a missing parent page fails the hit, as next() does in the reference (:1145) */
            PyErr_SetObject(PyExc_KeyError, page);
            goto next;
        }
        r = round4((double)rw.scores[j]);
        if (!r) goto next;
        text = getitem(chunk, k_text);
        if (!text) goto next;
        rec = new_record(tmpl, r, page, text, query_text);
        if (!rec || PyList_Append(out, rec) < 0) goto next;
        failed = 0;
    next:
        Py_XDECREF(chunk);
        Py_XDECREF(page);
        Py_XDECREF(text);
        Py_XDECREF(r);
        Py_XDECREF(rec);
    }
    Py_XDECREF(tmpl);
    rows_close(&rw);
    if (failed) return NULL;
    Py_RETURN_NONE;
}

/* append_page_records(out, D, I, row, chunks, pages, sha1, query_text) -> None
 * Parent-page mode: the chunks of one page merge into one record (first-hit order, best similarity), the
 * aggregation of the reference with key (sha1, "page", page) (:1150-1157, :1552-1599). */
static PyObject *append_page_records(PyObject *self, PyObject *args) {
    PyObject *out, *D, *I, *chunks, *pages, *sha1, *query_text;
    Py_ssize_t row;
    if (!PyArg_ParseTuple(args, "O!OOnOOOO", &PyList_Type, &out, &D, &I, &row, &chunks, &pages, &sha1, &query_text))
        return NULL;
    Rows rw;
    if (rows_open(&rw, D, I, row) < 0) return NULL;
    PyObject *by_page = PyDict_New(), *tmpl = new_template(sha1);
    int failed = by_page == NULL || tmpl == NULL;
    signed char known[PREFETCH_MAX];
    prefetch_hits(chunks, rw.ids, rw.k, NULL, known);
    for (Py_ssize_t j = 0; j < rw.k && !failed; ++j) {
        PyObject *chunk = NULL, *page = NULL, *parent = NULL, *ppage = NULL, *text = NULL, *r = NULL, *rec = NULL;
        failed = 1;
        chunk = seq_item(chunks, rw.ids[j]);
        if (!chunk) goto next;
        page = getitem(chunk, k_page);
        if (!page) goto next;
        parent = getitem(pages, page);
        if (!parent) goto next;
        ppage = getitem(parent, k_page);
        if (!ppage) goto next;
        r = round4((double)rw.scores[j]);
        if (!r) goto next;
        PyObject *seen = PyDict_GetItemWithError(by_page, ppage); /* borrowed */
        if (!seen && PyErr_Occurred()) goto next;
        if (!seen) {
            text = getitem(parent, k_text);
            if (!text) goto next;
            rec = new_record(tmpl, r, ppage, text, query_text);
            if (!rec || PyDict_SetItem(by_page, ppage, rec) < 0 || PyList_Append(out, rec) < 0) goto next;
        } else {
            PyObject *cur = PyDict_GetItemWithError(seen, k_vector_similarity); /* borrowed */
            if (!cur) {
                if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, k_vector_similarity);
                goto next;
            }
            int gt = PyObject_RichCompareBool(r, cur, Py_GT);
            if (gt < 0) goto next;
            if (gt && (PyDict_SetItem(seen, k_vector_similarity, r) < 0 ||
                       PyDict_SetItem(seen, k_max_original_similarity, r) < 0))
                goto next;
        }
        failed = 0;
    next:
        Py_XDECREF(chunk);
        Py_XDECREF(page);
        Py_XDECREF(parent);
        Py_XDECREF(ppage);
        Py_XDECREF(text);
        Py_XDECREF(r);
        Py_XDECREF(rec);
    }
    Py_XDECREF(by_page);
    Py_XDECREF(tmpl);
    rows_close(&rw);
    if (failed) return NULL;
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"round4", py_round4, METH_O, "round(x, 4) as float.__round__ computes it"},
    {"append_chunk_records", append_chunk_records, METH_VARARGS, "records of one result row, chunk mode"},
    {"append_page_records", append_page_records, METH_VARARGS, "records of one result row, parent-page mode"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_b2r_records", "record construction for batched retrieval", -1,
                                    methods};

PyMODINIT_FUNC PyInit__b2r_records(void) {
    PyObject *m = PyModule_Create(&moddef);
    if (!m) return NULL;
#define INTERN(var, s)                               \
    if (!(var = PyUnicode_InternFromString(s))) return NULL
    INTERN(k_vector_similarity, "vector_similarity");
    INTERN(k_max_original_similarity, "max_original_similarity");
    INTERN(k_page, "page");
    INTERN(k_text, "text");
    INTERN(k_hit_count, "hit_count");
    INTERN(k_retrieval_sources, "retrieval_sources");
    INTERN(k_source_sha1, "source_sha1");
    INTERN(k_query_sources, "query_sources");
    INTERN(k_basic, "basic");
#undef INTERN
    PyObject *builtins = PyImport_ImportModule("builtins");
    if (!builtins) return NULL;
    g_round = PyObject_GetAttrString(builtins, "round");
    Py_DECREF(builtins);
    g_four = PyLong_FromLong(4);
    g_one = PyLong_FromLong(1);
    if (!g_round || !g_four || !g_one) return NULL;
    return m;
}
