/* gdb_fastcols.c -- CPython helper for batch packing: concatenate one column
 * of many tiny per-graph tables into one array without per-array Python or
 * numpy overhead (20 000 graphs x 7 columns: np.concatenate spends ~1.5 us per
 * array, this loop ~0.1 us).  Host-side plumbing of
 * B200Backend.pack_graphs; the packing itself is gdb_graphs_pack_batch
 * (gdb_pack.cpp).  Replaces the per-graph Python of reference
 * graphdot/kernel/marginalized/_octilegraph.py:37-99.
 *
 *   lengths(tables, key, counts)        counts[i] = len(tables[i][key])
 *   gather(tables, key, out, counts)    out = concatenation of tables[i][key]
 *
 * `tables` is a list of dicts {column name: array}; arrays are read through
 * the buffer protocol: one-dimensional (possibly strided), one element type.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static PyObject *column_of(PyObject *table, PyObject *key) {
    PyObject *col;
    if (PyDict_Check(table)) {
        col = PyDict_GetItemWithError(table, key); /* borrowed */
        if (!col) {
            if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, key);
            return NULL;
        }
        Py_INCREF(col);
        return col;
    }
    return PyObject_GetItem(table, key);
}

static PyObject *lengths(PyObject *self, PyObject *args) {
    PyObject *tables, *key;
    Py_buffer counts;
    if (!PyArg_ParseTuple(args, "OOw*", &tables, &key, &counts)) return NULL;
    PyObject *seq = PySequence_Fast(tables, "tables must be a sequence");
    if (!seq) {
        PyBuffer_Release(&counts);
        return NULL;
    }
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    if (counts.len < (Py_ssize_t)(n * sizeof(int64_t))) {
        PyErr_SetString(PyExc_ValueError, "counts buffer too small");
        goto fail;
    }
    int64_t *cnt = (int64_t *)counts.buf;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *col = column_of(PySequence_Fast_GET_ITEM(seq, i), key);
        if (!col) goto fail;
        const Py_ssize_t len = PyObject_Length(col);
        Py_DECREF(col);
        if (len < 0) goto fail;
        cnt[i] = (int64_t)len;
    }
    Py_DECREF(seq);
    PyBuffer_Release(&counts);
    Py_RETURN_NONE;
fail:
    Py_DECREF(seq);
    PyBuffer_Release(&counts);
    return NULL;
}

static PyObject *gather(PyObject *self, PyObject *args) {
    PyObject *tables, *key;
    Py_buffer out, counts;
    Py_ssize_t itemsize;
    if (!PyArg_ParseTuple(args, "OOw*y*n", &tables, &key, &out, &counts, &itemsize)) return NULL;
    PyObject *seq = PySequence_Fast(tables, "tables must be a sequence");
    if (!seq) {
        PyBuffer_Release(&out);
        PyBuffer_Release(&counts);
        return NULL;
    }
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    const int64_t *cnt = (const int64_t *)counts.buf;
    char *dst = (char *)out.buf;
    char fmt0[32] = "";  /* element format of the first table's column */
    Py_ssize_t at = 0;
    if (counts.len < (Py_ssize_t)(n * sizeof(int64_t))) {
        PyErr_SetString(PyExc_ValueError, "counts buffer too small");
        goto fail;
    }
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *col = column_of(PySequence_Fast_GET_ITEM(seq, i), key);
        if (!col) goto fail;
        Py_buffer v;
        if (PyObject_GetBuffer(col, &v, PyBUF_STRIDES | PyBUF_FORMAT) < 0) {
            Py_DECREF(col);
            goto fail;
        }
        const char *fmt = v.format ? v.format : "B";
        if (i == 0) strncpy(fmt0, fmt, sizeof fmt0 - 1);
        const int ok = v.ndim == 1 && v.itemsize == itemsize && v.len == cnt[i] * itemsize && at + v.len <= out.len &&
                       strncmp(fmt, fmt0, sizeof fmt0 - 1) == 0;
        if (ok) {
            if (!v.strides || v.strides[0] == v.itemsize) {
                memcpy(dst + at, v.buf, (size_t)v.len);
            } else { /* a strided 1-D view, e.g. one column of an (m, 2) edge array */
                const char *src = (const char *)v.buf;
                for (Py_ssize_t k = 0; k < v.shape[0]; ++k) memcpy(dst + at + k * itemsize, src + k * v.strides[0], (size_t)itemsize);
            }
            at += v.len;
        }
        PyBuffer_Release(&v);
        Py_DECREF(col);
        if (!ok) {
            PyErr_Format(PyExc_TypeError,
                         "table %zd: column %R does not match the first table's element type or the table's "
                         "length (all nodes/edges must be of the same type; try Graph.unify_datatype)",
                         i, key);
            goto fail;
        }
    }
    Py_DECREF(seq);
    PyBuffer_Release(&out);
    PyBuffer_Release(&counts);
    return PyLong_FromSsize_t(at);
fail:
    Py_DECREF(seq);
    PyBuffer_Release(&out);
    PyBuffer_Release(&counts);
    return NULL;
}

static PyMethodDef methods[] = {
    {"lengths", lengths, METH_VARARGS, "counts[i] = len(tables[i][key])"},
    {"gather", gather, METH_VARARGS, "concatenate tables[i][key] into out; returns the bytes written"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fastcols", NULL, -1, methods};

PyMODINIT_FUNC PyInit__fastcols(void) { return PyModule_Create(&module); }
