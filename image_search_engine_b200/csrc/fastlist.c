/* _fastlist: walks the reference's ragged input contract -- a Python list with one (n_i, d) ndarray per image
 * (descriptors.py:104-139) -- in C: data pointer and row count of every array, with the dtype / layout checks the
 * packer needs, at ~20 ns per image instead of ~1.6 us in the interpreter (10 000 images: 0.2 ms instead of 16 ms).
 * Pure host plumbing for ise_pack_rows; no arithmetic. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <stdint.h>

/* walk(seq, ptrs_addr, counts_addr) -> (dtype_code, d) with dtype_code 0 = float32, 1 = uint8 (ise.h ISE_DTYPE_*),
 * or None when the sequence is not made of C-contiguous 2-D arrays of one of those dtypes with a common d.
 * ptrs_addr / counts_addr: addresses of caller-owned uint64[n] / int64[n] buffers. */
static PyObject* walk(PyObject* self, PyObject* args) {
    PyObject* seq;
    unsigned long long ptrs_addr, counts_addr;
    if (!PyArg_ParseTuple(args, "OKK", &seq, &ptrs_addr, &counts_addr)) return NULL;
    PyObject* fast = PySequence_Fast(seq, "expected a sequence of arrays");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject** items = PySequence_Fast_ITEMS(fast);
    uint64_t* ptrs = (uint64_t*)(uintptr_t)ptrs_addr;
    int64_t* counts = (int64_t*)(uintptr_t)counts_addr;
    int code = -1;
    npy_intp d = -1;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* o = items[i];
        if (!PyArray_Check(o)) goto unsupported;
        PyArrayObject* a = (PyArrayObject*)o;
        if (PyArray_NDIM(a) != 2 || !PyArray_IS_C_CONTIGUOUS(a)) goto unsupported;
        const int t = PyArray_TYPE(a);
        const int c = t == NPY_FLOAT32 ? 0 : (t == NPY_UINT8 ? 1 : -1);
        if (c < 0) goto unsupported;
        if (code < 0) { code = c; d = PyArray_DIM(a, 1); }
        if (c != code || PyArray_DIM(a, 1) != d) goto unsupported;
        ptrs[i] = (uint64_t)(uintptr_t)PyArray_DATA(a);
        counts[i] = (int64_t)PyArray_DIM(a, 0);
    }
    Py_DECREF(fast);
    if (code < 0) Py_RETURN_NONE;
    return Py_BuildValue("(in)", code, (Py_ssize_t)d);
unsupported:
    Py_DECREF(fast);
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {{"walk", walk, METH_VARARGS, "pointers and row counts of a list of 2-D arrays"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_fastlist", NULL, -1, methods};

PyMODINIT_FUNC PyInit__fastlist(void) {
    import_array();
    return PyModule_Create(&moduledef);
}
