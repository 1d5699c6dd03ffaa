/*
 * refrt.h — internal interface between the tiny CPU "OpenCL runtime"
 * (refrt.cpp) and the per-variant kernel translation units (ref_kernels.cpp).
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 */
#ifndef REFRT_H
#define REFRT_H

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

struct RefArg {
    std::vector<unsigned char> bytes;  /* by-value bytes, or the cl_mem handle */
    size_t local_size = 0;             /* >0: __local buffer of that many bytes */
    bool is_local = false;
    bool set = false;
};

struct _cl_mem {
    void *data;
    size_t size;
};

/* What a kernel trampoline sees for one work-item. */
struct RefLaunch {
    const std::vector<RefArg> *args;
    unsigned char *const *locals; /* per-arg base of this work-group's __local buffers */

    template <class T> T val(int i) const {
        T v;
        std::memcpy(&v, (*args)[i].bytes.data(), sizeof(T));
        return v;
    }
    void *mem(int i) const {
        _cl_mem *m;
        std::memcpy(&m, (*args)[i].bytes.data(), sizeof(m));
        return m ? m->data : nullptr;
    }
    void *local(int i) const { return locals[i]; }
};

typedef void (*ref_kernel_fn)(const RefLaunch &);

struct RefKernelDesc {
    const char *name;
    int nargs;
    ref_kernel_fn fn;
};

/* Each variant's kernel TU defines this table (terminated by name == nullptr). */
extern const RefKernelDesc ref_kernel_table[];

/* PT_SEEDS=a,b,c,d overrides the wall-clock seeds the reference host derives
 * (CLSuperPathTracer.c:209) — applied by the trampolines to the `seeds` kernel
 * argument, so the reference sources stay untouched. */
void refrt_override_seeds(uint32_t seeds[4]);

#endif
