/*
 * refrt.cpp — a minimal single-device CPU implementation of the OpenCL 1.2 host
 * entry points the reference's CLSuperPathTracer hosts + ocl_boiler.h call.
 *
 * TEST INFRASTRUCTURE ONLY.  It lets the UNMODIFIED reference programs
 * (host .c and kernel .ocl, compiled from /root/reference where they lie) run
 * in a container that has no OpenCL runtime, so that
 *   (a) oracle/oracle.c can be pinned bit-for-bit against the real reference, and
 *   (b) bench.py --impl reference can time the reference's own code on host cores.
 *
 * "Device" model: kernels are ordinary C++ functions (ref_kernels.cpp); an
 * NDRange runs work-groups in parallel over OpenMP threads; inside a work-group
 * every work-item is a fiber (hand-rolled x86-64 context switch) so that
 * barrier() has real work-group semantics (needed by the __local staging of the
 * _lmem variants and by reduce4img_lmem's tree reduction).
 *
 * When lws == NULL (all pathTracer launches, e.g. CLSuperPathTracer.c:179) the
 * runtime picks, per dimension, the largest divisor of gws that is <= 16
 * (2-D) or <= 256 (1-D) — the same freedom a real runtime has.
 */
#include <CL/cl.h>

#include <omp.h>
#include <time.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "refrt.h"

namespace ocl {
struct WorkItem {
    size_t gid[3], lid[3], grp[3];
    const size_t *gsz, *lsz, *ngrp;
};
thread_local WorkItem *refrt_wi = nullptr;
void refrt_barrier();
}  // namespace ocl

/* ------------------------------------------------------------------ fibers */
extern "C" void refrt_swap(void **save_sp, void *load_sp);
__asm__(
    ".text\n"
    ".globl refrt_swap\n"
    ".type refrt_swap,@function\n"
    "refrt_swap:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  subq $8, %rsp\n  stmxcsr (%rsp)\n  fnstcw 4(%rsp)\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  ldmxcsr (%rsp)\n  fldcw 4(%rsp)\n  addq $8, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size refrt_swap,.-refrt_swap\n");

namespace {

const size_t kFiberStack = 96 * 1024;

struct Fiber {
    void *sp = nullptr;
    unsigned char *stack = nullptr;
    bool done = false;
    ocl::WorkItem wi;
};

struct GroupRunner {
    std::vector<Fiber> fibers;
    void *sched_sp = nullptr;
    Fiber *current = nullptr;
    const RefLaunch *launch = nullptr;
    ref_kernel_fn fn = nullptr;
    std::vector<std::vector<unsigned char>> local_store;
    std::vector<unsigned char *> local_ptrs;
};

thread_local GroupRunner *tl_runner = nullptr;

void fiber_entry() {
    GroupRunner *r = tl_runner;
    Fiber *f = r->current;
    r->fn(*r->launch);
    f->done = true;
    refrt_swap(&f->sp, r->sched_sp);
    abort(); /* a finished fiber is never resumed */
}

void fiber_prepare(Fiber &f) {
    if (!f.stack) {
        void *p = nullptr;
        if (posix_memalign(&p, 64, kFiberStack)) abort();
        f.stack = (unsigned char *)p;
    }
    uintptr_t top = ((uintptr_t)f.stack + kFiberStack) & ~(uintptr_t)15;
    uint64_t *s = (uint64_t *)top;
    s[-1] = 0;                        /* fake return address of fiber_entry */
    s[-2] = (uint64_t)&fiber_entry;   /* popped by `ret` in refrt_swap */
    for (int i = 3; i <= 8; ++i) s[-i] = 0; /* rbp rbx r12 r13 r14 r15 */
    uint32_t *cw = (uint32_t *)&s[-9];
    cw[0] = 0x1F80;                   /* MXCSR default: all masked, RN, no FTZ/DAZ */
    cw[1] = 0x037F;                   /* x87 control word default */
    f.sp = &s[-9];
    f.done = false;
}

double now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e9 + (double)ts.tv_nsec;
}

size_t largest_divisor_le(size_t n, size_t cap) {
    for (size_t d = cap; d > 1; --d)
        if (n % d == 0) return d;
    return 1;
}

}  // namespace

void ocl::refrt_barrier() {
    GroupRunner *r = tl_runner;
    Fiber *f = r->current;
    refrt_swap(&f->sp, r->sched_sp);
}

/* ----------------------------------------------------------- runtime objects */
struct _cl_platform_id { int dummy; };
struct _cl_device_id { int dummy; };
struct _cl_context { int dummy; };
struct _cl_command_queue { int dummy; };
struct _cl_program { int dummy; };
struct _cl_kernel {
    const RefKernelDesc *desc;
    std::vector<RefArg> args;
};
struct _cl_event {
    cl_ulong start_ns, end_ns;
};

static _cl_platform_id g_platform;
static _cl_device_id g_device;
static const char kBuildLog[] = "refrt: kernels were compiled ahead of time from the reference .ocl with g++\n";

void refrt_override_seeds(uint32_t seeds[4]) {
    const char *env = getenv("PT_SEEDS");
    if (!env || !env[0]) return;
    unsigned long v[4];
    if (sscanf(env, "%lu,%lu,%lu,%lu", &v[0], &v[1], &v[2], &v[3]) == 4)
        for (int i = 0; i < 4; ++i) seeds[i] = (uint32_t)v[i];
}

static void set_err(cl_int *e, cl_int v) {
    if (e) *e = v;
}

static cl_int copy_out(const void *src, size_t n, size_t size, void *dst, size_t *ret) {
    if (ret) *ret = n;
    if (dst) {
        if (size < n) return CL_INVALID_VALUE;
        memcpy(dst, src, n);
    }
    return CL_SUCCESS;
}

extern "C" {

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id *p, cl_uint *np) {
    if (np) *np = 1;
    if (p && n > 0) p[0] = &g_platform;
    return CL_SUCCESS;
}

cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t size, void *dst, size_t *ret) {
    static const char name[] = "refrt (reference kernels compiled for the host CPU)";
    return copy_out(name, sizeof(name), size, dst, ret);
}

cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint n, cl_device_id *d, cl_uint *nd) {
    if (nd) *nd = 1;
    if (d && n > 0) d[0] = &g_device;
    return CL_SUCCESS;
}

cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t size, void *dst, size_t *ret) {
    char name[128];
    snprintf(name, sizeof(name), "host CPU, %d OpenMP threads", omp_get_max_threads());
    return copy_out(name, strlen(name) + 1, size, dst, ret);
}

cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *,
                           void (*)(const char *, const void *, size_t, void *), void *, cl_int *err) {
    set_err(err, CL_SUCCESS);
    return new _cl_context();
}

cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *err) {
    set_err(err, CL_SUCCESS);
    return new _cl_command_queue();
}

cl_program clCreateProgramWithSource(cl_context, cl_uint, const char **, const size_t *, cl_int *err) {
    set_err(err, CL_SUCCESS);
    return new _cl_program();
}

cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id *, const char *, void (*)(cl_program, void *), void *) {
    return CL_SUCCESS;
}

cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t size, void *dst, size_t *ret) {
    /* ocl_boiler.h:188-200 strips trailing '\n'/'\0' and writes 2 bytes after the
     * text, so the log must end in "\n\0" to stay inside its malloc. */
    return copy_out(kBuildLog, sizeof(kBuildLog), size, dst, ret);
}

cl_kernel clCreateKernel(cl_program, const char *name, cl_int *err) {
    for (const RefKernelDesc *d = ref_kernel_table; d->name; ++d) {
        if (!strcmp(d->name, name)) {
            _cl_kernel *k = new _cl_kernel();
            k->desc = d;
            k->args.resize(d->nargs);
            set_err(err, CL_SUCCESS);
            return k;
        }
    }
    set_err(err, CL_INVALID_KERNEL_NAME);
    return nullptr;
}

cl_int clGetKernelWorkGroupInfo(cl_kernel, cl_device_id, cl_kernel_work_group_info, size_t size, void *dst, size_t *ret) {
    size_t v = 256;
    return copy_out(&v, sizeof(v), size, dst, ret);
}

cl_mem clCreateBuffer(cl_context, cl_mem_flags flags, size_t size, void *host, cl_int *err) {
    _cl_mem *m = new _cl_mem();
    m->size = size;
    void *p = nullptr;
    if (posix_memalign(&p, 128, size ? size : 1)) abort();
    m->data = p;
    if ((flags & CL_MEM_COPY_HOST_PTR) && host) memcpy(m->data, host, size);
    set_err(err, CL_SUCCESS);
    return m;
}

cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t size, const void *value) {
    if ((int)idx >= k->desc->nargs) return CL_INVALID_ARG_INDEX;
    RefArg &a = k->args[idx];
    a.set = true;
    if (value == nullptr) {
        a.is_local = true;
        a.local_size = size;
    } else {
        a.is_local = false;
        a.bytes.assign((const unsigned char *)value, (const unsigned char *)value + size);
    }
    return CL_SUCCESS;
}

cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel k, cl_uint dim, const size_t *, const size_t *gws_in,
                              const size_t *lws_in, cl_uint, const cl_event *, cl_event *evt) {
    size_t gsz[3] = {1, 1, 1}, lsz[3] = {1, 1, 1}, ngrp[3] = {1, 1, 1};
    for (cl_uint d = 0; d < dim; ++d) gsz[d] = gws_in[d];
    for (cl_uint d = 0; d < dim; ++d) {
        if (lws_in) {
            lsz[d] = lws_in[d];
            if (gsz[d] % lsz[d]) return CL_INVALID_WORK_GROUP_SIZE;
        } else {
            lsz[d] = largest_divisor_le(gsz[d], dim == 1 ? 256 : 16);
        }
        ngrp[d] = gsz[d] / lsz[d];
    }
    const size_t nlocal = lsz[0] * lsz[1] * lsz[2];
    const long ngroups = (long)(ngrp[0] * ngrp[1] * ngrp[2]);
    const int nargs = k->desc->nargs;
    const std::vector<RefArg> *args = &k->args;
    ref_kernel_fn fn = k->desc->fn;

    double t0 = now_ns();
#pragma omp parallel
    {
        GroupRunner runner;
        runner.fibers.resize(nlocal);
        runner.fn = fn;
        runner.local_store.resize(nargs);
        runner.local_ptrs.assign(nargs, nullptr);
        for (int a = 0; a < nargs; ++a)
            if ((*args)[a].is_local) {
                runner.local_store[a].assign((*args)[a].local_size + 64, 0xCD);
                runner.local_ptrs[a] = runner.local_store[a].data();
            }
        RefLaunch launch{args, runner.local_ptrs.data()};
        runner.launch = &launch;
        tl_runner = &runner;
#pragma omp for schedule(dynamic, 1)
        for (long g = 0; g < ngroups; ++g) {
            size_t gx = (size_t)g % ngrp[0], gy = ((size_t)g / ngrp[0]) % ngrp[1], gz = (size_t)g / (ngrp[0] * ngrp[1]);
            size_t n = 0;
            for (size_t lz = 0; lz < lsz[2]; ++lz)
                for (size_t ly = 0; ly < lsz[1]; ++ly)
                    for (size_t lx = 0; lx < lsz[0]; ++lx, ++n) {
                        Fiber &f = runner.fibers[n];
                        fiber_prepare(f);
                        f.wi.lid[0] = lx; f.wi.lid[1] = ly; f.wi.lid[2] = lz;
                        f.wi.grp[0] = gx; f.wi.grp[1] = gy; f.wi.grp[2] = gz;
                        f.wi.gid[0] = gx * lsz[0] + lx; f.wi.gid[1] = gy * lsz[1] + ly; f.wi.gid[2] = gz * lsz[2] + lz;
                        f.wi.gsz = gsz; f.wi.lsz = lsz; f.wi.ngrp = ngrp;
                    }
            size_t remaining = nlocal;
            while (remaining) {
                /* one scheduling round = run every live fiber up to its next barrier */
                for (size_t i = 0; i < nlocal; ++i) {
                    Fiber &f = runner.fibers[i];
                    if (f.done) continue;
                    runner.current = &f;
                    ocl::refrt_wi = &f.wi;
                    refrt_swap(&runner.sched_sp, f.sp);
                    if (f.done) --remaining;
                }
            }
        }
        for (Fiber &f : runner.fibers) free(f.stack);
        tl_runner = nullptr;
    }
    double t1 = now_ns();
    if (evt) {
        _cl_event *e = new _cl_event();
        e->start_ns = (cl_ulong)t0;
        e->end_ns = (cl_ulong)t1;
        *evt = e;
    }
    return CL_SUCCESS;
}

void *clEnqueueMapBuffer(cl_command_queue, cl_mem m, cl_bool, cl_map_flags, size_t off, size_t, cl_uint,
                         const cl_event *, cl_event *evt, cl_int *err) {
    double t0 = now_ns();
    if (evt) {
        _cl_event *e = new _cl_event();
        e->start_ns = (cl_ulong)t0;
        e->end_ns = (cl_ulong)now_ns() + 1;
        *evt = e;
    }
    set_err(err, CL_SUCCESS);
    return (unsigned char *)m->data + off;
}

cl_int clEnqueueUnmapMemObject(cl_command_queue, cl_mem, void *, cl_uint, const cl_event *, cl_event *) { return CL_SUCCESS; }

cl_int clGetEventProfilingInfo(cl_event e, cl_profiling_info what, size_t size, void *dst, size_t *ret) {
    cl_ulong v = (what == CL_PROFILING_COMMAND_START) ? e->start_ns : e->end_ns;
    return copy_out(&v, sizeof(v), size, dst, ret);
}

cl_int clReleaseMemObject(cl_mem m) {
    if (m) {
        free(m->data);
        delete m;
    }
    return CL_SUCCESS;
}
cl_int clReleaseKernel(cl_kernel k) { delete k; return CL_SUCCESS; }
cl_int clReleaseProgram(cl_program p) { delete p; return CL_SUCCESS; }
cl_int clReleaseCommandQueue(cl_command_queue q) { delete q; return CL_SUCCESS; }
cl_int clReleaseContext(cl_context c) { delete c; return CL_SUCCESS; }

}  // extern "C"
