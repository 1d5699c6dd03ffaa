/*
 * Minimal OpenCL 1.2 host API declarations — TEST INFRASTRUCTURE ONLY.
 *
 * This header exists so that the UNMODIFIED reference host programs
 * (the CLSuperPathTracer.c of each variant + ocl_boiler.h, which do
 * `#include <CL/cl.h>`, ocl_boiler.h:27) can be compiled in a container that
 * ships no OpenCL SDK.  It declares only the types, constants and entry
 * points those programs use; the entry points are implemented by
 * oracle/refrt/refrt.cpp (a tiny single-device CPU "OpenCL runtime" that runs
 * the reference kernels compiled from their own .ocl sources through
 * oracle/refrt/clshim.h).  Written from the public Khronos OpenCL 1.2
 * specification; nothing in the product path includes it.
 */
#ifndef REFRT_CL_H
#define REFRT_CL_H

#include <stddef.h>
#include <stdint.h>
#include <float.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int8_t   cl_char;
typedef uint8_t  cl_uchar;
typedef int16_t  cl_short;
typedef uint16_t cl_ushort;
typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef int64_t  cl_long;
typedef uint64_t cl_ulong;
typedef float    cl_float;
typedef double   cl_double;

typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_map_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef intptr_t cl_context_properties;
typedef cl_uint  cl_platform_info;
typedef cl_uint  cl_device_info;
typedef cl_uint  cl_program_build_info;
typedef cl_uint  cl_kernel_work_group_info;
typedef cl_uint  cl_profiling_info;

typedef struct _cl_platform_id   *cl_platform_id;
typedef struct _cl_device_id     *cl_device_id;
typedef struct _cl_context       *cl_context;
typedef struct _cl_command_queue *cl_command_queue;
typedef struct _cl_mem           *cl_mem;
typedef struct _cl_program       *cl_program;
typedef struct _cl_kernel        *cl_kernel;
typedef struct _cl_event         *cl_event;

/* 4-wide host vector types: 16-byte aligned, addressable as .x/.y/.z/.w,
 * .s0-.s3 and .s[i] (OpenCL 1.2 spec, appendix on cl_platform.h types). */
#define REFRT_VEC4(T, NAME)                                            \
    typedef union {                                                    \
        T s[4] __attribute__((aligned(sizeof(T) * 4)));                \
        struct { T x, y, z, w; };                                      \
        struct { T s0, s1, s2, s3; };                                  \
    } NAME

REFRT_VEC4(cl_float, cl_float4);
REFRT_VEC4(cl_int,   cl_int4);
REFRT_VEC4(cl_uint,  cl_uint4);

#define CL_SUCCESS                 0
#define CL_INVALID_VALUE           (-30)
#define CL_INVALID_KERNEL_NAME     (-46)
#define CL_INVALID_ARG_INDEX       (-49)
#define CL_INVALID_WORK_GROUP_SIZE (-54)
#define CL_FALSE 0
#define CL_TRUE  1

#define CL_FLT_MAX FLT_MAX
#define CL_FLT_MIN FLT_MIN

#define CL_DEVICE_TYPE_ALL          0xFFFFFFFF
#define CL_PLATFORM_NAME            0x0902
#define CL_DEVICE_NAME              0x102B
#define CL_CONTEXT_PLATFORM         0x1084
#define CL_QUEUE_PROFILING_ENABLE   (1 << 1)
#define CL_MEM_READ_WRITE           (1 << 0)
#define CL_MEM_WRITE_ONLY           (1 << 1)
#define CL_MEM_READ_ONLY            (1 << 2)
#define CL_MEM_ALLOC_HOST_PTR       (1 << 4)
#define CL_MEM_COPY_HOST_PTR        (1 << 5)
#define CL_MAP_READ                 (1 << 0)
#define CL_PROGRAM_BUILD_LOG        0x1183
#define CL_KERNEL_WORK_GROUP_SIZE   0x11B0
#define CL_PROFILING_COMMAND_START  0x1282
#define CL_PROFILING_COMMAND_END    0x1283

cl_int clGetPlatformIDs(cl_uint, cl_platform_id *, cl_uint *);
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t, void *, size_t *);
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint, cl_device_id *, cl_uint *);
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void *, size_t *);
cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *,
                           void (*)(const char *, const void *, size_t, void *), void *, cl_int *);
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *);
cl_program clCreateProgramWithSource(cl_context, cl_uint, const char **, const size_t *, cl_int *);
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id *, const char *,
                      void (*)(cl_program, void *), void *);
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t, void *, size_t *);
cl_kernel clCreateKernel(cl_program, const char *, cl_int *);
cl_int clGetKernelWorkGroupInfo(cl_kernel, cl_device_id, cl_kernel_work_group_info, size_t, void *, size_t *);
cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void *, cl_int *);
cl_int clSetKernelArg(cl_kernel, cl_uint, size_t, const void *);
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel, cl_uint, const size_t *, const size_t *,
                              const size_t *, cl_uint, const cl_event *, cl_event *);
void *clEnqueueMapBuffer(cl_command_queue, cl_mem, cl_bool, cl_map_flags, size_t, size_t,
                         cl_uint, const cl_event *, cl_event *, cl_int *);
cl_int clEnqueueUnmapMemObject(cl_command_queue, cl_mem, void *, cl_uint, const cl_event *, cl_event *);
cl_int clGetEventProfilingInfo(cl_event, cl_profiling_info, size_t, void *, size_t *);
cl_int clReleaseMemObject(cl_mem);
cl_int clReleaseKernel(cl_kernel);
cl_int clReleaseProgram(cl_program);
cl_int clReleaseCommandQueue(cl_command_queue);
cl_int clReleaseContext(cl_context);

#ifdef __cplusplus
}
#endif
#endif /* REFRT_CL_H */
