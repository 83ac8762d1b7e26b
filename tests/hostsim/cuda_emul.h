/*
 * cuda_emul.h - TEST INFRASTRUCTURE ONLY.
 *
 * A tiny single-process stand-in for the CUDA execution model so that the kernels under
 * sla_b200/csrc/ can be unit-tested in a container without a GPU: every CUDA thread of a CTA
 * runs as a ucontext fibre, __syncthreads()/warp collectives are real rendezvous points, shared
 * memory is static storage (one CTA runs at a time).  It is compiled ONLY into
 * tests/hostsim/libsla_hostsim.so by tests/hostsim/Makefile (g++ -DSLAB_EMUL) and is never part
 * of libsla_b200.so: the product has no CPU path and fails loudly without a CUDA device.
 */
#ifndef SLAB_CUDA_EMUL_H
#define SLAB_CUDA_EMUL_H

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = {x, y, z, w}; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
struct uint2 { unsigned x, y; };
struct double2 { double x, y; };
struct longlong2 { long long x, y; };

namespace emu {
extern uint3 g_tid, g_bid;
extern dim3 g_bdim, g_gdim;
extern unsigned char* g_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void sync_threads();
void sync_warp();
unsigned long long warp_exchange(unsigned long long v, int src_lane);   /* full-mask only */
unsigned ballot(int pred);
}  // namespace emu

#define threadIdx (emu::g_tid)
#define blockIdx (emu::g_bid)
#define blockDim (emu::g_bdim)
#define gridDim (emu::g_gdim)
#define warpSize 32

#define __global__
#define __device__
#define __host__
#define __constant__
#define __shared__ static
#define __restrict__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

static inline void __syncthreads() { emu::sync_threads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T> static inline T emu_xchg(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload too large");
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  raw = emu::warp_exchange(raw, src);
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
static inline int emu_lane() { return (int)(emu::g_tid.x & 31u); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = emu_lane(), base = lane & ~(width - 1);
  return emu_xchg(v, base + (src & (width - 1)));
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = emu_lane(), src = lane + (int)d;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return emu_xchg(v, src);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = emu_lane(), src = lane - (int)d;
  if (src < (lane & ~(width - 1))) src = lane;
  return emu_xchg(v, src);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  (void)width;
  return emu_xchg(v, emu_lane() ^ m);
}
static inline unsigned __ballot_sync(unsigned, int p) { return emu::ballot(p); }
static inline int __any_sync(unsigned, int p) { return emu::ballot(p) != 0; }
static inline int __all_sync(unsigned, int p) { return emu::ballot(!p) == 0; }

static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  unsigned long long v = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
  s &= 31;
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
  s &= 31;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline long long __double2ll_rz(double d) { return (long long)d; }
static inline double __ll2double_rn(long long v) { return (double)v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }

template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

/* ---- the sliver of the runtime API the host-side launch code uses ---- */
typedef int cudaError_t;
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
/* SLAB_EMUL_POISON=1: fresh device memory is filled with 0xCD instead of zeros (finds reads of memory no
 * kernel has written: a real device hands out whatever the pages held) */
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  const char* poison = getenv("SLAB_EMUL_POISON");
  *p = calloc(n ? n : 1, 1);
  if (*p && poison && poison[0] == '1') memset(*p, 0xCD, n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
struct cudaFuncAttributes { size_t sharedSizeBytes; };
template <typename F> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { a->sharedSizeBytes = 0; return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };

#endif
