/*
 * slab_lanestream.cuh - warp-transposed streaming for the sequential kernels.
 *
 * The sequential stages of the codec (sign-LMS, Rice parameter trace, entropy decode, synthesis) give
 * every lane of a warp its own row: one block x channel slot of a block-padded plane.  Read directly,
 * every load instruction of such a warp touches 32 different cache lines and its latency lands on the
 * loop-carried path.  Here the warp instead moves TILE_BYTES of all 32 rows per step with cp.async
 * (16 bytes per lane per operation, whole 128-byte lines per row => coalesced), keeps STAGES tiles in
 * flight ahead of the consumer, and each lane then reads its own row from shared memory with 128-bit
 * loads (row stride = TILE_BYTES + 16 bytes => the four quarter-warp phases of an LDS.128 hit disjoint
 * banks).  Results go the opposite way: lanes write their row of an output tile, the warp stores the
 * tile with coalesced 128-bit global stores.
 *
 * Rows are described by their global byte address and byte length (a multiple of 16: slots in the
 * padded planes start on multiples of 8 samples and are padded to multiples of 8); length 0 = lane
 * idle.  All functions must be called by the full warp.
 */
#ifndef SLAB_LANESTREAM_CUH
#define SLAB_LANESTREAM_CUH

#include "slab_cuda.h"

struct __align__(16) SlabLsRows {
  unsigned long long ptr[32];
  uint32_t bytes[32];
};

template <int TILE_BYTES> struct SlabLsGeom {
  static constexpr int ROW = TILE_BYTES + 16;
  static constexpr int STAGE = 32 * ROW;
  static constexpr int SEGS = TILE_BYTES / 16;            /* 16-byte segments per row tile */
  static constexpr int ROWS_PER_OP = 32 / SEGS;           /* rows covered by one warp-wide operation */
  static constexpr int OPS = 32 / ROWS_PER_OP;
  static_assert(TILE_BYTES >= 16 && TILE_BYTES <= 512 && (TILE_BYTES & (TILE_BYTES - 1)) == 0, "tile size");
};

__device__ __forceinline__ void slab_cp_async16(void* smem_dst, const void* gsrc)
{
#ifdef SLAB_EMUL
  memcpy(smem_dst, gsrc, 16);
#else
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void slab_cp_async_commit()
{
#ifndef SLAB_EMUL
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
/* wait until at most PENDING of this thread's committed groups are still in flight */
template <int PENDING> __device__ __forceinline__ void slab_cp_async_wait()
{
#ifndef SLAB_EMUL
  asm volatile("cp.async.wait_group %0;" :: "n"(PENDING) : "memory");
#endif
}

/* start the copy of tile `tile` of all 32 rows into `stage` (no commit) */
template <int TILE_BYTES>
__device__ __forceinline__ void slab_ls_load(const SlabLsRows* rows, unsigned char* stage, uint32_t tile, uint32_t lane)
{
  typedef SlabLsGeom<TILE_BYTES> G;
  const uint32_t sub = lane / G::SEGS, seg = lane % G::SEGS;
  const uint32_t off = tile * (uint32_t)TILE_BYTES + seg * 16u;
#pragma unroll
  for (int q = 0; q < G::OPS; q++) {
    const uint32_t row = (uint32_t)q * G::ROWS_PER_OP + sub;
    if (off < rows->bytes[row])
      slab_cp_async16(stage + row * G::ROW + seg * 16u, reinterpret_cast<const unsigned char*>(rows->ptr[row]) + off);
  }
}

/* write tile `tile` of all 32 rows from `stage` to global memory; the caller has done __syncwarp()
 * after the lanes filled their rows and must __syncwarp() again before the stage is refilled */
template <int TILE_BYTES>
__device__ __forceinline__ void slab_ls_store(const SlabLsRows* rows, const unsigned char* stage, uint32_t tile, uint32_t lane)
{
  typedef SlabLsGeom<TILE_BYTES> G;
  const uint32_t sub = lane / G::SEGS, seg = lane % G::SEGS;
  const uint32_t off = tile * (uint32_t)TILE_BYTES + seg * 16u;
#pragma unroll
  for (int q = 0; q < G::OPS; q++) {
    const uint32_t row = (uint32_t)q * G::ROWS_PER_OP + sub;
    if (off < rows->bytes[row])
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(rows->ptr[row]) + off) =
          *reinterpret_cast<const uint4*>(stage + row * G::ROW + seg * 16u);
  }
}

__device__ __forceinline__ uint32_t slab_warp_max(uint32_t v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t w = __shfl_xor_sync(SLAB_FULL_MASK, v, o);
    v = w > v ? w : v;
  }
  return v;
}

#endif
