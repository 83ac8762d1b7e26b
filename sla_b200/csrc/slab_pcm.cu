/*
 * slab_pcm.cu - interleaved little-endian PCM <-> planar left-justified int32 on the device
 * (north_star kernel 1; replaces the reference's WAV sample loops, src/wav.c:208-252 with the
 * conversions of :392-417 on the way in, :630-668 on the way out).
 *
 * PCM is the layout of a WAV data chunk: frames of num_channels samples, 8-bit unsigned or
 * 16/24/32-bit signed little-endian.  A CTA moves a tile of 256 frames: the interleaved side is read
 * or written as one contiguous run of 128-bit words, the planar side as one 1 KiB run per channel,
 * and the (de)interleave itself happens in shared memory - both HBM sides are fully coalesced whatever
 * the channel count or sample width.
 */
#include "slab_common.cuh"
#include <stdlib.h>
#include "slab_ctx.cuh"

#define PCM_TILE 256u                      /* frames per CTA; 256 * frame_bytes is a multiple of 16 */

/* src/wav.c:392-417 */
__device__ __forceinline__ int32_t pcm_load_sample(const unsigned char* p, uint32_t bytes)
{
  switch (bytes) {
    case 1: return (int32_t)(((uint32_t)p[0] - 128u) << 24);
    case 2: return (int32_t)(((uint32_t)p[0] | ((uint32_t)p[1] << 8)) << 16);
    case 3: return (int32_t)(((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16)) << 8);
    default: return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
  }
}
/* src/wav.c:630-668 (the decoder's samples are left-justified: nothing is lost) */
__device__ __forceinline__ void pcm_store_sample(unsigned char* p, uint32_t bytes, int32_t v)
{
  const uint32_t u = (uint32_t)v;
  switch (bytes) {
    case 1: p[0] = (unsigned char)((u >> 24) + 128u); break;
    case 2: p[0] = (unsigned char)(u >> 16); p[1] = (unsigned char)(u >> 24); break;
    case 3: p[0] = (unsigned char)(u >> 8); p[1] = (unsigned char)(u >> 16); p[2] = (unsigned char)(u >> 24); break;
    default: p[0] = (unsigned char)u; p[1] = (unsigned char)(u >> 8); p[2] = (unsigned char)(u >> 16); p[3] = (unsigned char)(u >> 24); break;
  }
}

__global__ void __launch_bounds__(256) k_pcm_to_planar(int32_t* __restrict__ planes, size_t plane_stride,
    const unsigned char* __restrict__ pcm, uint32_t nch, uint32_t bytes, uint32_t nframes, int aligned)
{
  __shared__ __align__(16) unsigned char tile[PCM_TILE * SLAB_MAX_CH * 4u];
  const uint32_t tid = threadIdx.x;
  const size_t f0 = (size_t)blockIdx.x * PCM_TILE;
  const uint32_t frames = (nframes - f0 < PCM_TILE) ? (uint32_t)(nframes - f0) : PCM_TILE;
  const uint32_t fb = nch * bytes, total = frames * fb;
  const unsigned char* src = pcm + f0 * fb;
  if (aligned && frames == PCM_TILE) {
    for (uint32_t i = tid; i < total / 16u; i += 256u)
      reinterpret_cast<uint4*>(tile)[i] = __ldcs(reinterpret_cast<const uint4*>(src) + i);
  } else {
    for (uint32_t i = tid; i < total; i += 256u) tile[i] = src[i];
  }
  __syncthreads();
  if (tid < frames)
    for (uint32_t c = 0; c < nch; c++)
      planes[(size_t)c * plane_stride + f0 + tid] = pcm_load_sample(tile + (tid * nch + c) * bytes, bytes);
}

/* Many files in one launch: tile t covers plane positions [256 t, 256 t + 256), which belong to file
 * chunk_file[t / 4] (file starts are multiples of 1024); positions past the end of the file are zero-filled.
 * tab: start | len | pcm byte offset (low, high) per file. */
__global__ void __launch_bounds__(256) k_pcm_to_planar_files(int32_t* __restrict__ planes, size_t plane_stride,
    const unsigned char* __restrict__ pcm, uint32_t nch, uint32_t bytes,
    const uint32_t* __restrict__ chunk_file, const uint32_t* __restrict__ tab)
{
  __shared__ __align__(16) unsigned char tile[PCM_TILE * SLAB_MAX_CH * 4u];
  const uint32_t tid = threadIdx.x;
  const uint32_t f = chunk_file[blockIdx.x >> 2];
  const uint32_t start = tab[4u * f], len = tab[4u * f + 1u];
  const size_t pcm_off = (size_t)tab[4u * f + 2u] | ((size_t)tab[4u * f + 3u] << 32);
  const size_t p0 = (size_t)blockIdx.x * PCM_TILE;           /* plane position of the tile */
  const uint32_t f0 = (uint32_t)(p0 - start);                /* first frame of the file in this tile */
  const uint32_t frames = (f0 >= len) ? 0u : ((len - f0 < PCM_TILE) ? len - f0 : PCM_TILE);
  const uint32_t fb = nch * bytes, total = frames * fb;
  const unsigned char* src = pcm + pcm_off + (size_t)f0 * fb;
  if (frames == PCM_TILE) {
    for (uint32_t i = tid; i < total / 16u; i += 256u)
      reinterpret_cast<uint4*>(tile)[i] = __ldcs(reinterpret_cast<const uint4*>(src) + i);
  } else {
    for (uint32_t i = tid; i < total; i += 256u) tile[i] = src[i];
  }
  __syncthreads();
  for (uint32_t c = 0; c < nch; c++)
    planes[(size_t)c * plane_stride + p0 + tid] = (tid < frames) ? pcm_load_sample(tile + (tid * nch + c) * bytes, bytes) : 0;
}

/* the file of every 1024-frame chunk of the planes, from the table above */
__global__ void __launch_bounds__(128) k_pcm_chunk_map(const uint32_t* __restrict__ tab, uint32_t nfiles, uint32_t nchunks,
    uint32_t* __restrict__ chunk_file)
{
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfiles) return;
  const uint32_t c0 = (f == 0u) ? 0u : (tab[4u * f] >> 10);
  const uint32_t c1 = (f + 1u < nfiles) ? (tab[4u * (f + 1u)] >> 10) : nchunks;
  for (uint32_t ch = c0; ch < c1; ch++) chunk_file[ch] = f;
}

__global__ void __launch_bounds__(256) k_planar_to_pcm(unsigned char* __restrict__ pcm,
    const int32_t* __restrict__ planes, size_t plane_stride, uint32_t nch, uint32_t bytes, uint32_t nframes, int aligned)
{
  __shared__ __align__(16) unsigned char tile[PCM_TILE * SLAB_MAX_CH * 4u];
  const uint32_t tid = threadIdx.x;
  const size_t f0 = (size_t)blockIdx.x * PCM_TILE;
  const uint32_t frames = (nframes - f0 < PCM_TILE) ? (uint32_t)(nframes - f0) : PCM_TILE;
  const uint32_t fb = nch * bytes, total = frames * fb;
  if (tid < frames)
    for (uint32_t c = 0; c < nch; c++)
      pcm_store_sample(tile + (tid * nch + c) * bytes, bytes, planes[(size_t)c * plane_stride + f0 + tid]);
  __syncthreads();
  unsigned char* dst = pcm + f0 * fb;
  if (aligned && frames == PCM_TILE) {
    for (uint32_t i = tid; i < total / 16u; i += 256u)
      __stcs(reinterpret_cast<uint4*>(dst) + i, reinterpret_cast<const uint4*>(tile)[i]);
  } else {
    for (uint32_t i = tid; i < total; i += 256u) dst[i] = tile[i];
  }
}

extern "C" int slab_pcm_to_planar(SlabCtx* ctx, int32_t* d_planes, size_t plane_stride, const void* d_pcm,
    uint32_t nch, uint32_t bytes, uint32_t nframes)
{
  if (nframes == 0) return 0;
  if (nch < 1 || nch > SLAB_MAX_CH || bytes < 1 || bytes > 4) { slab_set_error("sla_b200: unsupported PCM layout"); return -1; }
  const int aligned = (((uintptr_t)d_pcm) & 15u) == 0;
  SLAB_RUN(ctx, "E1 k_pcm_to_planar", k_pcm_to_planar, slab_div_up(nframes, PCM_TILE), 256, 0, d_planes, plane_stride,
           (const unsigned char*)d_pcm, nch, bytes, nframes, aligned);
  return 0;
}

extern "C" int slab_pcm_to_planar_files(SlabCtx* ctx, int32_t* d_planes, size_t plane_stride, uint32_t plane_len, const void* d_pcm,
    uint32_t nch, uint32_t bytes, uint32_t num_files, const uint32_t* file_start, const uint32_t* file_len, const uint64_t* pcm_off)
{
  if (num_files == 0 || plane_len == 0) return 0;
  if (nch < 1 || nch > SLAB_MAX_CH || bytes < 1 || bytes > 4 || (plane_len & 1023u) != 0 || (((uintptr_t)d_pcm) & 15u) != 0) {
    slab_set_error("sla_b200: unsupported PCM layout");
    return -1;
  }
  const uint32_t nchunks = plane_len >> 10;
  const size_t words = (size_t)nchunks + 4u * (size_t)num_files;
  uint32_t* tab = (uint32_t*)malloc(sizeof(uint32_t) * 4u * (size_t)num_files);
  uint32_t* d = (uint32_t*)slab_user_buffer(ctx, 6, sizeof(uint32_t) * words);        /* chunk map | table */
  if (!tab || !d) { free(tab); return -1; }
  for (uint32_t f = 0; f < num_files; f++) {
    const uint32_t end = (f + 1u < num_files) ? file_start[f + 1u] : plane_len;
    if ((file_start[f] & 1023u) != 0 || (pcm_off[f] & 15u) != 0 || end < file_start[f] || end - file_start[f] < file_len[f] || end > plane_len) {
      free(tab);
      slab_set_error("sla_b200: merged PCM layout: file %u is misplaced", f);
      return -1;
    }
    tab[4u * f] = file_start[f]; tab[4u * f + 1u] = file_len[f];
    tab[4u * f + 2u] = (uint32_t)pcm_off[f]; tab[4u * f + 3u] = (uint32_t)(pcm_off[f] >> 32);
  }
  /* only the table crosses PCIe (pageable source: staged before the call returns); the map is built on the device */
  cudaError_t e = cudaMemcpyAsync(d + nchunks, tab, sizeof(uint32_t) * 4u * num_files, cudaMemcpyHostToDevice, ctx->stream);
  free(tab);
  SLAB_CUDA_TRY(e);
  SLAB_RUN(ctx, "E1 k_pcm_chunk_map", k_pcm_chunk_map, slab_div_up(num_files, 128), 128, 0, (const uint32_t*)(d + nchunks), num_files, nchunks, d);
  SLAB_RUN(ctx, "E1 k_pcm_to_planar_files", k_pcm_to_planar_files, plane_len / PCM_TILE, 256, 0, d_planes, plane_stride,
           (const unsigned char*)d_pcm, nch, bytes, (const uint32_t*)d, (const uint32_t*)(d + nchunks));
  return 0;
}

extern "C" int slab_planar_to_pcm(SlabCtx* ctx, void* d_pcm, const int32_t* d_planes, size_t plane_stride,
    uint32_t nch, uint32_t bytes, uint32_t nframes)
{
  if (nframes == 0) return 0;
  if (nch < 1 || nch > SLAB_MAX_CH || bytes < 1 || bytes > 4) { slab_set_error("sla_b200: unsupported PCM layout"); return -1; }
  const int aligned = (((uintptr_t)d_pcm) & 15u) == 0;
  SLAB_RUN(ctx, "D4 k_planar_to_pcm", k_planar_to_pcm, slab_div_up(nframes, PCM_TILE), 256, 0, (unsigned char*)d_pcm, d_planes,
           plane_stride, nch, bytes, nframes, aligned);
  return 0;
}
