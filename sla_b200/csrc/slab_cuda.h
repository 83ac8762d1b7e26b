/*
 * slab_cuda.h - one include for every .cu file: the CUDA runtime when built by nvcc (the product),
 * or the fibre-based host simulator when tests/hostsim/Makefile builds the kernel unit-test library
 * with -DSLAB_EMUL (never shipped, see tests/hostsim/cuda_emul.h).
 */
#ifndef SLAB_CUDA_H
#define SLAB_CUDA_H

#ifdef SLAB_EMUL
#include "cuda_emul.h"
#define SLAB_LAUNCH(kp, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [=]() { kp(__VA_ARGS__); })
#define SLAB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_dyn_smem)
#else
#include <cuda_runtime.h>
#define SLAB_LAUNCH(kp, grid, block, smem, stream, ...) \
  kp<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SLAB_DYN_SMEM(type, name)                              \
  extern __shared__ __align__(16) unsigned char name##_raw_[]; \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif

#include <stdint.h>

#define SLAB_FULL_MASK 0xffffffffu

#endif
