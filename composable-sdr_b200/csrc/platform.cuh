// platform.cuh -- the one place that knows whether kernel sources are being compiled by nvcc (product,
// sm_100a) or by g++ inside the TEST-ONLY thread emulator (tests/emu/cuda_emu.h, -DCSDR_EMU).  The emulator
// exists so that tile/index arithmetic of the real kernel sources can be checked in the CPU-only test suite;
// it is never built into libcsdr_b200.so and the package never loads it.
#pragma once

#ifdef CSDR_EMU
#include "cuda_emu.h"
#define CSDR_DYN_SMEM(name) unsigned char *name = ::csdr_emu::dyn_smem()
#define CSDR_GRID_CONSTANT
#else
#include <cuda_runtime.h>
#define CSDR_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define CSDR_GRID_CONSTANT __grid_constant__
#endif

#include <stdint.h>

namespace csdr {

constexpr int kMaxStages = 12;   // half-band stages (2^12 decimation) supported by the fused front end
constexpr int kMaxHbM    = 16;   // max half-band semi-length m (2m taps)
constexpr int kHsub      = 14;   // taps per polyphase branch of the arbitrary resampler (2*7, msresamp.c)
constexpr int kHcPad     = 16;   // c-rate history carried into every tile (>= kHsub-1, multiple of 8)

__host__ __device__ inline float2 cf(float re, float im) { float2 z; z.x = re; z.y = im; return z; }

}  // namespace csdr
