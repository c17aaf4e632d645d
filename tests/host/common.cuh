// Stub of dalle_pytorch_b200/csrc/common.cuh for the host-side geometry check (tests/host/attn_geometry_check.cpp): the CUDA
// function-space keywords become no-ops so that g++ compiles attn_common.cuh as plain C++.
#pragma once
#include <stdint.h>
#include "../../include/dalle_b200.h"
#define __device__
#define __host__
#define __forceinline__ inline
namespace db200 {}
