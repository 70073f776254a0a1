// sdeb_internal.h -- shared between the translation units of libsdeb.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdeb.h"
#include "sde_engine.cuh"

struct ModelInfo {
    const void* fn;        // general kernel
    const void* fn_lean;   // hot-configuration kernel (may be NULL)
    int nw, ndw, nx, npc, ncnt, jumps;
};

template <class M>
static ModelInfo info_of() {
    ModelInfo mi;
    mi.fn = (const void*)&sdeb::integrate_kernel<M>;
    mi.fn_lean = (const void*)&sdeb::integrate_lean_kernel<M>;
    mi.nw = M::NW; mi.ndw = M::NDW; mi.nx = M::NX; mi.npc = M::NPC;
    mi.ncnt = M::NCNT; mi.jumps = M::JUMPS;
    return mi;
}

// one registry function per translation unit (the kernels are instantiated
// where they are registered, so the units compile in parallel)
bool sdeb_lookup_linear(int64_t model, int64_t n, ModelInfo& mi);
bool sdeb_lookup_meanrev(int64_t model, int64_t n, ModelInfo& mi);
bool sdeb_lookup_heston(int64_t model, int64_t n, ModelInfo& mi);
