// sdeb_internal.h -- shared between the translation units of libsdeb.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdeb.h"
#include "sde_engine.cuh"

struct ModelInfo {
    const void* fn;        // general kernel
    const void* fn_lean;   // hot-configuration kernel (may be NULL)
    const void* fn_stream[4];   // full-path stream kernel [2*replay + time-dependent records]
                                // (jump models: replay variants only)
    int nw, ndw, nx, npc, ncnt, jumps;
};

template <class M>
static ModelInfo info_of() {
    ModelInfo mi;
    mi.fn = (const void*)&sdeb::integrate_kernel<M>;
    mi.fn_lean = (const void*)&sdeb::integrate_lean_kernel<M>;
    for (int k = 0; k < 4; ++k) mi.fn_stream[k] = 0;
    if constexpr (M::JUMPS == 0) {
        mi.fn_stream[0] = (const void*)&sdeb::stream_kernel<M, sdeb::NOISE_PHILOX, false>;
        mi.fn_stream[1] = (const void*)&sdeb::stream_kernel<M, sdeb::NOISE_PHILOX, true>;
    }
    mi.fn_stream[2] = (const void*)&sdeb::stream_kernel<M, sdeb::NOISE_REPLAY, false>;
    mi.fn_stream[3] = (const void*)&sdeb::stream_kernel<M, sdeb::NOISE_REPLAY, true>;
    mi.nw = M::NW; mi.ndw = M::NDW; mi.nx = M::NX; mi.npc = M::NPC;
    mi.ncnt = M::NCNT; mi.jumps = M::JUMPS;
    return mi;
}

// one registry function per compiled unit of sdeb_models.cu (the kernels are
// instantiated where they are registered, so the units compile in parallel)
#define SDEB_N_UNITS 11
#define SDEB_DECL_UNIT(k) bool sdeb_lookup_unit_##k(int64_t model, int64_t n, ModelInfo& mi);
SDEB_DECL_UNIT(0) SDEB_DECL_UNIT(1) SDEB_DECL_UNIT(2) SDEB_DECL_UNIT(3) SDEB_DECL_UNIT(4)
SDEB_DECL_UNIT(5) SDEB_DECL_UNIT(6) SDEB_DECL_UNIT(7) SDEB_DECL_UNIT(8) SDEB_DECL_UNIT(9)
SDEB_DECL_UNIT(10)
