// kernels_fast.cuh -- FAST-mode phase 1 (placeholder until the register-resident kernel lands).
#pragma once
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/mktfhe_params.h"

struct FastKeys { int dummy = 0; };
static inline bool fast_supported(const mktfhe_params &) { return false; }
static inline void fast_free(FastKeys &) {}
static inline int fast_build(FastKeys &, const mktfhe_params &, const std::vector<cplx *> &, cudaStream_t, std::string &) { return 0; }
static inline int fast_phase1(FastKeys &, const mktfhe_params &, const uint32_t *, cplx *, size_t, cudaStream_t, int *, std::string &err) {
    err = "FAST mode not built"; return -1;
}
static inline int fast_cmux_step(FastKeys &, const mktfhe_params &, int, int, const uint32_t *, void *, size_t, cudaStream_t, int *, std::string &err) {
    err = "FAST mode not built"; return -1;
}
