// Stage-1 (k-mer prefilter) — placeholder until the kernels land; see prefilter design in DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <functional>
#include <string>
#include <vector>
#include "../../include/fsb200.h"
#include "host_core.h"
namespace fsb { namespace prefilter {
struct State {
    uint64_t stats[8] = {0};
    void release() {}
    void invalidate() {}
};
template <typename F>
inline int run(State &, const fsb200_config &, const host::ScoreMatrix &, const host::ScoreMatrix &, cudaStream_t, int,
               const uint8_t *, const uint64_t *, const uint32_t *, const std::vector<uint64_t> &, const std::vector<uint32_t> &, uint32_t,
               const std::vector<uint8_t> &, size_t, F, fsb200_hit *, uint32_t *, std::string &err) {
    err = "prefilter kernels not built yet";
    return FSB200_E_STATE;
}
}}
