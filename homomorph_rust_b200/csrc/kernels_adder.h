// kernels_adder.h — host interface of kernels_adder.cu (the dynamically scheduled thread-per-value adder chain).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "gf2_blocks.cuh"

namespace hmk {

constexpr uint32_t ADDER_MAX_PHASES = 8;

// Work queue of one launch.  `counter` and `done[ngroups]` live in device memory and must be zero when the kernel starts.
struct AdderSched {
    uint32_t *counter;                 // next work unit (unit = phase * ngroups + group)
    uint32_t *done;                    // done[g] = number of phases of group g that are complete
    uint32_t ngroups;                  // ceil(n / 32)
    uint32_t nphases;                  // phases actually used (<= ADDER_MAX_PHASES)
    uint32_t kb[ADDER_MAX_PHASES + 1]; // phase p covers chain iterations [kb[p], kb[p+1])
};

// u32 words of scheduler state needed for n values (counter + one flag per group of 32 values)
size_t adder_chain_sched_words(uint64_t n);
// splits the L iterations of the chain into at most want_phases ranges of roughly equal work (fills kb and nphases)
void adder_chain_plan(uint32_t L, int wd, uint32_t want_phases, AdderSched *sc);
// wd = D / 32 in {4, 8}; variant = 10 * (window in shared memory) + CTAs per SM
cudaError_t launch_adder_chain(int wd, int variant, const uint64_t *A, const uint64_t *B, uint64_t *O, uint64_t n, uint32_t L, const Layout &lo,
                               const AdderSched &sc, int sm_count, cudaStream_t stream);

} // namespace hmk
