// probes.h — host interface of probes.cu (pipe-rate probes for the integer rooflines).
#pragma once
#include <cuda_runtime.h>

namespace hmk {
struct PipeProbe {
    double warp_instr_per_s;  // whole device, from CUDA-event time
    double cycle_counter_mhz; // clock64() ticks / elapsed time: a diagnostic, NOT the SM clock (see probes.cu)
    double ms;                // duration of the measured launch (fastest of three)
    int warps_per_sm;                 // resident warps per SM during the probe
};
// out[0] = IMAD.WIDE (register x register), out[1] = LOP3 (three registers), out[2] = 1 IMAD.WIDE : 2 LOP3 mix
cudaError_t measure_pipe_peaks(int sm_count, cudaStream_t stream, double min_ms, PipeProbe out[3]);
} // namespace hmk
