// BSS Eval v4 (reference: bsseval_v4.py:79-300 bss_eval and its helpers :449-617) and the ideal-mask oracle systems on
// mel spectrograms (reference: oracle_systems.py:264-350) on the device.  See bsseval.cu.
#pragma once
#include "common.cuh"

namespace asep {

// refs / ests: [nsrc, nsampl] float64 (mono images).  Distortion filters of length L are estimated on samples
// [f0, f1) (bsseval_v4.py:219-242 compute_GsfC / compute_Cj), the decomposition and the energy ratios are taken on each
// window [w0[t], w1[t]) (:245-279).  out [4, nsrc, nsrc, nwin] float64 = (SDR, ISR, SIR, SAR)[jtrue][jest][t]
// (s_r of :214).  sources_version: the bss_eval_sources criteria (:575-585).  win0 / win1: HOST arrays.
void bss_eval_core(const double* refs, const double* ests, int nsrc, long long nsampl, int L, long long f0, long long f1,
                   const long long* win0, const long long* win1, int nwin, int sources_version, double* out, cudaStream_t s);

// IRM_melspec / IBM_melspec (oracle_systems.py:264-350): mixture [P], sources [nsrc, P] -> estimates [nsrc, P]
void launch_ideal_mask(const float* mixture, const float* sources, float* estimates, int nsrc, long long P, int binary,
                       float theta, cudaStream_t s);

}  // namespace asep
