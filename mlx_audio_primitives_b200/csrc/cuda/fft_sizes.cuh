// The n_fft values with a compiled Stockham plan -- the ONE list every dispatcher, the build script and the host
// emulation read (the plans themselves: fft_plans_list.cuh).  Powers of two 32..8192 plus the 2^a 3^b 5^c sizes
// speech / music front ends use (10 / 25 / 30 / 50 / 75 ms windows at 16 / 24 / 32 / 48 kHz and friends).
#pragma once

// X(n_fft) over every planned size
#define MLXA_FOR_EACH_NFFT(X) \
    X(32) X(64) X(128) X(256) X(400) X(480) X(512) X(600) X(800) X(1000) X(1024) X(1200) X(1600) X(2000) X(2048) X(3072) X(4096) X(8192)
// the power-of-two sizes the autocorrelation pitch kernels are instantiated for (acf_inst.cu)
#define MLXA_FOR_EACH_ACF_NFFT(X) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)
