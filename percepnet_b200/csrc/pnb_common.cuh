// pnb_common.cuh -- shared constants and device-side table block of the B200 PercepNet hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pnb {

constexpr int kFrame = 480;        // FRAME_SIZE              (ref src/denoise.cpp:19)
constexpr int kWin = 960;          // WINDOW_SIZE             (:20)
constexpr int kBands = 34;         // NB_BANDS                (:35)
constexpr int kFeat = 70;          // NB_FEATURES             (:40)
constexpr int kHist = 5760;        // COMB_BUF_SIZE           (:32)
constexpr int kKeep = kHist - kFrame;  // samples of history carried between calls (5280)
constexpr int kBins = 400;         // bins below the last ERB border; bins 400..480 never reach the output
                                   // (SURVEY.md App. C.1), so spectra are stored/processed for 0..399 only
constexpr int kOffAnalysis = 2400; // analysis window   = line[2400 .. 3360)   
constexpr int kOffPitch = 1632;    // pitch buffer      = line[1632 .. 3360)
constexpr int kOffLook = 4800;     // look-ahead window = line[4800 .. 5760)
constexpr int kLineCalls = 8;      // at most this many calls between two moves of the history to the start of its row
constexpr int kLp = 864;           // decimated pitch buffer length
constexpr int kMaxPeriod = 768, kMinPeriod = 60;

// Constant tables, built on the host with the reference's expressions (pnb_engine.cu) and kept in
// global memory; kernels stage what they need into shared memory.
struct Tables {
  float half_window[kFrame];   // denoise.cpp:191-192
  float comb_w[8];             // denoise.cpp:200-206 (7 used)
  float2 tw[kWin];             // kiss_fft.cpp:415-419
  float frac[kBins];           // (float)j/band_size of bin's position in its band (denoise.cpp:99)
  float omf[kBins];            // 1 - frac
  short band_of[kBins];        // band index b such that border[b] <= bin < border[b+1]
  short border[kBands + 2];
  float tansig[208];           // tansig_table.h (201 used)
};

}  // namespace pnb
