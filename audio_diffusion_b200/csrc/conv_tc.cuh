// Parameter block of the tcgen05 implicit-GEMM convolution (conv_tc.cu). Host + device.
#pragma once
#include "common.cuh"

namespace b200ad {

constexpr int CONV_NT = 128;        // output-channel tile = MMA N
constexpr int CONV_TM = 128;        // pixels per MMA tile = MMA M
constexpr int CONV_MAXG = 4;        // pixel tiles per work item (4 x 128 fp32 columns = all of TMEM)
constexpr int CONV_STAGES = 3;
constexpr int CONV_MAXSEG = 4;
constexpr int CONV_MAXTAPS = 9;
constexpr int CONV_A_STAGE = 25088;              // >= 6 windows * 2 planes * 130 px * 16 B
constexpr int CONV_B_TAP = 16 * CONV_NT * 2;     // one tap, 16 input channels: 4096 B
constexpr int CONV_B_STAGE = CONV_MAXTAPS * CONV_B_TAP;
constexpr int CONV_STAGE_BYTES = CONV_A_STAGE + CONV_B_STAGE;
constexpr int CONV_SMEM_BYTES = CONV_STAGES * CONV_STAGE_BYTES + 1024;

// One K-segment: a source tensor (PF8) with its tap set and packed weights. A 3x3 conv is one
// segment with 9 taps; a fused 1x1 shortcut adds one 1-tap segment per shortcut source; a stride-2
// conv is four segments (one per input parity plane).
struct ConvSeg {
  const __nv_bfloat16* src;    // PF8 activations, same geometry as the output
  const __nv_bfloat16* wpack;  // [cout/128][ksteps][ntaps][2 k8][16 n8][8][8] bf16
  long long img_stride;        // elements between images of src (= C/8 * PL * 8)
  int ksteps;                  // input channels / 16
  int ntaps;
  int ht, hb, hl, hr;          // halo rows above / below, pixels left / right
  signed char dh[CONV_MAXTAPS];
  signed char dw[CONV_MAXTAPS];
};

struct ConvParams {
  ConvSeg seg[CONV_MAXSEG];
  int nseg;
  int N, H, W, Wp, lead, PL;
  int wide;             // 1: W % 128 == 0 (tiles = 128-px row pieces stacked over 4 rows); 0: flat
  int groups_per_img;
  int ntiles_n;         // cout / 128
  int total_work;       // N * groups_per_img * ntiles_n
  int cout;
  __nv_bfloat16* out;           // PF8, cout channels
  const float* bias;            // [cout]
  const float* temb;            // optional per-sample additive term [N][temb_stride] (offset applied)
  int temb_stride;
  const __nv_bfloat16* res;     // optional residual, PF8 with cout channels
  stat_t* stats;                // optional [N][cout/4][2] running (sum, sumsq) of the stored output
  int dbg;                      // B200AD_CONV_DBG bit flags (timing experiments only): 1 no stats, 2 no stores, 4 no tmem ld, 8 no epilogue
};

cudaError_t launch_conv_tc(const ConvParams& p, int num_sms, cudaStream_t stream);

}  // namespace b200ad
