// Geometry of the tensor-core schedule kernel shared by the kernel (vbn_schedule_tc.cuh) and the host launch
// code (vbn_cuda.cu).  Keep in sync with cpds.py pack_mlp_tc.
#pragma once

#ifdef __CUDACC__
#define VBN_HD __host__ __device__
#else
#define VBN_HD
#endif

namespace vbn {
namespace tc {

constexpr int kWgThreads = 128;
constexpr int kMaxTiles = 5;       // 128-row tiles resident per CTA (TMEM: 96 columns each + 8)
constexpr int kTmemCols = 512;
constexpr int kColsPerTile = 96;
constexpr int kColD = 0, kColAhi = 32, kColAlo = 64;
constexpr int kHidden = 32;
constexpr int kWbufBytes = 30720;  // largest weight image (K1 = 32, N3 = 32): 2 * 4 * 3 * 32 * 40 B; the ring's slot size is
                                   // the largest image of the PROGRAM (weights + descriptor tail), rounded up to 128 B
constexpr int kBiasK = 8;          // every MMA layer carries its bias as one extra K = 8 step (column K of the image)
constexpr int kL1PlainBytes = 640; // first layer on the FP32 pipe (Dp <= 4): W1^T[4][32], b1[32] fp32
constexpr int kMaxBufs = 8;
constexpr int kCtrlBytes = 256;    // tmem address, up to 13 mbarriers, the ring producer's counters

// bytes of one weight image.  k1 > 0: W1 hi/lo [32][K1+8], W2 hi/lo [32][40], W3 hi/lo [N3][40] (bias = column K);
// k1 == 0: the plain first-layer block instead of the W1 images.
VBN_HD constexpr int blob_bytes(int k1, int n3) {
  return (k1 == 0 ? kL1PlainBytes : 4 * 2 * kHidden * (k1 + kBiasK)) +
         4 * 2 * (kHidden * (kHidden + kBiasK) + n3 * (kHidden + kBiasK));
}

}  // namespace tc
}  // namespace vbn
