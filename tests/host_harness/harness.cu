// TEST ONLY: runs the serial (one-thread) pieces of the device source on the host so that the
// divide-and-conquer control flow can be checked against the oracle without a GPU.  The library
// itself never executes these functions on the CPU.
#include <vector>

#include "../../indelminer_b200/csrc/kernels.cuh"
#include "../../indelminer_b200/csrc/band_dp.cuh"

using namespace indelgpu;

extern "C" int hh_band_align(const int* prm, const uint8_t* read, int M, const uint8_t* win, int N,
                             int low, int up, int* out10, uint32_t* cig, int* script, int script_cap)
{
    DevParams P;
    P.k = prm[0]; P.g = prm[1]; P.maxdel = prm[2]; P.ethr = prm[3];
    P.match = prm[4]; P.mismatch = prm[5]; P.G = prm[6]; P.H = prm[7]; P.kmask = 0;
    const int lo = low > -M ? low : -M, hi = up < N ? up : N;
    const int band = hi - lo + 1;
    if (band < 2) return -1;
    const int mb = 2 * band;
    std::vector<int> scratch((size_t)band_scratch_ints(mb, M));
    DcFrame st[kDcFrames];
    align_banded_serial<1>(P, IArr<1>{scratch.data()}, IArr<1>{scratch.data() + 4 * (mb + 4)}, mb, M, st, read, M, win, N, lo, hi, cig, out10);
    const int* S = scratch.data() + 4 * (mb + 4) + 4 * (M + 2);
    for (int t = 0; t < out10[9] && t < script_cap; t++) script[t] = S[t];
    return 0;
}

// the same through the lane-interleaved view the thread-per-alignment kernel uses (lane `lane` of 32)
extern "C" int hh_band_align_interleaved(const int* prm, const uint8_t* read, int M, const uint8_t* win, int N,
                                         int low, int up, int* out10, uint32_t* cig, int* script, int script_cap, int lane)
{
    DevParams P;
    P.k = prm[0]; P.g = prm[1]; P.maxdel = prm[2]; P.ethr = prm[3];
    P.match = prm[4]; P.mismatch = prm[5]; P.G = prm[6]; P.H = prm[7]; P.kmask = 0;
    const int lo = low > -M ? low : -M, hi = up < N ? up : N;
    const int band = hi - lo + 1;
    if (band < 1) return -1;
    const int mb = 2 * band;
    std::vector<int> scratch((size_t)band_scratch_ints(mb, M) * 32, 0x5A5A5A5A);
    DcFrame st[kDcFrames];
    const IArr<32> base{scratch.data() + lane};
    align_banded_serial<32>(P, base, base + 4 * (mb + 4), mb, M, st, read, M, win, N, lo, hi, cig, out10);
    const IArr<32> S = base + 4 * (mb + 4) + 4 * (M + 2);
    for (int t = 0; t < out10[9] && t < script_cap; t++) script[t] = S[t];
    // the other 31 lanes' elements must be untouched
    for (size_t i = 0; i < scratch.size(); i++) if ((int)(i & 31) != lane && scratch[i] != 0x5A5A5A5A) return -2;
    return 0;
}
