// TEST ONLY: runs the serial (one-thread) pieces of the device source on the host so that the
// divide-and-conquer control flow can be checked against the oracle without a GPU.  The library
// itself never executes these functions on the CPU.
#include <cstring>
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

// ---------------------------------------------------------------------------------------------------
// the 16-bit two-pass support check (indel_support_pack.cuh): one segment of SEG lanes emulated in lockstep,
// shuffles replaced by array reads, the per-lane row step and the walk back being the device source itself
// ---------------------------------------------------------------------------------------------------
#include "../../indelminer_b200/csrc/indel_support_pack.cuh"

template <int CPL>
static void hh_pack_run(int SEG, const uint8_t* t1a, int len1a, const uint8_t* t2a, int len2a,
                        const uint8_t* t1b, int len1b, const uint8_t* t2b, int len2b, int* out6)
{
    std::vector<PackLane<CPL>> L((size_t)SEG);
    for (int li = 0; li < SEG; li++) L[li].init(t1a, len1a, t1b, len1b, li * CPL);
    const int len1 = std::max(len1a, len1b), len2 = std::max(len2a, len2b);
    const int nl = (len1 + CPL - 1) / CPL;
    const int steps = (len1 > 0 && len2 > 0) ? len2 + nl - 1 : 0;
    const int steps_pad = (steps + 8) & ~7;
    const size_t plane = (size_t)steps_pad * 32;
    std::vector<uint32_t> dirs(plane * 3, 0xDEADBEEFu);
    std::vector<unsigned> qcur((size_t)SEG, 0), qprev((size_t)SEG, 0), give((size_t)SEG), ov((size_t)SEG), oe((size_t)SEG);
    for (int s = 0; s < steps; s++) {
        const int sl = s & (SEG - 1);
        for (int li = 0; li < SEG; li++) {
            if (sl == 0) {
                const int r = s + li;
                const unsigned qa = r < len2a ? (unsigned)up_case(t2a[r]) << 4 : (unsigned)kPkPastQuery;
                const unsigned qb = r < len2b ? (unsigned)up_case(t2b[r]) << 4 : (unsigned)kPkPastQuery;
                qprev[li] = qcur[li];
                qcur[li] = pk2(0u - qa, 0u - qb);
            }
            give[li] = li <= sl ? qcur[li] : qprev[li];
            ov[li] = L[li].outV; oe[li] = L[li].outE;
        }
        for (int li = 0; li < SEG; li++) {
            const unsigned nb = give[(s - li) & (SEG - 1)];
            const unsigned inV = li ? ov[li - 1] : ov[li], inE = li ? oe[li - 1] : oe[li];   // shfl_up keeps lane 0's own value
            const int i = s - li + 1;
            if (i >= 1 && i <= len2 && li < nl) {
                unsigned w[3];
                L[li].row(i, li * CPL, li == 0, inV, inE, nb, 1u, w);
                for (int p = 0; p < pack_planes(CPL); p++) dirs[p * plane + (size_t)li * steps_pad + s] = w[p];
            }
        }
    }
    unsigned ka = 0, kb = 0;
    for (int li = 0; li < SEG; li++) { ka = std::max(ka, L[li].best_key(0, li * CPL)); kb = std::max(kb, L[li].best_key(1, li * CPL)); }
    out6[0] = out6[1] = out6[3] = out6[4] = 0; out6[2] = out6[5] = 1;
    if (ka) pack_walk_back(dirs.data(), plane, CPL, steps_pad, 0, 0x7FFFFu - (ka & 0x7FFFFu), t1a, t2a, out6[0], out6[1], out6[2]);
    if (kb) pack_walk_back(dirs.data(), plane, CPL, steps_pad, 1, 0x7FFFFu - (kb & 0x7FFFFu), t1b, t2b, out6[3], out6[4], out6[5]);
}

// out6 = subs, indels, aligned of pair A, then of pair B.  Returns the columns per lane used, < 0 if the pair does not fit.
extern "C" int hh_support_pack(int SEG, const uint8_t* t1a, int len1a, const uint8_t* t2a, int len2a,
                               const uint8_t* t1b, int len1b, const uint8_t* t2b, int len2b, int* out6)
{
    const int need = (std::max(len1a, len1b) + SEG - 1) / SEG;
    if (need > 16 || (SEG != 8 && SEG != 16 && SEG != 32)) return -1;
#define PACK(C) hh_pack_run<C>(SEG, t1a, len1a, t2a, len2a, t1b, len1b, t2b, len2b, out6); return C
    switch ((need + 1) >> 1) {
        case 0: case 1: PACK(2);
        case 2: PACK(4);
        case 3: PACK(6);
        case 4: PACK(8);
        case 5: PACK(10);
        case 6: PACK(12);
        case 7: PACK(14);
        default: PACK(16);
    }
#undef PACK
}

// ---------------------------------------------------------------------------------------------------
// the serial one-diagonal alignment of align1_tasks_kernel (task_kernels.cuh)
// ---------------------------------------------------------------------------------------------------
#include "../../indelminer_b200/csrc/task_kernels.cuh"

extern "C" int hh_diag1(const int* prm, const uint8_t* read, int M, const uint8_t* win, int N, int low, int up, int* out9,
                        uint32_t* cig, int cig_cap)
{
    DevParams P;
    P.k = prm[0]; P.g = prm[1]; P.maxdel = prm[2]; P.ethr = prm[3];
    P.match = prm[4]; P.mismatch = prm[5]; P.G = prm[6]; P.H = prm[7]; P.kmask = 0;
    const int lo = low > -M ? low : -M, hi = up < N ? up : N;
    if (hi - lo + 1 != 1) return -1;
    // padded copies at every byte alignment (the kernel reads whole words around the rows it needs; device buffers are padded)
    std::vector<uint32_t> rb((size_t)M / 4 + 16, 0xA5A5A5A5u), wb((size_t)N / 4 + 16, 0x5A5A5A5Au);
    uint8_t* r = reinterpret_cast<uint8_t*>(rb.data()) + 4 + (M % 4);
    uint8_t* w = reinterpret_cast<uint8_t*>(wb.data()) + 4 + ((N + M) % 4);
    memcpy(r, read, (size_t)M); memcpy(w, win, (size_t)N);
    align_diag1_serial(P, r, M, w, N, lo, cig, cig_cap, out9);
    return 0;
}
