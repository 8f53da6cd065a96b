// gat_codes.cpp -- host PRN generators behind gat_gen_code().
//
// The reference takes its chip tables from GNSSSignals.jl (`GPSL1().codes`, `GPSL5().codes`,
// /root/reference/src/benchmarks.jl:92-93, :829-837); that package is not vendored, so the
// product regenerates the tables from the public ICDs (IS-GPS-200 sec. 3.3.2.3, IS-GPS-705
// sec. 3.3.2.2).  Written as bit-packed Galois-free Fibonacci registers, independently of the
// array-based generator in oracle/oracle.c, so the two implementations cross-check each other.
#include <cstdint>
#include <initializer_list>

#include "../../include/gat.h"

namespace {

// G2 output-tap selections (1-based stage numbers), IS-GPS-200 Table 3-Ia, PRN 1..37
const uint8_t kCaTaps[37][2] = {
    {2, 6},  {3, 7},  {4, 8},  {5, 9},  {1, 9},  {2, 10}, {1, 8},  {2, 9},  {3, 10}, {2, 3},
    {3, 4},  {5, 6},  {6, 7},  {7, 8},  {8, 9},  {9, 10}, {1, 4},  {2, 5},  {3, 6},  {4, 7},
    {5, 8},  {6, 9},  {1, 3},  {4, 6},  {5, 7},  {6, 8},  {7, 9},  {8, 10}, {1, 6},  {2, 7},
    {3, 8},  {4, 9},  {5, 10}, {4, 10}, {1, 7},  {2, 8},  {4, 10}};

// XB code advance of the I5 codes, IS-GPS-705 Table 3-Ia, PRN 1..37
const uint16_t kL5IAdvance[37] = {266,  365,  804,  1138, 1509, 1559, 1756, 2084, 2170, 2303,
                                  2527, 2687, 2930, 3471, 3940, 4132, 4332, 4924, 5343, 5443,
                                  5641, 5816, 5898, 5918, 5955, 6243, 6345, 6477, 6518, 6875,
                                  7168, 7187, 7329, 7577, 7720, 7777, 8057};

// Fibonacci LFSR with stage i held in bit (i-1).  `taps` has bit (i-1) set for every stage
// that feeds the XOR going into stage 1; the output is stage `n`.
struct Lfsr {
    uint32_t state, taps;
    int n;
    int stage(int i) const { return (state >> (i - 1)) & 1u; }
    int out() const { return stage(n); }
    void clock()
    {
        const uint32_t fb = __builtin_parity(state & taps);
        state = ((state << 1) | fb) & ((1u << n) - 1u);
    }
};

uint32_t stage_mask(std::initializer_list<int> stages)
{
    uint32_t m = 0;
    for (int s : stages) m |= 1u << (s - 1);
    return m;
}

int gen_l1(int prn, int8_t *out)
{
    Lfsr g1{0x3FFu, stage_mask({3, 10}), 10};
    Lfsr g2{0x3FFu, stage_mask({2, 3, 6, 8, 9, 10}), 10};
    const int s1 = kCaTaps[prn - 1][0], s2 = kCaTaps[prn - 1][1];
    for (int k = 0; k < 1023; ++k) {
        const int bit = g1.out() ^ g2.stage(s1) ^ g2.stage(s2);
        out[k] = static_cast<int8_t>(1 - 2 * bit);
        g1.clock();
        g2.clock();
    }
    return 1023;
}

int gen_l5i(int prn, int8_t *out)
{
    Lfsr xa{0x1FFFu, stage_mask({9, 10, 12, 13}), 13};
    Lfsr xb{0x1FFFu, stage_mask({1, 3, 4, 6, 7, 8, 12, 13}), 13};
    for (int k = 0; k < kL5IAdvance[prn - 1]; ++k) xb.clock();
    for (int k = 0; k < 10230; ++k) {
        if (k == 8190) xa.state = 0x1FFFu;  // XA is short-cycled by one chip
        const int bit = xa.out() ^ xb.out();
        out[k] = static_cast<int8_t>(1 - 2 * bit);
        xa.clock();
        xb.clock();  // XB runs its natural 8191 period through the 1 ms epoch
    }
    return 10230;
}

}  // namespace

extern "C" int gat_gen_code(int system_id, int prn, int8_t *out, int cap)
{
    if (!out || prn < 1 || prn > 37) return GAT_ERR_INVALID;
    if (system_id == GAT_GPSL1) return cap >= 1023 ? gen_l1(prn, out) : GAT_ERR_INVALID;
    if (system_id == GAT_GPSL5) return cap >= 10230 ? gen_l5i(prn, out) : GAT_ERR_INVALID;
    return GAT_ERR_UNSUPPORTED;
}
