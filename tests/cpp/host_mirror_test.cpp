// host_mirror_test.cpp -- the reference's own core unit tests for this path, re-expressed against the C++ host mirror
// (openfhe-development_b200/host/dcrtpoly_b200.h).  Compiled by the CPU test tier (link check), run by the GPU tier.
//   UTTransform.CRT_polynomial_mult            src/core/unittest/UnitTestTransform.cpp:52-90
//   UTNTT.switch_format_simple_single_crt      src/core/unittest/UnitTestNTT.cpp:53-90
//   UTNTT.switch_format_simple_double_crt      src/core/unittest/UnitTestNTT.cpp:92-119
//   UTDCRTPoly arithmetic / error behaviour    src/core/unittest/UnitTestDCRTElements.cpp
#include <cstdio>
#include <memory>
#include <vector>

#include "dcrtpoly_b200.h"

using namespace b200;

static int failures = 0;
#define EXPECT(cond)                                                       \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);    \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

static std::shared_ptr<Params> make_params(uint32_t m, uint32_t bits, uint32_t towers) {
    std::vector<uint64_t> q{ofhe_last_prime(bits, m)}, r;
    for (uint32_t i = 1; i < towers; ++i)
        q.push_back(ofhe_previous_prime(q.back(), m));
    for (auto x : q)
        r.push_back(ofhe_root_of_unity(m, x));
    return std::make_shared<Params>(0, m / 2, q, r);
}
static DCRTPoly from_values(std::shared_ptr<Params> p, const std::vector<uint64_t>& v, Format f) {
    std::vector<uint32_t> t(p->GetModuli().size());
    for (size_t i = 0; i < t.size(); ++i)
        t[i] = (uint32_t)i;
    DCRTPoly x(p, t, f, true);
    for (size_t i = 0; i < t.size(); ++i)
        x.SetElementAtIndex(i, v);  // DCRTPoly = {list} puts the same small integers in every tower
    return x;
}

int main() {
    {  // CRT_polynomial_mult: q = 113, m = 8
        auto p = std::make_shared<Params>(0, 4, std::vector<uint64_t>{113}, std::vector<uint64_t>{ofhe_root_of_unity(8, 113)});
        DCRTPoly a = from_values(p, {1, 2, 4, 1}, COEFFICIENT), b = a;
        a.SwitchFormat();
        b.SwitchFormat();
        DCRTPoly c = a * b;
        c.SwitchFormat();
        EXPECT((c.GetElementAtIndex(0) == std::vector<uint64_t>{94, 109, 11, 18}));
    }
    {  // switch_format_simple_single_crt
        struct { uint32_t m; std::vector<uint64_t> v; } cases[] = {
            {16, {431, 3414, 1234, 7845, 2145, 7415, 5471, 8452}}, {8, {4127, 9647, 1987, 5410}}, {4, {9647, 1987}}};
        for (auto& cs : cases) {
            auto p     = make_params(cs.m, 16, 1);
            DCRTPoly x = from_values(p, cs.v, COEFFICIENT), clone = x;
            x.SwitchFormat();
            EXPECT(x.GetFormat() == EVALUATION);
            x.SwitchFormat();
            EXPECT(x == clone);
        }
    }
    {  // switch_format_simple_double_crt + arithmetic identities
        auto p      = make_params(16, 28, 2);
        DCRTPoly x1 = from_values(p, {431, 3414, 1234, 7845, 2145, 7415, 5471, 8452}, COEFFICIENT);
        DCRTPoly x2 = from_values(p, {4127, 9647, 1987, 5410, 6541, 7014, 9741, 1256}, COEFFICIENT);
        DCRTPoly c1 = x1, c2 = x2;
        x1.SwitchFormat();
        x2.SwitchFormat();
        DCRTPoly s = x1 + x2, d = s - x2;
        EXPECT(d == x1);
        EXPECT((x1 + x1.Negate()) == (x2 - x2));
        DCRTPoly pr = x1 * x2, pr2 = x2;
        pr2 *= x1;
        EXPECT(pr == pr2);
        x1.SwitchFormat();
        x2.SwitchFormat();
        EXPECT(x1 == c1 && x2 == c2);
        // error behaviour mirrors OPENFHE_THROW sites: Times in COEFFICIENT format, tower mismatch
        bool threw = false;
        try {
            (void)(x1 * x2);
        }
        catch (const OpenFHEException&) {
            threw = true;
        }
        EXPECT(threw);
        DCRTPoly y = x1;
        y.DropLastElement();
        threw = false;
        try {
            (void)(x1 + y);
        }
        catch (const OpenFHEException&) {
            threw = true;
        }
        EXPECT(threw);
        threw = false;
        try {
            y.DropLastElement();
        }
        catch (const OpenFHEException&) {
            threw = true;
        }
        EXPECT(threw);
    }
    {  // hybrid key switch objects construct and run on a small CKKS-shaped ring; KeySwitchCore(0) == 0
        uint64_t moduli[96], roots[96];
        uint32_t shape[4];
        check(ofhe_ckks_paramgen(1 << 12, 4, 50, 60, 2, 0, moduli, roots, shape));
        uint32_t sizeQ = shape[0], sizeP = shape[1];
        std::vector<uint64_t> q(moduli, moduli + sizeQ + sizeP), r(roots, roots + sizeQ + sizeP);
        auto p = std::make_shared<Params>(0, 1 << 12, q, r);
        KeySwitchHYBRID ks(p, sizeQ, sizeP, shape[2]);
        EXPECT(ks.GetNumPerPartQ() == shape[3]);
        ofhe_evk* evk = nullptr;
        check(ofhe_evk_create_random(ks.handle(), 7, &evk));
        std::vector<uint32_t> t(sizeQ);
        for (uint32_t i = 0; i < sizeQ; ++i)
            t[i] = i;
        DCRTPoly zero(p, t, EVALUATION, true);
        auto digits = ks.EvalKeySwitchPrecomputeCore(zero);
        EXPECT(digits.size() == 2 && digits[0].GetNumOfElements() == sizeQ + sizeP);
        auto ba = ks.KeySwitchCore(zero, evk);
        EXPECT(ba.first == zero && ba.second == zero);
        // KeySwitchCore(a) == EvalFastKeySwitchCore(EvalKeySwitchPrecomputeCore(a)) == KeySwitchDown(EvalFastKeySwitchCoreExt(..))
        // on a non-zero polynomial (the reference composes them exactly this way, keyswitch-hybrid.cpp:308-312,381-400)
        DCRTPoly a(p, t, EVALUATION, true);
        {
            std::vector<uint64_t> h(1u << 12);
            for (uint32_t i = 0; i < sizeQ; ++i) {
                for (uint32_t j = 0; j < (1u << 12); ++j)
                    h[j] = (0x9E3779B97F4A7C15ull * (i * 4099ull + j + 1)) % q[i];
                a.SetElementAtIndex(i, h);
            }
        }
        auto dg   = ks.EvalKeySwitchPrecomputeCore(a);
        auto core = ks.KeySwitchCore(a, evk);
        auto fast = ks.EvalFastKeySwitchCore(dg, evk, sizeQ);
        EXPECT(core.first == fast.first && core.second == fast.second);
        auto ext  = ks.EvalFastKeySwitchCoreExt(dg, evk, sizeQ);
        EXPECT(ext.first.GetNumOfElements() == sizeQ + sizeP);
        auto down = ks.KeySwitchDown(ext.first, ext.second);
        EXPECT(core.first == down.first && core.second == down.second);
        auto lifted = ks.KeySwitchExt(a, a, false);
        auto back   = ks.KeySwitchDown(lifted.first, lifted.second);   // ModDown([c * P | 0]) == c, and 0 for the skipped c0
        EXPECT(back.second == a && back.first == zero);
        ofhe_evk_destroy(evk);
    }
    if (failures == 0)
        std::printf("host mirror: all checks passed\n");
    return failures ? 1 : 0;
}
