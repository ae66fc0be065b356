// keyswitch-hybrid-b200.cpp -- the four hot members of lbcrypto::KeySwitchHYBRID rebound to the fused device entry
// points of libofhe_b200.so (SURVEY.md section 8 b: "src/pke KeySwitchHYBRID rebound to the device path").
//
// Replaces, with identical results (bit for bit) and identical signatures,
//   KeySwitchHYBRID::KeySwitchCore               src/pke/lib/keyswitch/keyswitch-hybrid.cpp:308-312
//   KeySwitchHYBRID::EvalKeySwitchPrecomputeCore  :314-379   (digit decomposition + ModUp)
//   KeySwitchHYBRID::EvalFastKeySwitchCore        :381-400   (inner product + ModDown)
//   KeySwitchHYBRID::EvalFastKeySwitchCoreExt     :402-435   (inner product with the evaluation key, in QlP)
// Everything else of the class (key generation, KeySwitchInPlace/Ext/Down glue) stays the reference's object code:
// oracle/Makefile.hal compiles the reference's keyswitch-hybrid.cpp untouched, marks exactly these four symbols weak
// in a COPY of its object file (objcopy --weaken-symbol) and links this translation unit next to it, so every caller
// inside src/pke -- EvalMult's relinearisation, EvalRotate, EvalFastRotation, bootstrapping -- lands here.
//
// With the DCRTPoly backend alone each of these functions is a dozen separate DCRTPoly calls plus a tower-by-tower inner
// product ON THE HOST (the reference multiplies NativePoly towers there); here a key switch is the fused device pipeline
// (ofhe_hks_*: INTT -> base conversion -> NTT -> inner product with the DEVICE-RESIDENT evaluation key -> ModDown) from
// the operand's device slab to device-resident results -- the digits of a hoisted rotation are views of one slab that
// the next call consumes in place.  Evaluation keys are uploaded once per EvalKey object and stay resident (a weak_ptr
// notices when the host object dies).
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "key/evalkeyrelin.h"
#include "keyswitch/keyswitch-hybrid.h"
#include "lattice/hal/b200/dcrtpoly-b200-impl.h"
#include "schemerns/rns-cryptoparameters.h"

namespace lbcrypto {

namespace {
using b200::Check;
using b200::detail::Op;

struct HksCtx {
    std::weak_ptr<CryptoParametersRNS> alive;
    ofhe_ring* ring = nullptr;
    ofhe_hks* hks   = nullptr;
    uint32_t N = 0, sizeQ = 0, sizeP = 0;
};
struct EvkCtx {
    std::weak_ptr<EvalKeyImpl<DCRTPoly>> alive;
    ofhe_evk* evk = nullptr;
};
std::mutex g_mu;
std::map<const CryptoParametersRNS*, HksCtx> g_hks;
std::map<const EvalKeyImpl<DCRTPoly>*, EvkCtx> g_evk;

// the key-switch context of a crypto-parameter object: ring over Q u P, digit layout, per-level tables (the library
// derives every table of CryptoParametersRNS::PrecomputeCRTTables from the moduli; tests pin them to the getters)
const HksCtx& HksOf(const std::shared_ptr<CryptoParametersRNS>& cp) {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_hks.find(cp.get());
    if (it != g_hks.end() && it->second.alive.lock().get() == cp.get())
        return it->second;
    if (it != g_hks.end()) {  // a dead object's address was reused
        ofhe_hks_destroy(it->second.hks);
        g_hks.erase(it);
    }
    HksCtx h;
    h.alive           = cp;
    const auto paramsQ = cp->GetElementParams();
    const auto paramsP = cp->GetParamsP();
    h.N                = paramsQ->GetRingDimension();
    h.sizeQ            = (uint32_t)paramsQ->GetParams().size();
    h.sizeP            = (uint32_t)paramsP->GetParams().size();
    std::vector<uint64_t> mod, root;
    b200::detail::CollectParams(paramsQ, h.sizeQ, mod, root);
    b200::detail::CollectParams(paramsP, h.sizeP, mod, root);
    h.ring = b200::Runtime::Get().Ring(h.N, mod, root);
    Check(ofhe_hks_create(h.ring, h.sizeQ, h.sizeP, cp->GetNumberOfQPartitions(), &h.hks), "ofhe_hks_create");
    if (ofhe_hks_num_per_part(h.hks) != cp->GetNumPerPartQ())
        OPENFHE_THROW("b200 backend: digit size differs from CryptoParametersRNS::GetNumPerPartQ()");
    return g_hks.emplace(cp.get(), h).first->second;
}

// device-resident copy of an evaluation key (A and B vectors, [dnum][sizeQ+sizeP][N] EVALUATION)
const ofhe_evk* EvkOf(const EvalKey<DCRTPoly>& ek, const HksCtx& h) {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_evk.find(ek.get());
    if (it != g_evk.end() && it->second.alive.lock().get() == ek.get())
        return it->second.evk;
    if (it != g_evk.end()) {
        ofhe_evk_destroy(it->second.evk);
        g_evk.erase(it);
    }
    const auto& av      = ek->GetAVector();
    const auto& bv      = ek->GetBVector();
    const uint32_t dnum = (uint32_t)av.size(), L = h.sizeQ + h.sizeP;
    std::vector<uint64_t> a((size_t)dnum * L * h.N), b(a.size());
    for (uint32_t j = 0; j < dnum; ++j) {
        if (av[j].GetNumOfElements() != L || bv[j].GetNumOfElements() != L || av[j].GetFormat() != Format::EVALUATION)
            OPENFHE_THROW("b200 backend: evaluation key is not a HYBRID key over the full QP basis in EVALUATION format");
        for (uint32_t i = 0; i < L; ++i) {
            std::memcpy(&a[((size_t)j * L + i) * h.N], &av[j].GetElementAtIndex(i).GetValues()[0], (size_t)h.N * 8);
            std::memcpy(&b[((size_t)j * L + i) * h.N], &bv[j].GetElementAtIndex(i).GetValues()[0], (size_t)h.N * 8);
        }
    }
    EvkCtx e;
    e.alive = ek;
    Check(ofhe_evk_create(h.hks, a.data(), b.data(), &e.evk), "ofhe_evk_create");
    return g_evk.emplace(ek.get(), e).first->second.evk;
}

// n device-resident polynomials over `params` (EVALUATION): views [k][towers][N] of ONE slab (they keep it alive)
std::shared_ptr<std::vector<DCRTPoly>> Views(Op& op, const b200::detail::Slab& slab, uint32_t n,
                                             const std::shared_ptr<DCRTPoly::Params>& params) {
    auto result = std::make_shared<std::vector<DCRTPoly>>();
    result->reserve(n);
    const size_t stride = params->GetParams().size() * (size_t)op.N;
    for (uint32_t k = 0; k < n; ++k)
        result->push_back(DCRTPoly::FromDevice(params, Format::EVALUATION,
                                               b200::detail::Slab(slab, slab.get() + k * stride)));
    return result;
}
// the digits as one device array [nd][sizeQlP][N]: in place when they are the views EvalKeySwitchPrecomputeCore made
// (the usual case), gathered into scratch otherwise
const uint64_t* DigitArray(Op& op, const std::vector<DCRTPoly>& digits, uint32_t sizeQlP,
                           std::vector<b200::detail::Slab>& keep) {
    const size_t E = (size_t)sizeQlP * op.N;
    keep.clear();
    for (const auto& d : digits) {
        if (d.GetNumOfElements() != sizeQlP)
            OPENFHE_THROW("b200 backend: digit with an unexpected number of towers");
        keep.push_back(d.DeviceSlab());
    }
    bool contiguous = true;
    for (size_t j = 1; j < keep.size(); ++j)
        contiguous = contiguous && keep[j].get() == keep[0].get() + j * E;
    if (contiguous)
        return keep[0].get();
    uint64_t* dig = op.Scratch(1, digits.size() * (size_t)sizeQlP);
    for (size_t j = 0; j < keep.size(); ++j)
        op.Copy(dig + j * E, keep[j].get(), sizeQlP);
    return dig;
}

PlaintextModulus NoiseScaleT(const std::shared_ptr<CryptoParametersRNS>& cp) {
    return (cp->GetNoiseScale() == 1) ? 0 : cp->GetPlaintextModulus();  // keyswitch-hybrid.cpp:386
}
}  // namespace

std::shared_ptr<std::vector<DCRTPoly>> KeySwitchHYBRID::EvalKeySwitchPrecomputeCore(
    const DCRTPoly& c, std::shared_ptr<CryptoParametersBase<DCRTPoly>> cryptoParamsBase) const {
    const auto cp = std::dynamic_pointer_cast<CryptoParametersRNS>(cryptoParamsBase);
    if (c.GetFormat() != Format::EVALUATION)
        OPENFHE_THROW("b200 backend: EvalKeySwitchPrecomputeCore expects EVALUATION format");
    const HksCtx& h        = HksOf(cp);
    const uint32_t sizeQl  = c.GetNumOfElements();
    const uint32_t nd      = ofhe_hks_num_digits(h.hks, sizeQl);
    const auto paramsQlP   = c.GetExtendedCRTBasis(cp->GetParamsP());
    Op op(h.N);
    const auto cd = c.DeviceSlab();
    auto dig      = op.New((size_t)nd * (sizeQl + h.sizeP));
    Check(ofhe_hks_modup(h.hks, sizeQl, cd.get(), dig.get(), op.l.stream), "ofhe_hks_modup");
    op.Sync();
    op.rt.n.modup += nd;
    return Views(op, dig, nd, paramsQlP);
}

std::shared_ptr<std::vector<DCRTPoly>> KeySwitchHYBRID::EvalFastKeySwitchCoreExt(
    const std::shared_ptr<std::vector<DCRTPoly>> digits, const EvalKey<DCRTPoly> evalKey,
    const std::shared_ptr<ParmType> paramsQl) const {
    const auto cp          = std::dynamic_pointer_cast<CryptoParametersRNS>(evalKey->GetCryptoParameters());
    const HksCtx& h        = HksOf(cp);
    const uint32_t sizeQl  = (uint32_t)paramsQl->GetParams().size();
    const uint32_t sizeQlP = sizeQl + h.sizeP;
    const uint32_t nd      = (uint32_t)digits->size();
    const auto paramsQlP   = (*digits)[0].GetParams();
    Op op(h.N);
    std::vector<b200::detail::Slab> keep;
    const uint64_t* dig = DigitArray(op, *digits, sizeQlP, keep);
    auto e01            = op.New((size_t)2 * sizeQlP);
    (void)nd;
    Check(ofhe_hks_inner_product(h.hks, sizeQl, dig, EvkOf(evalKey, h), e01.get(), op.l.stream), "ofhe_hks_inner_product");
    op.Sync();
    return Views(op, e01, 2, paramsQlP);
}

std::shared_ptr<std::vector<DCRTPoly>> KeySwitchHYBRID::EvalFastKeySwitchCore(
    const std::shared_ptr<std::vector<DCRTPoly>> digits, const EvalKey<DCRTPoly> evalKey,
    const std::shared_ptr<ParmType> paramsQl) const {
    const auto cp            = std::dynamic_pointer_cast<CryptoParametersRNS>(evalKey->GetCryptoParameters());
    const PlaintextModulus t = NoiseScaleT(cp);
    if (t != 0) {
        // BGV: the t-scaled ModDown (dcrtpoly-impl.h:982-983,999-1000) runs through DCRTPoly::ApproxModDown (device)
        auto result = EvalFastKeySwitchCoreExt(digits, evalKey, paramsQl);
        for (int k = 0; k < 2; ++k)
            (*result)[k] = (*result)[k].ApproxModDown(
                paramsQl, cp->GetParamsP(), cp->GetPInvModq(), cp->GetPInvModqPrecon(), cp->GetPHatInvModp(),
                cp->GetPHatInvModpPrecon(), cp->GetPHatModq(), cp->GetModqBarrettMu(), cp->GettInvModp(),
                cp->GettInvModpPrecon(), t, cp->GettModqPrecon());
        return result;
    }
    const HksCtx& h        = HksOf(cp);
    const uint32_t sizeQl  = (uint32_t)paramsQl->GetParams().size();
    const uint32_t sizeQlP = sizeQl + h.sizeP;
    const uint32_t nd      = (uint32_t)digits->size();
    Op op(h.N);
    std::vector<b200::detail::Slab> keep;
    const uint64_t* dig = DigitArray(op, *digits, sizeQlP, keep);
    uint64_t* e01       = op.Scratch(2, (size_t)2 * sizeQlP);
    auto out            = op.New((size_t)2 * sizeQl);
    (void)nd;
    Check(ofhe_hks_inner_product(h.hks, sizeQl, dig, EvkOf(evalKey, h), e01, op.l.stream), "ofhe_hks_inner_product");
    Check(ofhe_hks_approx_mod_down(h.hks, sizeQl, e01, out.get(), 2, op.l.stream), "ofhe_hks_approx_mod_down");
    op.Sync();
    op.rt.n.moddown += 2;
    return Views(op, out, 2, paramsQl);
}

std::shared_ptr<std::vector<DCRTPoly>> KeySwitchHYBRID::KeySwitchCore(const DCRTPoly& a,
                                                                      const EvalKey<DCRTPoly> evalKey) const {
    const auto cp = std::dynamic_pointer_cast<CryptoParametersRNS>(evalKey->GetCryptoParameters());
    if (NoiseScaleT(cp) != 0)
        return EvalFastKeySwitchCore(EvalKeySwitchPrecomputeCore(a, evalKey->GetCryptoParameters()), evalKey,
                                     a.GetParams());
    if (a.GetFormat() != Format::EVALUATION)
        OPENFHE_THROW("b200 backend: KeySwitchCore expects EVALUATION format");
    const HksCtx& h       = HksOf(cp);
    const uint32_t sizeQl = a.GetNumOfElements();
    Op op(h.N);
    const auto ad = a.DeviceSlab();
    auto out      = op.New((size_t)2 * sizeQl);
    Check(ofhe_hks_keyswitch_core(h.hks, sizeQl, ad.get(), EvkOf(evalKey, h), out.get(), op.l.stream),
          "ofhe_hks_keyswitch_core");
    op.Sync();
    op.rt.n.modup += ofhe_hks_num_digits(h.hks, sizeQl);
    op.rt.n.moddown += 2;
    return Views(op, out, 2, a.GetParams());
}

}  // namespace lbcrypto
