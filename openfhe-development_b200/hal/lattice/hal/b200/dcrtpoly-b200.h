// dcrtpoly-b200.h -- the B200 lattice backend of OpenFHE's HAL: lbcrypto::DCRTPoly bound to libofhe_b200.so.
//
// This is the class SURVEY.md section 8(b) asks for: it derives from the reference's DCRTPolyInterface
// (src/core/include/lattice/hal/dcrtpoly-interface.h:83-85, contract :153-1417) and is selected as `DCRTPoly` by the
// overlay lattice/hal/lat-backend.h of this directory (the reference's own selection point,
// src/core/include/lattice/hal/lat-backend.h:39-60), so src/pke (CKKS / BFV / BGV, KeySwitchHYBRID, ...) compiles
// and links UNCHANGED against it.  Recipe: oracle/Makefile.hal.
//
// Division of labour
//   * device (libofhe_b200.so through the C ABI of include/ofhe_b200.h), i.e. everything on the hot path of
//     SURVEY.md section 8(a):  SwitchFormat (a6-a8), Plus/Minus/Times/Negate and their Eq forms, Times(vector<scalar>)
//     (a9), AutomorphismTransform in EVALUATION (a15), ApproxSwitchCRTBasis (a10), ApproxModUp (a16's core),
//     ApproxModDown incl. the BGV t-branches (a11), DropLastElementAndScale (a14), ModReduce (f4), exact
//     SwitchCRTBasis / ExpandCRTBasis / FastExpandCRTBasisPloverQ (a12), ScaleAndRound DCRT->DCRT (a13).
//   * residency: a polynomial is a device slab [towers x N] (immutable once published, shared by copies, recycled
//     through a pool) AND / OR a host mirror -- an object of the reference's default implementation
//     (lattice/hal/default/dcrtpoly.h, class DCRTPolyImpl).  Device methods read their operands' slabs (uploading a
//     host-only operand once) and leave the result ON THE DEVICE; the mirror of such a result holds parameters and
//     formats only.  The host code reads and writes towers through GetAllElements() / GetElementAtIndex() /
//     SetElementAtIndex() all over src/pke: the const accessors download the slab on first use
//     (thread-safe, once), the non-const ones additionally invalidate it.  Everything that is NOT on the hot path
//     (samplers, CRT interpolation, decryption rounding, serialisation, arbitrary cyclotomics) is served by the mirror
//     exactly as the reference computes it.  Chains of device methods -- EvalMult, rescale, rotations, the BSGS loops of
//     bootstrapping -- therefore never cross PCIe.
//   * the mirror is never used to COMPUTE a device-listed method: there is no CPU fallback; without a CUDA device the
//     runtime's constructor throws.
#ifndef LBCRYPTO_INC_LATTICE_HAL_B200_DCRTPOLY_B200_H
#define LBCRYPTO_INC_LATTICE_HAL_B200_DCRTPOLY_B200_H

#include <atomic>
#include <functional>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "lattice/hal/default/dcrtpoly.h"  // the host mirror type (reference's default implementation, unmodified)
#include "lattice/hal/dcrtpoly-interface.h"
#include "lattice/hal/b200/b200-runtime.h"

namespace lbcrypto {
namespace b200 {

template <typename VecType>
class DCRTPolyB200 final : public DCRTPolyInterface<DCRTPolyB200<VecType>, VecType, NativeVector, PolyImpl> {
public:
    using Host                  = DCRTPolyImpl<VecType>;
    using Vector                = VecType;
    using Integer               = typename VecType::Integer;
    using Params                = ILDCRTParams<Integer>;
    using PolyType              = PolyImpl<NativeVector>;
    using PolyLargeType         = PolyImpl<VecType>;
    using DCRTPolyType          = DCRTPolyB200<VecType>;
    using DCRTPolyInterfaceType = DCRTPolyInterface<DCRTPolyB200<VecType>, VecType, NativeVector, PolyImpl>;
    using Precomputations       = typename DCRTPolyInterfaceType::CRTBasisExtensionPrecomputations;
    using DggType               = typename DCRTPolyInterfaceType::DggType;
    using DugType               = typename DCRTPolyInterfaceType::DugType;
    using TugType               = typename DCRTPolyInterfaceType::TugType;
    using BugType               = typename DCRTPolyInterfaceType::BugType;
    static_assert(sizeof(NativeInteger) == sizeof(uint64_t), "towers are exchanged with the device as raw 64-bit words");

    // ---- construction / assignment: as lattice/hal/default/dcrtpoly.h:71-127 ---------------------------------
    DCRTPolyB200() = default;
    DCRTPolyB200(const DCRTPolyType& e) {
        CopyFrom(e);
    }
    DCRTPolyB200(DCRTPolyType&& e) noexcept : m_h(std::move(e.m_h)), m_d(std::move(e.m_d)), m_w(e.m_w.load()) {}
    DCRTPolyB200(const Host& h) : m_h(h) {}              // NOLINT: the mirror type converts freely
    DCRTPolyB200(Host&& h) noexcept : m_h(std::move(h)) {}  // NOLINT
    DCRTPolyB200(const PolyLargeType& e, const std::shared_ptr<Params>& params) : m_h(e, params) {}
    DCRTPolyB200(const PolyType& e, const std::shared_ptr<Params>& params) : m_h(e, params) {}
    explicit DCRTPolyB200(const std::vector<PolyType>& elements) : m_h(elements) {}
    DCRTPolyB200(const std::shared_ptr<Params>& params, Format format = Format::EVALUATION,
                 bool initializeElementToZero = false)
        : m_h(params, format, initializeElementToZero) {}
    DCRTPolyB200(const DggType& dgg, const std::shared_ptr<Params>& p, Format f = Format::EVALUATION) : m_h(dgg, p, f) {}
    DCRTPolyB200(const BugType& bug, const std::shared_ptr<Params>& p, Format f = Format::EVALUATION) : m_h(bug, p, f) {}
    DCRTPolyB200(const TugType& tug, const std::shared_ptr<Params>& p, Format f = Format::EVALUATION, uint32_t h = 0)
        : m_h(tug, p, f, h) {}
    DCRTPolyB200(DugType& dug, const std::shared_ptr<Params>& p, Format f = Format::EVALUATION) : m_h(dug, p, f) {}

    DCRTPolyType& operator=(const DCRTPolyType& rhs) noexcept override {
        if (this != &rhs)
            CopyFrom(rhs);
        return *this;
    }
    DCRTPolyType& operator=(DCRTPolyType&& rhs) noexcept override {
        m_h = std::move(rhs.m_h);
        m_d = std::move(rhs.m_d);
        m_w = rhs.m_w.load();
        return *this;
    }
    DCRTPolyType& operator=(const PolyLargeType& rhs) noexcept {
        HW() = rhs;
        return *this;
    }
    DCRTPolyType& operator=(const PolyType& rhs) noexcept {
        HW() = rhs;
        return *this;
    }
    DCRTPolyType& operator=(std::initializer_list<uint64_t> rhs) noexcept override {
        HW() = rhs;
        return *this;
    }
    DCRTPolyType& operator=(uint64_t val) noexcept {
        HW() = val;
        return *this;
    }
    DCRTPolyType& operator=(const std::vector<int64_t>& rhs) noexcept {
        HW() = rhs;
        return *this;
    }
    DCRTPolyType& operator=(const std::vector<int32_t>& rhs) noexcept {
        HW() = rhs;
        return *this;
    }
    DCRTPolyType& operator=(std::initializer_list<std::string> rhs) noexcept {
        HW() = rhs;
        return *this;
    }

    // ---- accessors (host mirror) ---------------------------------------------------------------------------------
    Format GetFormat() const final {
        return m_h.GetFormat();
    }
    void OverrideFormat(const Format f) final {
        m_h.OverrideFormat(f);
    }
    const std::shared_ptr<Params>& GetParams() const {
        return m_h.GetParams();
    }
    // tower accessors: the const form downloads a device-only polynomial once; the non-const form hands the towers to
    // the host code for writing, so the device copy is dropped
    const std::vector<PolyType>& GetAllElements() const {
        return H().GetAllElements();
    }
    std::vector<PolyType>& GetAllElements() {
        return HW().GetAllElements();
    }
    void SetElementAtIndex(usint index, const PolyType& element) {
        HW().SetElementAtIndex(index, element);
    }
    void SetElementAtIndex(usint index, PolyType&& element) {
        HW().SetElementAtIndex(index, std::move(element));
    }
    // metadata only: never moves data (the interface's version goes through GetAllElements())
    usint GetNumOfElements() const {
        return m_h.GetAllElements().size();
    }
    const Host& HostMirror() const {
        return H();
    }
    // the towers as PARAMETER carriers (moduli, roots, formats): valid without any data movement
    const std::vector<PolyType>& TowerMeta() const {
        return m_h.GetAllElements();
    }
    bool OnDevice() const {
        return (m_w.load() & kDev) != 0;
    }
    static const std::string GetElementName() {
        return "DCRTPolyImpl";
    }
    std::string SerializedObjectName() const override {
        return "DCRTPoly";
    }
    static uint32_t SerializedVersion() {
        return 1;
    }
    template <class Archive>
    void save(Archive& ar, std::uint32_t const version) const {
        H().save(ar, version);  // byte-identical archives: the same three fields in the same order (dcrtpoly.h:337-341)
    }
    template <class Archive>
    void load(Archive& ar, std::uint32_t const version) {
        HW().load(ar, version);
    }

    // ---- not on the hot path: served by the mirror exactly as the reference does ---------------------------------
    DCRTPolyType CloneWithNoise(const DiscreteGaussianGeneratorImpl<VecType>& dgg, Format format) const override {
        return H().CloneWithNoise(dgg, format);
    }
    DCRTPolyType CloneTowers(uint32_t startTower, uint32_t endTower) const {
        return H().CloneTowers(startTower, endTower);
    }
    bool operator==(const DCRTPolyType& rhs) const override {
        return H() == rhs.H();
    }
    DCRTPolyType& operator+=(const Integer& rhs) override {
        HW() += rhs;
        return *this;
    }
    DCRTPolyType& operator+=(const NativeInteger& rhs) override {
        HW() += rhs;
        return *this;
    }
    DCRTPolyType& operator-=(const Integer& rhs) override {
        HW() -= rhs;
        return *this;
    }
    DCRTPolyType& operator-=(const NativeInteger& rhs) override {
        HW() -= rhs;
        return *this;
    }
    DCRTPolyType& operator*=(const Integer& rhs) override {
        HW() *= rhs;
        return *this;
    }
    DCRTPolyType& operator*=(const NativeInteger& rhs) override {
        HW() *= rhs;
        return *this;
    }
    std::vector<DCRTPolyType> BaseDecompose(usint baseBits, bool evalModeAnswer) const override {
        return Wrap(H().BaseDecompose(baseBits, evalModeAnswer));
    }
    std::vector<DCRTPolyType> PowersOfBase(usint baseBits) const override {
        return Wrap(H().PowersOfBase(baseBits));
    }
    std::vector<DCRTPolyType> CRTDecompose(uint32_t baseBits) const {
        return Wrap(H().CRTDecompose(baseBits));
    }
    DCRTPolyType Plus(const Integer& rhs) const override {
        return H().Plus(rhs);
    }
    DCRTPolyType Plus(const std::vector<Integer>& rhs) const {
        return H().Plus(rhs);
    }
    DCRTPolyType Minus(const Integer& rhs) const override {
        return H().Minus(rhs);
    }
    DCRTPolyType Minus(const std::vector<Integer>& rhs) const {
        return H().Minus(rhs);
    }
    DCRTPolyType Times(const Integer& rhs) const override {
        return H().Times(rhs);
    }
    DCRTPolyType Times(const std::vector<Integer>& rhs) const {
        return H().Times(rhs);
    }
    DCRTPolyType Times(NativeInteger::SignedNativeInt rhs) const override {
        return H().Times(rhs);
    }
#if NATIVEINT != 64
    DCRTPolyType Times(int64_t rhs) const {
        return H().Times(rhs);
    }
#endif
    DCRTPolyType MultiplicativeInverse() const override {
        return H().MultiplicativeInverse();
    }
    bool InverseExists() const override {
        return H().InverseExists();
    }
    bool IsEmpty() const override {
        return (m_w.load() & kDev) ? m_h.GetAllElements().empty() : m_h.IsEmpty();
    }
    void SetValuesToZero() override {
        HW().SetValuesToZero();
    }
    void AddILElementOne() override {
        HW().AddILElementOne();
    }
    void DropLastElement() override {
        m_h.DropLastElement();
    }
    void DropLastElements(size_t i) override {
        m_h.DropLastElements(i);
    }
    PolyLargeType CRTInterpolate() const override {
        return H().CRTInterpolate();
    }
    PolyType DecryptionCRTInterpolate(PlaintextModulus ptm) const override {
        return H().DecryptionCRTInterpolate(ptm);
    }
    PolyType ToNativePoly() const override {
        return H().ToNativePoly();
    }
    PolyLargeType CRTInterpolateIndex(usint i) const override {
        return H().CRTInterpolateIndex(i);
    }
    Integer GetWorkingModulus() const override {
        return m_h.GetWorkingModulus();
    }
    void SetValuesModSwitch(const DCRTPolyType& element, const NativeInteger& modulus) override {
        HW().SetValuesModSwitch(element.H(), modulus);
    }
    std::shared_ptr<Params> GetExtendedCRTBasis(const std::shared_ptr<Params>& paramsP) const override {
        return m_h.GetExtendedCRTBasis(paramsP);
    }
    void TimesQovert(const std::shared_ptr<Params>& paramsQ, const std::vector<NativeInteger>& tInvModq,
                     const NativeInteger& t, const NativeInteger& NegQModt,
                     const NativeInteger& NegQModtPrecon) override {
        HW().TimesQovert(paramsQ, tInvModq, t, NegQModt, NegQModtPrecon);
    }
    void ExpandCRTBasisReverseOrder(const std::shared_ptr<Params>& paramsQP, const std::shared_ptr<Params>& paramsP,
                                    const std::vector<NativeInteger>& QHatInvModq,
                                    const std::vector<NativeInteger>& QHatInvModqPrecon,
                                    const std::vector<std::vector<NativeInteger>>& QHatModp,
                                    const std::vector<std::vector<NativeInteger>>& alphaQModp,
                                    const std::vector<DoubleNativeInt>& modpBarrettMu, const std::vector<double>& qInv,
                                    Format resultFormat) override {
        HW().ExpandCRTBasisReverseOrder(paramsQP, paramsP, QHatInvModq, QHatInvModqPrecon, QHatModp, alphaQModp,
                                       modpBarrettMu, qInv, resultFormat);
    }
    PolyType ScaleAndRound(const NativeInteger& t, const std::vector<NativeInteger>& tQHatInvModqDivqModt,
                           const std::vector<NativeInteger>& tQHatInvModqDivqModtPrecon,
                           const std::vector<NativeInteger>& tQHatInvModqBDivqModt,
                           const std::vector<NativeInteger>& tQHatInvModqBDivqModtPrecon,
                           const std::vector<double>& tQHatInvModqDivqFrac,
                           const std::vector<double>& tQHatInvModqBDivqFrac) const override {
        return H().ScaleAndRound(t, tQHatInvModqDivqModt, tQHatInvModqDivqModtPrecon, tQHatInvModqBDivqModt,
                                 tQHatInvModqBDivqModtPrecon, tQHatInvModqDivqFrac, tQHatInvModqBDivqFrac);
    }
    PolyType ScaleAndRound(const std::vector<NativeInteger>& moduliQ, const NativeInteger& t,
                           const NativeInteger& tgamma, const std::vector<NativeInteger>& tgammaQHatModq,
                           const std::vector<NativeInteger>& tgammaQHatModqPrecon,
                           const std::vector<NativeInteger>& negInvqModtgamma,
                           const std::vector<NativeInteger>& negInvqModtgammaPrecon) const override {
        return H().ScaleAndRound(moduliQ, t, tgamma, tgammaQHatModq, tgammaQHatModqPrecon, negInvqModtgamma,
                                 negInvqModtgammaPrecon);
    }
    void SwitchModulusAtIndex(size_t index, const Integer& modulus, const Integer& rootOfUnity) override {
        HW().SwitchModulusAtIndex(index, modulus, rootOfUnity);
    }

    // ---- the hot path: DEVICE (definitions in dcrtpoly-b200-impl.h) -----------------------------------------------
    void SwitchFormat(uint32_t thread_limit = 0) override;
    DCRTPolyType& operator+=(const DCRTPolyType& rhs) override;
    DCRTPolyType& operator-=(const DCRTPolyType& rhs) override;
    DCRTPolyType& operator*=(const DCRTPolyType& rhs) override;
    DCRTPolyType Plus(const DCRTPolyType& rhs) const override;
    DCRTPolyType Minus(const DCRTPolyType& rhs) const override;
    DCRTPolyType Times(const DCRTPolyType& rhs) const override;
    DCRTPolyType Negate() const override;
    DCRTPolyType operator-() const override;
    DCRTPolyType Times(const std::vector<NativeInteger>& rhs) const;
    DCRTPolyType TimesNoCheck(const std::vector<NativeInteger>& rhs) const;
    DCRTPolyType AutomorphismTransform(uint32_t i) const override;
    DCRTPolyType AutomorphismTransform(uint32_t i, const std::vector<uint32_t>& vec) const override;
    void DropLastElementAndScale(const std::vector<NativeInteger>& QlQlInvModqlDivqlModq,
                                 const std::vector<NativeInteger>& qlInvModq) override;
    void ModReduce(const NativeInteger& t, const std::vector<NativeInteger>& tModqPrecon,
                   const NativeInteger& negtInvModq, const NativeInteger& negtInvModqPrecon,
                   const std::vector<NativeInteger>& qlInvModq,
                   const std::vector<NativeInteger>& qlInvModqPrecon) override;
    DCRTPolyType ApproxSwitchCRTBasis(const std::shared_ptr<Params>& paramsQ, const std::shared_ptr<Params>& paramsP,
                                      const std::vector<NativeInteger>& QHatInvModq,
                                      const std::vector<NativeInteger>& QHatInvModqPrecon,
                                      const std::vector<std::vector<NativeInteger>>& QHatModp,
                                      const std::vector<DoubleNativeInt>& modpBarrettMu) const override;
    void ApproxModUp(const std::shared_ptr<Params>& paramsQ, const std::shared_ptr<Params>& paramsP,
                     const std::shared_ptr<Params>& paramsQP, const std::vector<NativeInteger>& QHatInvModq,
                     const std::vector<NativeInteger>& QHatInvModqPrecon,
                     const std::vector<std::vector<NativeInteger>>& QHatModp,
                     const std::vector<DoubleNativeInt>& modpBarrettMu) override;
    DCRTPolyType ApproxModDown(
        const std::shared_ptr<Params>& paramsQ, const std::shared_ptr<Params>& paramsP,
        const std::vector<NativeInteger>& PInvModq, const std::vector<NativeInteger>& PInvModqPrecon,
        const std::vector<NativeInteger>& PHatInvModp, const std::vector<NativeInteger>& PHatInvModpPrecon,
        const std::vector<std::vector<NativeInteger>>& PHatModq, const std::vector<DoubleNativeInt>& modqBarrettMu,
        const std::vector<NativeInteger>& tInvModp, const std::vector<NativeInteger>& tInvModpPrecon,
        const NativeInteger& t, const std::vector<NativeInteger>& tModqPrecon) const override;
    DCRTPolyType SwitchCRTBasis(const std::shared_ptr<Params>& paramsP, const std::vector<NativeInteger>& QHatInvModq,
                                const std::vector<NativeInteger>& QHatInvModqPrecon,
                                const std::vector<std::vector<NativeInteger>>& QHatModp,
                                const std::vector<std::vector<NativeInteger>>& alphaQModp,
                                const std::vector<DoubleNativeInt>& modpBarrettMu,
                                const std::vector<double>& qInv) const override;
    void ExpandCRTBasis(const std::shared_ptr<Params>& paramsQP, const std::shared_ptr<Params>& paramsP,
                        const std::vector<NativeInteger>& QHatInvModq,
                        const std::vector<NativeInteger>& QHatInvModqPrecon,
                        const std::vector<std::vector<NativeInteger>>& QHatModp,
                        const std::vector<std::vector<NativeInteger>>& alphaQModp,
                        const std::vector<DoubleNativeInt>& modpBarrettMu, const std::vector<double>& qInv,
                        Format resultFormat) override;
    void FastExpandCRTBasisPloverQ(const Precomputations& precomputed) override;
    DCRTPolyType ApproxScaleAndRound(const std::shared_ptr<Params>& paramsP,
                                     const std::vector<std::vector<NativeInteger>>& tPSHatInvModsDivsModp,
                                     const std::vector<DoubleNativeInt>& modpBarretMu) const override;
    DCRTPolyType ScaleAndRound(const std::shared_ptr<Params>& paramsOutput,
                               const std::vector<std::vector<NativeInteger>>& tOSHatInvModsDivsModo,
                               const std::vector<double>& tOSHatInvModsDivsFrac,
                               const std::vector<DoubleNativeInt>& modoBarretMu) const override;
    void ScaleAndRoundPOverQ(const std::shared_ptr<Params>& paramsQ,
                             const std::vector<NativeInteger>& pInvModq) override;
    void ExpandCRTBasisQlHat(const std::shared_ptr<Params>& paramsQ, const std::vector<NativeInteger>& QlHatModq,
                             const std::vector<NativeInteger>& QlHatModqPrecon, const usint sizeQ) override;
    // BEHZ BFV (f4): dcrtpoly-impl.h:1694-1921
    void FastBaseConvqToBskMontgomery(
        const std::shared_ptr<Params>& paramsQBsk, const std::vector<NativeInteger>& moduliQ,
        const std::vector<NativeInteger>& moduliBsk, const std::vector<DoubleNativeInt>& modbskBarrettMu,
        const std::vector<NativeInteger>& mtildeQHatInvModq, const std::vector<NativeInteger>& mtildeQHatInvModqPrecon,
        const std::vector<std::vector<NativeInteger>>& QHatModbsk, const std::vector<uint64_t>& QHatModmtilde,
        const std::vector<NativeInteger>& QModbsk, const std::vector<NativeInteger>& QModbskPrecon,
        const uint64_t& negQInvModmtilde, const std::vector<NativeInteger>& mtildeInvModbsk,
        const std::vector<NativeInteger>& mtildeInvModbskPrecon) override;
    void FastRNSFloorq(const NativeInteger& t, const std::vector<NativeInteger>& moduliQ,
                       const std::vector<NativeInteger>& moduliBsk, const std::vector<DoubleNativeInt>& modbskBarrettMu,
                       const std::vector<NativeInteger>& tQHatInvModq,
                       const std::vector<NativeInteger>& tQHatInvModqPrecon,
                       const std::vector<std::vector<NativeInteger>>& QHatModbsk,
                       const std::vector<std::vector<NativeInteger>>& qInvModbsk,
                       const std::vector<NativeInteger>& tQInvModbsk,
                       const std::vector<NativeInteger>& tQInvModbskPrecon) override;
    void FastBaseConvSK(const std::shared_ptr<Params>& paramsQ, const std::vector<DoubleNativeInt>& modqBarrettMu,
                        const std::vector<NativeInteger>& moduliBsk,
                        const std::vector<DoubleNativeInt>& modbskBarrettMu,
                        const std::vector<NativeInteger>& BHatInvModb,
                        const std::vector<NativeInteger>& BHatInvModbPrecon,
                        const std::vector<NativeInteger>& BHatModmsk, const NativeInteger& BInvModmsk,
                        const NativeInteger& BInvModmskPrecon, const std::vector<std::vector<NativeInteger>>& BHatModq,
                        const std::vector<NativeInteger>& BModq,
                        const std::vector<NativeInteger>& BModqPrecon) override;

    // ---- residency plumbing (public for the rebound KeySwitchHYBRID, hal/keyswitch-hybrid-b200.cpp) ---------------
    // device slab of this polynomial [GetNumOfElements()][N] (uploads a host-only polynomial once)
    std::shared_ptr<uint64_t> DeviceSlab() const;
    // a device-resident polynomial over (params, format) whose rows are `slab` (EVALUATION / COEFFICIENT as stated)
    static DCRTPolyType FromDevice(const std::shared_ptr<Params>& params, Format format, std::shared_ptr<uint64_t> slab);

private:
    static constexpr uint8_t kHost = 1, kDev = 2;
    mutable Host m_h;                      // towers (values valid iff m_w & kHost), parameters and formats (always valid)
    mutable std::shared_ptr<uint64_t> m_d;  // device rows (valid iff m_w & kDev); immutable once published
    mutable std::atomic<uint8_t> m_w{kHost};

    void EnsureHost() const;  // download once, thread-safe
    const Host& H() const {
        if (!(m_w.load(std::memory_order_acquire) & kHost))
            EnsureHost();
        return m_h;
    }
    Host& HW() {  // the host code is about to write the towers: the device copy dies
        if (!(m_w.load(std::memory_order_acquire) & kHost))
            EnsureHost();
        m_d.reset();
        m_w.store(kHost, std::memory_order_release);
        return m_h;
    }
    void CopyFrom(const DCRTPolyType& e);
    // this polynomial becomes the device-resident (params, format, slab); towers carry parameters and formats only
    void Adopt(Host&& meta, std::shared_ptr<uint64_t> slab) {
        m_h = std::move(meta);
        m_d = std::move(slab);
        m_w.store(kDev, std::memory_order_release);
    }
    static Host Meta(const std::shared_ptr<Params>& params, Format format) {
        return Host(params, format, false);
    }

    static std::vector<DCRTPolyType> Wrap(std::vector<Host>&& v) {
        std::vector<DCRTPolyType> r;
        r.reserve(v.size());
        for (auto& h : v)
            r.emplace_back(std::move(h));
        return r;
    }
};

}  // namespace b200
}  // namespace lbcrypto

#endif
