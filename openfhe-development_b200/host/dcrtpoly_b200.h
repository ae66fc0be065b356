// dcrtpoly_b200.h -- host-side C++ mirror of the reference's DCRTPoly / KeySwitchHYBRID interface for the hot path,
// implemented over the C ABI of include/ofhe_b200.h (header only; link with -lofhe_b200).
//
// Same method names, argument meaning and error behaviour as the reference so that code and tests written against
//   lbcrypto::DCRTPolyImpl      src/core/include/lattice/hal/default/dcrtpoly.h, dcrtpoly-impl.h
//   lbcrypto::KeySwitchHYBRID   src/pke/include/keyswitch/keyswitch-hybrid.h, src/pke/lib/keyswitch/keyswitch-hybrid.cpp
// read the same here.  Differences that follow from device residency:
//   * tables are derived from the moduli by the library (ofhe_hks_create), so the long table argument lists of
//     ApproxSwitchCRTBasis / ApproxModDown collapse to the basis description;
//   * GetElementAtIndex hands out host copies (the "tower accessor leak" of SURVEY.md section 8 b);
//   * failures throw b200::OpenFHEException carrying the C ABI's error string (the reference throws
//     lbcrypto::OpenFHEException via OPENFHE_THROW, src/core/include/utils/exception.h).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ofhe_b200.h"

namespace b200 {

enum Format { EVALUATION = 0, COEFFICIENT = 1 };  // src/core/include/utils/inttypes.h

struct OpenFHEException : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc) {
    if (rc != 0)
        throw OpenFHEException(ofhe_last_error());
}

// one CUDA device + one ring (cyclotomic order 2N, registered moduli with their roots): ILDCRTParams
class Params {
public:
    Params(int device, uint32_t ringDim, const std::vector<uint64_t>& moduli, const std::vector<uint64_t>& roots)
        : m_ringDim(ringDim), m_moduli(moduli), m_roots(roots) {
        check(ofhe_ctx_create(device, &m_ctx));
        if (ofhe_ring_create(m_ctx, ringDim, (uint32_t)moduli.size(), moduli.data(), roots.data(), &m_ring) != 0) {
            std::string msg = ofhe_last_error();
            ofhe_ctx_destroy(m_ctx);
            throw OpenFHEException(msg);
        }
    }
    ~Params() {
        ofhe_ring_destroy(m_ring);
        ofhe_ctx_destroy(m_ctx);
    }
    Params(const Params&)            = delete;
    Params& operator=(const Params&) = delete;
    uint32_t GetRingDimension() const { return m_ringDim; }
    uint32_t GetCyclotomicOrder() const { return 2 * m_ringDim; }
    const std::vector<uint64_t>& GetModuli() const { return m_moduli; }
    ofhe_ctx* ctx() const { return m_ctx; }
    ofhe_ring* ring() const { return m_ring; }

private:
    uint32_t m_ringDim;
    std::vector<uint64_t> m_moduli, m_roots;
    ofhe_ctx* m_ctx   = nullptr;
    ofhe_ring* m_ring = nullptr;
};

// value-semantics polynomial over a subset of the ring's moduli, device resident
class DCRTPoly {
public:
    DCRTPoly() = default;
    DCRTPoly(std::shared_ptr<Params> params, std::vector<uint32_t> towers, Format format, bool initializeElementToZero = false)
        : m_params(std::move(params)), m_towers(std::move(towers)), m_format(format) {
        allocate();
        if (initializeElementToZero) {
            std::vector<uint64_t> z(words(), 0);
            check(ofhe_memcpy_h2d(m_params->ctx(), m_data, z.data(), bytes(), nullptr));
            check(ofhe_stream_sync(m_params->ctx(), nullptr));
        }
    }
    DCRTPoly(const DCRTPoly& o) : m_params(o.m_params), m_towers(o.m_towers), m_format(o.m_format) {
        if (o.m_data) {
            allocate();
            check(ofhe_memcpy_d2d(m_params->ctx(), m_data, o.m_data, bytes(), nullptr));
        }
    }
    DCRTPoly(DCRTPoly&& o) noexcept { swap(o); }
    DCRTPoly& operator=(DCRTPoly o) noexcept {
        swap(o);
        return *this;
    }
    ~DCRTPoly() {
        if (m_data)
            ofhe_free(m_params->ctx(), m_data);
    }

    // ---- accessors (dcrtpoly.h:378-392, dcrtpoly-interface.h:287-333)
    Format GetFormat() const { return m_format; }
    uint32_t GetRingDimension() const { return m_params->GetRingDimension(); }
    size_t GetNumOfElements() const { return m_towers.size(); }
    const std::shared_ptr<Params>& GetParams() const { return m_params; }
    const std::vector<uint32_t>& GetTowers() const { return m_towers; }
    uint64_t* device_data() const { return m_data; }
    // host copy of tower i (GetElementAtIndex(i).GetValues())
    std::vector<uint64_t> GetElementAtIndex(size_t i) const {
        require(i < m_towers.size(), "DCRTPolyImpl is empty or index out of range");
        std::vector<uint64_t> v(GetRingDimension());
        check(ofhe_memcpy_d2h(m_params->ctx(), v.data(), m_data + i * GetRingDimension(), v.size() * 8, nullptr));
        check(ofhe_stream_sync(m_params->ctx(), nullptr));
        return v;
    }
    void SetElementAtIndex(size_t i, const std::vector<uint64_t>& values) {
        require(i < m_towers.size() && values.size() == GetRingDimension(), "SetElementAtIndex: bad tower");
        check(ofhe_memcpy_h2d(m_params->ctx(), m_data + i * GetRingDimension(), values.data(), values.size() * 8, nullptr));
        check(ofhe_stream_sync(m_params->ctx(), nullptr));
    }

    // ---- format (dcrtpoly-impl.h:1932-1940)
    void SwitchFormat() {
        auto f = m_format == COEFFICIENT ? ofhe_ntt_fwd : ofhe_ntt_inv;
        check(f(m_params->ring(), m_data, m_data, (uint32_t)m_towers.size(), m_towers.data(), nullptr));
        m_format = m_format == COEFFICIENT ? EVALUATION : COEFFICIENT;
    }
    void SetFormat(Format f) {
        if (f != m_format)
            SwitchFormat();
    }

    // ---- limb-wise arithmetic (dcrtpoly.h:130-190); operands must share towers and format, as in the reference
    DCRTPoly Plus(const DCRTPoly& rhs) const { return binop(0, rhs); }
    DCRTPoly Minus(const DCRTPoly& rhs) const { return binop(1, rhs); }
    DCRTPoly Times(const DCRTPoly& rhs) const {
        require(m_format == EVALUATION && rhs.m_format == EVALUATION, "operator* for DCRTPolyImpl supported only in Format::EVALUATION");
        return binop(2, rhs);
    }
    DCRTPoly Negate() const {
        DCRTPoly r(m_params, m_towers, m_format);
        check(ofhe_ew_binop(m_params->ring(), 3, r.m_data, m_data, m_data, (uint32_t)m_towers.size(), m_towers.data(), nullptr));
        return r;
    }
    DCRTPoly& operator+=(const DCRTPoly& rhs) { return binop_eq(0, rhs); }
    DCRTPoly& operator-=(const DCRTPoly& rhs) { return binop_eq(1, rhs); }
    DCRTPoly& operator*=(const DCRTPoly& rhs) {
        require(m_format == EVALUATION && rhs.m_format == EVALUATION, "operator* for DCRTPolyImpl supported only in Format::EVALUATION");
        return binop_eq(2, rhs);
    }
    DCRTPoly operator+(const DCRTPoly& rhs) const { return Plus(rhs); }
    DCRTPoly operator-(const DCRTPoly& rhs) const { return Minus(rhs); }
    DCRTPoly operator*(const DCRTPoly& rhs) const { return Times(rhs); }
    // Times(const std::vector<NativeInteger>&): one scalar per tower
    DCRTPoly Times(const std::vector<uint64_t>& rhs) const {
        require(rhs.size() == m_towers.size(), "tower size mismatch; cannot multiply by vector of integers");
        DCRTPoly r(m_params, m_towers, m_format);
        check(ofhe_ew_mul_scalars(m_params->ring(), r.m_data, m_data, rhs.data(), (uint32_t)m_towers.size(), m_towers.data(), nullptr));
        return r;
    }
    // AutomorphismTransform(k, precomp): the map is built (and cached) on the device (dcrtpoly-impl.h:325-333)
    DCRTPoly AutomorphismTransform(uint32_t k) const {
        require(m_format == EVALUATION, "AutomorphismTransform: only Format::EVALUATION is accelerated");
        DCRTPoly r(m_params, m_towers, m_format);
        check(ofhe_automorphism(m_params->ring(), r.m_data, m_data, (uint32_t)m_towers.size(), k, nullptr));
        return r;
    }
    void DropLastElement() {  // dcrtpoly-impl.h:669-678
        require(!m_towers.empty(), "Input has no elements to drop.");
        require(m_towers.size() > 1, "Removing last element of DCRTPoly renders it invalid.");
        m_towers.pop_back();  // the slab keeps its allocation; the last row simply falls out of the limb map
    }
    // ApproxSwitchCRTBasis(paramsQ, paramsP, tables...) -> polynomial over `targetTowers` (dcrtpoly-impl.h:888-932)
    DCRTPoly ApproxSwitchCRTBasis(const std::vector<uint32_t>& targetTowers) const {
        require(m_format == COEFFICIENT, "ApproxSwitchCRTBasis expects Format::COEFFICIENT");
        DCRTPoly r(m_params, targetTowers, COEFFICIENT);
        check(ofhe_approx_switch_crt_basis(m_params->ring(), (uint32_t)m_towers.size(), m_towers.data(),
                                           (uint32_t)targetTowers.size(), targetTowers.data(), m_data, r.m_data, nullptr));
        return r;
    }
    bool operator==(const DCRTPoly& rhs) const {
        if (m_towers != rhs.m_towers || m_format != rhs.m_format)
            return false;
        for (size_t i = 0; i < m_towers.size(); ++i)
            if (GetElementAtIndex(i) != rhs.GetElementAtIndex(i))
                return false;
        return true;
    }

private:
    static void require(bool ok, const char* msg) {
        if (!ok)
            throw OpenFHEException(msg);
    }
    size_t words() const { return m_towers.size() * (size_t)m_params->GetRingDimension(); }
    size_t bytes() const { return words() * sizeof(uint64_t); }
    void allocate() {
        void* p = nullptr;
        check(ofhe_malloc(m_params->ctx(), bytes() ? bytes() : 8, &p));
        m_data = static_cast<uint64_t*>(p);
    }
    void swap(DCRTPoly& o) noexcept {
        std::swap(m_params, o.m_params);
        std::swap(m_towers, o.m_towers);
        std::swap(m_format, o.m_format);
        std::swap(m_data, o.m_data);
    }
    DCRTPoly binop(int op, const DCRTPoly& rhs) const {
        require(m_towers == rhs.m_towers, "tower size mismatch; cannot combine");
        require(m_format == rhs.m_format, "format mismatch");
        DCRTPoly r(m_params, m_towers, m_format);
        check(ofhe_ew_binop(m_params->ring(), op, r.m_data, m_data, rhs.m_data, (uint32_t)m_towers.size(), m_towers.data(), nullptr));
        return r;
    }
    DCRTPoly& binop_eq(int op, const DCRTPoly& rhs) {
        require(m_towers == rhs.m_towers, "tower size mismatch; cannot combine");
        require(m_format == rhs.m_format, "format mismatch");
        check(ofhe_ew_binop(m_params->ring(), op, m_data, m_data, rhs.m_data, (uint32_t)m_towers.size(), m_towers.data(), nullptr));
        return *this;
    }
    std::shared_ptr<Params> m_params;
    std::vector<uint32_t> m_towers;
    Format m_format  = EVALUATION;
    uint64_t* m_data = nullptr;
};

// EvalKeyRelin: A/B vectors on the device (src/pke/include/key/evalkeyrelin.h)
class EvalKey;

// KeySwitchHYBRID over a ring laid out [Q | P] (src/pke/lib/keyswitch/keyswitch-hybrid.cpp)
class KeySwitchHYBRID {
public:
    KeySwitchHYBRID(std::shared_ptr<Params> params, uint32_t sizeQ, uint32_t sizeP, uint32_t numPartQ)
        : m_params(std::move(params)), m_sizeQ(sizeQ), m_sizeP(sizeP) {
        check(ofhe_hks_create(m_params->ring(), sizeQ, sizeP, numPartQ, &m_hks));
    }
    ~KeySwitchHYBRID() { ofhe_hks_destroy(m_hks); }
    KeySwitchHYBRID(const KeySwitchHYBRID&)            = delete;
    KeySwitchHYBRID& operator=(const KeySwitchHYBRID&) = delete;
    ofhe_hks* handle() const { return m_hks; }
    uint32_t GetNumPerPartQ() const { return ofhe_hks_num_per_part(m_hks); }

    // EvalKeySwitchPrecomputeCore(c, cryptoParams) -> digits (:314-379); c over the first sizeQl towers, EVALUATION
    std::vector<DCRTPoly> EvalKeySwitchPrecomputeCore(const DCRTPoly& c) const {
        const uint32_t sizeQl = (uint32_t)c.GetNumOfElements(), nd = ofhe_hks_num_digits(m_hks, sizeQl);
        const size_t N = c.GetRingDimension(), sizeQlP = sizeQl + m_sizeP;
        DCRTPoly slab(m_params, repeat(towersQlP(sizeQl), nd), EVALUATION);
        check(ofhe_hks_modup(m_hks, sizeQl, c.device_data(), slab.device_data(), nullptr));
        std::vector<DCRTPoly> digits;
        for (uint32_t j = 0; j < nd; ++j) {
            DCRTPoly d(m_params, towersQlP(sizeQl), EVALUATION);
            check(ofhe_memcpy_d2d(m_params->ctx(), d.device_data(), slab.device_data() + j * sizeQlP * N,
                                  sizeQlP * N * sizeof(uint64_t), nullptr));
            digits.push_back(std::move(d));
        }
        return digits;
    }
    // KeySwitchCore(a, evalKey) -> (ba[0], ba[1]) (:308-312)
    std::pair<DCRTPoly, DCRTPoly> KeySwitchCore(const DCRTPoly& a, const ofhe_evk* evalKey) const {
        const uint32_t sizeQl = (uint32_t)a.GetNumOfElements();
        const size_t N        = a.GetRingDimension();
        DCRTPoly both(m_params, repeat(towersQl(sizeQl), 2), EVALUATION);
        check(ofhe_hks_keyswitch_core(m_hks, sizeQl, a.device_data(), evalKey, both.device_data(), nullptr));
        return {slice(both, 0, sizeQl, N), slice(both, 1, sizeQl, N)};
    }
    // EvalFastKeySwitchCoreExt(digits, evalKey, paramsQl) -> (cTilda[0], cTilda[1]) over QlP (:402-435)
    std::pair<DCRTPoly, DCRTPoly> EvalFastKeySwitchCoreExt(const std::vector<DCRTPoly>& digits, const ofhe_evk* evalKey,
                                                           uint32_t sizeQl) const {
        const size_t N = digits.at(0).GetRingDimension(), sizeQlP = sizeQl + m_sizeP;
        DCRTPoly slab = pack(digits, sizeQl);
        DCRTPoly both(m_params, repeat(towersQlP(sizeQl), 2), EVALUATION);
        check(ofhe_hks_inner_product(m_hks, sizeQl, slab.device_data(), evalKey, both.device_data(), nullptr));
        return {sliceExt(both, 0, sizeQl, N, sizeQlP), sliceExt(both, 1, sizeQl, N, sizeQlP)};
    }
    // EvalFastKeySwitchCore(digits, evalKey, paramsQl) = the above + ApproxModDown of both (:381-400)
    std::pair<DCRTPoly, DCRTPoly> EvalFastKeySwitchCore(const std::vector<DCRTPoly>& digits, const ofhe_evk* evalKey,
                                                        uint32_t sizeQl) const {
        const size_t N = digits.at(0).GetRingDimension(), sizeQlP = sizeQl + m_sizeP;
        DCRTPoly slab = pack(digits, sizeQl);
        DCRTPoly ext(m_params, repeat(towersQlP(sizeQl), 2), EVALUATION);
        check(ofhe_hks_inner_product(m_hks, sizeQl, slab.device_data(), evalKey, ext.device_data(), nullptr));
        DCRTPoly both(m_params, repeat(towersQl(sizeQl), 2), EVALUATION);
        check(ofhe_hks_approx_mod_down(m_hks, sizeQl, ext.device_data(), both.device_data(), 2, nullptr));
        (void)sizeQlP;
        return {slice(both, 0, sizeQl, N), slice(both, 1, sizeQl, N)};
    }
    // KeySwitchExt(ciphertext, addFirst) (:217-243): (c0, c1) over Ql -> over QlP as [c_k * (P mod q_i) | 0]
    std::pair<DCRTPoly, DCRTPoly> KeySwitchExt(const DCRTPoly& c0, const DCRTPoly& c1, bool addFirst) const {
        const uint32_t sizeQl = (uint32_t)c0.GetNumOfElements();
        const size_t N = c0.GetRingDimension(), sizeQlP = sizeQl + m_sizeP;
        DCRTPoly in(m_params, repeat(towersQl(sizeQl), 2), EVALUATION);
        check(ofhe_memcpy_d2d(m_params->ctx(), in.device_data(), c0.device_data(), sizeQl * N * sizeof(uint64_t), nullptr));
        check(ofhe_memcpy_d2d(m_params->ctx(), in.device_data() + (size_t)sizeQl * N, c1.device_data(),
                              sizeQl * N * sizeof(uint64_t), nullptr));
        DCRTPoly both(m_params, repeat(towersQlP(sizeQl), 2), EVALUATION);
        check(ofhe_hks_keyswitch_ext(m_hks, sizeQl, addFirst ? 1 : 0, in.device_data(), both.device_data(), nullptr));
        return {sliceExt(both, 0, sizeQl, N, sizeQlP), sliceExt(both, 1, sizeQl, N, sizeQlP)};
    }
    // KeySwitchDown(ciphertext) (:246-277): ApproxModDown of both polynomials of an extended ciphertext
    std::pair<DCRTPoly, DCRTPoly> KeySwitchDown(const DCRTPoly& e0, const DCRTPoly& e1) const {
        const uint32_t sizeQl = (uint32_t)e0.GetNumOfElements() - m_sizeP;
        const size_t N = e0.GetRingDimension(), sizeQlP = sizeQl + m_sizeP;
        DCRTPoly in(m_params, repeat(towersQlP(sizeQl), 2), EVALUATION);
        check(ofhe_memcpy_d2d(m_params->ctx(), in.device_data(), e0.device_data(), sizeQlP * N * sizeof(uint64_t), nullptr));
        check(ofhe_memcpy_d2d(m_params->ctx(), in.device_data() + sizeQlP * N, e1.device_data(),
                              sizeQlP * N * sizeof(uint64_t), nullptr));
        DCRTPoly both(m_params, repeat(towersQl(sizeQl), 2), EVALUATION);
        check(ofhe_hks_approx_mod_down(m_hks, sizeQl, in.device_data(), both.device_data(), 2, nullptr));
        return {slice(both, 0, sizeQl, N), slice(both, 1, sizeQl, N)};
    }

private:
    DCRTPoly pack(const std::vector<DCRTPoly>& digits, uint32_t sizeQl) const {  // digits -> one [nd][QlP][N] slab
        const size_t N = digits.at(0).GetRingDimension(), sizeQlP = sizeQl + m_sizeP;
        DCRTPoly slab(m_params, repeat(towersQlP(sizeQl), (uint32_t)digits.size()), EVALUATION);
        for (size_t j = 0; j < digits.size(); ++j)
            check(ofhe_memcpy_d2d(m_params->ctx(), slab.device_data() + j * sizeQlP * N, digits[j].device_data(),
                                  sizeQlP * N * sizeof(uint64_t), nullptr));
        return slab;
    }
    DCRTPoly sliceExt(const DCRTPoly& both, uint32_t k, uint32_t sizeQl, size_t N, size_t sizeQlP) const {
        DCRTPoly r(m_params, towersQlP(sizeQl), EVALUATION);
        check(ofhe_memcpy_d2d(m_params->ctx(), r.device_data(), both.device_data() + (size_t)k * sizeQlP * N,
                              sizeQlP * N * sizeof(uint64_t), nullptr));
        return r;
    }
    std::vector<uint32_t> towersQl(uint32_t sizeQl) const {
        std::vector<uint32_t> t(sizeQl);
        for (uint32_t i = 0; i < sizeQl; ++i)
            t[i] = i;
        return t;
    }
    std::vector<uint32_t> towersQlP(uint32_t sizeQl) const {
        auto t = towersQl(sizeQl);
        for (uint32_t j = 0; j < m_sizeP; ++j)
            t.push_back(m_sizeQ + j);
        return t;
    }
    static std::vector<uint32_t> repeat(const std::vector<uint32_t>& t, uint32_t n) {
        std::vector<uint32_t> r;
        for (uint32_t k = 0; k < n; ++k)
            r.insert(r.end(), t.begin(), t.end());
        return r;
    }
    DCRTPoly slice(const DCRTPoly& both, uint32_t k, uint32_t sizeQl, size_t N) const {
        DCRTPoly r(m_params, towersQl(sizeQl), EVALUATION);
        check(ofhe_memcpy_d2d(m_params->ctx(), r.device_data(), both.device_data() + (size_t)k * sizeQl * N,
                              (size_t)sizeQl * N * sizeof(uint64_t), nullptr));
        return r;
    }
    std::shared_ptr<Params> m_params;
    uint32_t m_sizeQ, m_sizeP;
    ofhe_hks* m_hks = nullptr;
};

}  // namespace b200
