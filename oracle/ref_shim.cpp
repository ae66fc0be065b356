// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" wrapper that lets the Python test/bench harness call the
// UNMODIFIED reference (openfhe-development v1.5.1, compiled from
// /root/reference by oracle/Makefile.ref into oracle/_ref/) on raw uint64
// arrays.  Nothing here re-implements arithmetic: every function forwards to the
// reference's own classes.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the resulting libofhe_ref.so.
//
// Layout convention shared with the product: a DCRTPoly is a row-major
// [limbs x N] slab of uint64 residues; EVALUATION format is the reference's
// bit-reversed order.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "openfhe.h"
#include <sstream>

// serialisation of the lattice layer (cereal): DCRTPoly / key objects to portable-binary and JSON
#include "ciphertext-ser.h"
#include "cryptocontext-ser.h"
#include "key/key-ser.h"
#include "scheme/ckksrns/ckksrns-ser.h"
#include "utils/utilities-int.h"

using namespace lbcrypto;

namespace {

thread_local std::string g_err;

std::shared_ptr<ILDCRTParams<BigInteger>> make_params(uint32_t N, uint32_t L, const uint64_t* moduli,
                                                      const uint64_t* roots) {
    std::vector<NativeInteger> m(L), r(L);
    for (uint32_t i = 0; i < L; ++i) {
        m[i] = NativeInteger(moduli[i]);
        r[i] = NativeInteger(roots[i]);
    }
    return std::make_shared<ILDCRTParams<BigInteger>>(2 * N, m, r);
}

DCRTPoly make_poly(const std::shared_ptr<ILDCRTParams<BigInteger>>& params, Format fmt, const uint64_t* data) {
    DCRTPoly p(params, fmt, true);
    const uint32_t N = params->GetRingDimension();
    const uint32_t L = params->GetParams().size();
    for (uint32_t i = 0; i < L; ++i) {
        NativePoly t = p.GetElementAtIndex(i);
        NativeVector v(N, t.GetModulus());
        for (uint32_t j = 0; j < N; ++j)
            v[j] = NativeInteger(data[(size_t)i * N + j]);
        t.SetValues(std::move(v), fmt);
        p.SetElementAtIndex(i, std::move(t));
    }
    return p;
}

void dump_poly(const DCRTPoly& p, uint64_t* out) {
    const uint32_t N = p.GetRingDimension();
    const uint32_t L = p.GetNumOfElements();
    for (uint32_t i = 0; i < L; ++i) {
        const auto& v = p.GetElementAtIndex(i).GetValues();
        for (uint32_t j = 0; j < N; ++j)
            out[(size_t)i * N + j] = v[j].ConvertToInt<uint64_t>();
    }
}

struct CkksCtx {
    CryptoContext<DCRTPoly> cc;
    KeyPair<DCRTPoly> keys;
    std::shared_ptr<CryptoParametersCKKSRNS> cp;
    Ciphertext<DCRTPoly> tmpl;  // an honest encryption, used as metadata template
    uint32_t N, sizeQ, sizeP;
    // first-l-limb parameter sets
    std::shared_ptr<ILDCRTParams<BigInteger>> paramsQl(uint32_t sizeQl) const {
        const auto& ps = cp->GetElementParams()->GetParams();
        std::vector<NativeInteger> m(sizeQl), r(sizeQl);
        for (uint32_t i = 0; i < sizeQl; ++i) {
            m[i] = ps[i]->GetModulus();
            r[i] = ps[i]->GetRootOfUnity();
        }
        return std::make_shared<ILDCRTParams<BigInteger>>(2 * N, m, r);
    }
    std::shared_ptr<ILDCRTParams<BigInteger>> paramsQlP(uint32_t sizeQl) const {
        const auto& ps = cp->GetElementParams()->GetParams();
        const auto& pp = cp->GetParamsP()->GetParams();
        std::vector<NativeInteger> m, r;
        for (uint32_t i = 0; i < sizeQl; ++i) {
            m.push_back(ps[i]->GetModulus());
            r.push_back(ps[i]->GetRootOfUnity());
        }
        for (auto& p : pp) {
            m.push_back(p->GetModulus());
            r.push_back(p->GetRootOfUnity());
        }
        return std::make_shared<ILDCRTParams<BigInteger>>(2 * N, m, r);
    }
};

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    }
    catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

void copy_vec(const std::vector<NativeInteger>& v, uint64_t* out) {
    for (size_t i = 0; i < v.size(); ++i)
        out[i] = v[i].ConvertToInt<uint64_t>();
}

}  // namespace

extern "C" {

const char* ref_last_error() {
    return g_err.c_str();
}

void ref_set_threads(int n) {
    // utils/parallel.h: OpenFHEParallelControls caps every omp region of the reference
    omp_set_num_threads(n);
    OpenFHEParallelControls.Enable();
    OpenFHEParallelControls.SetNumThreads(n);
}
int ref_get_threads() {
    return OpenFHEParallelControls.GetNumThreads();
}

// ---------------------------------------------------------------- number theory (nbtheory-impl.h)
uint64_t ref_root_of_unity(uint32_t m, uint64_t q) {
    return RootOfUnity<NativeInteger>(m, NativeInteger(q)).ConvertToInt<uint64_t>();
}
uint64_t ref_first_prime(uint32_t bits, uint64_t m) {
    return FirstPrime<NativeInteger>(bits, m).ConvertToInt<uint64_t>();
}
uint64_t ref_last_prime(uint32_t bits, uint64_t m) {
    return LastPrime<NativeInteger>(bits, m).ConvertToInt<uint64_t>();
}
uint64_t ref_previous_prime(uint64_t q, uint64_t m) {
    return PreviousPrime<NativeInteger>(NativeInteger(q), m).ConvertToInt<uint64_t>();
}
uint64_t ref_next_prime(uint64_t q, uint64_t m) {
    return NextPrime<NativeInteger>(NativeInteger(q), m).ConvertToInt<uint64_t>();
}

// ---------------------------------------------------------------- scalar modular arithmetic (ubintnat.h)
uint64_t ref_modmul_barrett(uint64_t a, uint64_t b, uint64_t q) {
    NativeInteger Q(q);
    return NativeInteger(a).ModMulFast(NativeInteger(b), Q, Q.ComputeMu()).ConvertToInt<uint64_t>();
}
uint64_t ref_prep_modmul_const(uint64_t w, uint64_t q) {
    return NativeInteger(w).PrepModMulConst(NativeInteger(q)).ConvertToInt<uint64_t>();
}
uint64_t ref_modmul_const(uint64_t x, uint64_t w, uint64_t q) {
    NativeInteger W(w), Q(q);
    return NativeInteger(x).ModMulFastConst(W, Q, W.PrepModMulConst(Q)).ConvertToInt<uint64_t>();
}
uint64_t ref_barrett128(uint64_t hi, uint64_t lo, uint64_t q) {
    // utilities-int.h: BarrettUint128ModUint64 with mu = floor(2^128 / q)
    DoubleNativeInt a  = (DoubleNativeInt(hi) << 64) | lo;
    DoubleNativeInt mu = (BigInteger(1).LShiftEq(128) / BigInteger(NativeInteger(q))).ConvertToInt<DoubleNativeInt>();
    return BarrettUint128ModUint64(a, q, mu);
}

// ---------------------------------------------------------------- NTT (transformnat-impl.h)
int ref_ntt_fwd(uint64_t q, uint64_t root, uint32_t N, uint64_t* x) {
    return guarded([&] {
        NativeVector v(N, NativeInteger(q));
        for (uint32_t i = 0; i < N; ++i)
            v[i] = NativeInteger(x[i]);
        ChineseRemainderTransformFTT<NativeVector>().ForwardTransformToBitReverseInPlace(NativeInteger(root), 2 * N, &v);
        for (uint32_t i = 0; i < N; ++i)
            x[i] = v[i].ConvertToInt<uint64_t>();
    });
}
int ref_ntt_inv(uint64_t q, uint64_t root, uint32_t N, uint64_t* x) {
    return guarded([&] {
        NativeVector v(N, NativeInteger(q));
        for (uint32_t i = 0; i < N; ++i)
            v[i] = NativeInteger(x[i]);
        ChineseRemainderTransformFTT<NativeVector>().InverseTransformFromBitReverseInPlace(NativeInteger(root), 2 * N,
                                                                                           &v);
        for (uint32_t i = 0; i < N; ++i)
            x[i] = v[i].ConvertToInt<uint64_t>();
    });
}
// the twiddle tables as the reference builds them (a5): T, T', Tinv, Tinv' each length N
int ref_twiddles(uint64_t q, uint64_t root, uint32_t N, uint64_t* T, uint64_t* Tp, uint64_t* Ti, uint64_t* Tip,
                 uint64_t* nInv, uint64_t* nInvPrecon) {
    return guarded([&] {
        NativeInteger Q(q);
        ChineseRemainderTransformFTT<NativeVector>().PreCompute(NativeInteger(root), 2 * N, Q);
        using X       = ChineseRemainderTransformFTT<NativeVector>;
        const auto& t = X::m_rootOfUnityReverseTableByModulus[Q];
        const auto& tp = X::m_rootOfUnityPreconReverseTableByModulus[Q];
        const auto& ti = X::m_rootOfUnityInverseReverseTableByModulus[Q];
        const auto& tip = X::m_rootOfUnityInversePreconReverseTableByModulus[Q];
        for (uint32_t i = 0; i < N; ++i) {
            T[i]   = t[i].ConvertToInt<uint64_t>();
            Tp[i]  = tp[i].ConvertToInt<uint64_t>();
            Ti[i]  = ti[i].ConvertToInt<uint64_t>();
            Tip[i] = tip[i].ConvertToInt<uint64_t>();
        }
        uint32_t msb = GetMSB(N - 1);
        *nInv        = X::m_cycloOrderInverseTableByModulus[Q][msb].ConvertToInt<uint64_t>();
        *nInvPrecon  = X::m_cycloOrderInversePreconTableByModulus[Q][msb].ConvertToInt<uint64_t>();
    });
}

// ---------------------------------------------------------------- DCRTPoly limb-wise ops (dcrtpoly.h / dcrtpoly-impl.h)
// op: 0 add, 1 sub, 2 mul (Hadamard), 3 negate(a)
int ref_dcrt_binop(uint32_t N, uint32_t L, const uint64_t* moduli, const uint64_t* roots, int op, const uint64_t* a,
                   const uint64_t* b, uint64_t* out) {
    return guarded([&] {
        auto params = make_params(N, L, moduli, roots);
        DCRTPoly A  = make_poly(params, Format::EVALUATION, a);
        if (op == 3) {
            dump_poly(A.Negate(), out);
            return;
        }
        DCRTPoly B = make_poly(params, Format::EVALUATION, b);
        if (op == 0)
            dump_poly(A + B, out);
        else if (op == 1)
            dump_poly(A - B, out);
        else
            dump_poly(A * B, out);
    });
}
// Times(std::vector<NativeInteger>) : per-limb scalar
int ref_dcrt_times_scalars(uint32_t N, uint32_t L, const uint64_t* moduli, const uint64_t* roots, const uint64_t* a,
                           const uint64_t* scalars, uint64_t* out) {
    return guarded([&] {
        auto params = make_params(N, L, moduli, roots);
        DCRTPoly A  = make_poly(params, Format::EVALUATION, a);
        std::vector<NativeInteger> s(L);
        for (uint32_t i = 0; i < L; ++i)
            s[i] = NativeInteger(scalars[i]);
        dump_poly(A.Times(s), out);
    });
}
int ref_dcrt_switch_format(uint32_t N, uint32_t L, const uint64_t* moduli, const uint64_t* roots, int from_eval,
                           uint64_t* data) {
    return guarded([&] {
        auto params = make_params(N, L, moduli, roots);
        DCRTPoly A  = make_poly(params, from_eval ? Format::EVALUATION : Format::COEFFICIENT, data);
        A.SwitchFormat();
        dump_poly(A, data);
    });
}
// Serial::Serialize(DCRTPoly, BINARY / JSON) of a poly built from caller-supplied residues: returns the byte count
// (0 on failure, or if `cap` is too small); format EVALUATION if eval != 0.   dcrtpoly.h:337-351, mubintvecnat.h:656-705
size_t ref_dcrt_serialize(uint32_t N, uint32_t L, const uint64_t* moduli, const uint64_t* roots, int eval,
                          const uint64_t* data, int json, uint8_t* out, size_t cap) {
    size_t n = 0;
    guarded([&] {
        auto params = make_params(N, L, moduli, roots);
        DCRTPoly A  = make_poly(params, eval ? Format::EVALUATION : Format::COEFFICIENT, data);
        std::stringstream ss;
        if (json)
            Serial::Serialize(A, ss, SerType::JSON);
        else
            Serial::Serialize(A, ss, SerType::BINARY);
        std::string b = ss.str();
        if (b.size() <= cap) {
            std::memcpy(out, b.data(), b.size());
            n = b.size();
        }
    });
    return n;
}
int ref_dcrt_automorphism(uint32_t N, uint32_t L, const uint64_t* moduli, const uint64_t* roots, uint32_t k,
                          const uint64_t* a, uint64_t* out) {
    return guarded([&] {
        auto params = make_params(N, L, moduli, roots);
        DCRTPoly A  = make_poly(params, Format::EVALUATION, a);
        std::vector<uint32_t> vec(N);
        PrecomputeAutoMap(N, k, &vec);
        dump_poly(A.AutomorphismTransform(k, vec), out);
    });
}
// COEFFICIENT-format automorphism: DCRTPolyImpl::AutomorphismTransform(k) -> PolyImpl::AutomorphismTransform(k)
int ref_dcrt_automorphism_coeff(uint32_t N, uint32_t L, const uint64_t* moduli, const uint64_t* roots, uint32_t k,
                                const uint64_t* a, uint64_t* out) {
    return guarded([&] {
        auto params = make_params(N, L, moduli, roots);
        DCRTPoly A  = make_poly(params, Format::COEFFICIENT, a);
        dump_poly(A.AutomorphismTransform(k), out);
    });
}
int ref_automorphism_map(uint32_t N, uint32_t k, uint32_t* map) {
    return guarded([&] {
        std::vector<uint32_t> vec(N);
        PrecomputeAutoMap(N, k, &vec);
        std::copy(vec.begin(), vec.end(), map);
    });
}
// NativeVector::SwitchModulus (mubintvecnat.cpp:109-122) on one limb
int ref_switch_modulus(uint32_t N, uint64_t q_from, uint64_t q_to, uint64_t* x) {
    return guarded([&] {
        NativeVector v(N, NativeInteger(q_from));
        for (uint32_t i = 0; i < N; ++i)
            v[i] = NativeInteger(x[i]);
        v.SwitchModulus(NativeInteger(q_to));
        for (uint32_t i = 0; i < N; ++i)
            x[i] = v[i].ConvertToInt<uint64_t>();
    });
}

// time L-limb DCRTPoly::SwitchFormat round trips (config 5 / bench reference arm); returns seconds per call
double ref_time_switch_format(uint32_t N, uint32_t L, const uint64_t* moduli, const uint64_t* roots, uint64_t* data,
                              int reps) {
    auto params = make_params(N, L, moduli, roots);
    DCRTPoly A  = make_poly(params, Format::COEFFICIENT, data);
    A.SwitchFormat();
    A.SwitchFormat();  // warm twiddle cache
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r)
        A.SwitchFormat();
    auto t1 = std::chrono::steady_clock::now();
    if (A.GetFormat() != Format::COEFFICIENT)
        A.SwitchFormat();
    dump_poly(A, data);
    return std::chrono::duration<double>(t1 - t0).count() / reps;
}

// ---------------------------------------------------------------- CKKS context (config 2 and friends)
// scal_tech: 0 FIXEDMANUAL, 1 FIXEDAUTO, 2 FLEXIBLEAUTO, 3 FLEXIBLEAUTOEXT
void* ref_ckks_create(uint32_t ring_dim, uint32_t depth, uint32_t scale_bits, uint32_t first_bits, uint32_t dnum,
                      int scal_tech) {
    CkksCtx* h = nullptr;
    int rc     = guarded([&] {
        CCParams<CryptoContextCKKSRNS> p;
        p.SetRingDim(ring_dim);
        p.SetMultiplicativeDepth(depth);
        p.SetScalingModSize(scale_bits);
        p.SetFirstModSize(first_bits);
        p.SetKeySwitchTechnique(HYBRID);
        p.SetSecurityLevel(HEStd_NotSet);
        if (dnum)
            p.SetNumLargeDigits(dnum);
        static const ScalingTechnique st[4] = {FIXEDMANUAL, FIXEDAUTO, FLEXIBLEAUTO, FLEXIBLEAUTOEXT};
        p.SetScalingTechnique(st[scal_tech]);
        auto cc = GenCryptoContext(p);
        cc->Enable(PKE);
        cc->Enable(KEYSWITCH);
        cc->Enable(LEVELEDSHE);
        h        = new CkksCtx();
        h->cc    = cc;
        h->keys  = cc->KeyGen();
        cc->EvalMultKeyGen(h->keys.secretKey);
        h->cp    = std::dynamic_pointer_cast<CryptoParametersCKKSRNS>(cc->GetCryptoParameters());
        h->N     = cc->GetRingDimension();
        h->sizeQ = h->cp->GetElementParams()->GetParams().size();
        h->sizeP = h->cp->GetParamsP()->GetParams().size();
        std::vector<double> msg(std::min<uint32_t>(8, h->N / 2), 0.25);
        auto pt  = cc->MakeCKKSPackedPlaintext(msg);
        h->tmpl  = cc->Encrypt(h->keys.publicKey, pt);
    });
    return rc == 0 ? h : nullptr;
}
void ref_ckks_destroy(void* hv) {
    auto* h = static_cast<CkksCtx*>(hv);
    if (h) {
        h->cc->ClearEvalMultKeys();
        h->cc->ClearEvalAutomorphismKeys();
        delete h;
    }
}
// shape[0..5] = N, sizeQ, sizeP, numPartQ (dnum), numPerPartQ (alpha), scaling technique
void ref_ckks_shape(void* hv, uint32_t* shape) {
    auto* h  = static_cast<CkksCtx*>(hv);
    shape[0] = h->N;
    shape[1] = h->sizeQ;
    shape[2] = h->sizeP;
    shape[3] = h->cp->GetNumPartQ();
    shape[4] = h->cp->GetNumPerPartQ();
    shape[5] = (uint32_t)h->cp->GetScalingTechnique();
}
// moduli/roots of Q (sizeQ) followed by P (sizeP)
void ref_ckks_moduli(void* hv, uint64_t* moduli, uint64_t* roots) {
    auto* h        = static_cast<CkksCtx*>(hv);
    const auto& ps = h->cp->GetParamsQP()->GetParams();
    for (size_t i = 0; i < ps.size(); ++i) {
        moduli[i] = ps[i]->GetModulus().ConvertToInt<uint64_t>();
        roots[i]  = ps[i]->GetRootOfUnity().ConvertToInt<uint64_t>();
    }
}

// ---- table dumpers (rns-cryptoparameters.h getters, a20) -------------------
void ref_ckks_PInvModq(void* hv, uint64_t* out) {
    copy_vec(static_cast<CkksCtx*>(hv)->cp->GetPInvModq(), out);
}
void ref_ckks_PModq(void* hv, uint64_t* out) {
    copy_vec(static_cast<CkksCtx*>(hv)->cp->GetPModq(), out);
}
void ref_ckks_PHatInvModp(void* hv, uint64_t* out) {
    copy_vec(static_cast<CkksCtx*>(hv)->cp->GetPHatInvModp(), out);
}
// out[j*sizeQ + i] = [P/p_j]_{q_i}
void ref_ckks_PHatModq(void* hv, uint64_t* out) {
    auto* h       = static_cast<CkksCtx*>(hv);
    const auto& t = h->cp->GetPHatModq();
    for (size_t j = 0; j < t.size(); ++j)
        copy_vec(t[j], out + j * h->sizeQ);
}
// digit `part`, truncated to `digit_size` limbs: out[i], i < digit_size
void ref_ckks_PartQlHatInvModq(void* hv, uint32_t part, uint32_t digit_size, uint64_t* out) {
    copy_vec(static_cast<CkksCtx*>(hv)->cp->GetPartQlHatInvModq(part, digit_size - 1), out);
}
// level l = sizeQl-1, digit `part`: out[i*sizeCompl + j]; returns sizeCompl; compl_moduli gets the target basis
uint32_t ref_ckks_PartQlHatModp(void* hv, uint32_t sizeQl, uint32_t part, uint64_t* out, uint64_t* compl_moduli) {
    auto* h       = static_cast<CkksCtx*>(hv);
    const auto& t = h->cp->GetPartQlHatModp(sizeQl - 1, part);
    const auto& c = h->cp->GetParamsComplPartQ(sizeQl - 1, part)->GetParams();
    uint32_t sc   = c.size();
    if (out)
        for (size_t i = 0; i < t.size(); ++i)
            copy_vec(t[i], out + i * sc);
    if (compl_moduli)
        for (uint32_t j = 0; j < sc; ++j)
            compl_moduli[j] = c[j]->GetModulus().ConvertToInt<uint64_t>();
    return sc;
}
// rescale tables for dropping limb l = sizeQl-1 (rns-cryptoparameters.h:314-345)
void ref_ckks_rescale_tables(void* hv, uint32_t sizeQl, uint64_t* QlQlInvModqlDivqlModq, uint64_t* qlInvModq) {
    auto* h = static_cast<CkksCtx*>(hv);
    copy_vec(h->cp->GetQlQlInvModqlDivqlModq(h->sizeQ - sizeQl), QlQlInvModqlDivqlModq);
    copy_vec(h->cp->GetqlInvModq(h->sizeQ - sizeQl), qlInvModq);
}

// ---- keys -----------------------------------------------------------------
// relinearisation key: a, b each [dnum][sizeQ+sizeP][N], EVALUATION
int ref_ckks_evalmult_key(void* hv, uint64_t* a, uint64_t* b) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        const auto& ek = h->cc->GetEvalMultKeyVector(h->keys.secretKey->GetKeyTag())[0];
        const auto& av = ek->GetAVector();
        const auto& bv = ek->GetBVector();
        size_t stride  = (size_t)(h->sizeQ + h->sizeP) * h->N;
        for (size_t j = 0; j < av.size(); ++j) {
            dump_poly(av[j], a + j * stride);
            dump_poly(bv[j], b + j * stride);
        }
    });
}
// rotation key for slot index `index`; returns the automorphism index k (0 on failure)
uint32_t ref_ckks_rotate_key(void* hv, int32_t index, uint64_t* a, uint64_t* b) {
    auto* h    = static_cast<CkksCtx*>(hv);
    uint32_t k = 0;
    guarded([&] {
        h->cc->EvalRotateKeyGen(h->keys.secretKey, {index});
        k              = h->cc->FindAutomorphismIndex(index);
        const auto& km = h->cc->GetEvalAutomorphismKeyMap(h->keys.secretKey->GetKeyTag());
        const auto& ek = km.at(k);
        const auto& av = ek->GetAVector();
        const auto& bv = ek->GetBVector();
        size_t stride  = (size_t)(h->sizeQ + h->sizeP) * h->N;
        for (size_t j = 0; j < av.size(); ++j) {
            dump_poly(av[j], a + j * stride);
            dump_poly(bv[j], b + j * stride);
        }
    });
    return k;
}

// ---- DCRTPoly-level pieces of the key switch on caller-supplied residues ----
// ModUp of one digit: a10 as called at keyswitch-hybrid.cpp:359-366.  in: [digit_size][N] COEFFICIENT limbs of
// digit `part` at level sizeQl; out: [sizeCompl][N] COEFFICIENT.
int ref_ckks_approx_switch_crt_basis_digit(void* hv, uint32_t sizeQl, uint32_t part, const uint64_t* in,
                                           uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        const uint32_t alpha = h->cp->GetNumPerPartQ();
        uint32_t digit_size  = std::min(alpha, sizeQl - alpha * part);
        const auto& pp       = h->cp->GetParamsPartQ(part)->GetParams();
        std::vector<NativeInteger> m(digit_size), r(digit_size);
        for (uint32_t i = 0; i < digit_size; ++i) {
            m[i] = pp[i]->GetModulus();
            r[i] = pp[i]->GetRootOfUnity();
        }
        auto params  = std::make_shared<ILDCRTParams<BigInteger>>(2 * h->N, m, r);
        DCRTPoly x   = make_poly(params, Format::COEFFICIENT, in);
        DCRTPoly res = x.ApproxSwitchCRTBasis(
            h->cp->GetParamsPartQ(part), h->cp->GetParamsComplPartQ(sizeQl - 1, part),
            h->cp->GetPartQlHatInvModq(part, digit_size - 1), h->cp->GetPartQlHatInvModqPrecon(part, digit_size - 1),
            h->cp->GetPartQlHatModp(sizeQl - 1, part), h->cp->GetmodComplPartqBarrettMu(sizeQl - 1, part));
        dump_poly(res, out);
    });
}
// EvalKeySwitchPrecomputeCore (a16): c [sizeQl][N] EVALUATION -> digits [numPartQl][sizeQl+sizeP][N]; returns numPartQl
int ref_ckks_modup(void* hv, uint32_t sizeQl, const uint64_t* c, uint64_t* digits) {
    auto* h = static_cast<CkksCtx*>(hv);
    int n   = -1;
    guarded([&] {
        DCRTPoly C  = make_poly(h->paramsQl(sizeQl), Format::EVALUATION, c);
        auto d      = h->cc->GetScheme()->EvalKeySwitchPrecomputeCore(C, h->cc->GetCryptoParameters());
        size_t strd = (size_t)(sizeQl + h->sizeP) * h->N;
        for (size_t j = 0; j < d->size(); ++j)
            dump_poly((*d)[j], digits + j * strd);
        n = d->size();
    });
    return n;
}
// ApproxModDown (a11): in [sizeQl+sizeP][N] EVALUATION -> out [sizeQl][N] EVALUATION
int ref_ckks_approx_mod_down(void* hv, uint32_t sizeQl, const uint64_t* in, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        DCRTPoly X = make_poly(h->paramsQlP(sizeQl), Format::EVALUATION, in);
        auto cp    = h->cp;
        DCRTPoly R = X.ApproxModDown(h->paramsQl(sizeQl), cp->GetParamsP(), cp->GetPInvModq(), cp->GetPInvModqPrecon(),
                                     cp->GetPHatInvModp(), cp->GetPHatInvModpPrecon(), cp->GetPHatModq(),
                                     cp->GetModqBarrettMu(), cp->GettInvModp(), cp->GettInvModpPrecon(), 0,
                                     cp->GettModqPrecon());
        dump_poly(R, out);
    });
}
// KeySwitchHYBRID::KeySwitchCore (a18) with the relinearisation key: c [sizeQl][N] -> out0,out1 [sizeQl][N]
int ref_ckks_keyswitch_core(void* hv, uint32_t sizeQl, const uint64_t* c, uint64_t* out0, uint64_t* out1) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        DCRTPoly C     = make_poly(h->paramsQl(sizeQl), Format::EVALUATION, c);
        const auto& ek = h->cc->GetEvalMultKeyVector(h->keys.secretKey->GetKeyTag())[0];
        auto ba        = h->cc->GetScheme()->KeySwitchCore(C, ek);
        dump_poly((*ba)[0], out0);
        dump_poly((*ba)[1], out1);
    });
}
// DropLastElementAndScale (a14) through the CKKS rescale tables: x [sizeQl][N] EVALUATION -> [sizeQl-1][N]
int ref_ckks_rescale_poly(void* hv, uint32_t sizeQl, const uint64_t* x, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        DCRTPoly X = make_poly(h->paramsQl(sizeQl), Format::EVALUATION, x);
        X.DropLastElementAndScale(h->cp->GetQlQlInvModqlDivqlModq(h->sizeQ - sizeQl),
                                  h->cp->GetqlInvModq(h->sizeQ - sizeQl));
        dump_poly(X, out);
    });
}

// ---- whole-operation calls through the CryptoContext API --------------------
static Ciphertext<DCRTPoly> make_ct(CkksCtx* h, uint32_t sizeQl, const uint64_t* data, uint32_t n_polys) {
    auto ct     = h->tmpl->Clone();
    auto params = h->paramsQl(sizeQl);
    std::vector<DCRTPoly> el;
    for (uint32_t k = 0; k < n_polys; ++k)
        el.push_back(make_poly(params, Format::EVALUATION, data + (size_t)k * sizeQl * h->N));
    ct->SetElements(std::move(el));
    ct->SetLevel(h->sizeQ - sizeQl);
    return ct;
}
// cc->EvalMult(ct1, ct2) (a19, cryptocontext.h:1871): ct* = 2 polys x [sizeQl][N]; out = 2 polys x [sizeQl][N].
// Valid for scaling techniques where EvalMult itself does not rescale its fresh inputs (all four at noiseScaleDeg 1).
int ref_ckks_evalmult(void* hv, uint32_t sizeQl, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto a = make_ct(h, sizeQl, ct1, 2);
        auto b = make_ct(h, sizeQl, ct2, 2);
        auto c = h->cc->EvalMult(a, b);
        if (c->GetElements().size() != 2 || c->GetElements()[0].GetNumOfElements() != sizeQl)
            OPENFHE_THROW("unexpected EvalMult output shape");
        dump_poly(c->GetElements()[0], out);
        dump_poly(c->GetElements()[1], out + (size_t)sizeQl * h->N);
    });
}
// times `reps` calls of cc->EvalMult on the given towers; returns seconds per op (negative on error)
double ref_ckks_time_evalmult(void* hv, uint32_t sizeQl, const uint64_t* ct1, const uint64_t* ct2, int warmup,
                              int reps) {
    auto* h  = static_cast<CkksCtx*>(hv);
    double s = -1;
    guarded([&] {
        auto a = make_ct(h, sizeQl, ct1, 2);
        auto b = make_ct(h, sizeQl, ct2, 2);
        for (int i = 0; i < warmup; ++i)
            auto c = h->cc->EvalMult(a, b);
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i)
            auto c = h->cc->EvalMult(a, b);
        auto t1 = std::chrono::steady_clock::now();
        s       = std::chrono::duration<double>(t1 - t0).count() / reps;
    });
    return s;
}
// the rescale every scaling technique ends up in (LeveledSHECKKSRNS::ModReduceInternalInPlace,
// ckksrns-leveledshe.cpp:172-191; cc->Rescale reaches it directly under FIXEDMANUAL): 2 polys [sizeQl][N] ->
// 2 polys [sizeQl-1][N]
int ref_ckks_rescale(void* hv, uint32_t sizeQl, const uint64_t* ct, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto a = make_ct(h, sizeQl, ct, 2);
        a->SetNoiseScaleDeg(2);
        auto c = h->cc->GetScheme()->ModReduceInternal(a, 1);
        if (c->GetElements()[0].GetNumOfElements() != sizeQl - 1)
            OPENFHE_THROW("unexpected rescale output shape");
        dump_poly(c->GetElements()[0], out);
        dump_poly(c->GetElements()[1], out + (size_t)(sizeQl - 1) * h->N);
    });
}
// cc->EvalRotate(ct, index) (needs ref_ckks_rotate_key(index) first): 2 polys [sizeQl][N] -> same shape
int ref_ckks_rotate(void* hv, uint32_t sizeQl, int32_t index, const uint64_t* ct, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto a = make_ct(h, sizeQl, ct, 2);
        auto c = h->cc->EvalRotate(a, index);
        dump_poly(c->GetElements()[0], out);
        dump_poly(c->GetElements()[1], out + (size_t)sizeQl * h->N);
    });
}
// ---- hoisted rotations: EvalFastRotationPrecompute + EvalFastRotation / EvalFastRotationExt /
// KeySwitchExt / KeySwitchDown on caller-supplied residues (rotation key of `index` must exist: ref_ckks_rotate_key)
// cc->EvalFastRotation(ct, index, M, cc->EvalFastRotationPrecompute(ct)): [2][sizeQl][N] -> same shape
int ref_ckks_fast_rotate(void* hv, uint32_t sizeQl, int32_t index, const uint64_t* ct, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto a      = make_ct(h, sizeQl, ct, 2);
        auto digits = h->cc->EvalFastRotationPrecompute(a);
        auto c      = h->cc->EvalFastRotation(a, index, 2 * h->N, digits);
        dump_poly(c->GetElements()[0], out);
        dump_poly(c->GetElements()[1], out + (size_t)sizeQl * h->N);
    });
}
// cc->EvalFastRotationExt(ct, index, digits, addFirst): [2][sizeQl][N] -> [2][sizeQl+sizeP][N] (extended basis QlP)
int ref_ckks_fast_rotate_ext(void* hv, uint32_t sizeQl, int32_t index, int add_first, const uint64_t* ct,
                             uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto a      = make_ct(h, sizeQl, ct, 2);
        auto digits = h->cc->EvalFastRotationPrecompute(a);
        auto c      = h->cc->EvalFastRotationExt(a, index, digits, add_first != 0);
        dump_poly(c->GetElements()[0], out);
        dump_poly(c->GetElements()[1], out + (size_t)(sizeQl + h->sizeP) * h->N);
    });
}
// cc->KeySwitchExt(ct, addFirst): [2][sizeQl][N] -> [2][sizeQl+sizeP][N]
int ref_ckks_keyswitch_ext(void* hv, uint32_t sizeQl, int add_first, const uint64_t* ct, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto a = make_ct(h, sizeQl, ct, 2);
        auto c = h->cc->KeySwitchExt(a, add_first != 0);
        dump_poly(c->GetElements()[0], out);
        dump_poly(c->GetElements()[1], out + (size_t)(sizeQl + h->sizeP) * h->N);
    });
}
// cc->KeySwitchDown on an extended ciphertext: [2][sizeQl+sizeP][N] -> [2][sizeQl][N]
int ref_ckks_keyswitch_down(void* hv, uint32_t sizeQl, const uint64_t* ct_ext, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto a = make_ct(h, sizeQl, ct_ext, 2);  // template for metadata
        auto e = h->cc->KeySwitchExt(a, true);   // right shape / params; towers overwritten below
        auto& ev = e->GetElements();
        for (uint32_t k = 0; k < 2; ++k)
            ev[k] = make_poly(ev[k].GetParams(), Format::EVALUATION, ct_ext + (size_t)k * (sizeQl + h->sizeP) * h->N);
        auto c = h->cc->KeySwitchDown(e);
        dump_poly(c->GetElements()[0], out);
        dump_poly(c->GetElements()[1], out + (size_t)sizeQl * h->N);
    });
}
// One level of the double-hoisted BSGS linear transform, run with the reference's OWN primitives in the order of
// FHECKKSRNS::EvalCoeffsToSlots' loop body (ckksrns-fhe.cpp:1934-1987): cc->EvalFastRotationPrecompute,
// cc->EvalFastRotationExt, cc->KeySwitchExt, cc->KeySwitchDown, cc->KeySwitchDownFirstElement; EvalMultExt /
// EvalAddExtInPlace (protected members of FHECKKSRNS, :2267-2290) are the element-wise products / sums they consist of.
// rot_in[g] / rot_out[b]: ROTATION indices (0 = none; keys must exist: ref_ckks_rotate_key); A: [b*g][sizeQl+sizeP][N]
// diagonals in the extended basis (EVALUATION); ct [2][sizeQl][N] -> out [2][sizeQl][N].
int ref_ckks_bsgs_level(void* hv, uint32_t sizeQl, uint32_t g, uint32_t b, const int32_t* rot_in, const int32_t* rot_out,
                        const uint64_t* A, uint32_t skip, const uint64_t* ct, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto cc       = h->cc;
        const size_t E = (size_t)(sizeQl + h->sizeP) * h->N;
        auto result   = make_ct(h, sizeQl, ct, 2);
        auto digits   = cc->EvalFastRotationPrecompute(result);
        std::vector<Ciphertext<DCRTPoly>> fr(g);
        for (uint32_t j = 0; j < g; ++j)
            fr[j] = (rot_in[j] != 0) ? cc->EvalFastRotationExt(result, rot_in[j], digits, true) : cc->KeySwitchExt(result, true);
        const auto paramsQlP = fr[0]->GetElements()[0].GetParams();
        auto multExt = [&](ConstCiphertext<DCRTPoly> c, uint32_t idx) {
            auto r = c->Clone();
            DCRTPoly pt = make_poly(paramsQlP, Format::EVALUATION, A + (size_t)idx * E);
            for (auto& e : r->GetElements())
                e *= pt;
            return r;
        };
        auto addExt = [&](Ciphertext<DCRTPoly>& a, ConstCiphertext<DCRTPoly> c) {
            auto& av = a->GetElements();
            const auto& cv = c->GetElements();
            for (size_t i = 0; i < av.size(); ++i)
                av[i] += cv[i];
        };
        const uint32_t N = h->N;
        std::vector<uint32_t> map(N);
        Ciphertext<DCRTPoly> outer;
        DCRTPoly first;
        for (uint32_t i = 0; i < b; ++i) {
            const uint32_t G = g * i;
            auto inner = multExt(fr[0], G);
            for (uint32_t j = 1; j < g; ++j)
                if (G + j != skip)
                    addExt(inner, multExt(fr[j], G + j));
            if (i == 0) {
                first = cc->KeySwitchDownFirstElement(inner);
                outer = std::move(inner);
                outer->GetElements()[0].SetValuesToZero();
            }
            else if (rot_out[i] != 0) {
                inner = cc->KeySwitchDown(inner);
                uint32_t autoIndex = FindAutomorphismIndex2nComplex(rot_out[i], cc->GetCyclotomicOrder());
                PrecomputeAutoMap(N, autoIndex, &map);
                first += inner->GetElements()[0].AutomorphismTransform(autoIndex, map);
                auto innerDigits = cc->EvalFastRotationPrecompute(inner);
                addExt(outer, cc->EvalFastRotationExt(inner, rot_out[i], innerDigits, false));
            }
            else {
                first += cc->KeySwitchDownFirstElement(inner);
                inner->GetElements()[0].SetValuesToZero();
                addExt(outer, inner);
            }
        }
        result = cc->KeySwitchDown(outer);
        result->GetElements()[0] += first;
        dump_poly(result->GetElements()[0], out);
        dump_poly(result->GetElements()[1], out + (size_t)sizeQl * h->N);
    });
}
// honest end-to-end check material: encrypt a real vector, return its towers; decrypt towers back.
int ref_ckks_encrypt(void* hv, const double* vals, uint32_t n, uint64_t* out) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        std::vector<double> v(vals, vals + n);
        auto pt = h->cc->MakeCKKSPackedPlaintext(v);
        auto ct = h->cc->Encrypt(h->keys.publicKey, pt);
        dump_poly(ct->GetElements()[0], out);
        dump_poly(ct->GetElements()[1], out + (size_t)h->sizeQ * h->N);
    });
}
// decrypt 2 polys at level sizeQl whose noiseScaleDeg is `deg` (2 after a multiplication without rescale)
int ref_ckks_decrypt(void* hv, uint32_t sizeQl, uint32_t deg, const uint64_t* ct, double* vals, uint32_t n) {
    auto* h = static_cast<CkksCtx*>(hv);
    return guarded([&] {
        auto a = make_ct(h, sizeQl, ct, 2);
        a->SetNoiseScaleDeg(deg);
        double sf = h->tmpl->GetScalingFactor();
        a->SetScalingFactor(deg == 2 ? sf * sf : sf);
        Plaintext pt;
        h->cc->Decrypt(h->keys.secretKey, a, &pt);
        pt->SetLength(n);
        auto v = pt->GetRealPackedValue();
        for (uint32_t i = 0; i < n; ++i)
            vals[i] = v[i];
    });
}


// ---------------------------------------------------------------- BFV (HPSPOVERQ) context: rows a12 / a13 (config 3)
struct BfvCtx {
    CryptoContext<DCRTPoly> cc;
    KeyPair<DCRTPoly> keys;
    std::shared_ptr<CryptoParametersBFVRNS> cp;
    Ciphertext<DCRTPoly> tmpl;
    uint32_t N, sizeQ, sizeR;
};
static void dump_mat(const std::vector<std::vector<NativeInteger>>& m, uint64_t* out) {  // row-major as stored
    size_t k = 0;
    for (auto& row : m)
        for (auto& v : row)
            out[k++] = v.ConvertToInt<uint64_t>();
}
void* ref_bfv_create(uint32_t ring_dim, uint32_t depth, uint32_t scale_bits, uint64_t ptm, uint32_t dnum) {
    BfvCtx* h = nullptr;
    int rc    = guarded([&] {
        CCParams<CryptoContextBFVRNS> p;
        p.SetRingDim(ring_dim);
        p.SetPlaintextModulus(ptm);
        p.SetMultiplicativeDepth(depth);
        p.SetScalingModSize(scale_bits);
        p.SetMultiplicationTechnique(HPSPOVERQ);
        p.SetKeySwitchTechnique(HYBRID);
        p.SetSecurityLevel(HEStd_NotSet);
        if (dnum)
            p.SetNumLargeDigits(dnum);
        auto cc = GenCryptoContext(p);
        cc->Enable(PKE);
        cc->Enable(KEYSWITCH);
        cc->Enable(LEVELEDSHE);
        h       = new BfvCtx();
        h->cc   = cc;
        h->keys = cc->KeyGen();
        cc->EvalMultKeyGen(h->keys.secretKey);
        h->cp    = std::dynamic_pointer_cast<CryptoParametersBFVRNS>(cc->GetCryptoParameters());
        h->N     = cc->GetRingDimension();
        h->sizeQ = h->cp->GetElementParams()->GetParams().size();
        h->sizeR = h->cp->GetParamsRl(h->sizeQ - 1)->GetParams().size();
        std::vector<int64_t> msg{1, 2, 3, 4};
        h->tmpl = cc->Encrypt(h->keys.publicKey, cc->MakePackedPlaintext(msg));
    });
    return rc == 0 ? h : nullptr;
}
void ref_bfv_destroy(void* hv) {
    auto* h = static_cast<BfvCtx*>(hv);
    if (h) {
        h->cc->ClearEvalMultKeys();
        delete h;
    }
}
// shape: N, sizeQ, sizeR;  moduli/roots: Q then R (the auxiliary basis of HPSPOVERQ at full level)
void ref_bfv_shape(void* hv, uint32_t* shape, uint64_t* moduli, uint64_t* roots) {
    auto* h  = static_cast<BfvCtx*>(hv);
    shape[0] = h->N;
    shape[1] = h->sizeQ;
    shape[2] = h->sizeR;
    if (moduli) {
        const auto& ps = h->cp->GetParamsQlRl(h->sizeQ - 1)->GetParams();
        for (size_t i = 0; i < ps.size(); ++i) {
            moduli[i] = ps[i]->GetModulus().ConvertToInt<uint64_t>();
            roots[i]  = ps[i]->GetRootOfUnity().ConvertToInt<uint64_t>();
        }
    }
}
// tables of the full-level EvalMult (bfvrns-leveledshe.cpp:239-261,385-396), l = sizeQ-1
//  which: 0 QlHatInvModq[sizeQ]  1 QlHatModr[sizeR][sizeQ]  2 alphaQlModr[sizeQ+1][sizeR]  3 qInv (double)[sizeQ]
//         4 mNegRlQlHatInvModq[sizeQ]  5 qInvModr[sizeQ][sizeR]  6 RlHatInvModr[sizeR]  7 RlHatModq[sizeQ][sizeR]
//         8 alphaRlModq[sizeR+1][sizeQ]  9 rInv (double)[sizeR]  10 tQlSlHatInvModsDivsModq[sizeQ][sizeR+1]
//         11 tQlSlHatInvModsDivsFrac (double)[sizeR]
int ref_bfv_table(void* hv, int which, void* out) {
    auto* h = static_cast<BfvCtx*>(hv);
    return guarded([&] {
        auto cp       = h->cp;
        uint32_t l    = h->sizeQ - 1;
        uint64_t* u   = static_cast<uint64_t*>(out);
        double* d     = static_cast<double*>(out);
        auto dump_dbl = [&](const std::vector<double>& v) { std::copy(v.begin(), v.end(), d); };
        switch (which) {
            case 0: copy_vec(cp->GetQlHatInvModq(l), u); break;
            case 1: dump_mat(cp->GetQlHatModr(l), u); break;
            case 2: dump_mat(cp->GetalphaQlModr(l), u); break;
            case 3: dump_dbl(cp->GetqInv()); break;
            case 4: copy_vec(cp->GetmNegRlQlHatInvModq(l), u); break;
            case 5: dump_mat(cp->GetqInvModr(), u); break;
            case 6: copy_vec(cp->GetRlHatInvModr(l), u); break;
            case 7: dump_mat(cp->GetRlHatModq(l), u); break;
            case 8: dump_mat(cp->GetalphaRlModq(l), u); break;
            case 9: dump_dbl(cp->GetrInv()); break;
            case 10: dump_mat(cp->GettQlSlHatInvModsDivsModq(l), u); break;
            case 11: dump_dbl(cp->GettQlSlHatInvModsDivsFrac(l)); break;
            default: OPENFHE_THROW("unknown table");
        }
    });
}
// dimensions of table `which` as (rows, cols); cols == 0 for vectors
void ref_bfv_table_dims(void* hv, int which, uint32_t* rows, uint32_t* cols) {
    auto* h    = static_cast<BfvCtx*>(hv);
    auto cp    = h->cp;
    uint32_t l = h->sizeQ - 1;
    auto md    = [&](const std::vector<std::vector<NativeInteger>>& m) { *rows = m.size(); *cols = m.empty() ? 0 : m[0].size(); };
    *cols      = 0;
    switch (which) {
        case 0: *rows = cp->GetQlHatInvModq(l).size(); break;
        case 1: md(cp->GetQlHatModr(l)); break;
        case 2: md(cp->GetalphaQlModr(l)); break;
        case 3: *rows = cp->GetqInv().size(); break;
        case 4: *rows = cp->GetmNegRlQlHatInvModq(l).size(); break;
        case 5: md(cp->GetqInvModr()); break;
        case 6: *rows = cp->GetRlHatInvModr(l).size(); break;
        case 7: md(cp->GetRlHatModq(l)); break;
        case 8: md(cp->GetalphaRlModq(l)); break;
        case 9: *rows = cp->GetrInv().size(); break;
        case 10: md(cp->GettQlSlHatInvModsDivsModq(l)); break;
        case 11: *rows = cp->GettQlSlHatInvModsDivsFrac(l).size(); break;
        default: *rows = 0;
    }
}
static std::shared_ptr<ILDCRTParams<BigInteger>> bfv_params(BfvCtx* h, int which /*0 Q, 1 R, 2 QR*/) {
    uint32_t l = h->sizeQ - 1;
    return which == 0 ? h->cp->GetParamsQl(l) : (which == 1 ? h->cp->GetParamsRl(l) : h->cp->GetParamsQlRl(l));
}
// ExpandCRTBasis Q -> QR (a12; exact SwitchCRTBasis inside): in [sizeQ][N] EVALUATION -> out [sizeQ+sizeR][N] EVALUATION
int ref_bfv_expand_crt_basis(void* hv, const uint64_t* in, uint64_t* out) {
    auto* h = static_cast<BfvCtx*>(hv);
    return guarded([&] {
        auto cp    = h->cp;
        uint32_t l = h->sizeQ - 1;
        DCRTPoly x = make_poly(bfv_params(h, 0), Format::EVALUATION, in);
        x.ExpandCRTBasis(cp->GetParamsQlRl(l), cp->GetParamsRl(l), cp->GetQlHatInvModq(l), cp->GetQlHatInvModqPrecon(l),
                         cp->GetQlHatModr(l), cp->GetalphaQlModr(l), cp->GetModrBarrettMu(), cp->GetqInv(),
                         Format::EVALUATION);
        dump_poly(x, out);
    });
}
// SwitchCRTBasis Q -> R alone: in [sizeQ][N] COEFFICIENT -> out [sizeR][N] COEFFICIENT
int ref_bfv_switch_crt_basis(void* hv, const uint64_t* in, uint64_t* out) {
    auto* h = static_cast<BfvCtx*>(hv);
    return guarded([&] {
        auto cp    = h->cp;
        uint32_t l = h->sizeQ - 1;
        DCRTPoly x = make_poly(bfv_params(h, 0), Format::COEFFICIENT, in);
        dump_poly(x.SwitchCRTBasis(cp->GetParamsRl(l), cp->GetQlHatInvModq(l), cp->GetQlHatInvModqPrecon(l),
                                   cp->GetQlHatModr(l), cp->GetalphaQlModr(l), cp->GetModrBarrettMu(), cp->GetqInv()),
                  out);
    });
}
// FastExpandCRTBasisPloverQ: in [sizeQ][N] COEFFICIENT -> out [sizeQ+sizeR][N] COEFFICIENT (bfvrns-leveledshe.cpp:248-261)
int ref_bfv_fast_expand_p_over_q(void* hv, const uint64_t* in, uint64_t* out) {
    auto* h = static_cast<BfvCtx*>(hv);
    return guarded([&] {
        auto cp    = h->cp;
        uint32_t l = h->sizeQ - 1;
        DCRTPoly::CRTBasisExtensionPrecomputations basisPQ(
            cp->GetParamsQlRl(l), cp->GetParamsRl(l), cp->GetParamsQl(l), cp->GetmNegRlQlHatInvModq(l),
            cp->GetmNegRlQlHatInvModqPrecon(l), cp->GetqInvModr(), cp->GetModrBarrettMu(), cp->GetRlHatInvModr(l),
            cp->GetRlHatInvModrPrecon(l), cp->GetRlHatModq(l), cp->GetalphaRlModq(l), cp->GetModqBarrettMu(), cp->GetrInv());
        DCRTPoly x = make_poly(bfv_params(h, 0), Format::COEFFICIENT, in);
        x.FastExpandCRTBasisPloverQ(basisPQ);
        dump_poly(x, out);
    });
}
// ScaleAndRound QR -> Q by t/R (a13): in [sizeQ+sizeR][N] COEFFICIENT -> out [sizeQ][N] COEFFICIENT
int ref_bfv_scale_and_round(void* hv, const uint64_t* in, uint64_t* out) {
    auto* h = static_cast<BfvCtx*>(hv);
    return guarded([&] {
        auto cp    = h->cp;
        uint32_t l = h->sizeQ - 1;
        DCRTPoly x = make_poly(bfv_params(h, 2), Format::COEFFICIENT, in);
        dump_poly(x.ScaleAndRound(cp->GetParamsQl(l), cp->GettQlSlHatInvModsDivsModq(l), cp->GettQlSlHatInvModsDivsFrac(l),
                                  cp->GetModqBarrettMu()),
                  out);
    });
}
// cc->EvalMultNoRelin on caller-supplied towers: ct* = 2 polys [sizeQ][N] EVALUATION -> out = 3 polys [sizeQ][N] COEFFICIENT
int ref_bfv_evalmult_norelin(void* hv, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out) {
    auto* h = static_cast<BfvCtx*>(hv);
    return guarded([&] {
        auto mk = [&](const uint64_t* d) {
            auto ct = h->tmpl->Clone();
            std::vector<DCRTPoly> el;
            for (uint32_t k = 0; k < 2; ++k)
                el.push_back(make_poly(bfv_params(h, 0), Format::EVALUATION, d + (size_t)k * h->sizeQ * h->N));
            ct->SetElements(std::move(el));
            return ct;
        };
        auto c = h->cc->EvalMultNoRelin(mk(ct1), mk(ct2));
        if (c->GetElements().size() != 3)
            OPENFHE_THROW("unexpected BFV EvalMult output shape");
        for (uint32_t k = 0; k < 3; ++k) {
            DCRTPoly e = c->GetElements()[k];
            e.SetFormat(Format::COEFFICIENT);
            dump_poly(e, out + (size_t)k * h->sizeQ * h->N);
        }
    });
}

// key-switching side of the BFV context (HYBRID): shape = sizeP, dnum (numPartQ), alpha (numPerPartQ); P moduli / roots
void ref_bfv_ks_shape(void* hv, uint32_t* shape, uint64_t* moduliP, uint64_t* rootsP) {
    auto* h  = static_cast<BfvCtx*>(hv);
    const auto& pp = h->cp->GetParamsP()->GetParams();
    shape[0] = pp.size();
    shape[1] = h->cp->GetNumPartQ();
    shape[2] = h->cp->GetNumPerPartQ();
    if (moduliP)
        for (size_t i = 0; i < pp.size(); ++i) {
            moduliP[i] = pp[i]->GetModulus().ConvertToInt<uint64_t>();
            rootsP[i]  = pp[i]->GetRootOfUnity().ConvertToInt<uint64_t>();
        }
}
// relinearisation key: a, b as [dnum][sizeQ+sizeP][N] EVALUATION
int ref_bfv_evalmult_key(void* hv, uint64_t* a, uint64_t* b) {
    auto* h = static_cast<BfvCtx*>(hv);
    return guarded([&] {
        const auto& ek = h->cc->GetEvalMultKeyVector(h->keys.secretKey->GetKeyTag())[0];
        const auto& av = ek->GetAVector();
        const auto& bv = ek->GetBVector();
        size_t stride  = (size_t)(h->sizeQ + h->cp->GetParamsP()->GetParams().size()) * h->N;
        for (size_t j = 0; j < av.size(); ++j) {
            dump_poly(av[j], a + j * stride);
            dump_poly(bv[j], b + j * stride);
        }
    });
}
// cc->EvalMult(ct1, ct2) (HPSPOVERQ tensor + RelinearizeCore, bfvrns-leveledshe.cpp:735-741,888-935):
// ct* = 2 polys [sizeQ][N] EVALUATION -> out = 2 polys [sizeQ][N] EVALUATION
int ref_bfv_evalmult(void* hv, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out) {
    auto* h = static_cast<BfvCtx*>(hv);
    return guarded([&] {
        auto mk = [&](const uint64_t* d) {
            auto ct = h->tmpl->Clone();
            std::vector<DCRTPoly> el;
            for (uint32_t k = 0; k < 2; ++k)
                el.push_back(make_poly(bfv_params(h, 0), Format::EVALUATION, d + (size_t)k * h->sizeQ * h->N));
            ct->SetElements(std::move(el));
            return ct;
        };
        auto c = h->cc->EvalMult(mk(ct1), mk(ct2));
        if (c->GetElements().size() != 2)
            OPENFHE_THROW("unexpected BFV EvalMult output shape");
        for (uint32_t k = 0; k < 2; ++k) {
            DCRTPoly e = c->GetElements()[k];
            e.SetFormat(Format::EVALUATION);
            dump_poly(e, out + (size_t)k * h->sizeQ * h->N);
        }
    });
}

}  // extern "C"
