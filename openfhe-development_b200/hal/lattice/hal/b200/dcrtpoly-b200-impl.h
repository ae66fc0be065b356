// dcrtpoly-b200-impl.h -- the device methods of lbcrypto::b200::DCRTPolyB200 (see dcrtpoly-b200.h).
//
// Every method states the reference function whose results it reproduces bit for bit
// (src/core/include/lattice/hal/default/dcrtpoly-impl.h unless another file is named) and is a composition of
// calls into libofhe_b200.so: towers go up to the calling thread's device slabs, the kernels run on that thread's
// stream, the result towers come back into the host mirror.  No arithmetic on residues happens on the host here.
#ifndef LBCRYPTO_INC_LATTICE_HAL_B200_DCRTPOLY_B200_IMPL_H
#define LBCRYPTO_INC_LATTICE_HAL_B200_DCRTPOLY_B200_IMPL_H

#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "lattice/hal/b200/dcrtpoly-b200.h"

namespace lbcrypto {
namespace b200 {
namespace detail {

using Tower = PolyImpl<NativeVector>;

// one device operation of the calling thread
struct Op {
    Runtime& rt;
    Runtime::Local& l;
    uint32_t N;
    explicit Op(uint32_t n) : rt(Runtime::Get()), l(rt.Mine()), N(n) {}
    uint64_t* Slab(int which, size_t limbs) {
        return rt.Slab(l, which, limbs * N);
    }
    void Up(uint64_t* d, const std::vector<Tower>& t, size_t first, size_t count) {
        for (size_t i = 0; i < count; ++i)
            Check(ofhe_memcpy_h2d(rt.ctx, d + i * N, &t[first + i].GetValues()[0], (size_t)N * 8, l.stream), "h2d");
        rt.n.h2d_bytes += count * (size_t)N * 8;
    }
    void Up1(uint64_t* d, const Tower& t) {
        Check(ofhe_memcpy_h2d(rt.ctx, d, &t.GetValues()[0], (size_t)N * 8, l.stream), "h2d");
        rt.n.h2d_bytes += (size_t)N * 8;
    }
    // towers must own N-word value vectors (constructed with initializeElementToZero = true, or copies)
    void Down(const uint64_t* d, std::vector<Tower>& t, size_t first, size_t count) {
        for (size_t i = 0; i < count; ++i)
            Check(ofhe_memcpy_d2h(rt.ctx, &t[first + i][0], d + i * N, (size_t)N * 8, l.stream), "d2h");
        rt.n.d2h_bytes += count * (size_t)N * 8;
    }
    void Sync() {
        Check(ofhe_stream_sync(rt.ctx, l.stream), "sync");
    }
};

inline void Collect(const std::vector<Tower>& towers, size_t first, size_t count, std::vector<uint64_t>& mod,
                    std::vector<uint64_t>& root) {
    for (size_t i = 0; i < count; ++i) {
        mod.push_back(towers[first + i].GetModulus().ConvertToInt());
        root.push_back(Runtime::RootOr(towers[first + i].GetRootOfUnity().ConvertToInt(),
                                       towers[first + i].GetParams()->GetRingDimension(), mod.back()));
    }
}
template <typename P>
inline void CollectParams(const std::shared_ptr<P>& params, size_t count, std::vector<uint64_t>& mod,
                          std::vector<uint64_t>& root) {
    const auto& ps = params->GetParams();
    for (size_t i = 0; i < count; ++i) {
        mod.push_back(ps[i]->GetModulus().ConvertToInt());
        root.push_back(Runtime::RootOr(ps[i]->GetRootOfUnity().ConvertToInt(), ps[i]->GetRingDimension(), mod.back()));
    }
}
inline std::vector<uint32_t> Iota(uint32_t first, uint32_t count) {
    std::vector<uint32_t> v(count);
    std::iota(v.begin(), v.end(), first);
    return v;
}
inline std::vector<uint64_t> Words(const std::vector<NativeInteger>& v, size_t count) {
    std::vector<uint64_t> w(count);
    for (size_t i = 0; i < count; ++i)
        w[i] = v[i].ConvertToInt();
    return w;
}
inline void SetFormats(std::vector<Tower>& t, size_t first, size_t count, Format f) {
    for (size_t i = 0; i < count; ++i)
        t[first + i].OverrideFormat(f);
}
inline void AppendKey(std::string& k, const void* p, size_t bytes) {
    k.append(static_cast<const char*>(p), bytes);
}

// table-driven fast base conversion src -> dst on ring indices [0, nSrc) -> [nSrc, nSrc + nDst):
// a10 ApproxSwitchCRTBasis (:888-932) when alpha == nullptr, a12 exact SwitchCRTBasis (:1008-1085) otherwise
inline ofhe_bconv* ConvTable(ofhe_ring* ring, uint32_t nSrc, uint32_t nDst, const std::vector<NativeInteger>& QHatInvModq,
                             const std::vector<std::vector<NativeInteger>>& QHatModp, bool dstMajor,
                             const std::vector<std::vector<NativeInteger>>* alpha, const std::vector<double>* qInv) {
    std::vector<uint64_t> hinv = Words(QHatInvModq, nSrc);
    std::vector<uint64_t> mat((size_t)nSrc * nDst);
    if (dstMajor) {
        for (uint32_t j = 0; j < nDst; ++j)
            for (uint32_t i = 0; i < nSrc; ++i)
                mat[(size_t)j * nSrc + i] = QHatModp[j][i].ConvertToInt();
    }
    else {
        for (uint32_t i = 0; i < nSrc; ++i)
            for (uint32_t j = 0; j < nDst; ++j)
                mat[(size_t)i * nDst + j] = QHatModp[i][j].ConvertToInt();
    }
    std::vector<uint64_t> al;
    if (alpha) {
        al.resize((size_t)(nSrc + 1) * nDst);
        for (uint32_t k = 0; k <= nSrc; ++k)
            for (uint32_t j = 0; j < nDst; ++j)
                al[(size_t)k * nDst + j] = (*alpha)[k][j].ConvertToInt();
    }
    std::string key;
    AppendKey(key, &ring, sizeof ring);
    AppendKey(key, &nSrc, 4);
    AppendKey(key, &nDst, 4);
    AppendKey(key, hinv.data(), hinv.size() * 8);
    AppendKey(key, mat.data(), mat.size() * 8);
    if (alpha) {
        AppendKey(key, al.data(), al.size() * 8);
        AppendKey(key, qInv->data(), nSrc * sizeof(double));
    }
    return Runtime::Get().BConv(key, [&]() {
        const std::vector<uint32_t> src = Iota(0, nSrc), dst = Iota(nSrc, nDst);
        ofhe_bconv* t = nullptr;
        Check(ofhe_bconv_create(ring, nSrc, src.data(), hinv.data(), nDst, dst.data(), mat.data(), dstMajor ? 1 : 0,
                                alpha ? al.data() : nullptr, alpha ? qInv->data() : nullptr, &t),
              "ofhe_bconv_create");
        return t;
    });
}

}  // namespace detail

#define B200T template <typename VecType>
#define B200C DCRTPolyB200<VecType>

// ---- a6-a8  DCRTPolyImpl::SwitchFormat (:1932-1940): every tower through the forward / inverse negacyclic NTT
B200T void B200C::SwitchFormat(uint32_t) {
    auto& towers      = m_h.GetAllElements();
    const uint32_t L  = (uint32_t)towers.size();
    const Format from = m_h.GetFormat(), to = (from == Format::EVALUATION) ? Format::COEFFICIENT : Format::EVALUATION;
    if (L == 0) {
        m_h.OverrideFormat(to);
        return;
    }
    if (m_h.GetParams()->GetCyclotomicOrder() != 2 * m_h.GetParams()->GetRingDimension()) {
        // arbitrary (non power-of-two) cyclotomics are outside this backend's scope (the RNS
        // schemes only use power-of-two rings); the mirror's Bluestein transform serves them as in the reference
        m_h.SwitchFormat();
        return;
    }
    m_h.OverrideFormat(to);
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, 0, L, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    uint64_t* d     = op.Slab(0, L);
    op.Up(d, towers, 0, L);
    const auto idx = detail::Iota(0, L);
    if (to == Format::EVALUATION)
        Check(ofhe_ntt_fwd(ring, d, d, L, idx.data(), op.l.stream), "ofhe_ntt_fwd");
    else
        Check(ofhe_ntt_inv(ring, d, d, L, idx.data(), op.l.stream), "ofhe_ntt_inv");
    op.Down(d, towers, 0, L);
    op.Sync();
    detail::SetFormats(towers, 0, L, to);
    op.rt.n.ntt += L;
}

// ---- a9  Plus / Minus / Times / Negate and the Eq forms (dcrtpoly.h:130-190, dcrtpoly-impl.h:347-621):
// op 0 add, 1 sub, 2 mul, 3 negate; same argument checks and exceptions as the reference
namespace detail {
template <typename Host>
inline void CheckPair(const Host& a, const Host& b, bool mul) {
    if (a.GetParams()->GetRingDimension() != b.GetParams()->GetRingDimension())
        OPENFHE_THROW("RingDimension mismatch");
    if (mul) {
        if (a.GetFormat() != Format::EVALUATION || b.GetFormat() != Format::EVALUATION)
            OPENFHE_THROW("operator* for DCRTPolyImpl supported only in Format::EVALUATION");
    }
    else if (a.GetFormat() != b.GetFormat())
        OPENFHE_THROW("Format mismatch");
    if (a.GetAllElements().size() != b.GetAllElements().size())
        OPENFHE_THROW(mul ? "tower size mismatch; cannot multiply" : "tower size mismatch; cannot add");
    if (a.GetAllElements().size() && a.GetAllElements()[0].GetModulus() != b.GetAllElements()[0].GetModulus())
        OPENFHE_THROW("Modulus mismatch");
}
// dst towers (may be a's) <- a (op) b
template <typename Host>
inline void Binop(int opc, const Host& a, const Host* b, Host& dst) {
    const auto& ta   = a.GetAllElements();
    const uint32_t L = (uint32_t)ta.size();
    if (L == 0)
        return;
    Op op(a.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    Collect(ta, 0, L, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    uint64_t* da    = op.Slab(0, L);
    uint64_t* db    = da;
    op.Up(da, ta, 0, L);
    if (b) {
        if (b->GetAllElements().size() < L)
            OPENFHE_THROW("b200 backend: right-hand side has fewer towers than the left-hand side");
        db = op.Slab(1, L);
        op.Up(db, b->GetAllElements(), 0, L);
    }
    const auto idx = Iota(0, L);
    Check(ofhe_ew_binop(ring, opc, da, da, db, L, idx.data(), op.l.stream), "ofhe_ew_binop");
    op.Down(da, dst.GetAllElements(), 0, L);
    op.Sync();
    op.rt.n.ew += L;
}
}  // namespace detail

// The Eq forms perform NO checks in the reference (:374-380, :402-408, dcrtpoly.h:117-123): they combine this
// polynomial's towers with the FIRST as many towers of rhs, and callers rely on that (rhs with more towers:
// FHECKKSRNS::KeySwitchSparse, ckksrns-fhe.cpp:3650).  A shorter rhs is undefined behaviour there and an error here.
B200T B200C& B200C::operator+=(const DCRTPolyType& rhs) {
    detail::Binop(0, m_h, &rhs.m_h, m_h);
    return *this;
}
B200T B200C& B200C::operator-=(const DCRTPolyType& rhs) {
    detail::Binop(1, m_h, &rhs.m_h, m_h);
    return *this;
}
B200T B200C& B200C::operator*=(const DCRTPolyType& rhs) {
    detail::Binop(2, m_h, &rhs.m_h, m_h);
    return *this;
}
B200T B200C B200C::Plus(const DCRTPolyType& rhs) const {
    detail::CheckPair(m_h, rhs.m_h, false);
    Host r(m_h.GetParams(), m_h.GetFormat(), true);
    detail::Binop(0, m_h, &rhs.m_h, r);
    return r;
}
B200T B200C B200C::Minus(const DCRTPolyType& rhs) const {
    detail::CheckPair(m_h, rhs.m_h, false);
    Host r(m_h.GetParams(), m_h.GetFormat(), true);
    detail::Binop(1, m_h, &rhs.m_h, r);
    return r;
}
B200T B200C B200C::Times(const DCRTPolyType& rhs) const {
    detail::CheckPair(m_h, rhs.m_h, true);
    Host r(m_h.GetParams(), m_h.GetFormat(), true);
    detail::Binop(2, m_h, &rhs.m_h, r);
    return r;
}
B200T B200C B200C::Negate() const {
    Host r(m_h.GetParams(), m_h.GetFormat(), true);
    detail::Binop<Host>(3, m_h, nullptr, r);
    return r;
}
B200T B200C B200C::operator-() const {
    return Negate();
}

// ---- a9  Times(const std::vector<NativeInteger>&) (:553-567) and TimesNoCheck (:569-577)
namespace detail {
template <typename Host>
inline Host TimesScalars(const Host& a, const std::vector<NativeInteger>& s, uint32_t L) {
    Host r(a.GetParams(), a.GetFormat(), true);
    if (L == 0)
        return r;
    Op op(a.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    Collect(a.GetAllElements(), 0, L, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    uint64_t* d     = op.Slab(0, L);
    op.Up(d, a.GetAllElements(), 0, L);
    const auto idx = Iota(0, L);
    const auto w   = Words(s, L);
    Check(ofhe_ew_mul_scalars(ring, d, d, w.data(), L, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    op.Down(d, r.GetAllElements(), 0, L);
    op.Sync();
    op.rt.n.ew += L;
    return r;
}
}  // namespace detail
B200T B200C B200C::Times(const std::vector<NativeInteger>& rhs) const {
    if (m_h.GetAllElements().size() != rhs.size())
        OPENFHE_THROW("tower size mismatch; cannot multiply");
    return detail::TimesScalars(m_h, rhs, (uint32_t)rhs.size());
}
B200T B200C B200C::TimesNoCheck(const std::vector<NativeInteger>& rhs) const {
    const uint32_t L = (uint32_t)std::min(m_h.GetAllElements().size(), rhs.size());
    Host r           = detail::TimesScalars(m_h, rhs, L);
    return r;
}

// ---- a15  AutomorphismTransform (:312-333 -> poly-impl.h:344-376).  EVALUATION: gather through the bit-reversed
// index map (the device builds the same map as PrecomputeAutoMap, nbtheory2.cpp:264-275, so `vec` is not shipped);
// COEFFICIENT: signed scatter out[jk mod N] = +-in[j] (poly-impl.h:359-362).
B200T B200C B200C::AutomorphismTransform(uint32_t k, const std::vector<uint32_t>&) const {
    return AutomorphismTransform(k);
}
B200T B200C B200C::AutomorphismTransform(uint32_t k) const {
    if (k % 2 == 0)
        OPENFHE_THROW("automorphism index should be odd\n");
    const auto& towers = m_h.GetAllElements();
    const uint32_t L   = (uint32_t)towers.size();
    Host r(m_h);  // same params, format, tower parameters; values overwritten below
    if (L == 0)
        return r;
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, 0, L, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    uint64_t* src   = op.Slab(0, L);
    uint64_t* dst   = op.Slab(1, L);
    op.Up(src, towers, 0, L);
    if (m_h.GetFormat() == Format::EVALUATION)
        Check(ofhe_automorphism(ring, dst, src, L, k, op.l.stream), "ofhe_automorphism");
    else {
        const auto idx = detail::Iota(0, L);
        Check(ofhe_automorphism_coeff(ring, dst, src, L, idx.data(), k, op.l.stream), "ofhe_automorphism_coeff");
    }
    op.Down(dst, r.GetAllElements(), 0, L);
    op.Sync();
    op.rt.n.automorph += L;
    return r;
}

// ---- a14  DropLastElementAndScale (:693-712) and f4  ModReduce (:736-754): both are
//   z = INTT(last tower) [* c];  x_i <- f(x_i, NTT(SwitchModulus(z, q_l -> q_i)))  for the remaining towers
namespace detail {
// returns device slab 2 = [L-1][N]: the dropped tower (COEFFICIENT, times `pre` if not null) switched into every
// remaining modulus (mubintvecnat.cpp:109-122), times post[i] if given, in the polynomial's format
template <typename Host>
inline uint64_t* DroppedTowerEverywhere(Op& op, ofhe_ring* ring, const Host& h, uint32_t L, const uint64_t* pre,
                                        const std::vector<uint64_t>* post) {
    const auto& towers = h.GetAllElements();
    uint64_t* z        = op.Slab(1, 1);
    uint64_t* sw       = op.Slab(2, L - 1);
    const uint32_t last = L - 1;
    op.Up1(z, towers[last]);
    if (h.GetFormat() == Format::EVALUATION)
        Check(ofhe_ntt_inv(ring, z, z, 1, &last, op.l.stream), "ofhe_ntt_inv");
    if (pre)
        Check(ofhe_ew_mul_scalars(ring, z, z, pre, 1, &last, op.l.stream), "ofhe_ew_mul_scalars");
    for (uint32_t i = 0; i < last; ++i)
        Check(ofhe_switch_modulus(ring, sw + (size_t)i * op.N, z, last, i, op.l.stream), "ofhe_switch_modulus");
    const auto idx = Iota(0, last);
    if (post)
        Check(ofhe_ew_mul_scalars(ring, sw, sw, post->data(), last, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    if (h.GetFormat() == Format::EVALUATION)
        Check(ofhe_ntt_fwd(ring, sw, sw, last, idx.data(), op.l.stream), "ofhe_ntt_fwd");
    op.rt.n.ntt += L;
    return sw;
}
}  // namespace detail

B200T void B200C::DropLastElementAndScale(const std::vector<NativeInteger>& QlQlInvModqlDivqlModq,
                                          const std::vector<NativeInteger>& qlInvModq) {
    const uint32_t L = (uint32_t)m_h.GetAllElements().size();
    if (L < 2)
        OPENFHE_THROW("DropLastElementAndScale needs at least two towers");
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(m_h.GetAllElements(), 0, L, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    const auto post = detail::Words(QlQlInvModqlDivqlModq, L - 1);
    uint64_t* t     = detail::DroppedTowerEverywhere(op, ring, m_h, L, nullptr, &post);
    uint64_t* x     = op.Slab(0, L - 1);
    op.Up(x, m_h.GetAllElements(), 0, L - 1);
    const auto idx = detail::Iota(0, L - 1);
    const auto s   = detail::Words(qlInvModq, L - 1);
    Check(ofhe_ew_mul_scalars(ring, x, x, s.data(), L - 1, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    Check(ofhe_ew_binop(ring, 0, x, x, t, L - 1, idx.data(), op.l.stream), "ofhe_ew_binop");
    const Format f = m_h.GetFormat();
    if (f == Format::COEFFICIENT)  // the reference leaves the towers in EVALUATION in this case (:708-709)
        Check(ofhe_ntt_fwd(ring, x, x, L - 1, idx.data(), op.l.stream), "ofhe_ntt_fwd");
    m_h.DropLastElement();
    op.Down(x, m_h.GetAllElements(), 0, L - 1);
    op.Sync();
    if (f == Format::COEFFICIENT)
        detail::SetFormats(m_h.GetAllElements(), 0, L - 1, Format::EVALUATION);
    op.rt.n.rescale += 1;
}

B200T void B200C::ModReduce(const NativeInteger& t, const std::vector<NativeInteger>&, const NativeInteger& negtInvModq,
                            const NativeInteger&, const std::vector<NativeInteger>& qlInvModq,
                            const std::vector<NativeInteger>&) {
    const uint32_t L = (uint32_t)m_h.GetAllElements().size();
    if (L < 2)
        OPENFHE_THROW("ModReduce needs at least two towers");
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(m_h.GetAllElements(), 0, L, mod, root);
    ofhe_ring* ring    = op.rt.Ring(op.N, mod, root);
    const uint64_t pre = negtInvModq.ConvertToInt();
    uint64_t* d        = detail::DroppedTowerEverywhere<Host>(op, ring, m_h, L, &pre, nullptr);
    uint64_t* x        = op.Slab(0, L - 1);
    op.Up(x, m_h.GetAllElements(), 0, L - 1);
    const auto idx = detail::Iota(0, L - 1);
    const std::vector<uint64_t> tt(L - 1, t.ConvertToInt());
    Check(ofhe_ew_mul_scalars(ring, d, d, tt.data(), L - 1, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    Check(ofhe_ew_binop(ring, 0, x, x, d, L - 1, idx.data(), op.l.stream), "ofhe_ew_binop");
    const auto s = detail::Words(qlInvModq, L - 1);
    Check(ofhe_ew_mul_scalars(ring, x, x, s.data(), L - 1, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    m_h.DropLastElement();
    op.Down(x, m_h.GetAllElements(), 0, L - 1);
    op.Sync();
    op.rt.n.rescale += 1;
}

// ---- a10  ApproxSwitchCRTBasis (:888-932): [sizeQ x N] residues -> [sizeP x N], format label carried over
B200T B200C B200C::ApproxSwitchCRTBasis(const std::shared_ptr<Params>& paramsQ, const std::shared_ptr<Params>& paramsP,
                                        const std::vector<NativeInteger>& QHatInvModq, const std::vector<NativeInteger>&,
                                        const std::vector<std::vector<NativeInteger>>& QHatModp,
                                        const std::vector<DoubleNativeInt>&) const {
    Host ans(paramsP, m_h.GetFormat(), true);
    const auto& towers   = m_h.GetAllElements();
    const uint32_t sizeQ = (uint32_t)std::min(towers.size(), paramsQ->GetParams().size());
    const uint32_t sizeP = (uint32_t)ans.GetAllElements().size();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, 0, sizeQ, mod, root);
    detail::Collect(ans.GetAllElements(), 0, sizeP, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, sizeQ, sizeP, QHatInvModq, QHatModp, false, nullptr, nullptr);
    uint64_t* in    = op.Slab(0, sizeQ);
    uint64_t* out   = op.Slab(1, sizeP);
    op.Up(in, towers, 0, sizeQ);
    Check(ofhe_bconv_apply(tbl, in, out, op.l.stream), "ofhe_bconv_apply");
    op.Down(out, ans.GetAllElements(), 0, sizeP);
    op.Sync();
    op.rt.n.bconv += 1;
    return ans;
}

// ---- a16 core  ApproxModUp (:935-963): Q -> Q u P, result in EVALUATION; the Q towers keep their EVALUATION values
// when the input was in EVALUATION (no NTT round trip), exactly as the reference
B200T void B200C::ApproxModUp(const std::shared_ptr<Params>& paramsQ, const std::shared_ptr<Params>& paramsP,
                              const std::shared_ptr<Params>& paramsQP, const std::vector<NativeInteger>& QHatInvModq,
                              const std::vector<NativeInteger>&, const std::vector<std::vector<NativeInteger>>& QHatModp,
                              const std::vector<DoubleNativeInt>&) {
    auto& towers         = m_h.GetAllElements();
    const uint32_t sizeQ = (uint32_t)std::min(towers.size(), paramsQ->GetParams().size());
    const uint32_t sizeP = (uint32_t)paramsP->GetParams().size();
    const bool wasEval   = m_h.GetFormat() == Format::EVALUATION;
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, 0, sizeQ, mod, root);
    detail::CollectParams(paramsP, sizeP, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, sizeQ, sizeP, QHatInvModq, QHatModp, false, nullptr, nullptr);
    uint64_t* q     = op.Slab(0, sizeQ);
    uint64_t* c     = op.Slab(1, sizeQ);
    uint64_t* p     = op.Slab(2, sizeP);
    const auto qi = detail::Iota(0, sizeQ), pi = detail::Iota(sizeQ, sizeP);
    op.Up(q, towers, 0, sizeQ);
    if (wasEval)
        Check(ofhe_ntt_inv(ring, q, c, sizeQ, qi.data(), op.l.stream), "ofhe_ntt_inv");
    Check(ofhe_bconv_apply(tbl, wasEval ? c : q, p, op.l.stream), "ofhe_bconv_apply");
    Check(ofhe_ntt_fwd(ring, p, p, sizeP, pi.data(), op.l.stream), "ofhe_ntt_fwd");
    if (!wasEval)
        Check(ofhe_ntt_fwd(ring, q, q, sizeQ, qi.data(), op.l.stream), "ofhe_ntt_fwd");
    Host partP(paramsP, Format::EVALUATION, true);
    op.Down(p, partP.GetAllElements(), 0, sizeP);
    if (!wasEval)
        op.Down(q, towers, 0, sizeQ);
    op.Sync();
    // assemble Q u P in the mirror: the reference moves partP's towers behind its own (:952-955)
    std::vector<PolyType> all;
    all.reserve(sizeQ + sizeP);
    for (uint32_t i = 0; i < sizeQ; ++i) {
        towers[i].OverrideFormat(Format::EVALUATION);
        all.push_back(std::move(towers[i]));
    }
    for (auto& t : partP.GetAllElements())
        all.push_back(std::move(t));
    Host qp(paramsQP, Format::EVALUATION, false);
    qp.GetAllElements() = std::move(all);
    m_h                 = std::move(qp);
    op.rt.n.modup += 1;
    op.rt.n.ntt += sizeP + sizeQ;
}

// ---- a11  ApproxModDown (:966-1005), CKKS/BFV (t == 0) and BGV (t > 0) branches
B200T B200C B200C::ApproxModDown(const std::shared_ptr<Params>& paramsQ, const std::shared_ptr<Params>& paramsP,
                                 const std::vector<NativeInteger>& PInvModq, const std::vector<NativeInteger>&,
                                 const std::vector<NativeInteger>& PHatInvModp, const std::vector<NativeInteger>&,
                                 const std::vector<std::vector<NativeInteger>>& PHatModq,
                                 const std::vector<DoubleNativeInt>&, const std::vector<NativeInteger>& tInvModp,
                                 const std::vector<NativeInteger>&, const NativeInteger& t,
                                 const std::vector<NativeInteger>&) const {
    const auto& towers   = m_h.GetAllElements();
    const uint32_t sizeP = (uint32_t)paramsP->GetParams().size();
    const uint32_t sizeQ = (uint32_t)towers.size() - sizeP;
    // ring: P towers first (the conversion's source), then the Q towers
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, sizeQ, sizeP, mod, root);
    detail::Collect(towers, 0, sizeQ, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, sizeP, sizeQ, PHatInvModp, PHatModq, false, nullptr, nullptr);
    uint64_t* p     = op.Slab(0, sizeP);
    uint64_t* q     = op.Slab(1, sizeQ);
    uint64_t* c     = op.Slab(2, sizeQ);
    const auto pi = detail::Iota(0, sizeP), qi = detail::Iota(sizeP, sizeQ);
    op.Up(p, towers, sizeQ, sizeP);
    op.Up(q, towers, 0, sizeQ);
    if (m_h.GetFormat() == Format::EVALUATION)
        Check(ofhe_ntt_inv(ring, p, p, sizeP, pi.data(), op.l.stream), "ofhe_ntt_inv");
    const bool bgv = t > NativeInteger(0);
    if (bgv) {
        const auto w = detail::Words(tInvModp, sizeP);
        Check(ofhe_ew_mul_scalars(ring, p, p, w.data(), sizeP, pi.data(), op.l.stream), "ofhe_ew_mul_scalars");
    }
    Check(ofhe_bconv_apply(tbl, p, c, op.l.stream), "ofhe_bconv_apply");
    if (bgv) {
        const std::vector<uint64_t> tt(sizeQ, t.ConvertToInt());
        Check(ofhe_ew_mul_scalars(ring, c, c, tt.data(), sizeQ, qi.data(), op.l.stream), "ofhe_ew_mul_scalars");
    }
    Check(ofhe_ntt_fwd(ring, c, c, sizeQ, qi.data(), op.l.stream), "ofhe_ntt_fwd");
    Check(ofhe_ew_binop(ring, 1, q, q, c, sizeQ, qi.data(), op.l.stream), "ofhe_ew_binop");
    const auto s = detail::Words(PInvModq, sizeQ);
    Check(ofhe_ew_mul_scalars(ring, q, q, s.data(), sizeQ, qi.data(), op.l.stream), "ofhe_ew_mul_scalars");
    Host ans(paramsQ, Format::EVALUATION, true);
    const uint32_t diffQ = (uint32_t)paramsQ->GetParams().size() - sizeQ;
    if (diffQ > 0)
        ans.DropLastElements(diffQ);
    op.Down(q, ans.GetAllElements(), 0, sizeQ);
    op.Sync();
    op.rt.n.moddown += 1;
    op.rt.n.ntt += sizeP + sizeQ;
    return ans;
}

// ---- a12  exact SwitchCRTBasis (:1008-1085), ExpandCRTBasis (:1088-1116), FastExpandCRTBasisPloverQ (:1151-1164)
B200T B200C B200C::SwitchCRTBasis(const std::shared_ptr<Params>& paramsP, const std::vector<NativeInteger>& QHatInvModq,
                                  const std::vector<NativeInteger>&,
                                  const std::vector<std::vector<NativeInteger>>& QHatModp,
                                  const std::vector<std::vector<NativeInteger>>& alphaQModp,
                                  const std::vector<DoubleNativeInt>&, const std::vector<double>& qInv) const {
    Host ans(paramsP, m_h.GetFormat(), true);
    const auto& towers   = m_h.GetAllElements();
    const uint32_t sizeQ = (uint32_t)towers.size();
    const uint32_t sizeP = (uint32_t)ans.GetAllElements().size();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, 0, sizeQ, mod, root);
    detail::Collect(ans.GetAllElements(), 0, sizeP, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, sizeQ, sizeP, QHatInvModq, QHatModp, true, &alphaQModp, &qInv);
    uint64_t* in    = op.Slab(0, sizeQ);
    uint64_t* out   = op.Slab(1, sizeP);
    op.Up(in, towers, 0, sizeQ);
    Check(ofhe_bconv_apply(tbl, in, out, op.l.stream), "ofhe_bconv_apply");
    op.Down(out, ans.GetAllElements(), 0, sizeP);
    op.Sync();
    op.rt.n.bconv += 1;
    return ans;
}

B200T void B200C::ExpandCRTBasis(const std::shared_ptr<Params>& paramsQP, const std::shared_ptr<Params>& paramsP,
                                 const std::vector<NativeInteger>& QHatInvModq,
                                 const std::vector<NativeInteger>& QHatInvModqPrecon,
                                 const std::vector<std::vector<NativeInteger>>& QHatModp,
                                 const std::vector<std::vector<NativeInteger>>& alphaQModp,
                                 const std::vector<DoubleNativeInt>& modpBarrettMu, const std::vector<double>& qInv,
                                 Format resultFormat) {
    // composition of device methods in the reference's order (:1094-1115)
    std::vector<PolyType> polyInNTT;
    if (m_h.GetFormat() == Format::EVALUATION) {
        polyInNTT = m_h.GetAllElements();
        SwitchFormat();
    }
    DCRTPolyType partP = SwitchCRTBasis(paramsP, QHatInvModq, QHatInvModqPrecon, QHatModp, alphaQModp, modpBarrettMu, qInv);
    if (resultFormat == Format::EVALUATION && !polyInNTT.empty()) {
        m_h.GetAllElements() = std::move(polyInNTT);
        m_h.OverrideFormat(Format::EVALUATION);
    }
    // bring both parts to the requested format on the device, then splice the towers
    if (m_h.GetFormat() != resultFormat)
        SwitchFormat();
    if (partP.GetFormat() != resultFormat)
        partP.SwitchFormat();
    std::vector<PolyType> all = std::move(m_h.GetAllElements());
    for (auto& t : partP.GetAllElements())
        all.push_back(std::move(t));
    Host qp(paramsQP, resultFormat, false);
    qp.GetAllElements() = std::move(all);
    m_h                 = std::move(qp);
}

B200T void B200C::FastExpandCRTBasisPloverQ(const Precomputations& pc) {
    DCRTPolyType partPl = ApproxSwitchCRTBasis(m_h.GetParams(), pc.paramsPl, pc.mPlQHatInvModq, pc.mPlQHatInvModqPrecon,
                                               pc.qInvModp, pc.modpBarrettMu);
    DCRTPolyType partQl = partPl.SwitchCRTBasis(pc.paramsQl, pc.PlHatInvModp, pc.PlHatInvModpPrecon, pc.PlHatModq,
                                                pc.alphaPlModq, pc.modqBarrettMu, pc.pInv);
    std::vector<PolyType> all = std::move(partQl.GetAllElements());
    for (auto& t : partPl.GetAllElements())
        all.push_back(std::move(t));
    const Format f = m_h.GetFormat();
    Host qp(pc.paramsQlPl, f, false);
    qp.GetAllElements() = std::move(all);
    m_h                 = std::move(qp);
}

// ---- a12 remainder  ExpandCRTBasisQlHat (:1167-1187): towers times QlHatModq, then zero towers up to sizeQ
B200T void B200C::ExpandCRTBasisQlHat(const std::shared_ptr<Params>& paramsQ, const std::vector<NativeInteger>& QlHatModq,
                                      const std::vector<NativeInteger>&, const usint sizeQ) {
    const uint32_t sizeQl = (uint32_t)m_h.GetAllElements().size();
    Host scaled           = detail::TimesScalars(m_h, QlHatModq, sizeQl);
    const Format f        = m_h.GetFormat();
    std::vector<PolyType> all = std::move(scaled.GetAllElements());
    for (uint32_t i = sizeQl; i < sizeQ; ++i)
        all.emplace_back(paramsQ->GetParams()[i], f, true);
    Host q(paramsQ, f, false);
    q.GetAllElements() = std::move(all);
    m_h                = std::move(q);
}

// ---- a13  ScaleAndRound DCRT -> DCRT (:1513-1628), ApproxScaleAndRound (:1470-1510), ScaleAndRoundPOverQ (:1674-1689)
B200T B200C B200C::ScaleAndRound(const std::shared_ptr<Params>& paramsOutput,
                                 const std::vector<std::vector<NativeInteger>>& tOSHatInvModsDivsModo,
                                 const std::vector<double>& tOSHatInvModsDivsFrac,
                                 const std::vector<DoubleNativeInt>&) const {
    Host ans(paramsOutput, m_h.GetFormat(), true);
    const auto& towers    = m_h.GetAllElements();
    const uint32_t sizeQP = (uint32_t)towers.size();
    const uint32_t sizeO  = (uint32_t)ans.GetAllElements().size();
    const uint32_t sizeI  = sizeQP - sizeO;
    const bool outFirst =
        paramsOutput->GetParams()[0]->GetModulus() == m_h.GetParams()->GetParams()[0]->GetModulus();  // (:1528-1535)
    if (sizeI == 0) {
        // nothing is scaled away (the levelled BFV code calls this unconditionally, bfvrns-leveledshe.cpp:807): the
        // reference's loop then reduces to x_j * tOSHatInvModsDivsModo[j][0] mod o_j with alpha = (uint64)0.5 = 0
        std::vector<NativeInteger> s(sizeO);
        for (uint32_t j = 0; j < sizeO; ++j)
            s[j] = tOSHatInvModsDivsModo[j][0];
        Host r = detail::TimesScalars(m_h, s, sizeO);
        std::vector<PolyType> tw = std::move(r.GetAllElements());
        for (uint32_t j = 0; j < sizeO; ++j)
            ans.GetAllElements()[j].SetValues(tw[j].GetValues(), m_h.GetFormat());
        return ans;
    }
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, 0, sizeQP, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    const auto inIdx  = detail::Iota(outFirst ? sizeO : 0, sizeI);
    const auto outIdx = detail::Iota(outFirst ? 0 : sizeI, sizeO);
    std::vector<uint64_t> tbl((size_t)sizeO * (sizeI + 1));
    for (uint32_t j = 0; j < sizeO; ++j)
        for (uint32_t i = 0; i <= sizeI; ++i)
            tbl[(size_t)j * (sizeI + 1) + i] = tOSHatInvModsDivsModo[j][i].ConvertToInt();
    std::string key("sr");
    detail::AppendKey(key, &ring, sizeof ring);
    detail::AppendKey(key, &sizeI, 4);
    detail::AppendKey(key, &sizeO, 4);
    detail::AppendKey(key, &outFirst, sizeof outFirst);
    detail::AppendKey(key, tbl.data(), tbl.size() * 8);
    detail::AppendKey(key, tOSHatInvModsDivsFrac.data(), sizeI * sizeof(double));
    ofhe_bconv* t = op.rt.BConv(key, [&]() {
        ofhe_bconv* b = nullptr;
        Check(ofhe_scale_round_create(ring, sizeI, inIdx.data(), sizeO, outIdx.data(), outFirst ? 1 : 0, tbl.data(),
                                      tOSHatInvModsDivsFrac.data(), &b),
              "ofhe_scale_round_create");
        return b;
    });
    uint64_t* in  = op.Slab(0, sizeQP);
    uint64_t* out = op.Slab(1, sizeO);
    op.Up(in, towers, 0, sizeQP);
    Check(ofhe_bconv_apply(t, in, out, op.l.stream), "ofhe_bconv_apply");
    op.Down(out, ans.GetAllElements(), 0, sizeO);
    op.Sync();
    op.rt.n.bconv += 1;
    return ans;
}

B200T B200C B200C::ApproxScaleAndRound(const std::shared_ptr<Params>& paramsP,
                                       const std::vector<std::vector<NativeInteger>>& tPSHatInvModsDivsModp,
                                       const std::vector<DoubleNativeInt>&) const {
    Host ans(paramsP, m_h.GetFormat(), true);
    const auto& towers    = m_h.GetAllElements();
    const uint32_t sizeQP = (uint32_t)towers.size();
    const uint32_t sizeP  = (uint32_t)ans.GetAllElements().size();
    const uint32_t sizeQ  = sizeQP - sizeP;
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, 0, sizeQP, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    std::vector<uint64_t> tbl((size_t)sizeP * (sizeQ + 1));
    for (uint32_t j = 0; j < sizeP; ++j)
        for (uint32_t i = 0; i <= sizeQ; ++i)
            tbl[(size_t)j * (sizeQ + 1) + i] = tPSHatInvModsDivsModp[j][i].ConvertToInt();
    const auto inIdx = detail::Iota(0, sizeQ), outIdx = detail::Iota(sizeQ, sizeP);
    std::string key("asr");
    detail::AppendKey(key, &ring, sizeof ring);
    detail::AppendKey(key, &sizeQ, 4);
    detail::AppendKey(key, &sizeP, 4);
    detail::AppendKey(key, tbl.data(), tbl.size() * 8);
    // the approximate variant has no rounding term (:1484-1503): zero fractions give alpha = (uint64)0.5 = 0
    const std::vector<double> zeros(sizeQ, 0.0);
    ofhe_bconv* t = op.rt.BConv(key, [&]() {
        ofhe_bconv* b = nullptr;
        Check(ofhe_scale_round_create(ring, sizeQ, inIdx.data(), sizeP, outIdx.data(), 0, tbl.data(), zeros.data(), &b),
              "ofhe_scale_round_create");
        return b;
    });
    uint64_t* in  = op.Slab(0, sizeQP);
    uint64_t* out = op.Slab(1, sizeP);
    op.Up(in, towers, 0, sizeQP);
    Check(ofhe_bconv_apply(t, in, out, op.l.stream), "ofhe_bconv_apply");
    op.Down(out, ans.GetAllElements(), 0, sizeP);
    op.Sync();
    op.rt.n.bconv += 1;
    return ans;
}

B200T void B200C::ScaleAndRoundPOverQ(const std::shared_ptr<Params>& paramsQ, const std::vector<NativeInteger>& pInvModq) {
    // x_i <- (x_i - SwitchModulus(x_last, p -> q_i)) * pInvModq[i], last tower dropped; no format change (:1674-1689)
    const uint32_t L = (uint32_t)m_h.GetAllElements().size();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(m_h.GetAllElements(), 0, L, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    const auto& towers = m_h.GetAllElements();
    uint64_t* x  = op.Slab(0, L - 1);
    uint64_t* z  = op.Slab(1, 1);
    uint64_t* sw = op.Slab(2, L - 1);
    op.Up(x, towers, 0, L - 1);
    op.Up1(z, towers[L - 1]);
    for (uint32_t i = 0; i + 1 < L; ++i)
        Check(ofhe_switch_modulus(ring, sw + (size_t)i * op.N, z, L - 1, i, op.l.stream), "ofhe_switch_modulus");
    const auto idx = detail::Iota(0, L - 1);
    Check(ofhe_ew_binop(ring, 1, x, x, sw, L - 1, idx.data(), op.l.stream), "ofhe_ew_binop");
    const auto s = detail::Words(pInvModq, L - 1);
    Check(ofhe_ew_mul_scalars(ring, x, x, s.data(), L - 1, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    const Format f = m_h.GetFormat();
    std::vector<PolyType> keep(std::make_move_iterator(m_h.GetAllElements().begin()),
                               std::make_move_iterator(m_h.GetAllElements().end() - 1));
    Host q(paramsQ, f, false);
    q.GetAllElements() = std::move(keep);
    m_h                = std::move(q);
    op.Down(x, m_h.GetAllElements(), 0, L - 1);
    op.Sync();
    op.rt.n.ew += L - 1;
}

#undef B200T
#undef B200C

}  // namespace b200
}  // namespace lbcrypto

#endif
