// dcrtpoly-b200-impl.h -- the device methods of lbcrypto::b200::DCRTPolyB200 (see dcrtpoly-b200.h).
//
// Every method states the reference function whose results it reproduces bit for bit
// (src/core/include/lattice/hal/default/dcrtpoly-impl.h unless another file is named) and is a composition of calls
// into libofhe_b200.so on the calling thread's stream.  Operands are read from their device slabs (a host-only operand is
// uploaded once and stays resident), results are NEW device slabs published to the result polynomial -- slabs are
// immutable once published, so copies share them -- and the stream is synchronised before the method returns.  No
// arithmetic on residues happens on the host here.
#ifndef LBCRYPTO_INC_LATTICE_HAL_B200_DCRTPOLY_B200_IMPL_H
#define LBCRYPTO_INC_LATTICE_HAL_B200_DCRTPOLY_B200_IMPL_H

#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "lattice/hal/b200/dcrtpoly-b200.h"

namespace lbcrypto {
namespace b200 {
namespace detail {

using Tower = PolyImpl<NativeVector>;
using Slab  = std::shared_ptr<uint64_t>;

// one device operation of the calling thread
struct Op {
    Runtime& rt;
    Runtime::Local& l;
    uint32_t N;
    explicit Op(uint32_t n) : rt(Runtime::Get()), l(rt.Mine()), N(n) {}
    uint64_t* Scratch(int which, size_t limbs) {  // per-thread intermediates (never published)
        return rt.Slab(l, which, std::max<size_t>(1, limbs) * N);
    }
    Slab New(size_t limbs) {  // a result slab
        return rt.Alloc(std::max<size_t>(1, limbs) * N);
    }
    void Up(uint64_t* d, const std::vector<Tower>& t, size_t first, size_t count) {
        for (size_t i = 0; i < count; ++i)
            Check(ofhe_memcpy_h2d(rt.ctx, d + i * N, &t[first + i].GetValues()[0], (size_t)N * 8, l.stream), "h2d");
        rt.n.h2d_bytes += count * (size_t)N * 8;
    }
    void Down(const uint64_t* d, std::vector<Tower>& t, size_t first, size_t count) {
        for (size_t i = 0; i < count; ++i) {
            if (t[first + i].IsEmpty())
                t[first + i].SetValuesToZero();  // allocate the tower's N words
            Check(ofhe_memcpy_d2h(rt.ctx, &t[first + i][0], d + i * N, (size_t)N * 8, l.stream), "d2h");
        }
        rt.n.d2h_bytes += count * (size_t)N * 8;
    }
    void Copy(uint64_t* dst, const uint64_t* src, size_t limbs) {
        if (limbs)
            Check(ofhe_memcpy_d2d(rt.ctx, dst, src, limbs * (size_t)N * 8, l.stream), "d2d");
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
inline ofhe_ring* RingOfTowers(Op& op, const std::vector<Tower>& towers, size_t first, size_t count) {
    std::vector<uint64_t> mod, root;
    Collect(towers, first, count, mod, root);
    return op.rt.Ring(op.N, mod, root);
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
inline void SetFormats(std::vector<Tower>& t, Format f) {
    for (auto& x : t)
        x.OverrideFormat(f);
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

// ---- residency ---------------------------------------------------------------------------------------------------
B200T void B200C::EnsureHost() const {
    std::lock_guard<std::mutex> g(Runtime::Get().Stripe(this));
    if (m_w.load(std::memory_order_acquire) & kHost)
        return;
    auto& towers = m_h.GetAllElements();
    if (!towers.empty()) {
        detail::Op op(m_h.GetParams()->GetRingDimension());
        op.Down(m_d.get(), towers, 0, towers.size());
        op.Sync();
    }
    m_w.store(m_w.load() | kHost, std::memory_order_release);
}
B200T std::shared_ptr<uint64_t> B200C::DeviceSlab() const {
    if (!(m_w.load(std::memory_order_acquire) & kDev)) {
        std::lock_guard<std::mutex> g(Runtime::Get().Stripe(this));
        if (!(m_w.load(std::memory_order_acquire) & kDev)) {
            const auto& towers = m_h.GetAllElements();
            detail::Op op(m_h.GetParams()->GetRingDimension());
            auto slab = op.New(towers.size());
            op.Up(slab.get(), towers, 0, towers.size());
            op.Sync();
            m_d = std::move(slab);
            m_w.store(m_w.load() | kDev, std::memory_order_release);
        }
    }
    return m_d;
}
B200T B200C B200C::FromDevice(const std::shared_ptr<Params>& params, Format format, std::shared_ptr<uint64_t> slab) {
    DCRTPolyType r;
    r.Adopt(Meta(params, format), std::move(slab));
    return r;
}
B200T void B200C::CopyFrom(const DCRTPolyType& e) {
    std::lock_guard<std::mutex> g(Runtime::Get().Stripe(&e));  // e may be downloading in another thread
    const uint8_t w = e.m_w.load(std::memory_order_acquire);
    if (w & kHost)
        m_h = e.m_h;
    else {  // device only: parameters, formats and empty towers
        Host meta(e.m_h.GetParams(), e.m_h.GetFormat(), false);
        auto& t        = meta.GetAllElements();
        const auto& et = e.m_h.GetAllElements();
        t.resize(et.size(), detail::Tower());
        for (size_t i = 0; i < et.size(); ++i)
            t[i] = detail::Tower(et[i].GetParams(), et[i].GetFormat(), false);
        m_h = std::move(meta);
    }
    m_d = e.m_d;  // slabs are immutable: shared
    m_w.store(w, std::memory_order_release);
}

// ---- a6-a8  DCRTPolyImpl::SwitchFormat (:1932-1940): every tower through the forward / inverse negacyclic NTT
B200T void B200C::SwitchFormat(uint32_t) {
    const uint32_t L  = GetNumOfElements();
    const Format from = m_h.GetFormat(), to = (from == Format::EVALUATION) ? Format::COEFFICIENT : Format::EVALUATION;
    if (L == 0) {
        m_h.OverrideFormat(to);
        return;
    }
    if (m_h.GetParams()->GetCyclotomicOrder() != 2 * m_h.GetParams()->GetRingDimension()) {
        // arbitrary (non power-of-two) cyclotomics are outside this backend's scope (the RNS
        // schemes only use power-of-two rings); the mirror's Bluestein transform serves them as in the reference
        HW().SwitchFormat();
        return;
    }
    detail::Op op(m_h.GetParams()->GetRingDimension());
    ofhe_ring* ring = detail::RingOfTowers(op, m_h.GetAllElements(), 0, L);
    const auto in   = DeviceSlab();
    auto out        = op.New(L);
    const auto idx  = detail::Iota(0, L);
    if (to == Format::EVALUATION)
        Check(ofhe_ntt_fwd(ring, in.get(), out.get(), L, idx.data(), op.l.stream), "ofhe_ntt_fwd");
    else
        Check(ofhe_ntt_inv(ring, in.get(), out.get(), L, idx.data(), op.l.stream), "ofhe_ntt_inv");
    op.Sync();
    m_d = std::move(out);
    m_w.store(kDev, std::memory_order_release);
    m_h.OverrideFormat(to);
    detail::SetFormats(m_h.GetAllElements(), to);
    op.rt.n.ntt += L;
}

// ---- a9  Plus / Minus / Times / Negate and the Eq forms (dcrtpoly.h:130-190, dcrtpoly-impl.h:347-621):
// op 0 add, 1 sub, 2 mul, 3 negate; same argument checks and exceptions as the reference
namespace detail {
template <typename P>
inline void CheckPair(const P& a, const P& b, bool mul) {
    if (a.GetParams()->GetRingDimension() != b.GetParams()->GetRingDimension())
        OPENFHE_THROW("RingDimension mismatch");
    if (mul) {
        if (a.GetFormat() != Format::EVALUATION || b.GetFormat() != Format::EVALUATION)
            OPENFHE_THROW("operator* for DCRTPolyImpl supported only in Format::EVALUATION");
    }
    else if (a.GetFormat() != b.GetFormat())
        OPENFHE_THROW("Format mismatch");
    if (a.GetNumOfElements() != b.GetNumOfElements())
        OPENFHE_THROW(mul ? "tower size mismatch; cannot multiply" : "tower size mismatch; cannot add");
    if (a.GetNumOfElements() && a.TowerMeta()[0].GetModulus() != b.TowerMeta()[0].GetModulus())
        OPENFHE_THROW("Modulus mismatch");
}
// a (op) b over a's towers (b may have more: the reference's Eq forms take its first towers) -> new slab
template <typename P>
inline Slab Binop(int opc, const P& a, const P* b) {
    const uint32_t L = a.GetNumOfElements();
    Op op(a.GetParams()->GetRingDimension());
    auto out = op.New(L);
    if (L == 0)
        return out;
    if (b && b->GetNumOfElements() < L)
        OPENFHE_THROW("b200 backend: right-hand side has fewer towers than the left-hand side");
    ofhe_ring* ring = RingOfTowers(op, a.TowerMeta(), 0, L);
    const auto da   = a.DeviceSlab();
    const auto db   = b ? b->DeviceSlab() : da;
    const auto idx  = Iota(0, L);
    Check(ofhe_ew_binop(ring, opc, out.get(), da.get(), db.get(), L, idx.data(), op.l.stream), "ofhe_ew_binop");
    op.Sync();
    op.rt.n.ew += L;
    return out;
}
}  // namespace detail

// The Eq forms perform NO checks in the reference (:374-380, :402-408, dcrtpoly.h:117-123): they combine this
// polynomial's towers with the FIRST as many towers of rhs, and callers rely on that (rhs with more towers:
// FHECKKSRNS::KeySwitchSparse, ckksrns-fhe.cpp:3650).  A shorter rhs is undefined behaviour there and an error here.
B200T B200C& B200C::operator+=(const DCRTPolyType& rhs) {
    m_d = detail::Binop(0, *this, &rhs);
    m_w.store(kDev, std::memory_order_release);
    return *this;
}
B200T B200C& B200C::operator-=(const DCRTPolyType& rhs) {
    m_d = detail::Binop(1, *this, &rhs);
    m_w.store(kDev, std::memory_order_release);
    return *this;
}
B200T B200C& B200C::operator*=(const DCRTPolyType& rhs) {
    m_d = detail::Binop(2, *this, &rhs);
    m_w.store(kDev, std::memory_order_release);
    return *this;
}
B200T B200C B200C::Plus(const DCRTPolyType& rhs) const {
    detail::CheckPair(*this, rhs, false);
    return FromDevice(m_h.GetParams(), m_h.GetFormat(), detail::Binop(0, *this, &rhs));
}
B200T B200C B200C::Minus(const DCRTPolyType& rhs) const {
    detail::CheckPair(*this, rhs, false);
    return FromDevice(m_h.GetParams(), m_h.GetFormat(), detail::Binop(1, *this, &rhs));
}
B200T B200C B200C::Times(const DCRTPolyType& rhs) const {
    detail::CheckPair(*this, rhs, true);
    return FromDevice(m_h.GetParams(), m_h.GetFormat(), detail::Binop(2, *this, &rhs));
}
B200T B200C B200C::Negate() const {
    return FromDevice(m_h.GetParams(), m_h.GetFormat(), detail::Binop<DCRTPolyType>(3, *this, nullptr));
}
B200T B200C B200C::operator-() const {
    return Negate();
}

// ---- a9  Times(const std::vector<NativeInteger>&) (:553-567) and TimesNoCheck (:569-577)
namespace detail {
// first L towers of a times s[i]; the result slab has `rows` rows (>= L; the others zero, as the reference's
// zero-initialised towers)
template <typename P>
inline Slab TimesScalars(const P& a, const std::vector<NativeInteger>& s, uint32_t L, uint32_t rows) {
    Op op(a.GetParams()->GetRingDimension());
    auto out = op.New(rows);
    if (rows > L)
        Check(ofhe_memset0(op.rt.ctx, out.get() + (size_t)L * op.N, (size_t)(rows - L) * op.N * 8, op.l.stream), "memset");
    if (L) {
        ofhe_ring* ring = RingOfTowers(op, a.TowerMeta(), 0, L);
        const auto da   = a.DeviceSlab();
        const auto idx  = Iota(0, L);
        const auto w    = Words(s, L);
        Check(ofhe_ew_mul_scalars(ring, out.get(), da.get(), w.data(), L, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    }
    op.Sync();
    op.rt.n.ew += L;
    return out;
}
}  // namespace detail
B200T B200C B200C::Times(const std::vector<NativeInteger>& rhs) const {
    const uint32_t L = GetNumOfElements();
    if (L != rhs.size())
        OPENFHE_THROW("tower size mismatch; cannot multiply");
    return FromDevice(m_h.GetParams(), m_h.GetFormat(), detail::TimesScalars(*this, rhs, L, L));
}
B200T B200C B200C::TimesNoCheck(const std::vector<NativeInteger>& rhs) const {
    const uint32_t T = GetNumOfElements(), L = (uint32_t)std::min<size_t>(T, rhs.size());
    return FromDevice(m_h.GetParams(), m_h.GetFormat(), detail::TimesScalars(*this, rhs, L, T));
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
    const uint32_t L = GetNumOfElements();
    DCRTPolyType r(*this);  // parameters, formats (shares the slab until the new one is published)
    if (L == 0)
        return r;
    detail::Op op(m_h.GetParams()->GetRingDimension());
    ofhe_ring* ring = detail::RingOfTowers(op, m_h.GetAllElements(), 0, L);
    const auto src  = DeviceSlab();
    auto dst        = op.New(L);
    if (m_h.GetFormat() == Format::EVALUATION)
        Check(ofhe_automorphism(ring, dst.get(), src.get(), L, k, op.l.stream), "ofhe_automorphism");
    else {
        const auto idx = detail::Iota(0, L);
        Check(ofhe_automorphism_coeff(ring, dst.get(), src.get(), L, idx.data(), k, op.l.stream), "ofhe_automorphism_coeff");
    }
    op.Sync();
    r.m_d = std::move(dst);
    r.m_w.store(kDev, std::memory_order_release);
    op.rt.n.automorph += L;
    return r;
}

// ---- a14  DropLastElementAndScale (:693-712) and f4  ModReduce (:736-754): both are
//   z = INTT(last tower) [* c];  x_i <- f(x_i, NTT(SwitchModulus(z, q_l -> q_i)))  for the remaining towers
namespace detail {
// scratch 2 = [L-1][N]: the dropped tower (COEFFICIENT, times `pre` if not null) switched into every remaining modulus
// (mubintvecnat.cpp:109-122), times post[i] if given, in the polynomial's format.  x = the polynomial's slab [L][N].
inline uint64_t* DroppedTowerEverywhere(Op& op, ofhe_ring* ring, Format f, const uint64_t* x, uint32_t L,
                                        const uint64_t* pre, const std::vector<uint64_t>* post) {
    uint64_t* z         = op.Scratch(1, 1);
    uint64_t* sw        = op.Scratch(2, L - 1);
    const uint32_t last = L - 1;
    if (f == Format::EVALUATION)
        Check(ofhe_ntt_inv(ring, x + (size_t)last * op.N, z, 1, &last, op.l.stream), "ofhe_ntt_inv");
    else
        op.Copy(z, x + (size_t)last * op.N, 1);
    if (pre)
        Check(ofhe_ew_mul_scalars(ring, z, z, pre, 1, &last, op.l.stream), "ofhe_ew_mul_scalars");
    for (uint32_t i = 0; i < last; ++i)
        Check(ofhe_switch_modulus(ring, sw + (size_t)i * op.N, z, last, i, op.l.stream), "ofhe_switch_modulus");
    const auto idx = Iota(0, last);
    if (post)
        Check(ofhe_ew_mul_scalars(ring, sw, sw, post->data(), last, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    if (f == Format::EVALUATION)
        Check(ofhe_ntt_fwd(ring, sw, sw, last, idx.data(), op.l.stream), "ofhe_ntt_fwd");
    op.rt.n.ntt += L;
    return sw;
}
}  // namespace detail

B200T void B200C::DropLastElementAndScale(const std::vector<NativeInteger>& QlQlInvModqlDivqlModq,
                                          const std::vector<NativeInteger>& qlInvModq) {
    const uint32_t L = GetNumOfElements();
    if (L < 2)
        OPENFHE_THROW("DropLastElementAndScale needs at least two towers");
    detail::Op op(m_h.GetParams()->GetRingDimension());
    ofhe_ring* ring = detail::RingOfTowers(op, m_h.GetAllElements(), 0, L);
    const auto x    = DeviceSlab();
    const Format f  = m_h.GetFormat();
    const auto post = detail::Words(QlQlInvModqlDivqlModq, L - 1);
    uint64_t* t     = detail::DroppedTowerEverywhere(op, ring, f, x.get(), L, nullptr, &post);
    auto out        = op.New(L - 1);
    const auto idx  = detail::Iota(0, L - 1);
    const auto s    = detail::Words(qlInvModq, L - 1);
    Check(ofhe_ew_mul_scalars(ring, out.get(), x.get(), s.data(), L - 1, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    Check(ofhe_ew_binop(ring, 0, out.get(), out.get(), t, L - 1, idx.data(), op.l.stream), "ofhe_ew_binop");
    if (f == Format::COEFFICIENT)  // the reference leaves the TOWERS in EVALUATION in this case (:708-709)
        Check(ofhe_ntt_fwd(ring, out.get(), out.get(), L - 1, idx.data(), op.l.stream), "ofhe_ntt_fwd");
    op.Sync();
    m_h.DropLastElement();
    m_d = std::move(out);
    m_w.store(kDev, std::memory_order_release);
    if (f == Format::COEFFICIENT)
        detail::SetFormats(m_h.GetAllElements(), Format::EVALUATION);
    op.rt.n.rescale += 1;
}

B200T void B200C::ModReduce(const NativeInteger& t, const std::vector<NativeInteger>&, const NativeInteger& negtInvModq,
                            const NativeInteger&, const std::vector<NativeInteger>& qlInvModq,
                            const std::vector<NativeInteger>&) {
    const uint32_t L = GetNumOfElements();
    if (L < 2)
        OPENFHE_THROW("ModReduce needs at least two towers");
    detail::Op op(m_h.GetParams()->GetRingDimension());
    ofhe_ring* ring    = detail::RingOfTowers(op, m_h.GetAllElements(), 0, L);
    const auto x       = DeviceSlab();
    const uint64_t pre = negtInvModq.ConvertToInt();
    uint64_t* d        = detail::DroppedTowerEverywhere(op, ring, m_h.GetFormat(), x.get(), L, &pre, nullptr);
    auto out           = op.New(L - 1);
    const auto idx     = detail::Iota(0, L - 1);
    const std::vector<uint64_t> tt(L - 1, t.ConvertToInt());
    Check(ofhe_ew_mul_scalars(ring, d, d, tt.data(), L - 1, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    Check(ofhe_ew_binop(ring, 0, out.get(), x.get(), d, L - 1, idx.data(), op.l.stream), "ofhe_ew_binop");
    const auto s = detail::Words(qlInvModq, L - 1);
    Check(ofhe_ew_mul_scalars(ring, out.get(), out.get(), s.data(), L - 1, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    op.Sync();
    m_h.DropLastElement();
    m_d = std::move(out);
    m_w.store(kDev, std::memory_order_release);
    op.rt.n.rescale += 1;
}

// ---- a10  ApproxSwitchCRTBasis (:888-932): [sizeQ x N] residues -> [sizeP x N], format label carried over
B200T B200C B200C::ApproxSwitchCRTBasis(const std::shared_ptr<Params>& paramsQ, const std::shared_ptr<Params>& paramsP,
                                        const std::vector<NativeInteger>& QHatInvModq, const std::vector<NativeInteger>&,
                                        const std::vector<std::vector<NativeInteger>>& QHatModp,
                                        const std::vector<DoubleNativeInt>&) const {
    const uint32_t sizeQ = (uint32_t)std::min<size_t>(GetNumOfElements(), paramsQ->GetParams().size());
    const uint32_t sizeP = (uint32_t)paramsP->GetParams().size();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(m_h.GetAllElements(), 0, sizeQ, mod, root);
    detail::CollectParams(paramsP, sizeP, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, sizeQ, sizeP, QHatInvModq, QHatModp, false, nullptr, nullptr);
    const auto in   = DeviceSlab();
    auto out        = op.New(sizeP);
    Check(ofhe_bconv_apply(tbl, in.get(), out.get(), op.l.stream), "ofhe_bconv_apply");
    op.Sync();
    op.rt.n.bconv += 1;
    return FromDevice(paramsP, m_h.GetFormat(), std::move(out));
}

// ---- a16 core  ApproxModUp (:935-963): Q -> Q u P, result in EVALUATION; the Q towers keep their EVALUATION values
// when the input was in EVALUATION (no NTT round trip), exactly as the reference
B200T void B200C::ApproxModUp(const std::shared_ptr<Params>& paramsQ, const std::shared_ptr<Params>& paramsP,
                              const std::shared_ptr<Params>& paramsQP, const std::vector<NativeInteger>& QHatInvModq,
                              const std::vector<NativeInteger>&, const std::vector<std::vector<NativeInteger>>& QHatModp,
                              const std::vector<DoubleNativeInt>&) {
    const uint32_t sizeQ = (uint32_t)std::min<size_t>(GetNumOfElements(), paramsQ->GetParams().size());
    const uint32_t sizeP = (uint32_t)paramsP->GetParams().size();
    const bool wasEval   = m_h.GetFormat() == Format::EVALUATION;
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(m_h.GetAllElements(), 0, sizeQ, mod, root);
    detail::CollectParams(paramsP, sizeP, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, sizeQ, sizeP, QHatInvModq, QHatModp, false, nullptr, nullptr);
    const auto q    = DeviceSlab();
    auto out        = op.New(sizeQ + sizeP);
    uint64_t* p     = out.get() + (size_t)sizeQ * op.N;
    const auto qi = detail::Iota(0, sizeQ), pi = detail::Iota(sizeQ, sizeP);
    if (wasEval) {
        uint64_t* c = op.Scratch(1, sizeQ);
        Check(ofhe_ntt_inv(ring, q.get(), c, sizeQ, qi.data(), op.l.stream), "ofhe_ntt_inv");
        Check(ofhe_bconv_apply(tbl, c, p, op.l.stream), "ofhe_bconv_apply");
        op.Copy(out.get(), q.get(), sizeQ);
    }
    else {
        Check(ofhe_bconv_apply(tbl, q.get(), p, op.l.stream), "ofhe_bconv_apply");
        Check(ofhe_ntt_fwd(ring, q.get(), out.get(), sizeQ, qi.data(), op.l.stream), "ofhe_ntt_fwd");
    }
    Check(ofhe_ntt_fwd(ring, p, p, sizeP, pi.data(), op.l.stream), "ofhe_ntt_fwd");
    op.Sync();
    Adopt(Meta(paramsQP, Format::EVALUATION), std::move(out));
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
    const uint32_t sizeQ = GetNumOfElements() - sizeP;
    // ring: P towers first (the conversion's source), then the Q towers
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, sizeQ, sizeP, mod, root);
    detail::Collect(towers, 0, sizeQ, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, sizeP, sizeQ, PHatInvModp, PHatModq, false, nullptr, nullptr);
    const auto x    = DeviceSlab();  // [Q | P]
    uint64_t* p     = op.Scratch(0, sizeP);
    uint64_t* c     = op.Scratch(2, sizeQ);
    auto out        = op.New(sizeQ);
    const auto pi = detail::Iota(0, sizeP), qi = detail::Iota(sizeP, sizeQ);
    if (m_h.GetFormat() == Format::EVALUATION)
        Check(ofhe_ntt_inv(ring, x.get() + (size_t)sizeQ * op.N, p, sizeP, pi.data(), op.l.stream), "ofhe_ntt_inv");
    else
        op.Copy(p, x.get() + (size_t)sizeQ * op.N, sizeP);
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
    Check(ofhe_ew_binop(ring, 1, out.get(), x.get(), c, sizeQ, qi.data(), op.l.stream), "ofhe_ew_binop");
    const auto s = detail::Words(PInvModq, sizeQ);
    Check(ofhe_ew_mul_scalars(ring, out.get(), out.get(), s.data(), sizeQ, qi.data(), op.l.stream), "ofhe_ew_mul_scalars");
    op.Sync();
    Host meta            = Meta(paramsQ, Format::EVALUATION);
    const uint32_t diffQ = (uint32_t)paramsQ->GetParams().size() - sizeQ;
    if (diffQ > 0)
        meta.DropLastElements(diffQ);
    DCRTPolyType ans;
    ans.Adopt(std::move(meta), std::move(out));
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
    const uint32_t sizeQ = GetNumOfElements();
    const uint32_t sizeP = (uint32_t)paramsP->GetParams().size();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::Collect(m_h.GetAllElements(), 0, sizeQ, mod, root);
    detail::CollectParams(paramsP, sizeP, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, sizeQ, sizeP, QHatInvModq, QHatModp, true, &alphaQModp, &qInv);
    const auto in   = DeviceSlab();
    auto out        = op.New(sizeP);
    Check(ofhe_bconv_apply(tbl, in.get(), out.get(), op.l.stream), "ofhe_bconv_apply");
    op.Sync();
    op.rt.n.bconv += 1;
    return FromDevice(paramsP, m_h.GetFormat(), std::move(out));
}

namespace detail {
// [a | b] as one slab
inline Slab Splice(Op& op, const Slab& a, uint32_t na, const Slab& b, uint32_t nb) {
    auto out = op.New(na + nb);
    op.Copy(out.get(), a.get(), na);
    op.Copy(out.get() + (size_t)na * op.N, b.get(), nb);
    op.Sync();
    return out;
}
}  // namespace detail

B200T void B200C::ExpandCRTBasis(const std::shared_ptr<Params>& paramsQP, const std::shared_ptr<Params>& paramsP,
                                 const std::vector<NativeInteger>& QHatInvModq,
                                 const std::vector<NativeInteger>& QHatInvModqPrecon,
                                 const std::vector<std::vector<NativeInteger>>& QHatModp,
                                 const std::vector<std::vector<NativeInteger>>& alphaQModp,
                                 const std::vector<DoubleNativeInt>& modpBarrettMu, const std::vector<double>& qInv,
                                 Format resultFormat) {
    // composition of device methods in the reference's order (:1094-1115)
    DCRTPolyType inNTT;
    const bool hadEval = m_h.GetFormat() == Format::EVALUATION;
    if (hadEval) {
        inNTT = *this;  // shares the slab
        SwitchFormat();
    }
    DCRTPolyType partP = SwitchCRTBasis(paramsP, QHatInvModq, QHatInvModqPrecon, QHatModp, alphaQModp, modpBarrettMu, qInv);
    if (resultFormat == Format::EVALUATION && hadEval)
        *this = std::move(inNTT);
    if (m_h.GetFormat() != resultFormat)
        SwitchFormat();
    if (partP.GetFormat() != resultFormat)
        partP.SwitchFormat();
    const uint32_t nq = GetNumOfElements(), np = partP.GetNumOfElements();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    auto out = detail::Splice(op, DeviceSlab(), nq, partP.DeviceSlab(), np);
    Adopt(Meta(paramsQP, resultFormat), std::move(out));
}

B200T void B200C::FastExpandCRTBasisPloverQ(const Precomputations& pc) {
    DCRTPolyType partPl = ApproxSwitchCRTBasis(m_h.GetParams(), pc.paramsPl, pc.mPlQHatInvModq, pc.mPlQHatInvModqPrecon,
                                               pc.qInvModp, pc.modpBarrettMu);
    DCRTPolyType partQl = partPl.SwitchCRTBasis(pc.paramsQl, pc.PlHatInvModp, pc.PlHatInvModpPrecon, pc.PlHatModq,
                                                pc.alphaPlModq, pc.modqBarrettMu, pc.pInv);
    const Format f = m_h.GetFormat();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    auto out = detail::Splice(op, partQl.DeviceSlab(), partQl.GetNumOfElements(), partPl.DeviceSlab(),
                              partPl.GetNumOfElements());
    Adopt(Meta(pc.paramsQlPl, f), std::move(out));
}

// ---- a12 remainder  ExpandCRTBasisQlHat (:1167-1187): towers times QlHatModq, then zero towers up to sizeQ
B200T void B200C::ExpandCRTBasisQlHat(const std::shared_ptr<Params>& paramsQ, const std::vector<NativeInteger>& QlHatModq,
                                      const std::vector<NativeInteger>&, const usint sizeQ) {
    const uint32_t sizeQl = GetNumOfElements();
    const Format f        = m_h.GetFormat();
    auto out              = detail::TimesScalars(*this, QlHatModq, sizeQl, sizeQ);
    Adopt(Meta(paramsQ, f), std::move(out));
}

// ---- a13  ScaleAndRound DCRT -> DCRT (:1513-1628), ApproxScaleAndRound (:1470-1510), ScaleAndRoundPOverQ (:1674-1689)
B200T B200C B200C::ScaleAndRound(const std::shared_ptr<Params>& paramsOutput,
                                 const std::vector<std::vector<NativeInteger>>& tOSHatInvModsDivsModo,
                                 const std::vector<double>& tOSHatInvModsDivsFrac,
                                 const std::vector<DoubleNativeInt>&) const {
    const uint32_t sizeQP = GetNumOfElements();
    const uint32_t sizeO  = (uint32_t)paramsOutput->GetParams().size();
    const uint32_t sizeI  = sizeQP - sizeO;
    const bool outFirst =
        paramsOutput->GetParams()[0]->GetModulus() == m_h.GetParams()->GetParams()[0]->GetModulus();  // (:1528-1535)
    if (sizeI == 0) {
        // nothing is scaled away (the levelled BFV code calls this unconditionally, bfvrns-leveledshe.cpp:807): the
        // reference's loop then reduces to x_j * tOSHatInvModsDivsModo[j][0] mod o_j with alpha = (uint64)0.5 = 0
        std::vector<NativeInteger> s(sizeO);
        for (uint32_t j = 0; j < sizeO; ++j)
            s[j] = tOSHatInvModsDivsModo[j][0];
        return FromDevice(paramsOutput, m_h.GetFormat(), detail::TimesScalars(*this, s, sizeO, sizeO));
    }
    detail::Op op(m_h.GetParams()->GetRingDimension());
    ofhe_ring* ring   = detail::RingOfTowers(op, m_h.GetAllElements(), 0, sizeQP);
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
    const auto in = DeviceSlab();
    auto out      = op.New(sizeO);
    Check(ofhe_bconv_apply(t, in.get(), out.get(), op.l.stream), "ofhe_bconv_apply");
    op.Sync();
    op.rt.n.bconv += 1;
    return FromDevice(paramsOutput, m_h.GetFormat(), std::move(out));
}

B200T B200C B200C::ApproxScaleAndRound(const std::shared_ptr<Params>& paramsP,
                                       const std::vector<std::vector<NativeInteger>>& tPSHatInvModsDivsModp,
                                       const std::vector<DoubleNativeInt>&) const {
    const uint32_t sizeQP = GetNumOfElements();
    const uint32_t sizeP  = (uint32_t)paramsP->GetParams().size();
    const uint32_t sizeQ  = sizeQP - sizeP;
    detail::Op op(m_h.GetParams()->GetRingDimension());
    ofhe_ring* ring = detail::RingOfTowers(op, m_h.GetAllElements(), 0, sizeQP);
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
    const auto in = DeviceSlab();
    auto out      = op.New(sizeP);
    Check(ofhe_bconv_apply(t, in.get(), out.get(), op.l.stream), "ofhe_bconv_apply");
    op.Sync();
    op.rt.n.bconv += 1;
    return FromDevice(paramsP, m_h.GetFormat(), std::move(out));
}

B200T void B200C::ScaleAndRoundPOverQ(const std::shared_ptr<Params>& paramsQ, const std::vector<NativeInteger>& pInvModq) {
    // x_i <- (x_i - SwitchModulus(x_last, p -> q_i)) * pInvModq[i], last tower dropped; no format change (:1674-1689)
    const uint32_t L = GetNumOfElements();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    ofhe_ring* ring = detail::RingOfTowers(op, m_h.GetAllElements(), 0, L);
    const auto x    = DeviceSlab();
    uint64_t* sw    = op.Scratch(2, L - 1);
    auto out        = op.New(L - 1);
    for (uint32_t i = 0; i + 1 < L; ++i)
        Check(ofhe_switch_modulus(ring, sw + (size_t)i * op.N, x.get() + (size_t)(L - 1) * op.N, L - 1, i, op.l.stream),
              "ofhe_switch_modulus");
    const auto idx = detail::Iota(0, L - 1);
    Check(ofhe_ew_binop(ring, 1, out.get(), x.get(), sw, L - 1, idx.data(), op.l.stream), "ofhe_ew_binop");
    const auto s = detail::Words(pInvModq, L - 1);
    Check(ofhe_ew_mul_scalars(ring, out.get(), out.get(), s.data(), L - 1, idx.data(), op.l.stream), "ofhe_ew_mul_scalars");
    op.Sync();
    // the reference keeps the first L - 1 TOWERS (their own parameter objects) under the new parameter set
    const Format f = m_h.GetFormat();
    Host meta(paramsQ, f, false);
    auto& mt       = meta.GetAllElements();
    const auto& ot = m_h.GetAllElements();
    mt.resize(L - 1, detail::Tower());
    for (uint32_t i = 0; i + 1 < L; ++i)
        mt[i] = detail::Tower(ot[i].GetParams(), ot[i].GetFormat(), false);
    Adopt(std::move(meta), std::move(out));
    op.rt.n.ew += L - 1;
}

// ---- f4  BEHZ BFV ------------------------------------------------------------------------------------------------
// FastBaseConvqToBskMontgomery (:1694-1777): Q -> Q u Bsk with the Montgomery digit modulo m~ = 2^16; result EVALUATION
B200T void B200C::FastBaseConvqToBskMontgomery(
    const std::shared_ptr<Params>& paramsQBsk, const std::vector<NativeInteger>& moduliQ,
    const std::vector<NativeInteger>& moduliBsk, const std::vector<DoubleNativeInt>&,
    const std::vector<NativeInteger>& mtildeQHatInvModq, const std::vector<NativeInteger>&,
    const std::vector<std::vector<NativeInteger>>& QHatModbsk, const std::vector<uint64_t>& QHatModmtilde,
    const std::vector<NativeInteger>& QModbsk, const std::vector<NativeInteger>&, const uint64_t& negQInvModmtilde,
    const std::vector<NativeInteger>& mtildeInvModbsk, const std::vector<NativeInteger>&) {
    const uint32_t numQ = (uint32_t)moduliQ.size(), numBsk = (uint32_t)moduliBsk.size();
    DCRTPolyType inNTT;
    const bool hadEval = m_h.GetFormat() == Format::EVALUATION;
    if (hadEval) {
        inNTT = *this;
        SwitchFormat();
    }
    detail::Op op(m_h.GetParams()->GetRingDimension());
    std::vector<uint64_t> mod, root;
    detail::CollectParams(paramsQBsk, numQ + numBsk, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    ofhe_bconv* tbl = detail::ConvTable(ring, numQ, numBsk, mtildeQHatInvModq, QHatModbsk, false, nullptr, nullptr);
    const auto x    = DeviceSlab();  // Q towers, COEFFICIENT
    auto out        = op.New(numQ + numBsk);
    uint64_t* bsk   = out.get() + (size_t)numQ * op.N;
    uint64_t* rm    = op.Scratch(1, 1);
    const auto qi = detail::Iota(0, numQ), bi = detail::Iota(numQ, numBsk);
    Check(ofhe_bconv_apply(tbl, x.get(), bsk, op.l.stream), "ofhe_bconv_apply");
    const auto hat = detail::Words(mtildeQHatInvModq, numQ);
    Check(ofhe_behz_mtilde(ring, numQ, qi.data(), x.get(), hat.data(), QHatModmtilde.data(), negQInvModmtilde, rm,
                           op.l.stream),
          "ofhe_behz_mtilde");
    const auto qm = detail::Words(QModbsk, numBsk), mi = detail::Words(mtildeInvModbsk, numBsk);
    Check(ofhe_behz_mtilde_fix(ring, numBsk, bi.data(), bsk, rm, qm.data(), mi.data(), op.l.stream),
          "ofhe_behz_mtilde_fix");
    Check(ofhe_ntt_fwd(ring, bsk, bsk, numBsk, bi.data(), op.l.stream), "ofhe_ntt_fwd");
    if (hadEval)
        op.Copy(out.get(), inNTT.DeviceSlab().get(), numQ);
    else
        Check(ofhe_ntt_fwd(ring, x.get(), out.get(), numQ, qi.data(), op.l.stream), "ofhe_ntt_fwd");
    op.Sync();
    Adopt(Meta(paramsQBsk, Format::EVALUATION), std::move(out));
    op.rt.n.bconv += 1;
    op.rt.n.ntt += numQ + numBsk;
}

// FastRNSFloorq (:1780-1830): in place on [Q | Bsk]; the Q towers come out multiplied by tQHatInvModq, the Bsk towers
// as  x * tQInvModbsk - sum_i (x_i tQHatInvModq_i) qInvModbsk[i][j]
B200T void B200C::FastRNSFloorq(const NativeInteger&, const std::vector<NativeInteger>& moduliQ,
                                const std::vector<NativeInteger>& moduliBsk, const std::vector<DoubleNativeInt>&,
                                const std::vector<NativeInteger>& tQHatInvModq, const std::vector<NativeInteger>&,
                                const std::vector<std::vector<NativeInteger>>&,
                                const std::vector<std::vector<NativeInteger>>& qInvModbsk,
                                const std::vector<NativeInteger>& tQInvModbsk, const std::vector<NativeInteger>&) {
    const uint32_t numQ = (uint32_t)moduliQ.size(), numBsk = (uint32_t)moduliBsk.size();
    detail::Op op(m_h.GetParams()->GetRingDimension());
    ofhe_ring* ring = detail::RingOfTowers(op, m_h.GetAllElements(), 0, numQ + numBsk);
    ofhe_bconv* tbl = detail::ConvTable(ring, numQ, numBsk, tQHatInvModq, qInvModbsk, false, nullptr, nullptr);
    const auto x    = DeviceSlab();
    auto out        = op.New(numQ + numBsk);
    uint64_t* aq    = op.Scratch(1, numBsk);
    uint64_t* bsk   = out.get() + (size_t)numQ * op.N;
    const auto qi = detail::Iota(0, numQ), bi = detail::Iota(numQ, numBsk);
    const auto tq = detail::Words(tQHatInvModq, numQ), tb = detail::Words(tQInvModbsk, numBsk);
    Check(ofhe_ew_mul_scalars(ring, out.get(), x.get(), tq.data(), numQ, qi.data(), op.l.stream), "ofhe_ew_mul_scalars");
    Check(ofhe_bconv_apply(tbl, x.get(), aq, op.l.stream), "ofhe_bconv_apply");
    Check(ofhe_ew_mul_scalars(ring, bsk, x.get() + (size_t)numQ * op.N, tb.data(), numBsk, bi.data(), op.l.stream),
          "ofhe_ew_mul_scalars");
    Check(ofhe_ew_binop(ring, 1, bsk, bsk, aq, numBsk, bi.data(), op.l.stream), "ofhe_ew_binop");
    op.Sync();
    m_d = std::move(out);
    m_w.store(kDev, std::memory_order_release);
    op.rt.n.bconv += 1;
}

// FastBaseConvSK (:1833-1921): Bsk -> Q with the exact correction through m_sk (Shenoy-Kumaresan); the Bsk towers are
// erased, the parameter set becomes paramsQ
B200T void B200C::FastBaseConvSK(const std::shared_ptr<Params>& paramsQ, const std::vector<DoubleNativeInt>&,
                                 const std::vector<NativeInteger>& moduliBsk, const std::vector<DoubleNativeInt>&,
                                 const std::vector<NativeInteger>& BHatInvModb, const std::vector<NativeInteger>&,
                                 const std::vector<NativeInteger>& BHatModmsk, const NativeInteger& BInvModmsk,
                                 const NativeInteger&, const std::vector<std::vector<NativeInteger>>& BHatModq,
                                 const std::vector<NativeInteger>& BModq, const std::vector<NativeInteger>&) {
    const uint32_t sizeQ = (uint32_t)paramsQ->GetParams().size(), sizeBsk = (uint32_t)moduliBsk.size(), nb = sizeBsk - 1;
    const uint32_t T     = GetNumOfElements();
    if (T < sizeQ + sizeBsk)
        OPENFHE_THROW("b200 backend: FastBaseConvSK expects [Q | Bsk] towers");
    detail::Op op(m_h.GetParams()->GetRingDimension());
    // ring: the nb source towers of B, then the targets m_sk, q_0 .. q_{sizeQ-1}
    const auto& towers = m_h.GetAllElements();
    std::vector<uint64_t> mod, root;
    detail::Collect(towers, sizeQ, nb, mod, root);
    detail::Collect(towers, sizeQ + nb, 1, mod, root);
    detail::Collect(towers, 0, sizeQ, mod, root);
    ofhe_ring* ring = op.rt.Ring(op.N, mod, root);
    std::vector<std::vector<NativeInteger>> mat(nb, std::vector<NativeInteger>(1 + sizeQ));
    for (uint32_t i = 0; i < nb; ++i) {
        mat[i][0] = BHatModmsk[i];
        for (uint32_t j = 0; j < sizeQ; ++j)
            mat[i][1 + j] = BHatModq[i][j];
    }
    ofhe_bconv* tbl = detail::ConvTable(ring, nb, 1 + sizeQ, BHatInvModb, mat, false, nullptr, nullptr);
    const auto x    = DeviceSlab();
    uint64_t* conv  = op.Scratch(1, 1 + sizeQ);  // [alpha_sk | q_0 ..]
    Check(ofhe_bconv_apply(tbl, x.get() + (size_t)sizeQ * op.N, conv, op.l.stream), "ofhe_bconv_apply");
    // alpha_sk = (alpha_sk - x_msk) * BInvModmsk mod m_sk                                                 (:1886-1889)
    const uint32_t mskIdx = nb;
    Check(ofhe_ew_binop(ring, 1, conv, conv, x.get() + (size_t)(sizeQ + nb) * op.N, 1, &mskIdx, op.l.stream), "ofhe_ew_binop");
    const uint64_t binv = BInvModmsk.ConvertToInt();
    Check(ofhe_ew_mul_scalars(ring, conv, conv, &binv, 1, &mskIdx, op.l.stream), "ofhe_ew_mul_scalars");
    auto out      = op.New(T - sizeBsk);
    const auto qi = detail::Iota(nb + 1, sizeQ);
    op.Copy(out.get(), conv + op.N, sizeQ);
    const auto bm = detail::Words(BModq, sizeQ);
    Check(ofhe_behz_sk_fix(ring, sizeQ, qi.data(), out.get(), conv, moduliBsk[nb].ConvertToInt(), bm.data(), op.l.stream),
          "ofhe_behz_sk_fix");
    // towers behind [Q | Bsk] (none in the reference's callers) keep their values                          (:1913-1920)
    op.Copy(out.get() + (size_t)sizeQ * op.N, x.get() + (size_t)(sizeQ + sizeBsk) * op.N, T - sizeQ - sizeBsk);
    op.Sync();
    const Format f = m_h.GetFormat();
    Host meta(paramsQ, f, false);
    auto& mt = meta.GetAllElements();
    mt.resize(T - sizeBsk, detail::Tower());
    for (uint32_t i = 0; i < sizeQ; ++i)
        mt[i] = detail::Tower(towers[i].GetParams(), towers[i].GetFormat(), false);
    for (uint32_t i = sizeQ + sizeBsk; i < T; ++i)
        mt[i - sizeBsk] = detail::Tower(towers[i].GetParams(), towers[i].GetFormat(), false);
    Adopt(std::move(meta), std::move(out));
    op.rt.n.bconv += 1;
}

#undef B200T
#undef B200C

}  // namespace b200
}  // namespace lbcrypto

#endif
