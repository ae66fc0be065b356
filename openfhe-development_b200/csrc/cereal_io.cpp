// cereal_io.cpp -- the reference's wire / on-disk format for DCRTPoly, read and written WITHOUT a host-side DCRTPoly
// (SURVEY.md 8 f3).  Host code only.
//
// Format: cereal PortableBinary archives as produced by lbcrypto::Serial::Serialize(obj, stream, SerType::BINARY) for
//   DCRTPolyImpl::save           src/core/include/lattice/hal/default/dcrtpoly.h:337-341   (towers, format, params)
//   PolyImpl::save               src/core/include/lattice/hal/default/poly.h:336-340       (values, format, params)
//   NativeVectorT::save (binary) src/core/include/math/hal/intnat/mubintvecnat.h:656-665   (size, raw words, modulus)
//   ElemParams / ILParamsImpl / ILDCRTParams::save   lattice/hal/elemparams.h:229-236, default/ilparams.h:139-141,
//                                                    default/ildcrtparams.h:333-336
//   ubint::save                  src/core/include/math/hal/bigintdyn/ubintdyn.h:885-888    (64-bit limbs, MSB)
// cereal framing reproduced here: one leading endianness byte (1 = little endian); the uint32 class version of every
// versioned type the FIRST time that type is written; smart pointers as uint32 polymorphic id (0x40000000 = statically
// typed), then for unique_ptr a uint8 validity flag, for shared_ptr a uint32 object id whose top bit marks the first
// occurrence (followed by the object) -- later occurrences are the bare id.
//
// The tower payloads are raw little-endian uint64 words, so `scan` only returns their offsets: a caller maps / pins the
// serialized buffer and copies each tower straight into its device row (ofhe_dcrtpoly_upload_cereal), and `write` emits
// the same bytes the reference would, so its Deserialize accepts what the device produced.
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "cereal_io.h"

namespace ofhe {
namespace cio {

namespace {
constexpr uint32_t kStaticPtr = 0x40000000u, kFirstUse = 0x80000000u;

struct Reader {
    const uint8_t* b;
    size_t n, o = 0;
    bool seen[8] = {false, false, false, false, false, false, false, false};
    void need(size_t k) const {
        if (k > n - o)
            throw std::runtime_error("cereal: buffer ends inside a DCRTPoly record");
    }
    uint8_t u8() {
        need(1);
        return b[o++];
    }
    uint32_t u32() {
        need(4);
        uint32_t v;
        std::memcpy(&v, b + o, 4);
        o += 4;
        return v;
    }
    uint64_t u64() {
        need(8);
        uint64_t v;
        std::memcpy(&v, b + o, 8);
        o += 8;
        return v;
    }
    void version(int type) {  // class version, present only at the type's first occurrence in the archive
        if (!seen[type]) {
            seen[type] = true;
            if (u32() != 1)
                throw std::runtime_error("cereal: unsupported class version (written by a later library?)");
        }
    }
};
enum { T_DCRT, T_POLY, T_NVEC, T_NINT, T_ILP, T_EPN, T_ILDCRT, T_EPB_or_BINT };

struct Writer {
    std::vector<uint8_t>& out;
    bool seen[9] = {false, false, false, false, false, false, false, false, false};
    void raw(const void* p, size_t k) {
        const uint8_t* c = static_cast<const uint8_t*>(p);
        out.insert(out.end(), c, c + k);
    }
    void u8(uint8_t v) { raw(&v, 1); }
    void u32(uint32_t v) { raw(&v, 4); }
    void u64(uint64_t v) { raw(&v, 8); }
    void version(int type) {
        if (!seen[type]) {
            seen[type] = true;
            u32(1);
        }
    }
};
enum { W_DCRT, W_POLY, W_NVEC, W_NINT, W_ILP, W_EPN, W_ILDCRT, W_EPB, W_BINT };

uint64_t native_int(Reader& r) {
    r.version(T_NINT);
    return r.u64();
}
void native_int(Writer& w, uint64_t v) {
    w.version(W_NINT);
    w.u64(v);
}
// little-endian multi-limb integer times a word
void mul_word(std::vector<uint64_t>& a, uint64_t m) {
    unsigned __int128 carry = 0;
    for (auto& limb : a) {
        unsigned __int128 t = (unsigned __int128)limb * m + carry;
        limb  = (uint64_t)t;
        carry = t >> 64;
    }
    if (carry)
        a.push_back((uint64_t)carry);
}
void big_int(Writer& w, const std::vector<uint64_t>& limbs) {  // ubint: vector<limb> then the bit length
    w.version(W_BINT);
    w.u64(limbs.size());
    for (uint64_t l : limbs)
        w.u64(l);
    uint32_t msb = 0;
    for (size_t i = limbs.size(); i-- > 0;)
        if (limbs[i]) {
            msb = (uint32_t)(64 * i) + (64 - (uint32_t)__builtin_clzll(limbs[i]));
            break;
        }
    w.u32(msb);
}
}  // namespace

DcrtInfo scan_dcrtpoly(const void* buf, size_t len, std::vector<uint64_t>* moduli, std::vector<uint64_t>* roots,
                       std::vector<size_t>* offsets) {
    Reader r{static_cast<const uint8_t*>(buf), len};
    if (r.u8() != 1)
        throw std::runtime_error("cereal: not a little-endian portable-binary archive");
    r.version(T_DCRT);
    const uint64_t towers = r.u64();
    if (towers == 0 || towers > 4096)
        throw std::runtime_error("cereal: implausible tower count");
    DcrtInfo info{};
    info.n_towers = (uint32_t)towers;
    for (uint64_t t = 0; t < towers; ++t) {
        r.version(T_POLY);
        if (r.u32() != kStaticPtr || r.u8() != 1)
            throw std::runtime_error("cereal: tower without values");
        r.version(T_NVEC);
        const uint64_t size = r.u64();
        if (t == 0) {
            if (size == 0 || (size & (size - 1)) || size > (1u << 20))
                throw std::runtime_error("cereal: ring dimension is not a power of two");
            info.N = (uint32_t)size;
        }
        else if (size != info.N)
            throw std::runtime_error("cereal: towers of different length");
        const size_t off = r.o;
        r.need(8 * size);
        r.o += 8 * size;
        const uint64_t q = native_int(r);
        const uint32_t f = r.u32();
        if (t == 0)
            info.format = f;
        if (r.u32() != kStaticPtr)
            throw std::runtime_error("cereal: polymorphic parameter pointers are not expected here");
        const uint32_t id = r.u32();
        uint64_t root    = 0;
        if (id & kFirstUse) {  // ILNativeParams: base ElemParams<NativeInteger> = co, rd, cm, ru, bm, br
            r.version(T_ILP);
            r.version(T_EPN);
            const uint32_t co = r.u32(), rd = r.u32();
            const uint64_t cm = native_int(r);
            root              = native_int(r);
            native_int(r);
            native_int(r);
            if (rd != info.N || co != 2 * info.N || cm != q)
                throw std::runtime_error("cereal: tower parameters disagree with its values");
        }
        if (moduli)
            moduli->push_back(q);
        if (roots)
            roots->push_back(root);
        if (offsets)
            offsets->push_back(off);
    }
    const uint32_t f = r.u32();
    if (f != info.format || f > 1)
        throw std::runtime_error("cereal: inconsistent format flags");
    info.header_end = r.o;  // the ILDCRTParams record that follows is redundant with the per-tower data
    return info;
}

void write_dcrtpoly(uint32_t N, uint32_t n_towers, const uint64_t* moduli, const uint64_t* roots, uint32_t format,
                    const uint64_t* data, std::vector<uint8_t>& out) {
    if (N == 0 || (N & (N - 1)) || n_towers == 0 || format > 1)
        throw std::runtime_error("cereal: bad shape");
    out.clear();
    out.reserve((size_t)n_towers * (8 * (size_t)N + 96) + 64 + 16 * (size_t)n_towers);
    Writer w{out};
    w.u8(1);
    w.version(W_DCRT);
    w.u64(n_towers);
    for (uint32_t t = 0; t < n_towers; ++t) {
        w.version(W_POLY);
        w.u32(kStaticPtr);
        w.u8(1);
        w.version(W_NVEC);
        w.u64(N);
        w.raw(data + (size_t)t * N, 8 * (size_t)N);
        native_int(w, moduli[t]);
        w.u32(format);
        w.u32(kStaticPtr);
        w.u32(kFirstUse | (t + 1));
        w.version(W_ILP);
        w.version(W_EPN);
        w.u32(2 * N);
        w.u32(N);
        native_int(w, moduli[t]);
        native_int(w, roots[t]);
        native_int(w, 0);
        native_int(w, 0);
    }
    w.u32(format);
    w.u32(kStaticPtr);
    w.u32(kFirstUse | (n_towers + 1));
    w.version(W_ILDCRT);
    w.version(W_EPB);
    w.u32(2 * N);
    w.u32(N);
    std::vector<uint64_t> prod{1};
    for (uint32_t t = 0; t < n_towers; ++t)
        mul_word(prod, moduli[t]);
    big_int(w, prod);  // ciphertext modulus = product of the tower moduli
    const std::vector<uint64_t> zero{0};
    big_int(w, zero);  // composite root of unity, big modulus, big root: unset in ILDCRTParams
    big_int(w, zero);
    big_int(w, zero);
    w.u64(n_towers);  // m_params: references to the tower parameter objects written above
    for (uint32_t t = 0; t < n_towers; ++t) {
        w.u32(kStaticPtr);
        w.u32(t + 1);
    }
}

}  // namespace cio
}  // namespace ofhe
