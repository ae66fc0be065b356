// bfv.h -- BFV homomorphic multiplication, HPSPOVERQ variant (device composition of rows a12 / a13).  See bfv.cu.
#pragma once
#include "hks.h"
#include "internal.h"

namespace ofhe {

struct BfvMult {
    Ring* ring;  // registered moduli: Q (sizeQ) then R (sizeR), the auxiliary basis of the multiplication
    uint32_t sizeQ, sizeR;
    // conversions built by the caller from the reference's CryptoParametersBFVRNS tables (they embed long-double /
    // BigInteger quotients, so they are INPUTS): rows of `in` = source limbs in order, rows of `out` = target limbs
    const BConv* q_to_r;         // exact SwitchCRTBasis Q -> R       (ExpandCRTBasis, dcrtpoly-impl.h:1088-1116)
    const BConv* q_to_r_approx;  // ApproxSwitchCRTBasis Q -> R with mNegRlQlHatInvModq (FastExpandCRTBasisPloverQ :1151-1164)
    const BConv* r_to_q;         // exact SwitchCRTBasis R -> Q
    const BConv* scale_round;    // ScaleAndRound QR -> Q             (:1513-1628)
    LimbMap map_q2, map_r2, map_qr, map_qr2, map_qr3, map_q3;
    uint64_t* ws_d = nullptr;
    size_t ws_words = 0;

    BfvMult(Ring* r, uint32_t sizeQ, uint32_t sizeR, const BConv* q_to_r, const BConv* q_to_r_approx,
            const BConv* r_to_q, const BConv* scale_round);
    ~BfvMult();
    uint64_t* scratch();
    // LeveledSHEBFVRNS::EvalMult without relinearisation (bfvrns-leveledshe.cpp:198-439, HPSPOVERQ branch):
    // ct1, ct2 [2][sizeQ][N] EVALUATION -> out3 [3][sizeQ][N] COEFFICIENT
    void evalmult_norelin(const uint64_t* ct1, const uint64_t* ct2, uint64_t* out3, cudaStream_t st);
    // EvalMult(ct1, ct2, evalKey) = the above + RelinearizeCore (:735-741, 888-935): out [2][sizeQ][N] EVALUATION.
    // `hks` is the HYBRID key-switch context over Q + P of the same context (first sizeQ moduli identical).
    void evalmult(Hks& hks, const uint64_t* ct1, const uint64_t* ct2, const EvalKey& evk, uint64_t* out, cudaStream_t st);
};

}  // namespace ofhe
