// bfv.cu -- BFV EvalMult, HPSPOVERQ (BASELINE config 3), as a device-resident composition of the kernels of rows
// a6/a7 (NTT), a12 (exact / approximate base conversion), a19 (tensor) and a13 (scale-and-round).
//
// Replaces the DCRTPoly calls of LeveledSHEBFVRNS::EvalMult (pke/lib/scheme/bfvrns/bfvrns-leveledshe.cpp:198-439):
//   ct1:  ExpandCRTBasis Q -> QR                  (:239-246; INTT, exact SwitchCRTBasis, NTT of the R limbs)
//   ct2:  SetFormat(COEFFICIENT), FastExpandCRTBasisPloverQ, SetFormat(EVALUATION)   (:256-261; Q -> R approx, R -> Q exact)
//   tensor product over the 24 limbs of QR       (:353-364)
//   each of the 3 results: SetFormat(COEFFICIENT), ScaleAndRound QR -> Q   (:385-396)
// and of RelinearizeCore (:888-935): NTT of c2, hybrid KeySwitchCore, c0 += ab0, c1 += ab1.
// Every polynomial of a step goes through ONE batched launch (limb maps / batch strides), nothing returns to the host.
#include "bfv.h"

namespace ofhe {

namespace {
void fill(LimbMap& m, uint32_t copies, uint32_t first_idx, uint32_t n, uint32_t row0, uint32_t row_stride) {
    OFHE_REQUIRE(copies * n <= (uint32_t)kMaxLimbs, "too many limbs in one call");
    m.n = copies * n;
    for (uint32_t k = 0; k < copies; ++k)
        for (uint32_t i = 0; i < n; ++i) {
            m.idx[k * n + i] = (uint16_t)(first_idx + i);
            m.pos[k * n + i] = (uint16_t)(row0 + k * row_stride + i);
        }
}
}  // namespace

BfvMult::BfvMult(Ring* r, uint32_t sizeQ_, uint32_t sizeR_, const BConv* a, const BConv* b, const BConv* c,
                 const BConv* d)
    : ring(r), sizeQ(sizeQ_), sizeR(sizeR_), q_to_r(a), q_to_r_approx(b), r_to_q(c), scale_round(d) {
    OFHE_REQUIRE(sizeQ >= 1 && sizeR >= 1 && sizeQ + sizeR <= r->moduli.size(), "BFV basis sizes do not fit the ring");
    OFHE_REQUIRE(a && b && c && d, "null conversion table");
    OFHE_REQUIRE(a->sizeQ == sizeQ && a->sizeP == sizeR && b->sizeQ == sizeQ && b->sizeP == sizeR &&
                     c->sizeQ == sizeR && c->sizeP == sizeQ && d->sizeQ == sizeR && d->sizeP == sizeQ,
                 "conversion table shapes do not match (sizeQ, sizeR)");
    const uint32_t QR = sizeQ + sizeR;
    fill(map_q2, 2, 0, sizeQ, 0, sizeQ);        // two Q polynomials, rows packed [2][Q]
    fill(map_q3, 3, 0, sizeQ, 0, sizeQ);        // three Q polynomials, rows packed [3][Q]
    fill(map_r2, 2, sizeQ, sizeR, sizeQ, QR);   // the R rows of two QR polynomials
    fill(map_qr, 1, 0, QR, 0, QR);
    fill(map_qr2, 2, 0, QR, 0, QR);
    fill(map_qr3, 3, 0, QR, 0, QR);
}
BfvMult::~BfvMult() {
    if (ws_d)
        cudaFree(ws_d);
}
uint64_t* BfvMult::scratch() {
    const size_t N = ring->N, Q = sizeQ, QR = sizeQ + sizeR;
    const size_t need = N * (2 * QR + 2 * QR + 2 * Q + 3 * QR + 3 * Q);  // e1 | e2 | coef | cv | out3 (relin)
    if (need > ws_words) {
        if (ws_d)
            OFHE_CUDA(cudaFree(ws_d));
        OFHE_CUDA(cudaMalloc(&ws_d, sizeof(uint64_t) * need));
        ws_words = need;
    }
    return ws_d;
}

void BfvMult::evalmult_norelin(const uint64_t* ct1, const uint64_t* ct2, uint64_t* out3, cudaStream_t st) {
    const Ring& r  = *ring;
    const size_t N = r.N, Q = sizeQ, QR = sizeQ + sizeR;
    uint64_t* e1   = scratch();
    uint64_t* e2   = e1 + 2 * QR * N;
    uint64_t* coef = e2 + 2 * QR * N;
    uint64_t* cv   = coef + 2 * Q * N;
    // ---- ct1: ExpandCRTBasis, both polynomials per launch
    launch_ntt_inv(r, ct1, coef, map_q2, st);                                              // dcrtpoly-impl.h:1094-1096
    launch_bconv(r, *q_to_r, coef, e1 + Q * N, 2, Q * N, QR * N, false, st);               // :1097-1100 (a12)
    launch_ntt_fwd(r, e1, e1, map_r2, st);                                                 // :1104-1106
    for (uint32_t k = 0; k < 2; ++k)                                                       // the Q limbs are kept as they are
        OFHE_CUDA(cudaMemcpyAsync(e1 + k * QR * N, ct1 + k * Q * N, sizeof(uint64_t) * Q * N, cudaMemcpyDeviceToDevice, st));
    // ---- ct2: FastExpandCRTBasisPloverQ in COEFFICIENT format
    launch_ntt_inv(r, ct2, coef, map_q2, st);                                              // bfvrns-leveledshe.cpp:257
    launch_bconv(r, *q_to_r_approx, coef, e2 + Q * N, 2, Q * N, QR * N, false, st);        // dcrtpoly-impl.h:1154-1157
    launch_bconv(r, *r_to_q, e2 + Q * N, e2, 2, QR * N, QR * N, false, st);                // :1158-1163
    launch_ntt_fwd(r, e2, e2, map_qr2, st);                                                // bfvrns-leveledshe.cpp:260
    // ---- tensor product over QR (schoolbook; USE_KARATSUBA is not defined in the reference build)   :353-364
    launch_tensor(r, e1, e1 + QR * N, e2, e2 + QR * N, cv, cv + QR * N, cv + 2 * QR * N, map_qr, st);
    // ---- t/P scaling with rounding, QR -> Q                                                          :385-396
    launch_ntt_inv(r, cv, cv, map_qr3, st);
    launch_bconv(r, *scale_round, cv, out3, 3, QR * N, Q * N, false, st);                  // dcrtpoly-impl.h:1513-1628 (a13)
}

void BfvMult::evalmult(Hks& hks, const uint64_t* ct1, const uint64_t* ct2, const EvalKey& evk, uint64_t* out,
                       cudaStream_t st) {
    OFHE_REQUIRE(hks.ring->ctx == ring->ctx && hks.ring->N == ring->N && hks.sizeQ == sizeQ,
                 "key-switch context does not match the BFV context");
    for (uint32_t i = 0; i < sizeQ; ++i)
        OFHE_REQUIRE(hks.ring->moduli[i] == ring->moduli[i], "key-switch context has different Q moduli");
    const Ring& r  = *ring;
    const size_t N = r.N, Q = sizeQ, QR = sizeQ + sizeR;
    uint64_t* c3   = scratch() + N * (2 * QR + 2 * QR + 2 * Q + 3 * QR);
    evalmult_norelin(ct1, ct2, c3, st);
    launch_ntt_fwd(r, c3, c3, map_q3, st);  // RelinearizeCore: cv[sel].SetFormat(EVALUATION), cv[0/1].SetFormat (:915,926)
    uint64_t* ws = static_cast<uint64_t*>(r.ctx->get_scratch(sizeof(uint64_t) * hks.workspace_words(sizeQ), st));
    hks.keyswitch_core(sizeQ, c3 + 2 * Q * N, evk, out, c3, ws, st);  // ab = KeySwitchCore(c2); c0 += ab0; c1 += ab1
}

}  // namespace ofhe
