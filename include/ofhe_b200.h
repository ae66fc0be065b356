/* ofhe_b200.h -- C ABI of libofhe_b200.so: a Blackwell (sm_100a) RNS / NTT polynomial-arithmetic backend for
 * OpenFHE's DCRTPoly hot path and hybrid key switching.
 *
 * The reference (openfhe-development v1.5.1) has NO C ABI for polynomial arithmetic: its boundary is the compile-time
 * lattice HAL  src/core/include/lattice/hal/dcrtpoly-interface.h:83-85,153-1417  selected in
 * src/core/include/lattice/hal/lat-backend.h:39-60.  This header is the boundary a `b200::DCRTPolyImpl` HAL class and a
 * rebound KeySwitchHYBRID call into (INTEGRATION.md shows both bindings).  Each entry point names the reference
 * function it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - A polynomial is a row-major [limbs x N] slab of uint64 residues in DEVICE memory; limb l holds residues modulo
 *    the ring modulus with index mod_idx[l].  Format::EVALUATION is the reference's bit-reversed order, so slabs can
 *    be exchanged with NativeVector storage word for word (mubintvecnat.h:120-130).
 *  - Every function returns 0 on success, non-zero on failure; ofhe_last_error() gives the message (per thread).
 *    No C++ exception crosses this boundary (the reference throws lbcrypto::OpenFHEException, utils/exception.h).
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on that stream
 *    unless they take host pointers (suffix _host), which return after the result is on the host.
 *  - There is NO CPU fallback: without a CUDA device ofhe_ctx_create fails.
 */
#ifndef OFHE_B200_H
#define OFHE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ofhe_ctx ofhe_ctx;   /* one CUDA device */
typedef struct ofhe_ring ofhe_ring; /* ring dimension + registered moduli with their device tables */
typedef struct ofhe_hks ofhe_hks;   /* hybrid key-switch context over a ring: Q, P, digits, per-level tables */
typedef struct ofhe_evk ofhe_evk;   /* device-resident evaluation key */

const char* ofhe_last_error(void);
const char* ofhe_version(void);

/* ---- context / memory ------------------------------------------------------------------------------------------- */
int ofhe_ctx_create(int device, ofhe_ctx** out);
void ofhe_ctx_destroy(ofhe_ctx* ctx);
int ofhe_ctx_sm_count(const ofhe_ctx* ctx);
int ofhe_malloc(ofhe_ctx* ctx, size_t bytes, void** dptr);
int ofhe_free(ofhe_ctx* ctx, void* dptr);
int ofhe_memcpy_h2d(ofhe_ctx* ctx, void* dst_d, const void* src_h, size_t bytes, void* stream);
int ofhe_memcpy_d2h(ofhe_ctx* ctx, void* dst_h, const void* src_d, size_t bytes, void* stream);
int ofhe_memcpy_d2d(ofhe_ctx* ctx, void* dst_d, const void* src_d, size_t bytes, void* stream);
int ofhe_memset0(ofhe_ctx* ctx, void* dptr, size_t bytes, void* stream);
int ofhe_stream_sync(ofhe_ctx* ctx, void* stream);
/* One stream per calling thread: the reference's callers invoke DCRTPoly methods from inside their own
 * `omp parallel for` (pke/lib/keyswitch/keyswitch-hybrid.cpp:423, scheme/ckksrns/ckksrns-fhe.cpp:1943); every entry
 * point below is asynchronous on the stream it is given and keeps its workspace per stream. */
int ofhe_stream_create(ofhe_ctx* ctx, void** stream_out);
int ofhe_stream_destroy(ofhe_ctx* ctx, void* stream);

/* ---- ring: replaces ILDCRTParams + the static twiddle caches
 *      ChineseRemainderTransformFTTNat::PreCompute            math/hal/intnat/transformnat-impl.h:714-756
 *      (moduli/roots as ILDCRTParams holds them, lattice/hal/default/ildcrtparams.h:60-170)                        */
int ofhe_ring_create(ofhe_ctx* ctx, uint32_t N, uint32_t n_moduli, const uint64_t* moduli_h, const uint64_t* roots_h,
                     ofhe_ring** out);
void ofhe_ring_destroy(ofhe_ring* ring);
/* copies T, T' (forward) and Tinv, Tinv' (inverse) of one modulus to the host, each N words, plus N^-1 and its
 * Shoup constant: the same six quantities the reference caches per modulus                                        */
int ofhe_ring_twiddles_host(ofhe_ring* ring, uint32_t mod_idx, uint64_t* T, uint64_t* Tp, uint64_t* Ti, uint64_t* Tip,
                            uint64_t* nInv, uint64_t* nInvPrecon);

/* ---- a6-a8  DCRTPolyImpl::SwitchFormat                       lattice/hal/default/dcrtpoly-impl.h:1932-1940
 *             ChineseRemainderTransformFTTNat::ForwardTransformToBitReverseInPlace   transformnat-impl.h:648-658,303-374
 *             ChineseRemainderTransformFTTNat::InverseTransformFromBitReverseInPlace transformnat-impl.h:678-691,512-625
 * in -> out (may be the same slab); COEFFICIENT -> EVALUATION (fwd) or back (inv)                                 */
int ofhe_ntt_fwd(ofhe_ring* ring, const uint64_t* in_d, uint64_t* out_d, uint32_t n_limbs, const uint32_t* mod_idx_h,
                 void* stream);
int ofhe_ntt_inv(ofhe_ring* ring, const uint64_t* in_d, uint64_t* out_d, uint32_t n_limbs, const uint32_t* mod_idx_h,
                 void* stream);

/* ---- a9  DCRTPolyImpl::Plus / Minus / Times / Negate (+ the Eq forms)   lattice/hal/default/dcrtpoly.h:130-190,
 *          dcrtpoly-impl.h:347-621;  op: 0 add, 1 sub, 2 mul, 3 negate(a), 4 dst += a*b.  dst may alias a or b.   */
int ofhe_ew_binop(ofhe_ring* ring, int op, uint64_t* dst_d, const uint64_t* a_d, const uint64_t* b_d, uint32_t n_limbs,
                  const uint32_t* mod_idx_h, void* stream);
/*          DCRTPolyImpl::Times(const std::vector<NativeInteger>&)          dcrtpoly-impl.h:553-567                */
int ofhe_ew_mul_scalars(ofhe_ring* ring, uint64_t* dst_d, const uint64_t* a_d, const uint64_t* scalars_h,
                        uint32_t n_limbs, const uint32_t* mod_idx_h, void* stream);
/* ---- f4  BEHZ BFV: the per-coefficient corrections of DCRTPolyImpl::FastBaseConvqToBskMontgomery
 *          (dcrtpoly-impl.h:1694-1777) and ::FastBaseConvSK (:1833-1921); their multiply-accumulate parts (and all of
 *          ::FastRNSFloorq, :1780-1830) are table-driven base conversions: ofhe_bconv_create with the reference's tables.
 *      mtilde    : r[k] = ((sum_i (x_i[k] * mtildeQHatInvModq_i mod q_i) * QHatModmtilde_i) mod 2^16) * negQInvModmtilde
 *                  mod 2^16   -> r_out_d [N]                                                             (:1724-1738)
 *      mtilde_fix: s_j[k] <- ((centre(r[k]) * QModbsk_j + s_j[k]) * mtildeInvModbsk_j) mod bsk_j, in place on
 *                  s_d [n_bsk][N] (the converted residues)                                                (:1760-1767)
 *      sk_fix    : q_j[k] <- q_j[k] - (lift(alpha[k]) * BModq_j) mod q_j, lift = the reference's ModSubFast(msk, q_j)
 *                  for alpha > msk/2, in place on q_d [n_q][N]                                            (:1903-1907)  */
int ofhe_behz_mtilde(ofhe_ring* ring, uint32_t n_q, const uint32_t* q_idx_h, const uint64_t* in_d,
                     const uint64_t* mtildeQHatInvModq_h, const uint64_t* QHatModmtilde_h, uint64_t negQInvModmtilde,
                     uint64_t* r_out_d, void* stream);
int ofhe_behz_mtilde_fix(ofhe_ring* ring, uint32_t n_bsk, const uint32_t* bsk_idx_h, uint64_t* s_d, const uint64_t* r_d,
                         const uint64_t* QModbsk_h, const uint64_t* mtildeInvModbsk_h, void* stream);
int ofhe_behz_sk_fix(ofhe_ring* ring, uint32_t n_q, const uint32_t* q_idx_h, uint64_t* q_d, const uint64_t* alpha_d,
                     uint64_t msk, const uint64_t* BModq_h, void* stream);
/* ---- a4  NativeVectorT::SwitchModulus on one limb               math/hal/intnat/mubintvecnat.cpp:109-122        */
int ofhe_switch_modulus(ofhe_ring* ring, uint64_t* dst_d, const uint64_t* src_d, uint32_t from_idx, uint32_t to_idx,
                        void* stream);
/* ---- a15 DCRTPolyImpl::AutomorphismTransform(k, precomp) (EVALUATION)     dcrtpoly-impl.h:325-333,
 *          poly-impl.h:366-376, PrecomputeAutoMap core/lib/math/nbtheory2.cpp:264-275 (map built on the device)    */
int ofhe_automorphism(ofhe_ring* ring, uint64_t* dst_d, const uint64_t* src_d, uint32_t n_limbs, uint32_t k,
                      void* stream);
int ofhe_automorphism_map_host(ofhe_ring* ring, uint32_t k, uint32_t* map_h);
/*          the COEFFICIENT-format branch of PolyImpl::AutomorphismTransform       poly-impl.h:359-362
 *          out[jk mod N] = (jk div N odd) ? q - in[j] : in[j]  (plain subtraction: 0 maps to q, as in the reference) */
int ofhe_automorphism_coeff(ofhe_ring* ring, uint64_t* dst_d, const uint64_t* src_d, uint32_t n_limbs,
                            const uint32_t* mod_idx_h, uint32_t k, void* stream);

/* ---- a10 DCRTPolyImpl::ApproxSwitchCRTBasis                                 dcrtpoly-impl.h:888-932
 * generic source basis {src_idx} -> target basis {dst_idx} (QHatInvModq / QHatModp are derived from the moduli);
 * in_d [n_src][N] COEFFICIENT -> out_d [n_dst][N] COEFFICIENT                                                      */
int ofhe_approx_switch_crt_basis(ofhe_ring* ring, uint32_t n_src, const uint32_t* src_idx_h, uint32_t n_dst,
                                 const uint32_t* dst_idx_h, const uint64_t* in_d, uint64_t* out_d, void* stream);

/* ---- table-driven base conversions: the caller supplies the reference's precomputed tables (a20) verbatim.
 * a10 DCRTPolyImpl::ApproxSwitchCRTBasis with explicit tables      dcrtpoly-impl.h:888-932   (alphaQModp_h == NULL)
 *     e.g. the first half of FastExpandCRTBasisPloverQ              dcrtpoly-impl.h:1151-1164
 * a12 DCRTPolyImpl::SwitchCRTBasis (exact, HPS)                     dcrtpoly-impl.h:1008-1085 (alphaQModp_h, qInv_h given)
 *     used by ExpandCRTBasis                                        dcrtpoly-impl.h:1088-1116
 *   QHatModp_h is [n_dst][n_src] if table_is_dst_major (the order SwitchCRTBasis takes), else [n_src][n_dst]
 *   (the order ApproxSwitchCRTBasis takes); alphaQModp_h is [n_src+1][n_dst]; qInv_h is [n_src] doubles.
 *   The overflow estimate nu = 0.5 + sum double(y_i)*qInv[i] is evaluated in index order with separately rounded
 *   IEEE multiply and add (no FMA), so (size_t)nu equals the reference's on every coefficient.
 *   apply: in_d [n_src][N] COEFFICIENT -> out_d [n_dst][N] COEFFICIENT                                            */
typedef struct ofhe_bconv ofhe_bconv;
int ofhe_bconv_create(ofhe_ring* ring, uint32_t n_src, const uint32_t* src_idx_h, const uint64_t* QHatInvModq_h,
                      uint32_t n_dst, const uint32_t* dst_idx_h, const uint64_t* QHatModp_h, int table_is_dst_major,
                      const uint64_t* alphaQModp_h, const double* qInv_h, ofhe_bconv** out);
/* a13 DCRTPolyImpl::ScaleAndRound (DCRT -> DCRT)                    dcrtpoly-impl.h:1513-1628
 *   input slab in_d [sizeI + sizeO][N] COEFFICIENT: the sizeO limbs of the output basis come first if
 *   out_limbs_first (output modulus == first input modulus), else last.  in_idx_h[sizeI] / out_idx_h[sizeO] are the
 *   ring indices of the scaled-away and of the surviving limbs; tbl_h is tOSHatInvModsDivsModo [sizeO][sizeI+1],
 *   frac_h is tOSHatInvModsDivsFrac [sizeI].  apply: in_d -> out_d [sizeO][N]                                       */
int ofhe_scale_round_create(ofhe_ring* ring, uint32_t sizeI, const uint32_t* in_idx_h, uint32_t sizeO,
                            const uint32_t* out_idx_h, int out_limbs_first, const uint64_t* tbl_h, const double* frac_h,
                            ofhe_bconv** out);
int ofhe_bconv_apply(ofhe_bconv* t, const uint64_t* in_d, uint64_t* out_d, void* stream);
void ofhe_bconv_destroy(ofhe_bconv* t);

/* diagnostic: how many thread-block clusters of the forward / inverse NTT kernel (one cluster = one limb of a
 * 2^logN ring, 13 <= logN <= 16) the device holds at once: fwd_inv2[0], fwd_inv2[1]                                 */
int ofhe_ntt_cluster_occupancy(ofhe_ctx* ctx, uint32_t logN, int* fwd_inv2);

/* ---- hybrid key switching.  The ring must hold Q at indices [0,sizeQ) and P at [sizeQ, sizeQ+sizeP).
 *      Tables: CryptoParametersRNS::PrecomputeCRTTables      pke/lib/schemerns/rns-cryptoparameters.cpp:48-401
 *              CryptoParametersCKKSRNS::PrecomputeCRTTables  pke/lib/scheme/ckksrns/ckksrns-cryptoparameters.cpp:44-160 */
int ofhe_hks_create(ofhe_ring* ring, uint32_t sizeQ, uint32_t sizeP, uint32_t numPartQ, ofhe_hks** out);
void ofhe_hks_destroy(ofhe_hks* hks);
/* Concurrency and capture contract of every fused entry point below (modup .. bfv_evalmult): workspace comes from a
 * per-(context, stream) arena, so calls on DIFFERENT streams of one context / ring / hks / evk are independent and may
 * be issued from different host threads (the reference's callers sit inside `omp parallel for`,
 * keyswitch-hybrid.cpp:423, ckksrns-fhe.cpp:1943); calls on ONE stream are ordered by the stream.  The first call
 * at a level on a stream allocates (level tables: once per hks; arena: once per stream) -- that is a cudaMalloc, not
 * legal under stream capture.  ofhe_hks_reserve does both ahead of time for (sizeQl, stream): after it, every entry
 * point at that level is asynchronous and capturable on that stream.                                                */
int ofhe_hks_reserve(ofhe_hks* hks, uint32_t sizeQl, void* stream);
uint32_t ofhe_hks_num_per_part(const ofhe_hks* hks); /* alpha = GetNumPerPartQ()  rns-cryptoparameters.h:401 */
/* table read-back for parity checks against the reference getters (rns-cryptoparameters.h:314-560).
 * which: 0 PInvModq[sizeQ]  1 PModq[sizeQ]  2 PHatInvModp[sizeP]  3 PHatModq[sizeP][sizeQ]
 *        4 PartQlHatInvModq(part=a, digit_size=b)[b]   5 PartQlHatModp(sizeQl=a, part=b)[digit][compl]
 *        6 QlQlInvModqlDivqlModq(sizeQl=a)[a-1]        7 qlInvModq(sizeQl=a)[a-1]
 * returns the number of words written (or needed when out_h is NULL) through *n_words                              */
int ofhe_hks_table_host(ofhe_hks* hks, int which, uint32_t a, uint32_t b, uint64_t* out_h, size_t* n_words);

/* evaluation key upload: a_h, b_h are [dnum][sizeQ+sizeP][N] EVALUATION (EvalKeyRelinImpl A / B vectors,
 * pke/include/key/evalkeyrelin.h:141-171)                                                                          */
int ofhe_evk_create(ofhe_hks* hks, const uint64_t* a_h, const uint64_t* b_h, ofhe_evk** out);
/* synthetic key of uniform residues generated on the device (benchmarks: throughput does not depend on the values)  */
int ofhe_evk_create_random(ofhe_hks* hks, uint64_t seed, ofhe_evk** out);
/* copies the key's A (which = 0) or B (which = 1) vector, [dnum][sizeQ+sizeP][N], into a device buffer of the caller
 * (limb-sharded callers slice their rows out of it)                                                                 */
int ofhe_evk_copy_to_device(const ofhe_evk* evk, int which, uint64_t* dst_d, void* stream);
void ofhe_evk_destroy(ofhe_evk* evk);

/* a16 KeySwitchHYBRID::EvalKeySwitchPrecomputeCore              pke/lib/keyswitch/keyswitch-hybrid.cpp:314-379
 *     c_d [sizeQl][N] EVALUATION -> digits_d [numPartQl][sizeQl+sizeP][N] EVALUATION                                */
int ofhe_hks_modup(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* c_d, uint64_t* digits_d, void* stream);
uint32_t ofhe_hks_num_digits(const ofhe_hks* hks, uint32_t sizeQl);
/* a17 KeySwitchHYBRID::EvalFastKeySwitchCoreExt                 keyswitch-hybrid.cpp:402-435
 *     -> e01_d [2][sizeQl+sizeP][N]  (element 0 pairs with the key's B vector, element 1 with A)                    */
int ofhe_hks_inner_product(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* digits_d, const ofhe_evk* evk,
                           uint64_t* e01_d, void* stream);
/* a11 DCRTPolyImpl::ApproxModDown (t = 0, CKKS)                  dcrtpoly-impl.h:966-1005, on n_polys polynomials
 *     in_d [n_polys][sizeQl+sizeP][N] -> out_d [n_polys][sizeQl][N];  n_polys must be 1 or 2                         */
int ofhe_hks_approx_mod_down(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* in_d, uint64_t* out_d, uint32_t n_polys,
                             void* stream);
/* a18 KeySwitchHYBRID::KeySwitchCore                             keyswitch-hybrid.cpp:308-312,381-400
 *     c_d [sizeQl][N] -> out01_d [2][sizeQl][N]                                                                     */
int ofhe_hks_keyswitch_core(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* c_d, const ofhe_evk* evk,
                            uint64_t* out01_d, void* stream);

/* a19 CryptoContextImpl::EvalMult(ct1, ct2) = EvalMultCore + relinearise
 *     pke/include/cryptocontext.h:1871-1879, pke/lib/schemebase/base-leveledshe.cpp:202-214,608-646
 *     ct*_d [2][sizeQl][N] -> out_d [2][sizeQl][N]  (out_d must not alias the inputs)                               */
int ofhe_ckks_evalmult(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* ct1_d, const uint64_t* ct2_d,
                       const ofhe_evk* evk, uint64_t* out_d, void* stream);
/*     same operation on HOST buffers: copies the inputs to the device, runs, copies the result back, synchronises.  */
int ofhe_ckks_evalmult_host(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* ct1_h, const uint64_t* ct2_h,
                            const ofhe_evk* evk, uint64_t* out_h);
/*     n independent multiplications on HOST buffers, software-pipelined: while multiplication k runs, the inputs of
 *     k+1 stream in on the H2D copy engine and the result of k-1 streams out on the D2H engine (three staging slots,
 *     one compute stream).  Host buffers must be page-locked for the copies to overlap.  ct1_h[i], ct2_h[i] and
 *     out_h[i] are [2][sizeQl][N]; returns when every result is on the host.                                        */
int ofhe_ckks_evalmult_host_batch(ofhe_hks* hks, uint32_t sizeQl, uint32_t n, const uint64_t* const* ct1_h,
                                  const uint64_t* const* ct2_h, const ofhe_evk* evk, uint64_t* const* out_h);
/* a14 DCRTPolyImpl::DropLastElementAndScale over n_polys polynomials (CKKS rescale)
 *     dcrtpoly-impl.h:693-712, pke/lib/scheme/ckksrns/ckksrns-leveledshe.cpp:172-191
 *     in_d [n_polys][sizeQl][N] -> out_d [n_polys][sizeQl-1][N]                                                     */
int ofhe_ckks_rescale(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* in_d, uint64_t* out_d, uint32_t n_polys,
                      void* stream);
/* a15+a18 LeveledSHEBase::EvalAutomorphism (EvalRotate)          pke/lib/schemebase/base-leveledshe.cpp:382-422
 *     ct_d [2][sizeQl][N], automorphism index k (FindAutomorphismIndex2nComplex, nbtheory2.cpp:243-262)             */
int ofhe_ckks_rotate(ofhe_hks* hks, uint32_t sizeQl, uint32_t k, const uint64_t* ct_d, const ofhe_evk* evk,
                     uint64_t* out_d, void* stream);
uint32_t ofhe_find_automorphism_index(int32_t index, uint32_t m);
/* ---- hoisted rotations: one ModUp shared by many rotations of the same ciphertext.
 *      EvalFastRotationPrecompute (base-leveledshe.cpp:425-431)  = ofhe_hks_modup on c1 -> digits_d
 *      EvalFastRotation           (base-leveledshe.cpp:433-463)  = ofhe_ckks_fast_rotation:  (auto_k(ba0 + c0), auto_k(ba1))
 *      EvalFastRotationExt        (ckksrns-leveledshe.cpp:534-582) = ofhe_ckks_fast_rotation_ext: result stays in QlP,
 *                                  out_ext_d [2][sizeQl+sizeP][N]; add_first adds c0 * (P mod q_i)
 *      KeySwitchExt               (keyswitch-hybrid.cpp:217-243) = ofhe_hks_keyswitch_ext: [c_k * (P mod q_i) | 0]
 *      KeySwitchDown              (keyswitch-hybrid.cpp:246-277) = ofhe_hks_approx_mod_down(..., n_polys = 2)
 *      ct_d [2][sizeQl][N], digits_d [numPartQl][sizeQl+sizeP][N] (caller-owned), EVALUATION format throughout.       */
int ofhe_ckks_fast_rotation(ofhe_hks* hks, uint32_t sizeQl, uint32_t k, const uint64_t* ct_d, const uint64_t* digits_d,
                            const ofhe_evk* evk, uint64_t* out_d, void* stream);
int ofhe_ckks_fast_rotation_ext(ofhe_hks* hks, uint32_t sizeQl, uint32_t k, int add_first, const uint64_t* ct_d,
                                const uint64_t* digits_d, const ofhe_evk* evk, uint64_t* out_ext_d, void* stream);
int ofhe_hks_keyswitch_ext(ofhe_hks* hks, uint32_t sizeQl, int add_first, const uint64_t* ct_d, uint64_t* out_ext_d,
                           void* stream);

/* ---- f2  one level of the double-hoisted baby-step / giant-step linear transform: the body of the `for s` loops of
 *      FHECKKSRNS::EvalCoeffsToSlots / EvalSlotsToCoeffs   pke/lib/scheme/ckksrns/ckksrns-fhe.cpp:1934-1987, 2094-2139
 *      (EvalLinearTransform :1834-1882 is the same scheme with one hoisting level), every intermediate device resident:
 *        digits = EvalFastRotationPrecompute(ct);  fr[j] = EvalFastRotationExt(ct, k_in[j], digits, true) | KeySwitchExt(ct, true)
 *        for i < b: inner = sum_j fr[j] (.) A[g i + j]  (EvalMultExt / EvalAddExtInPlace; g i + j == skip left out, j > 0)
 *                   first (+)= KeySwitchDown(inner)_0 [rotated by k_out[i]];  outer (+)= EvalFastRotationExt(KeySwitchDown(inner), k_out[i]) | inner
 *        out = KeySwitchDown(outer); out_0 += first
 *      k_in_h[g] / k_out_h[b]: AUTOMORPHISM indices (ofhe_find_automorphism_index of the rotation; 0 = no rotation) with
 *      their keys; k_out_h[0] is ignored (the reference's i == 0 branch never rotates).  A_d [b*g][sizeQl+sizeP][N]: the
 *      pre-encoded diagonals in the extended basis QlP (EVALUATION).  ct_d, out_d [2][sizeQl][N].  ws_d: caller-owned
 *      workspace of ofhe_ckks_bsgs_workspace_words(hks, sizeQl, g) words -- the call allocates nothing, is asynchronous
 *      on `stream` and capturable once ofhe_hks_reserve(hks, sizeQl, stream) has run.                                   */
size_t ofhe_ckks_bsgs_workspace_words(const ofhe_hks* hks, uint32_t sizeQl, uint32_t g);
int ofhe_ckks_bsgs_level(ofhe_hks* hks, uint32_t sizeQl, uint32_t g, uint32_t b, const uint32_t* k_in_h,
                         const uint32_t* k_out_h, const ofhe_evk* const* evk_in_h, const ofhe_evk* const* evk_out_h,
                         const uint64_t* A_d, uint32_t skip, const uint64_t* ct_d, uint64_t* out_d, uint64_t* ws_d,
                         void* stream);

/* ---- a20 (host helper, setup time) CKKS modulus chain + auxiliary primes with the reference's selection rules
 *      ParameterGenerationCKKSRNS::SinglePrimeModuliGen  pke/lib/scheme/ckksrns/ckksrns-parametergeneration.cpp:409-518
 *      scal_tech: 0 FIXEDMANUAL 1 FIXEDAUTO 2 FLEXIBLEAUTO 3 FLEXIBLEAUTOEXT (sizeQ = depth + 2).  shape_out: sizeQ, sizeP, dnum, alpha.
 *      moduli_out / roots_out need room for (depth + 1) + 64 words.                                                */
int ofhe_ckks_paramgen(uint32_t N, uint32_t depth, uint32_t scale_bits, uint32_t first_bits, uint32_t dnum,
                       int scal_tech, uint64_t* moduli_out, uint64_t* roots_out, uint32_t* shape_out);
uint64_t ofhe_root_of_unity(uint64_t m, uint64_t q);         /* RootOfUnity   nbtheory-impl.h:183-231 */
uint64_t ofhe_first_prime(uint32_t bits, uint64_t m);        /* FirstPrime    nbtheory-impl.h:329-347 */
uint64_t ofhe_last_prime(uint32_t bits, uint64_t m);         /* LastPrime     :349-373 */
uint64_t ofhe_previous_prime(uint64_t q, uint64_t m);        /* PreviousPrime :385-393 */
uint64_t ofhe_next_prime(uint64_t q, uint64_t m);            /* NextPrime     :375-383 */

/* ---- limb-sharded EvalMult across `world` ranks (one GPU each), SURVEY.md 8(e).  Q limb i lives on rank i mod world
 *      (local row i div world), P limb j on rank (j + sizeQ) mod world (local row j div world).  Per rank:
 *        ct*_d / out_d   [2][nOwnQ][N]                local rows of the ciphertexts
 *        evk_a_d/evk_b_d [dnum][nKeyRows][N]          local rows of the key (owned Q rows of the FULL Q, then owned P rows)
 *        gatherQ_d       [world][slotsQ][N]           all-gather buffer (this rank writes rows rank*slotsQ ...)
 *        gatherP_d       [world][2][slotsP][N]        all-gather buffer (this rank writes rows (rank*2+b)*slotsP ...)
 *      Call a, all-gather gatherQ in place, call b, all-gather gatherP in place, call c.  The collectives are the
 *      caller's (NCCL): they replace nothing in the reference, which is single-process.        */
typedef struct ofhe_shard ofhe_shard;
int ofhe_shard_create(ofhe_ring* ring, uint32_t sizeQ, uint32_t sizeP, uint32_t numPartQ, uint32_t rank, uint32_t world,
                      ofhe_shard** out);
void ofhe_shard_destroy(ofhe_shard* s);
/* info_out: nOwnQ(sizeQl), nOwnP, slotsQ(sizeQl), slotsP, nKeyRows, numDigits(sizeQl) */
int ofhe_shard_info(ofhe_shard* s, uint32_t sizeQl, uint32_t* info_out);
/* owned limb indices at this level: q_idx_out[nOwnQ] (indices into Q), p_idx_out[nOwnP] (indices into P) */
int ofhe_shard_owned(ofhe_shard* s, uint32_t sizeQl, uint32_t* q_idx_out, uint32_t* p_idx_out);
int ofhe_shard_evalmult_a(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct1_d, const uint64_t* ct2_d, uint64_t* out_d,
                          uint64_t* gatherQ_d, void* stream);
int ofhe_shard_evalmult_b(ofhe_shard* s, uint32_t sizeQl, const uint64_t* gatherQ_d, const uint64_t* evk_a_d,
                          const uint64_t* evk_b_d, uint64_t* gatherP_d, void* stream);
int ofhe_shard_evalmult_c(ofhe_shard* s, uint32_t sizeQl, const uint64_t* gatherP_d, uint64_t* out_d, void* stream);

/* ---- the same exchange WITHOUT a collective call: the last pass of the inverse NTT that produces a rank's rows
 *      stores them into every rank's gather buffer directly (NVLink peer stores; the transfer overlaps the transform),
 *      and a one-CTA device barrier replaces the all-gather's completion.  `*_of_rank_h` are HOST arrays of `world`
 *      DEVICE pointers, entry r = rank r's buffer as mapped in THIS process (own allocation for r == rank, an
 *      ofhe_ipc_import()ed mapping otherwise; plain pointers when several virtual ranks share one process).
 *      Sequence per multiplication: a_push, peer_barrier(index 2*slot), b_push, peer_barrier(2*slot+1), c.
 *      Buffers may be reused for the next multiplication on the same stream without further synchronisation: every
 *      rank passed the second barrier only after all ranks finished reading gatherQ, and passes the next first barrier
 *      only after all ranks finished reading gatherP (multi-GPU).                                          */
int ofhe_shard_evalmult_a_push(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct1_d, const uint64_t* ct2_d,
                               uint64_t* out_d, uint64_t* const* gatherQ_of_rank_h, void* stream);
int ofhe_shard_evalmult_b_push(ofhe_shard* s, uint32_t sizeQl, const uint64_t* evk_a_d, const uint64_t* evk_b_d,
                               uint64_t* const* gatherQ_of_rank_h, uint64_t* const* gatherP_of_rank_h, void* stream);
/* ---- limb-sharded EvalAutomorphism / EvalRotate and rescale: the other two operations a bootstrapping-sized circuit
 *      (BASELINE config 4) needs besides EvalMult, on the same ownership plan and gather buffers.
 *      Rotation (LeveledSHEBase::EvalAutomorphism, base-leveledshe.cpp:382-422 = KeySwitchInPlace + automorphism):
 *        rotate_a, exchange gatherQ, ofhe_shard_evalmult_b with the ROTATION key's local rows, exchange gatherP,
 *        rotate_c.  ct_d / tmp_d / out_d: [2][nOwnQ][N]; k: the automorphism index (ofhe_find_automorphism_index).
 *      Rescale (DCRTPolyImpl::DropLastElementAndScale, dcrtpoly-impl.h:693-712, both polynomials): rescale_a on
 *        every rank (only the owner of limb sizeQl-1 works: it writes the limb's coefficient form into gatherL [2][N]),
 *        broadcast of gatherL from that owner (or the _push form + one peer barrier), rescale_b.
 *        out_d: [2][nOwnQ(sizeQl-1)][N].                                                                              */
int ofhe_shard_rotate_a(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, uint64_t* tmp_d, uint64_t* gatherQ_d,
                        void* stream);
int ofhe_shard_rotate_a_push(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, uint64_t* tmp_d,
                             uint64_t* const* gatherQ_of_rank_h, void* stream);
int ofhe_shard_rotate_c(ofhe_shard* s, uint32_t sizeQl, uint32_t k, const uint64_t* gatherP_d, uint64_t* tmp_d,
                        uint64_t* out_d, void* stream);
int ofhe_shard_rescale_a(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, uint64_t* gatherL_d, void* stream);
int ofhe_shard_rescale_a_push(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, uint64_t* const* gatherL_of_rank_h,
                              void* stream);
int ofhe_shard_rescale_b(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, const uint64_t* gatherL_d, uint64_t* out_d,
                         void* stream);
/* flags buffers: ofhe_peer_flags_words(world, n_barriers) zero-initialised words per rank.  The kernel bumps this
 * rank's epoch for barrier `index`, publishes it to every rank (release, system scope) and waits until every rank's
 * arrival for that epoch is visible (acquire).  A wait longer than 10 s of wall clock (%globaltimer) counts a timeout
 * (readable with ofhe_peer_barrier_timeouts) and then ABORTS this rank's CUDA context (__trap): the phases behind the
 * barrier would otherwise consume partially written gather buffers and return a silently wrong ciphertext; after the
 * abort every later CUDA call of the rank fails.  `index` must be below the n_barriers the flags were sized for.  The
 * spin needs the barrier kernels of all slots to be co-resident: one hardware queue per stream
 * (CUDA_DEVICE_MAX_CONNECTIONS >= number of streams in flight), as bench.py and tests/conftest.py set it.            */
size_t ofhe_peer_flags_words(uint32_t world, uint32_t n_barriers);
int ofhe_peer_barrier(ofhe_ctx* ctx, uint64_t* const* flags_of_rank_h, uint32_t world, uint32_t rank, uint32_t index,
                      void* stream);
int ofhe_peer_barrier_timeouts(ofhe_ctx* ctx, const uint64_t* flags_d, uint32_t world, uint32_t n_barriers,
                               uint64_t* count_out);
/* CUDA IPC for ofhe_malloc()ed buffers (one process per GPU): export a 64-byte handle, import it in another process
 * (enables peer access), release the mapping.                                                                        */
int ofhe_ipc_export(ofhe_ctx* ctx, void* dptr, void* handle64_out);
int ofhe_ipc_import(ofhe_ctx* ctx, const void* handle64, void** dptr_out);
int ofhe_ipc_release(ofhe_ctx* ctx, void* dptr);

/* ---- BFV homomorphic multiplication, HPSPOVERQ (BASELINE config 3): LeveledSHEBFVRNS::EvalMult
 *      pke/lib/scheme/bfvrns/bfvrns-leveledshe.cpp:198-439 (tensor + t/P scale-and-round) and :735-741, 888-935
 *      (RelinearizeCore).  `ring` registers Q (sizeQ limbs) then R (sizeR limbs).  The four conversions are built by
 *      the caller with ofhe_bconv_create / ofhe_scale_round_create from CryptoParametersBFVRNS getters at level
 *      l = sizeQ-1 (the tables embed long-double / BigInteger quotients, so they are inputs):
 *        q_to_r_exact   mode 1, Q -> R: GetQlHatInvModq, GetQlHatModr, GetalphaQlModr, GetqInv
 *        q_to_r_approx  mode 0, Q -> R: GetmNegRlQlHatInvModq, GetqInvModr
 *        r_to_q_exact   mode 1, R -> Q: GetRlHatInvModr, GetRlHatModq, GetalphaRlModq, GetrInv
 *        scale_round    mode 2, QR -> Q: GettQlSlHatInvModsDivsModq, GettQlSlHatInvModsDivsFrac
 *      ct*_d [2][sizeQ][N] EVALUATION.  norelin: out3_d [3][sizeQ][N] COEFFICIENT (the reference's cvMult).  With
 *      relinearisation: `hks` = HYBRID context over Q + P of the same crypto context, out_d [2][sizeQ][N] EVALUATION.   */
typedef struct ofhe_bfv ofhe_bfv;
/* The same tables from (Q, R, t) alone, for callers without a CryptoParametersBFVRNS at hand (host routine, word-size
 * arithmetic; bfvrns-cryptoparameters.cpp:140-600).  which: 0 QlHatInvModq[sizeQ]  1 QlHatModr[sizeR][sizeQ]
 * 2 alphaQlModr[sizeQ+1][sizeR]  3 qInv[sizeQ] (double)  4 negRlQlHatInvModq[sizeQ]  5 qInvModr[sizeQ][sizeR]
 * 6 RlHatInvModr[sizeR]  7 RlHatModq[sizeQ][sizeR]  8 alphaRlModq[sizeR+1][sizeQ]  9 rInv[sizeR] (double)
 * 10 tQlSlHatInvModsDivsModq[sizeQ][sizeR+1]  11 tQlSlHatInvModsDivsFrac[sizeR] (double); out_h needs room for the table */
int ofhe_bfv_tables_host(uint32_t sizeQ, const uint64_t* moduliQ_h, uint32_t sizeR, const uint64_t* moduliR_h, uint64_t t,
                         int which, void* out_h, size_t* count_out);
int ofhe_bfv_create(ofhe_ring* ring, uint32_t sizeQ, uint32_t sizeR, const ofhe_bconv* q_to_r_exact,
                    const ofhe_bconv* q_to_r_approx, const ofhe_bconv* r_to_q_exact, const ofhe_bconv* scale_round,
                    ofhe_bfv** out);
void ofhe_bfv_destroy(ofhe_bfv* b);
int ofhe_bfv_evalmult_norelin(ofhe_bfv* b, const uint64_t* ct1_d, const uint64_t* ct2_d, uint64_t* out3_d, void* stream);
int ofhe_bfv_evalmult(ofhe_bfv* b, ofhe_hks* hks, const uint64_t* ct1_d, const uint64_t* ct2_d, const ofhe_evk* evk,
                      uint64_t* out_d, void* stream);

/* ---- the reference's wire / on-disk format: cereal portable-binary archives of a DCRTPoly as written by
 *      Serial::Serialize(poly, stream, SerType::BINARY) -- DCRTPolyImpl::save dcrtpoly.h:337-341, PolyImpl::save
 *      poly.h:336-340, NativeVectorT::save mubintvecnat.h:656-665, ElemParams::save elemparams.h:229-236.
 *      scan : shape_out = (N, n_towers, format 0 EVALUATION / 1 COEFFICIENT); per tower its modulus, root of unity and
 *             the byte offset of its N raw words inside the buffer (any output array may be NULL).
 *      write: the byte-identical archive for [n_towers][N] host words (what the reference's Deserialize expects).
 *      upload: scan + one asynchronous host-to-device copy per tower, archive -> dst_d [n_towers][N] (no host DCRTPoly).
 *              The archive is UNTRUSTED: it is rejected unless its ring dimension equals the target ring's, it has at
 *              most max_towers towers and tower t's modulus equals the ring's modulus mod_idx_h[t] (t when NULL), so a
 *              slab sized for the ring can never be overrun; check_residues != 0 additionally scans the payload for
 *              residues >= modulus (the kernels assume canonical inputs).                                            */
int ofhe_cereal_dcrtpoly_scan(const void* buf_h, size_t len, uint32_t max_towers, uint32_t* shape_out,
                              uint64_t* moduli_out, uint64_t* roots_out, size_t* offsets_out);
int ofhe_cereal_dcrtpoly_write(uint32_t N, uint32_t n_towers, const uint64_t* moduli_h, const uint64_t* roots_h,
                               uint32_t format, const uint64_t* data_h, void* out_h, size_t cap, size_t* written_out);
int ofhe_dcrtpoly_upload_cereal(ofhe_ring* ring, const uint32_t* mod_idx_h, const void* buf_h, size_t len,
                                uint32_t max_towers, int check_residues, uint64_t* dst_d, uint32_t* shape_out,
                                void* stream);

/* fills dst_d [n_limbs][N] with uniform residues in [0, q_l) from a counter-based generator (synthetic inputs) */
int ofhe_fill_uniform(ofhe_ring* ring, uint64_t* dst_d, uint32_t n_limbs, const uint32_t* mod_idx_h, uint64_t seed,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFHE_B200_H */
