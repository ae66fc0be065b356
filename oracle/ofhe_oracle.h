/* oracle/ofhe_oracle.h -- TEST INFRASTRUCTURE ONLY (see ofhe_oracle.c).
 *
 * CPU restatement, in plain C, of the reference's RNS/NTT + hybrid key-switch
 * hot path.  Never linked into, imported by or
 * executed from the product path.  All polynomials are row-major [limbs x N]
 * uint64 slabs; EVALUATION format = the reference's bit-reversed order.
 */
#ifndef OFHE_ORACLE_H
#define OFHE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned __int128 orc_u128;

/* a1-a3 scalar arithmetic */
uint64_t orc_shoup_precon(uint64_t w, uint64_t q);
uint64_t orc_mulmod_shoup(uint64_t x, uint64_t w, uint64_t wp, uint64_t q);
uint64_t orc_barrett_mu(uint64_t q);
uint64_t orc_mulmod_barrett(uint64_t a, uint64_t b, uint64_t q, uint64_t mu);
void orc_barrett128_mu(uint64_t q, uint64_t* mu_hi, uint64_t* mu_lo);
uint64_t orc_barrett128(uint64_t a_hi, uint64_t a_lo, uint64_t q, uint64_t mu_hi, uint64_t mu_lo);
uint64_t orc_powmod(uint64_t b, uint64_t e, uint64_t q);
uint64_t orc_invmod(uint64_t a, uint64_t q);

/* a8(host) number theory */
int orc_is_prime(uint64_t n);
uint64_t orc_first_prime(uint32_t bits, uint64_t m);
uint64_t orc_last_prime(uint32_t bits, uint64_t m);
uint64_t orc_next_prime(uint64_t q, uint64_t m);
uint64_t orc_previous_prime(uint64_t q, uint64_t m);
uint64_t orc_root_of_unity(uint64_t m, uint64_t q);

/* a5-a7 NTT */
void orc_twiddles(uint64_t q, uint64_t root, uint32_t N, uint64_t* T, uint64_t* Tp, uint64_t* Ti, uint64_t* Tip,
                  uint64_t* nInv, uint64_t* nInvPrecon);
void orc_ntt_fwd(uint64_t q, uint64_t root, uint32_t N, uint64_t* x);
void orc_ntt_inv(uint64_t q, uint64_t root, uint32_t N, uint64_t* x);

/* a4/a9/a15 limb-wise ops; op: 0 add 1 sub 2 mul 3 neg */
void orc_limb_binop(uint64_t q, uint32_t N, int op, const uint64_t* a, const uint64_t* b, uint64_t* out);
void orc_limb_mul_scalar(uint64_t q, uint32_t N, const uint64_t* a, uint64_t s, uint64_t* out);
void orc_switch_modulus(uint32_t N, uint64_t q_from, uint64_t q_to, uint64_t* x);
void orc_automorphism_map(uint32_t N, uint32_t k, uint32_t* map);
void orc_automorphism_eval(uint32_t N, const uint32_t* map, const uint64_t* in, uint64_t* out);
void orc_automorphism_coeff(uint32_t N, uint32_t k, uint64_t q, const uint64_t* in, uint64_t* out);

/* a10 fast base conversion: in [sizeQ][N] COEFF -> out [sizeP][N] COEFF; QHatModp is [sizeQ][sizeP] */
void orc_approx_switch_crt_basis(uint32_t N, uint32_t sizeQ, const uint64_t* q, const uint64_t* QHatInvModq,
                                 uint32_t sizeP, const uint64_t* p, const uint64_t* QHatModp, const uint64_t* in,
                                 uint64_t* out);

/* a12 exact base conversion (QHatModp is [sizeP][sizeQ], alphaQModp [sizeQ+1][sizeP]) and a13 scale-and-round */
void orc_switch_crt_basis(uint32_t N, uint32_t sizeQ, const uint64_t* q, const uint64_t* QHatInvModq, uint32_t sizeP,
                          const uint64_t* p, const uint64_t* QHatModp, const uint64_t* alphaQModp, const double* qInv,
                          const uint64_t* in, uint64_t* out);
void orc_scale_and_round(uint32_t N, uint32_t sizeI, uint32_t sizeO, const uint64_t* o, int out_limbs_first,
                         const uint64_t* tbl, const double* frac, const uint64_t* in, uint64_t* out);

/* a20 CKKS / HYBRID tables */
typedef struct orc_ckks orc_ckks;
/* scal_tech: 0 FIXEDMANUAL, 1 FIXEDAUTO, 2 FLEXIBLEAUTO, 3 FLEXIBLEAUTOEXT (chain and tables; sizeQ = depth + 2) */
orc_ckks* orc_ckks_create(uint32_t N, uint32_t depth, uint32_t scale_bits, uint32_t first_bits, uint32_t dnum,
                          int scal_tech);
/* hybrid key-switch context over a caller-supplied chain (e.g. BFV's Q then P) */
orc_ckks* orc_hks_create_from_moduli(uint32_t N, uint32_t sizeQ, uint32_t sizeP, uint32_t dnum, const uint64_t* moduli,
                                     const uint64_t* roots);
void orc_ckks_destroy(orc_ckks* c);
void orc_ckks_shape(const orc_ckks* c, uint32_t* shape /* N,sizeQ,sizeP,dnum,alpha */);
void orc_ckks_moduli(const orc_ckks* c, uint64_t* moduli, uint64_t* roots);
void orc_ckks_PInvModq(const orc_ckks* c, uint64_t* out);
void orc_ckks_PModq(const orc_ckks* c, uint64_t* out);
void orc_ckks_PHatInvModp(const orc_ckks* c, uint64_t* out);
void orc_ckks_PHatModq(const orc_ckks* c, uint64_t* out /* [sizeP][sizeQ] */);
void orc_ckks_PartQlHatInvModq(const orc_ckks* c, uint32_t part, uint32_t digit_size, uint64_t* out);
uint32_t orc_ckks_PartQlHatModp(const orc_ckks* c, uint32_t sizeQl, uint32_t part, uint64_t* out,
                                uint64_t* compl_moduli);
void orc_ckks_rescale_tables(const orc_ckks* c, uint32_t sizeQl, uint64_t* QlQlInvModqlDivqlModq, uint64_t* qlInvModq);

/* a16-a19, a14, a11 on caller-supplied residues; evk a/b are [dnum][sizeQ+sizeP][N] */
int orc_ckks_modup(const orc_ckks* c, uint32_t sizeQl, const uint64_t* poly, uint64_t* digits);
void orc_ckks_inner_product(const orc_ckks* c, uint32_t sizeQl, uint32_t n_digits, const uint64_t* digits,
                            const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* e0, uint64_t* e1);
void orc_ckks_approx_mod_down(const orc_ckks* c, uint32_t sizeQl, const uint64_t* in, uint64_t* out);
void orc_ckks_keyswitch_core(const orc_ckks* c, uint32_t sizeQl, const uint64_t* poly, const uint64_t* evk_a,
                             const uint64_t* evk_b, uint64_t* out0, uint64_t* out1);
void orc_ckks_evalmult(const orc_ckks* c, uint32_t sizeQl, const uint64_t* ct1, const uint64_t* ct2,
                       const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* out);
void orc_ckks_rescale_poly(const orc_ckks* c, uint32_t sizeQl, const uint64_t* x, uint64_t* out);
/* EvalAutomorphism (a15 + a18): rotate ct by automorphism index k with key (evk_a, evk_b) */
void orc_ckks_automorphism_ct(const orc_ckks* c, uint32_t sizeQl, uint32_t k, const uint64_t* ct,
                              const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* out);

/* hoisted rotations: digits from orc_ckks_modup(c1); *_ext results are [2][sizeQl+sizeP][N] */
void orc_ckks_fast_rotation(const orc_ckks* c, uint32_t sizeQl, uint32_t k, const uint64_t* ct, const uint64_t* digits,
                            const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* out);
void orc_ckks_fast_rotation_ext(const orc_ckks* c, uint32_t sizeQl, uint32_t k, int add_first, const uint64_t* ct,
                                const uint64_t* digits, const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* out);
void orc_ckks_keyswitch_ext(const orc_ckks* c, uint32_t sizeQl, int add_first, const uint64_t* ct, uint64_t* out);
void orc_ckks_keyswitch_down(const orc_ckks* c, uint32_t sizeQl, const uint64_t* ct_ext, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif
