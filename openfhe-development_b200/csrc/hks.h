// hks.h -- hybrid key-switch context (device tables per level) and evaluation keys.  See hks.cu.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "internal.h"

namespace ofhe {

struct Hks;

// evaluation key: a_j, b_j for j < dnum, each over the full QP basis, EVALUATION format, device resident
// (EvalKeyRelinImpl::GetAVector/GetBVector, pke/include/key/evalkeyrelin.h:141-171)
struct EvalKey {
    Hks* hks      = nullptr;
    uint32_t dnum = 0;
    uint64_t* a_d = nullptr;  // [dnum][sizeQ+sizeP][N]
    uint64_t* b_d = nullptr;
    ~EvalKey();
};

struct Hks {
    Ring* ring;
    uint32_t sizeQ, sizeP, dnum, alpha;
    std::vector<uint64_t> PModq, PInvModq;

    struct Level {
        uint32_t n_digits = 0;
        LimbMap map_Ql, map_QlP;
        LimbMap map_modup_ntt;  // every converted row of every digit slab
        LimbMap map_P2, map_Ql2;
        PosMap gather_P2;  // output rows of the ModDown inverse NTT (p_tmp [2][sizeP][N]) in map_P2's launch order
        std::vector<std::unique_ptr<BConv>> modup;
        std::unique_ptr<BConv> moddown;
        uint64_t* fold_modup_d = nullptr;    // [sizeQl][4]: INTT constants with PartQlHatInvModq folded in
        uint64_t* fold_moddown_d = nullptr;  // [2*sizeP][4]: INTT constants with PHatInvModp folded in
        ~Level();
        ScalarVec pinv2;
        uint64_t* pinv2_d = nullptr;  // [2*sizeQl][2] in map_Ql2's launch order: P^-1 mod q_i and its Shoup constant
        PosMap tail_erow, tail_drow;  // rows of e01 [2][sizeQl+sizeP][N] / out01 [2][sizeQl][N] per launch limb of map_Ql2
        ScalarVec rescaleA, rescaleB;
    };
    std::map<uint32_t, std::unique_ptr<Level>> levels;
    std::mutex mu;

    Hks(Ring* r, uint32_t sizeQ, uint32_t sizeP, uint32_t dnum);
    ~Hks();
    Level& level(uint32_t sizeQl);
    size_t workspace_words(uint32_t sizeQl) const;
    size_t max_workspace_words(uint32_t sizeQl) const;  // the largest arena any entry point asks for at this level

    // c [sizeQl][N] EVALUATION -> digits [n_digits][sizeQl+sizeP][N] EVALUATION; coeff_tmp [sizeQl][N] scratch
    // own_rows = false: the digit slabs do NOT receive the digit's own (original) rows -- only for a caller that hands
    // the switched polynomial to inner_product() as c_own
    void modup(uint32_t sizeQl, const uint64_t* c, uint64_t* digits, uint64_t* coeff_tmp, cudaStream_t st,
               bool own_rows = true);
    // e01 [2][sizeQl+sizeP][N]
    void inner_product(uint32_t sizeQl, uint32_t n_digits, const uint64_t* digits, const EvalKey& evk, uint64_t* e01,
                       cudaStream_t st, const uint64_t* c_own = nullptr);
    // e01 [2][sizeQl+sizeP][N] -> out01 [2][sizeQl][N] (+ addend01 if given; may alias out01)
    void mod_down2(uint32_t sizeQl, const uint64_t* e01, uint64_t* out01, const uint64_t* addend01, uint64_t* p_tmp,
                   uint64_t* conv_tmp, cudaStream_t st);
    void keyswitch_core(uint32_t sizeQl, const uint64_t* c, const EvalKey& evk, uint64_t* out01,
                        const uint64_t* addend01, uint64_t* ws, cudaStream_t st);
    void evalmult(uint32_t sizeQl, const uint64_t* ct1, const uint64_t* ct2, const EvalKey& evk, uint64_t* out,
                  cudaStream_t st);
    void rescale(uint32_t sizeQl, const uint64_t* in, uint64_t* out, uint32_t n_polys, cudaStream_t st);
    void rotate(uint32_t sizeQl, uint32_t k, const uint64_t* ct, const EvalKey& evk, uint64_t* out, cudaStream_t st);
    // hoisted rotations: digits = modup(c1) computed ONCE by the caller and shared by several rotations
    void fast_rotation(uint32_t sizeQl, uint32_t k, const uint64_t* ct, const uint64_t* digits, const EvalKey& evk,
                       uint64_t* out, cudaStream_t st);
    // result stays in the extended basis QlP: out_ext [2][sizeQl+sizeP][N]
    void fast_rotation_ext(uint32_t sizeQl, uint32_t k, bool add_first, const uint64_t* ct, const uint64_t* digits,
                           const EvalKey& evk, uint64_t* out_ext, cudaStream_t st);
    void keyswitch_ext(uint32_t sizeQl, bool add_first, const uint64_t* ct, uint64_t* out_ext, cudaStream_t st);
    // One level of the double-hoisted baby-step / giant-step linear transform: the loop body of
    // FHECKKSRNS::EvalCoeffsToSlots / EvalSlotsToCoeffs (pke/lib/scheme/ckksrns/ckksrns-fhe.cpp:1934-1987, 2094-2139)
    // with every intermediate device resident.  See hks.cu.
    size_t bsgs_workspace_words(uint32_t sizeQl, uint32_t g) const;
    void bsgs_level(uint32_t sizeQl, uint32_t g, uint32_t b, const uint32_t* k_in, const uint32_t* k_out,
                    const EvalKey* const* evk_in, const EvalKey* const* evk_out, const uint64_t* A, uint32_t skip,
                    const uint64_t* ct, uint64_t* out, uint64_t* ws, cudaStream_t st);
    ScalarVec pmodq_sv;  // P mod q_i with Shoup constants (GetPModq, rns-cryptoparameters.h)
};

}  // namespace ofhe
