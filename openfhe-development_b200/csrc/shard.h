// shard.h -- limb-sharded CKKS EvalMult / hybrid key switch (one rank's share).  See shard.cu.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "internal.h"

namespace ofhe {

struct Shard {
    Ring* ring;
    uint32_t sizeQ, sizeP, dnum, alpha, rank, world;
    uint32_t slotsP = 0, nOwnQFull = 0;
    std::vector<uint32_t> ownP;  // owned P limb indices (0-based within P), ascending

    uint32_t owner_q(uint32_t i) const { return i % world; }
    uint32_t owner_p(uint32_t j) const { return (j + sizeQ) % world; }

    struct Level {
        uint32_t slotsQ = 0, nOwnQ = 0, n_digits = 0;
        std::vector<uint32_t> ownQ;
        LimbMap map_q, map_qp, map_q2, map_p2_src, map_modup_ntt;
        PosMap gatherq_pos, gatherp_pos;
        std::vector<std::unique_ptr<BConv>> modup;
        std::vector<uint32_t> own_first, own_count;  // owned rows of each digit (contiguous local rows)
        std::unique_ptr<BConv> moddown;
        uint64_t* fold_q_d = nullptr;
        uint64_t* fold_p_d = nullptr;
        ScalarVec pinv;
        // rescale (drop the last limb l = sizeQl-1): the owned limbs below l in local order, their constants
        // -q_l^-1 and q_l^-1 mod q_i (ckksrns-cryptoparameters.cpp:60-83), and where the last limb lives
        LimbMap map_low;
        ScalarVec rescaleA, rescaleB;
        int last_row = -1;  // local row of limb l if this rank owns it
        ~Level();
    };
    std::map<uint32_t, std::unique_ptr<Level>> levels;
    std::mutex mu;
    uint64_t* ws_d = nullptr;  // scratch owned by THIS object, so several shards (pipeline slots) can be in flight
    size_t ws_words = 0;
    uint64_t* scratch(uint32_t sizeQl);

    Shard(Ring* r, uint32_t sizeQ, uint32_t sizeP, uint32_t dnum, uint32_t rank, uint32_t world);
    ~Shard();
    Level& level(uint32_t sizeQl);
    size_t workspace_words(uint32_t sizeQl);
    // phase A: tensor product on the owned limbs (out = c0 | c1), INTT(c2) * QHatInv into this rank's rows of gatherQ
    // `peers` (optional): the same gather buffer of EVERY rank, peer-mapped; the finished rows are then written into all
    // of them by the INTT's last pass (the exchange rides on the kernel's own stores over NVLink, no collective call)
    void evalmult_a(uint32_t sizeQl, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out, uint64_t* gatherQ,
                    cudaStream_t st, const PeerOut* peers = nullptr);
    // phase B (after the all-gather of gatherQ): ModUp into the owned rows, NTT, evk inner product, INTT of the owned
    // P rows * PHatInv into this rank's rows of gatherP
    void evalmult_b(uint32_t sizeQl, const uint64_t* gatherQ, const uint64_t* evk_a, const uint64_t* evk_b,
                    uint64_t* gatherP, cudaStream_t st, const PeerOut* peers = nullptr);
    // phase C (after the all-gather of gatherP): ModDown of the owned Q rows and the final additions into out
    void evalmult_c(uint32_t sizeQl, const uint64_t* gatherP, uint64_t* out, cudaStream_t st);

    // ---- limb-sharded EvalAutomorphism / EvalRotate (base-leveledshe.cpp:382-422: KeySwitchInPlace, then the
    //      automorphism of both polynomials).  ct, tmp, out: [2][nOwnQ][N] local rows.  Sequence: rotate_a, exchange of
    //      gatherQ, evalmult_b with the rotation key's rows, exchange of gatherP, rotate_c.
    // phase A: c1 becomes the polynomial to switch (its INTT * QHatInv goes to gatherQ), tmp = (c0 | 0)
    void rotate_a(uint32_t sizeQl, const uint64_t* ct, uint64_t* tmp, uint64_t* gatherQ, cudaStream_t st,
                  const PeerOut* peers = nullptr);
    // phase C: tmp += ModDown(e0 | e1) on the owned rows, out = automorphism_k(tmp) (limb-local)
    void rotate_c(uint32_t sizeQl, uint32_t k, const uint64_t* gatherP, uint64_t* tmp, uint64_t* out, cudaStream_t st);

    // ---- limb-sharded rescale (DCRTPolyImpl::DropLastElementAndScale, dcrtpoly-impl.h:693-712) of a 2-element
    //      ciphertext: the owner of the last limb sends its coefficient form to everyone (gatherL [2][N], one exchange),
    //      every rank then finishes its own limbs.  out: [2][nOwnQ(sizeQl-1)][N].
    void rescale_a(uint32_t sizeQl, const uint64_t* ct, uint64_t* gatherL, cudaStream_t st, const PeerOut* peers = nullptr);
    void rescale_b(uint32_t sizeQl, const uint64_t* ct, const uint64_t* gatherL, uint64_t* out, cudaStream_t st);
};

}  // namespace ofhe
