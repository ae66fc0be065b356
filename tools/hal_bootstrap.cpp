// hal_bootstrap.cpp -- BASELINE config 4 through the reference's own public API: CKKS bootstrapping with the parameters of
// the reference's simple-ckks-bootstrapping example (src/pke/examples/simple-ckks-bootstrapping.cpp:63-156: UNIFORM_TERNARY
// secret, FLEXIBLEAUTO, scaling 59 / first modulus 60 bits, level budget {4, 4}, 10 levels after bootstrapping) at a ring
// dimension given on the command line (2^17 for config 4).
//
// The SAME source is linked twice by oracle/Makefile.hal:
//   oracle/_ref/hal/bootstrap_b200  against libOPENFHEpke_b200.so  (DCRTPoly = b200::DCRTPolyB200, KeySwitchHYBRID rebound:
//                                   every polynomial operation of EvalBootstrap runs on the GPU)
//   oracle/_ref/hal/bootstrap_ref   against the unmodified reference (oracle/_ref/libOPENFHEpke.so): the CPU baseline
// and prints one JSON line: seconds per EvalBootstrap (after one warm-up call), precision of the result against the
// clear-text input, levels, limb counts.  This is measurement / demonstration tooling, not a kernel: nothing here computes
// on residues.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "openfhe.h"

using namespace lbcrypto;

int main(int argc, char** argv) {
    uint32_t logn = 12, slots = 8, reps = 1, warm = 1;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "--logn"))
            logn = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--slots"))
            slots = std::atoi(argv[i + 1]);  // 0 = full packing (ring dimension / 2)
        else if (!std::strcmp(argv[i], "--reps"))
            reps = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--warm"))
            warm = std::atoi(argv[i + 1]);  // untimed EvalBootstrap calls first (device tables, key uploads)
    }
    CCParams<CryptoContextCKKSRNS> parameters;
    const SecretKeyDist secretKeyDist = UNIFORM_TERNARY;
    parameters.SetSecretKeyDist(secretKeyDist);
    parameters.SetSecurityLevel(HEStd_NotSet);
    parameters.SetRingDim(1u << logn);
    parameters.SetScalingModSize(59);
    parameters.SetScalingTechnique(FLEXIBLEAUTO);
    parameters.SetFirstModSize(60);
    const std::vector<uint32_t> levelBudget = {4, 4};
    const uint32_t levelsAfter               = 10;
    const uint32_t depth = levelsAfter + FHECKKSRNS::GetBootstrapDepth(levelBudget, secretKeyDist);
    parameters.SetMultiplicativeDepth(depth);

    auto t0 = std::chrono::steady_clock::now();
    CryptoContext<DCRTPoly> cc = GenCryptoContext(parameters);
    cc->Enable(PKE);
    cc->Enable(KEYSWITCH);
    cc->Enable(LEVELEDSHE);
    cc->Enable(ADVANCEDSHE);
    cc->Enable(FHE);
    const uint32_t ringDim  = cc->GetRingDimension();
    const uint32_t numSlots = slots ? slots : ringDim / 2;
    cc->EvalBootstrapSetup(levelBudget, {0, 0}, numSlots);
    auto keyPair = cc->KeyGen();
    cc->EvalMultKeyGen(keyPair.secretKey);
    cc->EvalBootstrapKeyGen(keyPair.secretKey, numSlots);
    const double setup_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::vector<double> x(std::min<uint32_t>(numSlots, 8));
    const double vals[8] = {0.25, 0.5, 0.75, 1.0, 2.0, 3.0, 4.0, 5.0};
    for (size_t i = 0; i < x.size(); ++i)
        x[i] = vals[i];
    Plaintext ptxt = cc->MakeCKKSPackedPlaintext(x, 1, depth - 1, nullptr, numSlots);  // a depleted ciphertext
    ptxt->SetLength(x.size());
    auto ciph = cc->Encrypt(keyPair.publicKey, ptxt);

    Ciphertext<DCRTPoly> after;
    for (uint32_t r = 0; r < warm; ++r)
        after = cc->EvalBootstrap(ciph);
    t0 = std::chrono::steady_clock::now();
    for (uint32_t r = 0; r < reps; ++r)
        after = cc->EvalBootstrap(ciph);
    const double boot_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;

    Plaintext result;
    cc->Decrypt(keyPair.secretKey, after, &result);
    result->SetLength(x.size());
    double err = 0;
    const auto& got = result->GetCKKSPackedValue();
    for (size_t i = 0; i < x.size(); ++i)
        err = std::max(err, std::abs(got[i].real() - x[i]));
    const auto& cp   = cc->GetCryptoParameters();
    const auto rns    = std::dynamic_pointer_cast<CryptoParametersRNS>(cp);
    std::printf(
        "{\"workload\": \"CKKS EvalBootstrap, simple-ckks-bootstrapping parameters\", \"backend\": \"%s\", \"logN\": %u, "
        "\"slots\": %u, \"depth\": %u, \"sizeQ\": %zu, \"sizeP\": %zu, \"levels_remaining_after\": %u, "
        "\"setup_and_keygen_s\": %.3f, \"s_per_bootstrap\": %.4f, \"bootstraps_per_s\": %.4f, \"max_abs_error\": %.3e, "
        "\"precision_bits\": %.1f, \"threads\": %d}\n",
#ifdef OFHE_B200_HAL
        "b200 (DCRTPoly = b200::DCRTPolyB200, KeySwitchHYBRID rebound)",
#else
        "reference CPU (unmodified)",
#endif
        logn, numSlots, depth, rns->GetElementParams()->GetParams().size(), rns->GetParamsP()->GetParams().size(),
        depth - after->GetLevel() - (after->GetNoiseScaleDeg() - 1), setup_s, boot_s, 1.0 / boot_s, err,
        err > 0 ? -std::log2(err) : 99.0, OpenFHEParallelControls.GetNumThreads());
    return err < 1e-2 ? 0 : 1;
}
