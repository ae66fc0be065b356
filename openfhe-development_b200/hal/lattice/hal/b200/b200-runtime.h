// b200-runtime.h -- process-wide state of the B200 lattice backend: the device context, the rings (device-resident
// twiddle tables) of every CRT basis the host code has used, cached base-conversion tables, and ONE CUDA stream plus
// grow-only device slabs PER CALLING THREAD.
//
// Everything here goes through the C ABI of include/ofhe_b200.h; this header is compiled by the reference's own
// translation units (g++, no CUDA headers) when they are built with the overlay include directory
// openfhe-development_b200/hal first on the include path (oracle/Makefile.hal).
//
// Threading contract (SURVEY.md section 8 b): the reference calls DCRTPoly methods from inside its own
// `#pragma omp parallel for` regions (src/pke/lib/scheme/ckksrns/ckksrns-fhe.cpp:1943,
// src/pke/lib/keyswitch/keyswitch-hybrid.cpp:423).  Each thread therefore owns a stream and its slabs
// (thread_local); the registries are guarded by a mutex and hand out immutable objects.
#ifndef LBCRYPTO_INC_LATTICE_HAL_B200_RUNTIME_H
#define LBCRYPTO_INC_LATTICE_HAL_B200_RUNTIME_H

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "ofhe_b200.h"
#include "utils/exception.h"

namespace lbcrypto {
namespace b200 {

inline void Check(int rc, const char* what) {
    if (rc != 0)
        OPENFHE_THROW(std::string("b200 backend: ") + what + ": " + ofhe_last_error());
}

// statistics the tests read back to prove that the device path (and not the host mirror) did the work
struct Counters {
    std::atomic<uint64_t> ntt{0}, ew{0}, automorph{0}, bconv{0}, rescale{0}, moddown{0}, modup{0}, h2d_bytes{0},
        d2h_bytes{0};
};

class Runtime {
public:
    // The runtime is IMMORTAL (never destroyed): polynomials held by objects with static storage duration -- the
    // reference keeps every CryptoContext and every evaluation key in static containers -- release their slabs during
    // static destruction, i.e. possibly after a function-local static would already be gone; returning a slab to the
    // pool of a destroyed runtime corrupted the heap at exit (tools/hal_bootstrap.cpp, which does not call
    // ReleaseAllContexts as the gtest suites do).  The tallies are printed by an atexit handler instead.
    static Runtime& Get() {
        static Runtime* r = new Runtime();
        return *r;
    }
    ofhe_ctx* ctx = nullptr;

    // ---- rings: one per ordered list of (modulus, root) pairs over a ring dimension -------------------------
    ofhe_ring* Ring(uint32_t N, const std::vector<uint64_t>& moduli, const std::vector<uint64_t>& roots) {
        std::string key(reinterpret_cast<const char*>(&N), sizeof N);
        key.append(reinterpret_cast<const char*>(moduli.data()), moduli.size() * 8);
        std::lock_guard<std::mutex> g(mu_);
        auto it = rings_.find(key);
        if (it != rings_.end())
            return it->second;
        ofhe_ring* r = nullptr;
        Check(ofhe_ring_create(ctx, N, (uint32_t)moduli.size(), moduli.data(), roots.data(), &r), "ofhe_ring_create");
        rings_.emplace(std::move(key), r);
        return r;
    }

    // ---- base-conversion tables, keyed by the bytes of the tables themselves ----------------------------------
    // (the reference passes its precomputed vectors by const reference on every call; equal bytes = equal table)
    ofhe_bconv* BConv(const std::string& key, const std::function<ofhe_bconv*()>& make) {
        std::lock_guard<std::mutex> g(mu_);
        auto it = bconvs_.find(key);
        if (it != bconvs_.end())
            return it->second;
        ofhe_bconv* t = make();
        bconvs_.emplace(key, t);
        return t;
    }

    // ---- per-thread stream and slabs --------------------------------------------------------------------------
    struct Local {
        void* stream = nullptr;
        uint64_t* slab[4] = {nullptr, nullptr, nullptr, nullptr};
        size_t words[4]   = {0, 0, 0, 0};
        ~Local() {
            // device memory and the stream die with the context at process exit; threads of an OpenMP pool outlive
            // every DCRTPoly, so nothing is released here
        }
    };
    Local& Mine() {
        thread_local Local l;
        if (!l.stream)
            Check(ofhe_stream_create(ctx, &l.stream), "ofhe_stream_create");
        return l;
    }
    uint64_t* Slab(Local& l, int which, size_t words) {
        if (l.words[which] < words) {
            if (l.slab[which]) {
                Check(ofhe_stream_sync(ctx, l.stream), "ofhe_stream_sync");
                Check(ofhe_free(ctx, l.slab[which]), "ofhe_free");
            }
            void* p = nullptr;
            Check(ofhe_malloc(ctx, words * sizeof(uint64_t), &p), "ofhe_malloc");
            l.slab[which]  = static_cast<uint64_t*>(p);
            l.words[which] = words;
        }
        return l.slab[which];
    }
    // ---- device slabs of resident polynomials: size-keyed free lists (a polynomial's slab is released when its last
    // owner dies; every device method synchronises its stream before it returns, so a released slab has no work in
    // flight and can be handed to any thread).  Slabs are IMMUTABLE once published to a polynomial: copies share them.
    std::shared_ptr<uint64_t> Alloc(size_t words) {
        uint64_t* p = nullptr;
        {
            std::lock_guard<std::mutex> g(pool_mu_);
            auto it = pool_.find(words);
            if (it != pool_.end() && !it->second.empty()) {
                p = it->second.back();
                it->second.pop_back();
                pooled_words_ -= words;
            }
        }
        if (!p) {
            void* v = nullptr;
            Check(ofhe_malloc(ctx, words * sizeof(uint64_t), &v), "ofhe_malloc");
            p = static_cast<uint64_t*>(v);
        }
        return std::shared_ptr<uint64_t>(p, [this, words](uint64_t* q) {
            std::lock_guard<std::mutex> g(pool_mu_);
            if (pooled_words_ + words > kPoolCapWords) {
                ofhe_free(ctx, q);  // keep the cache bounded (24 GiB)
                return;
            }
            pool_[words].push_back(q);
            pooled_words_ += words;
        });
    }
    // one of 256 mutexes by object address: const accessors of a polynomial may be called from several threads
    std::mutex& Stripe(const void* obj) {
        return stripes_[(reinterpret_cast<uintptr_t>(obj) >> 6) & 255u];
    }

    // process-wide tallies of the device work issued through this backend; OFHE_B200_HAL_STATS=1 prints them at exit
    // (the GPU tests read that line to prove the reference's suites ran on the device path, not on the mirror)
    Counters n;
    // primitive 2N-th root of unity of q when the caller's parameter object carries none (RootOfUnity, nbtheory-impl.h:183-231)
    static uint64_t RootOr(uint64_t root, uint32_t N, uint64_t q) {
        return root ? root : ofhe_root_of_unity(2ull * N, q);
    }

private:
    Runtime() {
        const char* dev = std::getenv("OFHE_B200_DEVICE");
        Check(ofhe_ctx_create(dev ? std::atoi(dev) : 0, &ctx), "ofhe_ctx_create (no CUDA device: this backend has no CPU path)");
        if (std::getenv("OFHE_B200_HAL_STATS"))
            std::atexit(&Runtime::PrintStats);
    }
    static void PrintStats() {
        const Counters& n = Get().n;
        std::fprintf(stderr,
                     "[ofhe_b200 hal] device work: limb_ntt=%llu ew_limbs=%llu automorph_limbs=%llu bconv=%llu "
                     "rescale=%llu modup=%llu moddown=%llu h2d_bytes=%llu d2h_bytes=%llu\n",
                     (unsigned long long)n.ntt, (unsigned long long)n.ew, (unsigned long long)n.automorph,
                     (unsigned long long)n.bconv, (unsigned long long)n.rescale, (unsigned long long)n.modup,
                     (unsigned long long)n.moddown, (unsigned long long)n.h2d_bytes, (unsigned long long)n.d2h_bytes);
    }
    ~Runtime() {}  // never runs, see Get(): process exit tears the CUDA context down
    static constexpr size_t kPoolCapWords = (size_t)3 << 30;  // 24 GiB of cached slabs
    std::mutex pool_mu_;
    std::map<size_t, std::vector<uint64_t*>> pool_;
    size_t pooled_words_ = 0;
    std::mutex stripes_[256];
    std::mutex mu_;
    std::map<std::string, ofhe_ring*> rings_;
    std::map<std::string, ofhe_bconv*> bconvs_;
};

}  // namespace b200
}  // namespace lbcrypto

#endif
