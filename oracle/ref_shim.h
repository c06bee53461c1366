/* TEST INFRASTRUCTURE ONLY -- forced into every reference translation unit with
 * `g++ -include oracle/ref_shim.h` when oracle/_ref/libbess_ref.so is built.
 *
 * The reference seeds its CV fold shuffle from std::random_device
 * (/root/reference/src/Metric.h:57-58), so two runs of the unmodified reference
 * disagree with each other under CV.  Parity on "CV-chosen s" is only definable
 * with the shuffle pinned.  Rather than copying/patching reference sources, this
 * shim renames the identifier `random_device` (after <random> has been fully
 * parsed) to a deterministic stand-in that returns BESS_CV_SEED (default 123).
 * Nothing else in the reference changes: same mt19937, same std::shuffle, same
 * fold chunking.
 */
#ifndef BESS_REF_SHIM_H
#define BESS_REF_SHIM_H
#include <random>
#include <algorithm>
#include <cstdlib>
namespace std {
struct bess_seeded_random_device {
    typedef unsigned int result_type;
    bess_seeded_random_device() {}
    result_type operator()() const {
        const char *s = std::getenv("BESS_CV_SEED");
        return s ? static_cast<result_type>(std::strtoul(s, nullptr, 10)) : 123u;
    }
};
}  // namespace std
#define random_device bess_seeded_random_device
#endif
