// Test-only shim: the transcript's SHA-256 (algoplonk_b200/csrc/sha256.hpp) through a C ABI.  `split` feeds the
// message in two update() calls at that offset, like the prover does when it binds several fields.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../algoplonk_b200/csrc/sha256.hpp"
extern "C" void ht_sha256(const uint8_t* msg, uint64_t len, uint64_t split, uint8_t* out) {
    b2p::Sha256 h;
    if (split > len) split = len;
    h.update(msg, split);
    h.update(msg + split, len - split);
    h.final(out);
    // reset() gives a fresh state: hashing again must reproduce the digest
    uint8_t again[32];
    h.reset();
    h.update(msg, len);
    h.final(again);
    if (memcmp(out, again, 32) != 0) memset(out, 0, 32);
}
