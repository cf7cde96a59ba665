/* ORACLE — TEST INFRASTRUCTURE ONLY (see fr.h).
 * Unkeyed BLAKE2b-512 per RFC 7693, restating the `blake2 0.10` crate's `Blake2b512` that
 * /root/reference/src/rng.rs:5,22-25 uses (crate not vendored; algorithm is the published RFC).
 * Incremental, with a by-value clonable state (rng.rs:62 clones the running digest).
 */
#ifndef ORACLE_BLAKE2B_H
#define ORACLE_BLAKE2B_H
#include <stddef.h>
#include <stdint.h>

typedef struct {
    uint64_t h[8];
    uint64_t t[2];
    uint8_t buf[128];
    size_t buflen;
} blake2b_state;

void blake2b_init(blake2b_state *s);                                 /* Blake2b512::new() */
void blake2b_update(blake2b_state *s, const void *in, size_t inlen); /* Digest::update     */
void blake2b_final(const blake2b_state *s, uint8_t out[64]);         /* clone().finalize() — does not disturb *s */
#endif
