// owshen_b200/csrc/host_math.hpp -- small host-side pieces the product needs around the kernels:
// Keccak-256 (to derive the circomlib MiMC7 round constants exactly as published: c_0 = 0,
// c_i = keccak256^(i+1)("mimc") mod r) and a few Fr helpers.  Written independently of oracle/.
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>
#include "fp.cuh"

namespace og {

inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

inline void keccak_f1600(uint64_t s[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    static const int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
        }
        uint64_t t = s[1];
        for (int i = 0; i < 24; i++) {
            int j = PILN[i];
            uint64_t b = s[j];
            s[j] = rotl64(t, ROTC[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            uint64_t row[5];
            for (int i = 0; i < 5; i++) row[i] = s[j + i];
            for (int i = 0; i < 5; i++) s[j + i] = row[i] ^ ((~row[(i + 1) % 5]) & row[(i + 2) % 5]);
        }
        s[0] ^= RC[round];
    }
}

// Keccak-256 with the original 0x01 padding (Ethereum), single-call, any length
inline void keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    const size_t rate = 136;
    uint64_t s[25];
    memset(s, 0, sizeof(s));
    std::vector<uint8_t> msg(data, data + len);
    msg.push_back(0x01);
    while (msg.size() % rate) msg.push_back(0x00);
    msg.back() |= 0x80;
    for (size_t off = 0; off < msg.size(); off += rate) {
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w = 0;
            for (int b = 7; b >= 0; b--) w = (w << 8) | msg[off + 8 * i + b];
            s[i] ^= w;
        }
        keccak_f1600(s);
    }
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(s[i] >> (8 * b));
}

// big-endian 32-byte integer reduced mod r, returned in Montgomery form
inline Fr fr_from_be_bytes_reduce(const uint8_t be[32]) {
    // value = hi * 2^128 + lo with hi, lo < 2^128 < r
    uint32_t hi[8] = {0}, lo[8] = {0};
    for (int i = 0; i < 16; i++) {
        hi[(15 - i) / 4] |= (uint32_t)be[i] << (8 * ((15 - i) % 4));
        lo[(15 - i) / 4] |= (uint32_t)be[16 + i] << (8 * ((15 - i) % 4));
    }
    uint32_t two128[8] = {0, 0, 0, 0, 1, 0, 0, 0};
    return Fr::from_canonical(hi) * Fr::from_canonical(two128) + Fr::from_canonical(lo);
}

constexpr int MIMC_ROUNDS = 91;

inline void mimc7_round_constants(Fr out[MIMC_ROUNDS]) {
    uint8_t c[32];
    const uint8_t seed[4] = {'m', 'i', 'm', 'c'};
    keccak256(seed, 4, c);
    out[0] = Fr::zero();
    for (int i = 1; i < MIMC_ROUNDS; i++) {
        uint8_t n[32];
        keccak256(c, 32, n);
        memcpy(c, n, 32);
        out[i] = fr_from_be_bytes_reduce(c);
    }
}

}  // namespace og
