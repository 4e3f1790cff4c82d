#!/usr/bin/env python3
"""The carry-less-multiply folding constants of tools/rg_crc32.hpp, derived from the gzip polynomial (no table copied
from anywhere): k = bitreflect(x^n mod P) << 1; mu = bitreflect(floor(x^64 / P)); P' = bitreflect(P)."""
P = 0x104C11DB7


def xpow_mod(n):
    r = 1
    for _ in range(n):
        r <<= 1
        if (r >> 32) & 1:
            r ^= P
    return r


def reflect(v, bits):
    return int(bin(v)[2:].zfill(bits)[::-1], 2)


def quotient(n):
    q = r = 0
    for i in range(n, -1, -1):
        r = (r << 1) | (1 if i == n else 0)
        q <<= 1
        if (r >> 32) & 1:
            r ^= P
            q |= 1
    return q


if __name__ == "__main__":
    for name, n in (("k1", 4 * 128 + 32), ("k2", 4 * 128 - 32), ("k3", 128 + 32), ("k4", 128 - 32), ("k5", 64)):
        print(f"{name} = x^{n} mod P reflected << 1 = {reflect(xpow_mod(n), 32) << 1:#x}")
    print(f"mu = {reflect(quotient(64), 33):#x}   P' = {reflect(P, 33):#x}")
