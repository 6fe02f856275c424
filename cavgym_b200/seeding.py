"""Seed -> numpy RandomState, compatible with what the reference obtains from `gym.utils.seeding.np_random`
(gym==0.17.2, used at config.py:275): the seed's decimal string is hashed with SHA-512, the first 8 bytes are read as
little-endian 32-bit words, and those words seed numpy's legacy MT19937 RandomState.  Same seed, same stream, so
single-environment runs reproduce the reference's spawns and epsilon draws."""
import hashlib
import os
import struct

import numpy as np


def _words(seed):
    digest = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    value = sum(word << (32 * i) for i, word in enumerate(struct.unpack("<2I", digest)))
    words = []
    while value > 0:
        value, word = divmod(value, 2 ** 32)
        words.append(word)
    return words or [0]


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, int) and seed >= 0):
        raise ValueError(f"Seed must be a non-negative integer or omitted, not {seed}")
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    seed %= 2 ** 64
    rng = np.random.RandomState()
    rng.seed(_words(seed))
    return rng, seed
