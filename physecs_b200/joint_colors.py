"""Greedy joint-graph colouring, exactly as the reference does it when a joint is added
(reference src/Physecs.cpp:690-710, JointGraph include/Physecs/Physecs.h:137-152).

Each entity carries an 8-bit set of colours in use; a new joint takes the lowest colour free on both entities.  The
union of two uint8 promotes to int, so bit 8 is always free: when all 8 colours are taken the joint lands in the
sequential overflow bucket (index 8) and consumes no colour bit (quirk Q21).  Static entities consume colours too.
"""
from __future__ import annotations

import numpy as np


def color_joints(entity_pairs):
    bits = {}
    out = np.zeros(len(entity_pairs), np.int32)
    for k, (e0, e1) in enumerate(entity_pairs):
        c0, c1 = bits.get(e0, 0), bits.get(e1, 0)
        free = ~(c0 | c1)
        i = (free & -free).bit_length() - 1   # count trailing zeros of ~union (bit 8 is always free)
        if i < 8:
            bits[e0] = c0 | (1 << i)
            bits[e1] = bits.get(e1, 0) | (1 << i)
        out[k] = i
    return out
