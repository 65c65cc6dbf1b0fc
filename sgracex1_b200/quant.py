"""Quantisation constants and host-side quantisers of the SGRACE driver.

Host mirror of demo/sgrace_lib/sgrace.py:53-265 (quantisers) and of the per-bit-width constant
tables in init_SGRACE (sgrace.py:1298-1538 clip ranges, 1645-1839 scales / scale_fea /
internal_quantization / deq_o).  Same names, same arithmetic; `w_qbits` is an argument instead
of the `config` global.
"""
from __future__ import annotations

import numpy as np
import torch


# ---- sgrace.py:53-92 -------------------------------------------------------------------
def quantization(x, s, z, alpha_q, beta_q):
    x_q = np.round(1 / s * x + z, decimals=0)
    return np.clip(x_q, a_min=alpha_q, a_max=beta_q)


def quantization_b(x, s, z, alpha_q, beta_q):
    x_q = (1 / s * x + z)
    x_q[x_q < 0] = -1
    x_q[x_q >= 0] = 1
    return x_q


def quantization_uqbits(x, s, z, qbits):
    return quantization(x, s, z, 0, (2 ** qbits - 1))


def quantization_qbits(x, s, z, qbits):
    if qbits == 1:
        return quantization_b(x, s, z, -1, 1)
    return quantization(x, s, z, (-2 ** (qbits - 1) + 1), (2 ** (qbits - 1) - 1))


# ---- sgrace.py:95-174 ------------------------------------------------------------------
def generate_quantization_constants(alpha, beta, alpha_q, beta_q, w_qbits):
    if w_qbits == 1:
        beta_o, alpha_o = beta_q / (2 ** 2), alpha_q / (2 ** 2)
    else:
        beta_o, alpha_o = beta_q / (2 ** w_qbits), alpha_q / (2 ** w_qbits)
    s_o = (beta - alpha) / (beta_o - alpha_o)
    s = (beta - alpha) / (beta_q - alpha_q)
    z = int((beta * alpha_q - alpha * beta_q) / (beta - alpha))
    return s_o, s, z


def generate_quantization_uqbits_constants(alpha, beta, qbits, w_qbits=None):
    return generate_quantization_constants(alpha, beta, 0, (2 ** qbits - 1), qbits if w_qbits is None else w_qbits)


def generate_quantization_qbits_constants(alpha, beta, qbits, w_qbits=None):
    if qbits == 1:
        alpha_q, beta_q = -1, 1
    else:
        alpha_q, beta_q = (-2 ** (qbits - 1) + 1), (2 ** (qbits - 1) - 1)
    return generate_quantization_constants(alpha, beta, alpha_q, beta_q, qbits if w_qbits is None else w_qbits)


# ---- sgrace.py:177-265 (torch "fake quantisation", the emulation's view of the hardware) ----
def fake_quantization_b(x, s, z, alpha_q, beta_q):
    x_q = (1 / s * x + z)
    x_q[x_q < 0] = -0.5
    x_q[x_q >= 0] = 0.5
    return x_q


def fake_quantization_b2(x, s, z, alpha_q, beta_q):
    x_r = torch.round(1 / s * x + z, decimals=0)
    return torch.clip(x_r, min=alpha_q, max=beta_q) / 2


def fake_quantization(x, s, z, alpha_q, beta_q, w_qbits):
    x_r = torch.round(1 / s * x + z, decimals=0)
    return torch.clip(x_r, min=alpha_q, max=beta_q) / (2 ** (w_qbits - 1))


def quantization_fbits(x, s, z, qbits):
    if qbits == 1:
        return fake_quantization_b(x, s, z, -1, 1)
    return fake_quantization(x, s, z, (-2 ** (qbits - 1) + 1), (2 ** (qbits - 1) - 1), qbits)


def quantization_ufbits(x, s, z, qbits):
    if qbits == 1:
        return fake_quantization_b2(x, s, z, 0, 1)
    return fake_quantization(x, s, z, 0, (2 ** qbits - 1), qbits)


# ---- init_SGRACE constant tables (sgrace.py:1298-1538, 1645-1839) -------------------------
_CLIP = {   # w_qbits: (w_min, w_max, a_min, a_max, f_min, f_max)
    8: (-1.0, 1.0, 0.0, 1.0, 0.0, 1.0),
    4: (-1.0, 1.0, 0.0, 1.0, 0.0, 1.0),
    2: (-0.1, 0.1, 0.0, 0.1, 0.0, 1.0),
    1: (-0.1, 0.1, 0.0, 0.1, 0.0, 1.0),
}
_SCALE_FEA = {8: 4, 4: 3, 2: 3, 1: 2}
_INTERNAL_Q = {8: 16, 4: 8, 2: 4, 1: 4}
_F_ALIGN = {8: 0, 4: 4, 2: 6, 1: 7}
_BETA_QU = {8: 255, 4: 15, 2: 2, 1: 1}


def layer_constants(w_qbits: int) -> dict:
    """(s_o, s, z) for weights / adjacency / features and the derived per-layer registers, as
    init_SGRACE computes them for `config.w_qbits` (layer 1 and layer 2 tables are identical in
    the shipped driver)."""
    if w_qbits not in _CLIP:
        raise ValueError("w_qbits must be 8, 4, 2 or 1")
    w_min, w_max, a_min, a_max, f_min, f_max = _CLIP[w_qbits]
    w_s_o, w_s, w_z = generate_quantization_qbits_constants(w_min, w_max, w_qbits)
    a_s_o, a_s, a_z = generate_quantization_uqbits_constants(a_min, a_max, w_qbits)
    f_s_o, f_s, f_z = generate_quantization_uqbits_constants(f_min, f_max, w_qbits)
    deq_o = w_s_o * f_s_o * a_s_o * pow(2, 1)           # sgrace.py:1681 and 1702/1755/1769/1781
    return dict(w_qbits=w_qbits, w_s_o=w_s_o, w_s=w_s, w_z=w_z, a_s_o=a_s_o, a_s=a_s, a_z=a_z,
                f_s_o=f_s_o, f_s=f_s, f_z=f_z, deq_o=deq_o, scale_fea=_SCALE_FEA[w_qbits],
                internal_quantization=_INTERNAL_Q[w_qbits], f_align=_F_ALIGN[w_qbits],
                beta_qu=_BETA_QU[w_qbits])


def float_bits(x) -> int:
    """IEEE-754 bits of float32(x) as the int the driver writes into a 32-bit register
    (sgrace.py:336: np.asarray(v, dtype=np.float32).view(np.int32).item())."""
    return int(np.asarray(x, dtype=np.float32).view(np.int32).item())
