"""Dual-quaternion helpers with the reference's names and semantics (nnutils/dual_quat.py), CUDA-backed.

Differences from the reference, on purpose: no host-synchronising singularity asserts
(dual_quat.py:11, :61 call torch.any(...) on the device and block the stream).
"""
from .ops import DqMulFn, DqUnaryFn, DQ_QCONJ, DQ_CCONJ, DQ_NORMALIZE, DQ_INVERSE, Q_NORMALIZE


def q_normalize(q):
    """dual_quat.py:4-12."""
    assert q.shape[-1] == 4
    return DqUnaryFn.apply(q, Q_NORMALIZE)


def q_mul(q1, q2):
    """dual_quat.py:14-31: Hamilton product q1 (x) q2, real part first."""
    assert q1.shape[-1] == 4 and q2.shape[-1] == 4
    return DqMulFn.apply(q1, q2, 4)


def dq_mul(dq1, dq2):
    """dual_quat.py:33-49."""
    assert dq1.shape[-1] == 8 and dq2.shape[-1] == 8
    return DqMulFn.apply(dq1, dq2, 8)


def dq_normalize(dq):
    """dual_quat.py:51-62."""
    assert dq.shape[-1] == 8
    return DqUnaryFn.apply(dq, DQ_NORMALIZE)


def dq_quaternion_conjugate(dq):
    """dual_quat.py:65-74."""
    assert dq.shape[-1] == 8
    return DqUnaryFn.apply(dq, DQ_QCONJ)


def dq_combined_conjugate(dq):
    """dual_quat.py:76-85."""
    assert dq.shape[-1] == 8
    return DqUnaryFn.apply(dq, DQ_CCONJ)


def dq_inverse(dq):
    """dual_quat.py:87-93."""
    assert dq.shape[-1] == 8
    return DqUnaryFn.apply(dq, DQ_INVERSE)
