import numpy as np


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    nb = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0)


def rand_matrix(dtype, m, n, seed, scale=None):
    """Philox matrix, identical on host (oracle/philox.py) and device."""
    from oracle.philox import philox_matrix, IH4
    scale = 1.0 / np.sqrt(m) if scale is None else scale
    return philox_matrix(dtype, m, n, seed, IH4, scale), scale


def rand_vector(dtype, n, seed, stream=1):
    from oracle.philox import philox_vector, IH4
    return philox_vector(dtype, n, seed, stream, IH4, 1.0)


def sparse_truth(dtype, n, seed, every=29):
    from oracle.philox import philox_vector, UNIFORM01
    x = np.zeros(n, dtype)
    v = philox_vector(dtype, n, seed, 2, UNIFORM01)
    x[::every] = v[::every]
    return x
