"""dict <-> flat (n_particles x n_params) layout of the particles.

Mirrors stein/utilities/converters.py:4-89 of the reference: variables are
ordered by the lexicographic sort of their `.name` (converters.py:40), each value
is flattened row-major to (n, prod(shape[1:])) (:47-48) and the blocks are
concatenated along columns (:51).  This fixes the column order of the matrices
the CUDA kernels see.  Pure host glue, float64 like the
reference.
"""
import numpy as np


def convert_dictionary_to_array(dictionary):
    """{variable: (n, *shape) array} -> ((n, n_params) float64 array, {variable: (start, stop)})."""
    keys = sorted(dictionary.keys(), key=lambda v: v.name)
    blocks, access_indices, start = [], {}, 0
    n_particles = next(iter(dictionary.values())).shape[0]
    for v in keys:
        block = np.asarray(dictionary[v]).reshape(n_particles, -1)
        access_indices[v] = (start, start + block.shape[1])
        start += block.shape[1]
        blocks.append(block)
    array = np.zeros((n_particles, start))
    if blocks:
        array[:, :] = np.concatenate(blocks, axis=1)
    return array, access_indices


def convert_array_to_dictionary(array, access_indices):
    """Inverse of convert_dictionary_to_array (converters.py:58-89)."""
    n_particles = array.shape[0]
    return {
        v: np.reshape(array[:, start:stop], [n_particles] + v.get_shape().as_list())
        for v, (start, stop) in access_indices.items()
    }
