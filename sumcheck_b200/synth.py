"""Deterministic synthetic inputs for benches and smoke runs (the generator is ours, not the reference's).

Counter-based SplitMix64: element e, attempt k (< 8), limb i is mix(seed + GAMMA * ((e*8 + k)*4 + i + 1)).  The first
attempt whose value (top bit cleared) is < p is used directly as the Montgomery representation — uniform over Fr.
If all 8 attempts fail (p ~ 6e-9 per element) the top two bits of attempt 7 are cleared.  Vectorised with numpy so a
2^24-element table takes about a second; oracle/sumcheck_oracle.c has the scalar twin used to cross-check it.
"""
import numpy as np

GAMMA = np.uint64(0x9E3779B97F4A7C15)
P_LIMBS = (0xFFFFFFFF00000001, 0x53BDA402FFFE5BFE, 0x3339D80809A1D805, 0x73EDA753299D7D48)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _geq_p(t):
    """t: [n,4] uint64 -> bool[n], value >= p"""
    ge = np.ones(t.shape[0], dtype=bool)    # equal so far -> >=
    decided = np.zeros(t.shape[0], dtype=bool)
    for i in (3, 2, 1, 0):
        pi = np.uint64(P_LIMBS[i])
        gt, lt = t[:, i] > pi, t[:, i] < pi
        ge = np.where(~decided & lt, False, ge)
        decided |= gt | lt
    return ge


def synth_table(n_elems, seed, out=None, chunk=1 << 20):
    """[n_elems, 4] uint64 Montgomery limbs; `out` may be a preallocated (e.g. pinned) array."""
    if out is None:
        out = np.empty((n_elems, 4), dtype=np.uint64)
    seed = np.uint64(seed)
    with np.errstate(over="ignore"):
        for s in range(0, n_elems, chunk):
            e = np.arange(s, min(s + chunk, n_elems), dtype=np.uint64)
            todo = np.arange(e.shape[0])
            for k in range(8):
                ctr = (e[todo] * np.uint64(8) + np.uint64(k)) * np.uint64(4)
                t = np.stack([_mix(seed + GAMMA * (ctr + np.uint64(i + 1))) for i in range(4)], axis=1)
                t[:, 3] &= np.uint64(0x7FFFFFFFFFFFFFFF)
                bad = _geq_p(t)
                if k == 7:
                    t[bad, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
                    bad[:] = False
                out[s + todo[~bad]] = t[~bad]
                todo = todo[bad]
                if todo.size == 0:
                    break
    return out


def synth_table_fast(n_elems, seed, out=None, first=0):
    """Same stream through the C helper sc_synth_table_at (scalar C, ~10x faster than the numpy twin); `first` selects
    the slice [first, first + n_elems) of the table, e.g. one rank's shard."""
    import ctypes as C

    from . import capi
    if out is None:
        out = np.empty((n_elems, 4), dtype=np.uint64)
    assert out.dtype == np.uint64 and out.flags["C_CONTIGUOUS"] and out.shape == (n_elems, 4)
    capi.lib().sc_synth_table_at(out.ctypes.data_as(capi.U64P), C.c_uint64(first), C.c_uint64(n_elems), C.c_uint64(seed))
    return out
