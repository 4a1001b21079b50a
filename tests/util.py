"""helpers shared by the parity tests"""
import numpy as np


def owner_and_shift(j, nlocal, src, shift):
    """resolve list index j (owned or ghost, possibly ghost-of-ghost in the oracle's staged ghosts) to
    (owner host index, shift code)"""
    j = np.asarray(j, dtype=np.int64)
    own = j.copy()
    sh = np.zeros((len(j), 3), np.int64)
    g = own >= nlocal
    if g.any():
        gi = own[g] - nlocal
        sh[g] = shift[gi]
        o = src[gi].astype(np.int64)
        # staged ghosts: src may itself be a ghost
        for _ in range(4):
            gg = o >= nlocal
            if not gg.any():
                break
            o[gg] = src[o[gg] - nlocal]
        own[g] = o
    code = (sh[:, 0] + 1) + 3 * (sh[:, 1] + 1) + 9 * (sh[:, 2] + 1)
    return own, code, sh


def pair_keys(nlocal, numneigh, entries, src, shift, symmetrize=False):
    """canonical int64 key per list entry: (i, owner(j), image)"""
    i = np.repeat(np.arange(nlocal, dtype=np.int64), numneigh)
    j = (np.asarray(entries, dtype=np.int64) & 0x3FFFFFFF)
    own, code, sh = owner_and_shift(j, nlocal, src, shift)
    keys = (i * nlocal + own) * 27 + code
    if symmetrize:
        mcode = (1 - sh[:, 0]) + 3 * (1 - sh[:, 1]) + 9 * (1 - sh[:, 2])
        keys = np.concatenate([keys, (own * nlocal + i) * 27 + mcode])
    return np.sort(keys)


def rel_force_err(f, fref):
    """max |df| over atoms relative to the largest force component of the reference"""
    scale = np.abs(fref).max()
    return float(np.abs(f - fref).max() / scale)
