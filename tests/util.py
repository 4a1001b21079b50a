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


def water_special_bits(nlocal, numneigh, entries, src, mol, type_):
    """special-bond bits (SBBITS = 30) for a neighbour list of a 3-site water system: an entry whose partner sits in the
    same molecule is a 1-2 pair (O-H, bits 01) or a 1-3 pair (H-H, bits 10) — what stock Neighbor puts into the list from
    the special arrays built off the bond topology of examples/data.spce.  `src` resolves ghost partners to their owners."""
    i = np.repeat(np.arange(nlocal, dtype=np.int64), numneigh)
    j = np.asarray(entries, dtype=np.int64) & 0x3FFFFFFF
    own = j.copy()
    g = own >= nlocal
    for _ in range(4):
        if not g.any():
            break
        own[g] = src[own[g] - nlocal]
        g = own >= nlocal
    same = mol[i] == mol[own]
    hh = same & (type_[i] == 2) & (type_[own] == 2)
    bits = np.where(same, np.where(hh, 2, 1), 0).astype(np.int64)
    return (j | (bits << 30)).astype(np.uint32).view(np.int32)
