"""TEST INFRASTRUCTURE — decode the device program blob (csrc/rrx_compile.h ProgHeader) in Python."""
import numpy as np

FIELDS = ("magic P NW flags n_terms n_bicl use_rows pad0 n_rterms n_rbicl use_rrows pad1 off_first off_accept off_all "
          "off_terms off_bicl off_rowsrc off_rterms off_rbicl off_rrowsrc off_B off_follow off_rfollow off_rows off_rrows "
          "total n_raw rwc nullable static_depth n_flat off_rfull").split()


def _int(words):
    return int.from_bytes(np.asarray(words, dtype=np.uint32).tobytes(), "little")


def decode_blob(b):
    h = dict(zip(FIELDS, [int(x) for x in b[:len(FIELDS)]]))
    NW, P = h["NW"], h["P"]
    g = lambda off: _int(b[off:off + NW])
    prog = dict(h)
    prog["first"], prog["accept"], prog["all"] = g(h["off_first"]), g(h["off_accept"]), g(h["off_all"])
    for pre, ot, ob, orow, orows, nt, nb, ur in (("", "off_terms", "off_bicl", "off_rowsrc", "off_rows", "n_terms", "n_bicl", "use_rows"),
                                                 ("r", "off_rterms", "off_rbicl", "off_rrowsrc", "off_rrows", "n_rterms", "n_rbicl", "use_rrows")):
        terms = []
        for t in range(h[nt]):
            o = h[ot] + t * (1 + NW)
            terms.append((int(b[o]) >> 8, int(b[o]) & 255, g(o + 1)))
        bic = []
        for t in range(h[nb]):
            o = h[ob] + t * 2 * NW
            bic.append((g(o), g(o + NW)))
        prog[pre + "terms"], prog[pre + "bicl"] = terms, bic
        prog[pre + "rowsrc"] = g(h[orow]) if h[ur] else 0
        prog[pre + "rows"] = [g(h[orows] + k * NW) for k in range(P)]
    prog["follow"] = [g(h["off_follow"] + k * NW) for k in range(P)]
    prog["rfollow"] = [g(h["off_rfollow"] + k * NW) for k in range(P)]
    prog["B"] = [g(h["off_B"] + c * NW) for c in range(256)]
    return prog


def apply_follow(prog, S, rev=False):
    """What the kernels' follow_terms computes (rrx_tile.cuh), on Python ints."""
    NW = prog["NW"]
    bits = 32 * NW
    full = (1 << bits) - 1
    pre = "r" if rev else ""
    f = 0
    for kind, r, mask in prog[pre + "terms"]:
        if NW == 1 or kind == 0:
            x = ((S << r) | (S >> (32 - r))) & 0xFFFFFFFF if r else S
        elif kind == 3:
            x = S
        elif kind == 1:
            x = (S << r) & full
        else:
            x = S >> r
        f |= x & mask
    for src, dst in prog[pre + "bicl"]:
        if S & src:
            f |= dst
    v = S & prog[pre + "rowsrc"]
    k = 0
    while v:
        if v & 1:
            f |= prog[pre + "rows"][k]
        v >>= 1
        k += 1
    return f
